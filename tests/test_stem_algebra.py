"""CPU checks of the algebra behind csrc/stem_kernels.cuh (student conv1 + train-mode BN + ReLU + max-pool by
linearity in the one-channel input), against the oracle operators.

The numpy functions below restate, index for index, what the device kernels accumulate (row-pair products per
boundary class -> assembled patch autocorrelation R and patch sum S -> BN statistics; pooled-resolution BN
reductions; dW = A*G1 - D*(R w + b S) + E*S) so that the identities and the index bookkeeping are pinned in fp64
without a GPU; tests/test_gpu_programs.py then checks the kernels themselves through the training step."""
import numpy as np
import pytest

from oracle import mcn_ops as M
from mcncrossmodalemotions_b200.programs import student_conv1_to_s2d, student_conv1_from_s2d


def s2d(spec, pad_t=1, pad_l=1):
    """hbm_kernels.cuh spec_s2d_from_hwcn_kernel: H x W x 1 x N -> [N][HP][OW][16]."""
    H, W, _, N = spec.shape
    OH, OW = (H + 2 - 7) // 2 + 1, (W + 2 - 7) // 2 + 1
    HP = OH + 3
    X = np.zeros((N, HP, OW, 16))
    for hp in range(HP):
        for dr in range(2):
            h = 2 * hp + dr - pad_t
            if not 0 <= h < H:
                continue
            for s in range(7):
                w = 2 * np.arange(OW) + s - pad_l
                ok = (w >= 0) & (w < W)
                X[:, hp, ok, dr * 8 + s] = spec[h, w[ok], 0, :].T
    return X, OH, OW


def autocorr_bins(X, OH):
    """stem_autocorr_kernel: acc[cls][(c'*4+d)*16+c] and column sums, classes {all, h1=0,1,2, h1=OH,OH+1,OH+2}."""
    N, HP, OW, _ = X.shape
    acc = np.zeros((7, 16 * 4 * 16 + 16))
    Xz = np.concatenate([X, np.zeros((N, 3, OW, 16))], axis=1)
    for h1 in range(HP):
        prod = np.zeros((16, 4, 16))
        for d in range(4):
            prod[:, d, :] = np.einsum("nwa,nwb->ab", Xz[:, h1], Xz[:, h1 + d])
        cs = Xz[:, h1].sum(axis=(0, 1))
        row = np.concatenate([prod.reshape(-1), cs])
        acc[0] += row
        cls = 1 + h1 if h1 < 3 else (4 + h1 - OH if h1 >= OH else 0)
        if cls:
            acc[cls] += row
    return acc


def assemble(acc):
    """stem_assemble_kernel."""
    R = np.zeros((64, 64))
    S = np.zeros(64)
    for tp in range(64):
        for tq in range(64):
            jp, cp, jq, cq = tp >> 4, tp & 15, tq >> 4, tq & 15
            if jq < jp:
                jp, jq, cp, cq = jq, jp, cq, cp
            e = (cp * 4 + (jq - jp)) * 16 + cq
            v = acc[0, e]
            for h in range(jp):
                v -= acc[1 + h, e]
            for h in range(jp, 3):
                v -= acc[4 + h, e]
            R[tp, tq] = v
    for i in range(64):
        j, c = i >> 4, i & 15
        v = acc[0, 1024 + c]
        for h in range(j):
            v -= acc[1 + h, 1024 + c]
        for h in range(j, 3):
            v -= acc[4 + h, 1024 + c]
        S[i] = v
    return R, S


def patches(X, OH):
    N, HP, OW, _ = X.shape
    return np.concatenate([X[:, j : j + OH] for j in range(4)], axis=-1).reshape(N * OH * OW, 64)


@pytest.fixture(scope="module")
def case():
    rng = np.random.default_rng(7)
    H, W, N, K = 40, 38, 3, 16
    spec = rng.standard_normal((H, W, 1, N))
    f = (rng.standard_normal((7, 7, 1, K)) * 0.2).astype(np.float32).astype(np.float64)   # the s2d helpers keep fp32
    b = rng.standard_normal(K) * 0.3
    g = rng.uniform(0.5, 1.5, K) * np.where(np.arange(K) % 5 == 0, -1.0, 1.0)   # a few negative scales
    beta = rng.standard_normal(K) * 0.2
    X, OH, OW = s2d(spec)
    return dict(spec=spec, f=f, b=b, g=g, beta=beta, X=X, OH=OH, OW=OW, K=K, N=N)


def test_s2d_patches_reproduce_conv1(case):
    X, OH, OW, K, N = case["X"], case["OH"], case["OW"], case["K"], case["N"]
    w = student_conv1_to_s2d(case["f"]).reshape(K, 64)
    y = patches(X, OH) @ w.T + case["b"]
    ref = M.vl_nnconv(case["spec"], case["f"], case["b"], pad=1, stride=2)
    assert ref.shape == (OH, OW, K, N)
    got = y.reshape(N, OH, OW, K).transpose(1, 2, 3, 0)
    assert np.abs(got - ref).max() < 1e-10


def test_assembled_autocorrelation_equals_brute_force(case):
    X, OH = case["X"], case["OH"]
    R, S = assemble(autocorr_bins(X, OH))
    xs = patches(X, OH)
    assert np.abs(R - xs.T @ xs).max() < 1e-8
    assert np.abs(S - xs.sum(0)).max() < 1e-9


def test_bn_statistics_from_autocorrelation(case):
    X, OH, K = case["X"], case["OH"], case["K"]
    R, S = assemble(autocorr_bins(X, OH))
    w = student_conv1_to_s2d(case["f"]).reshape(K, 64)
    P = patches(X, OH).shape[0]
    m0 = w @ S / P
    var = np.einsum("kt,tu,ku->k", w, R, w) / P - m0 ** 2
    mu, sigma = m0 + case["b"], np.sqrt(var + 1e-5)
    x = M.vl_nnconv(case["spec"], case["f"], case["b"], pad=1, stride=2)
    _, moments = M.vl_nnbnorm(x, case["g"], case["beta"], epsilon=1e-5)
    assert np.abs(mu - moments[:, 0]).max() < 1e-10
    assert np.abs(sigma - moments[:, 1]).max() < 1e-10


def test_stem_backward_by_linearity_matches_oracle_chain(case):
    spec, f, b, g, beta = case["spec"], case["f"], case["b"], case["g"], case["beta"]
    X, OH, OW, K, N = case["X"], case["OH"], case["OW"], case["K"], case["N"]
    rng = np.random.default_rng(11)
    # oracle chain: conv1 -> BN(train) -> ReLU -> max pool 3x3/2, random gradient at the pooled output
    x = M.vl_nnconv(spec, f, b, pad=1, stride=2)
    y, moments = M.vl_nnbnorm(x, g, beta, epsilon=1e-5)
    z = M.vl_nnrelu(y)
    pooled = M.vl_nnpool(z, (3, 3), stride=2, method="max")
    dpool = rng.standard_normal(pooled.shape)
    dz = M.vl_nnrelu(y, M.vl_nnpool(z, (3, 3), dpool, stride=2, method="max"))
    dx, dg_ref, dbeta_ref, _ = M.vl_nnbnorm(x, g, beta, dz, epsilon=1e-5)
    _, df_ref, db_ref = M.vl_nnconv(spec, f, b, dx, pad=1, stride=2)

    # device formulation.  (1) pooling forward records the raw winner of every window
    mu, sigma = moments[:, 0], moments[:, 1]
    a = g / sigma
    bb = beta - a * mu
    POH, POW = pooled.shape[:2]
    xw = np.zeros_like(pooled)
    for oh in range(POH):
        for ow in range(POW):
            win_x = x[2 * oh : 2 * oh + 3, 2 * ow : 2 * ow + 3].reshape(9, K, N)
            win_z = z[2 * oh : 2 * oh + 3, 2 * ow : 2 * ow + 3].reshape(9, K, N)
            idx = win_z.argmax(0)
            xw[oh, ow] = np.take_along_axis(win_x, idx[None], 0)[0]
    # (2) stem_pool_bn_reduce_kernel: mask + reductions at the pooled resolution
    alive = (a[None, None, :, None] * xw + bb[None, None, :, None]) > 0
    gm = dpool * alive
    s_dz = gm.sum(axis=(0, 1, 3))
    s_dzxhat = (gm * (xw - mu[None, None, :, None]) / sigma[None, None, :, None]).sum(axis=(0, 1, 3))
    assert np.abs(s_dz - dbeta_ref.ravel()).max() < 1e-9
    assert np.abs(s_dzxhat - dg_ref.ravel()).max() < 1e-9
    # (3) dz at the conv resolution from the masked pooled gradient (xemo_op_maxpool_bwd), G1 by the wgrad kernel
    dz_dev = M.vl_nnpool(z, (3, 3), gm, stride=2, method="max")
    assert np.abs(dz_dev - dz).max() < 1e-12
    xs = patches(X, OH)
    G1 = dz_dev.transpose(3, 0, 1, 2).reshape(-1, K).T @ xs          # [K][64]
    # (4) stem_wgrad_finalize_kernel
    R, S = assemble(autocorr_bins(X, OH))
    w = student_conv1_to_s2d(f).reshape(K, 64)
    P = xs.shape[0]
    D = a * s_dzxhat / (P * sigma)
    E = mu * D - a * s_dz / P
    dW = a[:, None] * G1 - D[:, None] * (w @ R + b[:, None] * S[None]) + E[:, None] * S[None]
    t = np.arange(64)
    dW[:, ((t & 15) & 7 == 7) | ((t >> 4 == 3) & ((t & 15) >= 8))] = 0
    got = student_conv1_from_s2d(dW.reshape(K, 4, 1, 16))                 # (casts to fp32)
    assert np.abs(got - df_ref).max() < 1e-6 * np.abs(df_ref).max()
    dW7 = np.stack([dW.reshape(K, 4, 2, 8)[:, r // 2, r % 2, :7] for r in range(7)], 0).transpose(0, 2, 1)[:, :, None, :]
    assert np.abs(dW7 - df_ref).max() < 1e-9 * max(1.0, np.abs(df_ref).max())   # the same comparison in fp64
    assert np.abs(db_ref).max() < 1e-8        # the bias ahead of train-mode BN has a zero gradient


def test_pixel_pair_form_is_the_same_convolution(case):
    """programs.pair_filter: the [N][HP][OW/2][32] view of the s2d tensor convolved with the block-diagonal
    [2K][4][1][32] filter, read back as [N][OH][OW][K], equals the 16-channel form (same bytes in memory)."""
    from mcncrossmodalemotions_b200.programs import pair_filter

    K, N = case["K"], case["N"]
    X, OH, OW = s2d(np.random.default_rng(3).standard_normal((40, 40, 1, N)))    # an even conv1 output width
    assert OW % 2 == 0
    w = student_conv1_to_s2d(case["f"]).astype(np.float64)                      # [K][4][1][16]
    y = patches(X, OH) @ w.reshape(K, 64).T                                       # [N*OH*OW][K]
    Xp = X.reshape(N, X.shape[1], OW // 2, 32)
    xs2 = np.concatenate([Xp[:, j : j + OH] for j in range(4)], axis=-1).reshape(N * OH * (OW // 2), 128)
    w2 = pair_filter(w)                                                           # [2K][4][1][32]
    assert w2.shape == (2 * K, 4, 1, 32)
    y2 = xs2 @ w2.reshape(2 * K, 128).T                                           # [N*OH*OW/2][2K]
    assert np.abs(y2.reshape(N * OH * OW, K) - y).max() < 1e-12
