"""CPU checks behind the (experimental, default-off) SE-by-linearity path of the teacher: the squeeze of the expand
convolution's output equals an affine function of the spatial mean of its input, and the per-tile image-slot bookkeeping of
the fused expand+excite epilogue (csrc/conv_fprop.cuh, kNC instantiation) addresses the right image for every GEMM row."""
import numpy as np

from oracle import mcn_ops as M


def test_squeeze_is_linear_in_the_mean_of_the_bottleneck_output():
    rng = np.random.default_rng(0)
    H = W = 7
    Cm, C, N = 16, 64, 3
    t2 = rng.standard_normal((H, W, Cm, N))
    w3 = rng.standard_normal((1, 1, Cm, C)) * 0.3
    g, beta = rng.uniform(0.5, 1.5, C), rng.standard_normal(C) * 0.1
    mom = np.stack([rng.standard_normal(C) * 0.1, rng.uniform(0.5, 1.5, C)], 1)          # [mu sigma]
    u, _ = M.vl_nnbnorm(M.vl_nnconv(t2, w3, None), g, beta, epsilon=1e-5, moments=mom)   # test-mode BN of the expand conv
    s = M.vl_nnglobalpool(u)                                                             # 1 x 1 x C x N
    a3 = g / mom[:, 1]
    b3 = beta - a3 * mom[:, 0]
    m2 = t2.mean(axis=(0, 1))                                                            # Cm x N
    s_lin = a3[:, None] * (w3[0, 0].T @ m2) + b3[:, None]
    assert np.abs(s_lin - s[0, 0]).max() < 1e-10
    # ... and the excite folds into the expand convolution's epilogue: relu(gate*(a3*acc + b3) + shortcut)
    gate = 1.0 / (1.0 + np.exp(-rng.standard_normal((C, N))))
    sc = rng.standard_normal(u.shape)
    y = M.vl_nnrelu(M.vl_nnaxpy(gate.reshape(1, 1, C, N), u, sc))
    acc = M.vl_nnconv(t2, w3, None)
    y_fused = np.maximum((gate * a3[:, None]).reshape(1, 1, C, N) * acc + (gate * b3[:, None]).reshape(1, 1, C, N) + sc, 0)
    assert np.abs(y_fused - y).max() < 1e-10


def test_epilogue_image_slots_cover_every_row_of_every_tile():
    """conv_fprop_kernel<64, true>: img0 = m0 / hw, nimg = last_row / hw - img0 + 1 <= 4 slots, slot(row) = row / hw - img0."""
    for hw in (49, 196, 784, 3136):
        for n in (1, 2, 5, 8, 9):
            Mrows = n * hw
            for m0 in range(0, Mrows, 128):
                img0 = m0 // hw
                last_row = min(m0 + 128, Mrows) - 1
                nimg = last_row // hw - img0 + 1
                assert 1 <= nimg <= 4, (hw, n, m0, nimg)
                assert (128 + hw - 2) // hw + 1 <= 4          # the host-side admission rule for this hw
                for row in range(m0, m0 + 128):
                    slot = min(row, Mrows - 1) // hw - img0
                    assert 0 <= slot < nimg
                    if row < Mrows:
                        assert img0 + slot == row // hw
