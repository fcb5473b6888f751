"""Parity of the MatConvNet-boundary operators (xemo_vl_* through the C ABI) against the CPU oracle.

Tolerances (north_star): 1e-3 relative (max|a-b| <= tol * max|ref|) for operators that run on the
fp16-operand tensor-core path (vl_nnconv); 1e-5 for the fp32 operators; pooling arg-max indices and
class-error counts bit-exact."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu

CONV_TOL = 1e-3
F32_TOL = 2e-5


@pytest.fixture(scope="module")
def vl():
    from mcncrossmodalemotions_b200 import vl_nn

    return vl_nn


@pytest.fixture(scope="module")
def M():
    from oracle import mcn_ops

    return mcn_ops


CONV_CASES = [
    # H, W, C, N, FH, FW, K, pad, stride, bias
    (12, 10, 3, 2, 3, 3, 8, 1, 1, True),
    (16, 16, 64, 2, 1, 1, 64, 0, 1, False),
    (14, 14, 64, 3, 3, 3, 64, 1, 1, True),
    (28, 28, 32, 2, 1, 1, 48, 0, 2, False),               # strided 1x1 (teacher stage transitions)
    (30, 21, 96, 2, 5, 5, 256, 1, 2, True),               # student conv2
    (15, 15, 16, 2, 3, 3, 24, (0, 1, 0, 1), 2, True),     # asymmetric pad
    (9, 8, 256, 3, 9, 1, 512, 0, 1, True),                # fc6-like
    (1, 1, 1024, 5, 1, 1, 8, 0, 1, True),                 # fc8 (K = 8 -> padded to 16)
    (40, 30, 1, 2, 7, 7, 96, 1, 2, True),                 # student conv1 (C = 1)
    (32, 32, 3, 2, 7, 7, 64, 3, 2, False),                # teacher conv1
    (17, 13, 24, 1, 4, 2, 40, (2, 1, 0, 1), (3, 2), True),
    (34, 30, 64, 2, 3, 3, 384, 1, 1, True),               # 3 x 128 kout: filter-gradient items own two kout tiles (mt = 2), odd count
    (40, 36, 96, 2, 5, 5, 256, 1, 2, False),              # student conv2 with enough pixels for the 64-pixel wgrad stages
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_vl_nnconv_forward_backward(vl, M, case):
    H, W, Cc, N, FH, FW, K, pad, stride, bias = case
    rng = np.random.default_rng(hash(case[:7]) % 2**32)
    x = rng.standard_normal((H, W, Cc, N)).astype(np.float32)
    f = (rng.standard_normal((FH, FW, Cc, K)) / np.sqrt(FH * FW * Cc)).astype(np.float32)
    b = rng.standard_normal(K).astype(np.float32) if bias else None
    y = vl.vl_nnconv(x, f, b, pad=pad, stride=stride)
    yr = M.vl_nnconv(x.astype(np.float64), f.astype(np.float64), b, pad=pad, stride=stride)
    assert y.shape == yr.shape
    assert rel_err(y, yr) < CONV_TOL
    dy = rng.standard_normal(yr.shape).astype(np.float32)
    dx, df, db = vl.vl_nnconv(x, f, b, dy, pad=pad, stride=stride)
    dxr, dfr, dbr = M.vl_nnconv(x.astype(np.float64), f.astype(np.float64), b, dy.astype(np.float64), pad=pad, stride=stride)
    assert rel_err(dx, dxr) < CONV_TOL
    assert rel_err(df, dfr) < CONV_TOL
    if bias:
        assert rel_err(db, dbr) < F32_TOL
    else:
        assert db is None


POOL_CASES = [
    # H, W, C, N, pool, pad, stride
    (14, 12, 8, 2, (3, 3), 0, 2),
    (15, 15, 64, 2, (3, 3), (0, 1, 0, 1), 2),    # teacher pool1
    (30, 17, 16, 2, (5, 3), 0, (3, 2)),          # student pool5
    (7, 7, 32, 3, (7, 7), 0, 1),                 # teacher pool5 window
    (1, 8, 24, 2, (1, 8), 0, 1),                 # student pool6
    (9, 11, 3, 2, (2, 3), (1, 0, 1, 1), (2, 1)),
]


@pytest.mark.parametrize("case", POOL_CASES)
@pytest.mark.parametrize("method", ["max", "avg"])
def test_vl_nnpool(vl, M, case, method):
    H, W, Cc, N, pool, pad, stride = case
    rng = np.random.default_rng(7)
    # quantised, ReLU-like data: many exact ties and all-zero windows (the arg-max tie-break case)
    x = np.maximum(np.round(rng.standard_normal((H, W, Cc, N)) * 2) / 2, 0).astype(np.float32)
    if method == "max":
        y, idx = vl.vl_nnpool(x, pool, pad=pad, stride=stride, method="max", return_index=True)
        yr, idxr = M.vl_nnpool(x, pool, pad=pad, stride=stride, method="max", return_index=True)
        assert np.array_equal(idx, idxr), "pooling arg-max indices must be bit-exact"
        assert np.array_equal(y, yr)
    else:
        y = vl.vl_nnpool(x, pool, pad=pad, stride=stride, method="avg")
        yr = M.vl_nnpool(x.astype(np.float64), pool, pad=pad, stride=stride, method="avg")
        assert rel_err(y, yr) < F32_TOL
    dy = rng.standard_normal(yr.shape).astype(np.float32)
    dx = vl.vl_nnpool(x, pool, dy, pad=pad, stride=stride, method=method)
    dxr = M.vl_nnpool(x.astype(np.float64), pool, dy.astype(np.float64), pad=pad, stride=stride, method=method)
    assert rel_err(dx, dxr) < F32_TOL


@pytest.mark.parametrize("shape", [(6, 5, 8, 3), (14, 14, 96, 2), (1, 1, 40, 16), (5, 3, 3, 4)])
@pytest.mark.parametrize("test_mode", [False, True])
def test_vl_nnbnorm(vl, M, shape, test_mode):
    rng = np.random.default_rng(11)
    Cc = shape[2]
    x = (rng.standard_normal(shape) * 2 + 0.5).astype(np.float32)
    g = rng.uniform(0.5, 1.5, Cc).astype(np.float32)
    b = rng.standard_normal(Cc).astype(np.float32)
    mom = np.stack([rng.standard_normal(Cc) * 0.1, rng.uniform(0.5, 1.5, Cc)], 1).astype(np.float32) if test_mode else None
    y, mo = vl.vl_nnbnorm(x, g, b, epsilon=1e-5, moments=mom)
    yr, mor = M.vl_nnbnorm(x.astype(np.float64), g, b, epsilon=1e-5, moments=mom)
    assert rel_err(y, yr) < F32_TOL
    assert rel_err(mo, mor) < F32_TOL
    dy = rng.standard_normal(shape).astype(np.float32)
    dx, dg, db, _ = vl.vl_nnbnorm(x, g, b, dy, epsilon=1e-5, moments=mom)
    dxr, dgr, dbr, _ = M.vl_nnbnorm(x.astype(np.float64), g, b, dy.astype(np.float64), epsilon=1e-5, moments=mom)
    assert rel_err(dx, dxr) < 1e-4
    assert rel_err(dg, dgr) < 1e-4
    assert rel_err(db, dbr) < 1e-4


def test_elementwise_and_se_ops(vl, M):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((5, 4, 16, 3)).astype(np.float32)
    dy = rng.standard_normal(x.shape).astype(np.float32)
    assert np.array_equal(vl.vl_nnrelu(x), M.vl_nnrelu(x))
    assert np.array_equal(vl.vl_nnrelu(x, dy), M.vl_nnrelu(x, dy))
    assert rel_err(vl.vl_nnrelu(x, leak=0.1), M.vl_nnrelu(x, leak=0.1)) < F32_TOL
    assert rel_err(vl.vl_nnsigmoid(x), M.vl_nnsigmoid(x.astype(np.float64))) < F32_TOL
    assert rel_err(vl.vl_nnsigmoid(x, dy), M.vl_nnsigmoid(x.astype(np.float64), dy)) < F32_TOL
    assert rel_err(vl.vl_nnglobalpool(x), M.vl_nnglobalpool(x.astype(np.float64))) < F32_TOL
    dz = rng.standard_normal((1, 1, 16, 3)).astype(np.float32)
    assert rel_err(vl.vl_nnglobalpool(x, dz), M.vl_nnglobalpool(x.astype(np.float64), dz)) < F32_TOL
    a = rng.uniform(0, 1, (1, 1, 16, 3)).astype(np.float32)
    y = rng.standard_normal(x.shape).astype(np.float32)
    assert rel_err(vl.vl_nnaxpy(a, x, y), M.vl_nnaxpy(a, x.astype(np.float64), y)) < F32_TOL
    assert rel_err(vl.vl_nnsoftmaxt(x, dim=3), M.vl_nnsoftmaxt(x.astype(np.float64), dim=3)) < F32_TOL


@pytest.mark.parametrize("n", [1, 37, 256])
@pytest.mark.parametrize("logit_targets", [True, False])
def test_vl_nnsoftmaxceloss(vl, M, n, logit_targets):
    rng = np.random.default_rng(n)
    x = (3 * rng.standard_normal((1, 1, 8, n))).astype(np.float32)
    t = (3 * rng.standard_normal((1, 1, 8, n))).astype(np.float32)
    if not logit_targets:
        t = M.vl_nnsoftmaxt(t.astype(np.float64), dim=3).astype(np.float32)
    w = rng.uniform(0.5, 2, n).astype(np.float32)
    for iw in (None, w):
        y = vl.vl_nnsoftmaxceloss(x, t, temperature=2, logitTargets=logit_targets, instanceWeights=iw)
        yr = M.vl_nnsoftmaxceloss(x.astype(np.float64), t.astype(np.float64), temperature=2, logitTargets=logit_targets,
                                  instanceWeights=iw, tol=1e-4)
        assert abs(y - yr) <= 1e-5 * max(1.0, abs(yr))
        dx = vl.vl_nnsoftmaxceloss(x, t, 1.5, temperature=2, logitTargets=logit_targets, instanceWeights=iw)
        dxr = M.vl_nnsoftmaxceloss(x.astype(np.float64), t.astype(np.float64), 1.5, temperature=2, logitTargets=logit_targets,
                                   instanceWeights=iw, tol=1e-4)
        assert rel_err(dx, dxr) < F32_TOL


def test_classerror_bit_exact(vl, M):
    rng = np.random.default_rng(5)
    x = np.round(rng.standard_normal((1, 1, 8, 300)) * 2).astype(np.float32)  # ties -> first maximum wins
    c = rng.integers(1, 9, 300)
    assert vl.vl_nnloss(x, c, loss="classerror") == M.vl_nnloss(x, c, loss="classerror")


def test_gpuarray_inputs_stay_on_device(vl, M):
    rng = np.random.default_rng(9)
    x = rng.standard_normal((14, 14, 32, 2)).astype(np.float32)
    f = (rng.standard_normal((3, 3, 32, 48)) / 17).astype(np.float32)
    y = vl.vl_nnconv(vl.gpuArray(x), vl.gpuArray(f), None, pad=1)
    assert isinstance(y, vl.GpuArray)
    assert rel_err(vl.gather(y), M.vl_nnconv(x.astype(np.float64), f.astype(np.float64), None, pad=1)) < CONV_TOL
    r = vl.vl_nnrelu(y)
    assert isinstance(r, vl.GpuArray)
    assert np.array_equal(vl.gather(r), np.maximum(vl.gather(y), 0))


def test_scratch_pool_is_kept_between_calls_and_can_be_trimmed(vl, M):
    """The boundary operators draw their scratch from the context's own stream-ordered pool (kept between calls);
    xemo_trim hands it back and the next call simply grows it again."""
    rng = np.random.default_rng(10)
    x = rng.standard_normal((20, 20, 32, 2)).astype(np.float32)
    f = (rng.standard_normal((3, 3, 32, 32)) / 17).astype(np.float32)
    ctx = vl.default_context()
    y0 = vl.vl_nnconv(x, f, None, pad=1)
    ctx.trim()
    y1 = vl.vl_nnconv(x, f, None, pad=1)
    assert np.array_equal(y0, y1)
    ctx.trim()


def test_error_behaviour(vl):
    from mcncrossmodalemotions_b200._lib import XemoError

    x = np.zeros((4, 4, 8, 1), np.float32)
    with pytest.raises(XemoError):
        vl.vl_nnconv(x, np.zeros((3, 3, 4, 8), np.float32))  # filter depth mismatch (groups unsupported)
    with pytest.raises(XemoError):
        vl.vl_nnconv(x, np.zeros((5, 5, 8, 8), np.float32))  # filter larger than the input
    with pytest.raises(ValueError):
        vl.vl_nnpool(x, 2, method="median")


@pytest.mark.parametrize("geom", [
    # H, W, C, N, FH, FW, K, pad, stride  -- several M tiles each, so that clusters of two CTAs have pairs to take
    (30, 17, 128, 3, 3, 3, 256, (1, 1, 1, 1), (1, 1)),       # conv3-like; odd number of M tiles (12): no, 1530 rows -> 12
    (29, 21, 96, 2, 5, 5, 128, (1, 1, 1, 1), (2, 2)),        # strided: the data gradient runs 4 parity classes through pairs
    (14, 14, 256, 5, 1, 1, 512, (0, 0, 0, 0), (1, 1)),       # 980 rows -> 8 M tiles, 1x1
    (9, 8, 64, 11, 9, 1, 160, (0, 0, 0, 0), (1, 1)),         # fc6-like, K = 160 -> N tile 160 (80 rows per CTA)
    (23, 23, 64, 1, 3, 3, 64, (1, 1, 1, 1), (1, 1)),         # 529 rows -> 5 M tiles: the last cluster's peer has no tile
])
def test_vl_nnconv_with_cta_pairs_is_bit_identical_to_single_ctas(geom):
    """conv_fprop_kernel<BK, false, 2> (tcgen05.mma.cta_group::2: a cluster of two CTAs executes two M tiles as one M = 256
    MMA, each staging half of the filter tile) against the single-CTA kernel -- same K order, same accumulators: the outputs
    must be bit-identical, forward and data gradient -- and against the fp64 oracle."""
    from oracle import mcn_ops as M
    from mcncrossmodalemotions_b200 import _lib, vl_nn

    H, W, C, N, FH, FW, K, pad, stride = geom
    rng = np.random.default_rng(17)
    x = rng.standard_normal((H, W, C, N)).astype(np.float32)
    f = (rng.standard_normal((FH, FW, C, K)) / np.sqrt(FH * FW * C)).astype(np.float32)
    b = rng.standard_normal(K).astype(np.float32)
    y64 = M.vl_nnconv(x.astype(np.float64), f.astype(np.float64), b.astype(np.float64), pad=pad, stride=stride)
    dy = rng.standard_normal(y64.shape).astype(np.float32)
    dx64, _, _ = M.vl_nnconv(x.astype(np.float64), f.astype(np.float64), b.astype(np.float64), dy.astype(np.float64), pad=pad, stride=stride)
    lib = _lib.load_library()
    out = {}
    try:
        for mode in (0, 2):
            assert lib.xemo_debug_set_conv_pair_mode(mode) == 0
            y = vl_nn.vl_nnconv(x, f, b, pad=pad, stride=stride)
            dx, df, db = vl_nn.vl_nnconv(x, f, b, dy, pad=pad, stride=stride)
            out[mode] = (y, dx)
    finally:
        lib.xemo_debug_set_conv_pair_mode(-1)
    assert np.array_equal(out[0][0], out[2][0]) and np.array_equal(out[0][1], out[2][1])
    assert rel_err(out[2][0], y64) < 1e-3 and rel_err(out[2][1], dx64) < 1e-3
