"""Pins the CPU oracle (oracle/mcn_ops.py): the reference ships no golden vectors for this path (SURVEY.md
section 8c, "parity unpinned"), so every operator is pinned by (1) hand-computed known-answer tests,
(2) numerical-gradient checks in fp64, (3) an independent torch-CPU cross-check."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import mcn_ops as M


def numgrad(f, x, dy, eps=1e-6):
    g = np.zeros_like(x)
    it = np.nditer(x, flags=["multi_index"])
    for _ in it:
        i = it.multi_index
        old = x[i]
        x[i] = old + eps
        a = (f(x) * dy).sum()
        x[i] = old - eps
        b = (f(x) * dy).sum()
        x[i] = old
        g[i] = (a - b) / (2 * eps)
    return g


def test_out_size_matches_vl_nnconv_rule():
    assert M.out_size(224, 224, 7, 7, 3, 2) == (112, 112)
    assert M.out_size(112, 112, 3, 3, (0, 1, 0, 1), 2) == (56, 56)
    assert M.out_size(512, 300, 7, 7, 1, 2) == (254, 148)
    assert M.out_size(30, 17, 5, 3, 0, (3, 2)) == (9, 8)


def test_conv_known_answer():
    # 3x3 input, 2x2 filter of ones, no pad: each output is the sum of a 2x2 window (cross-correlation)
    x = np.arange(9, dtype=np.float64).reshape(3, 3).T.reshape(3, 3, 1, 1)  # x[h,w] = h + 3w ... column-major fill
    f = np.ones((2, 2, 1, 1))
    y = M.vl_nnconv(x, f, np.array([10.0]))
    expect = np.array([[x[0, 0, 0, 0] + x[1, 0, 0, 0] + x[0, 1, 0, 0] + x[1, 1, 0, 0], x[0, 1, 0, 0] + x[1, 1, 0, 0] + x[0, 2, 0, 0] + x[1, 2, 0, 0]],
                       [x[1, 0, 0, 0] + x[2, 0, 0, 0] + x[1, 1, 0, 0] + x[2, 1, 0, 0], x[1, 1, 0, 0] + x[2, 1, 0, 0] + x[1, 2, 0, 0] + x[2, 2, 0, 0]]]) + 10
    assert np.array_equal(y[:, :, 0, 0], expect)
    # no flip: an asymmetric filter picks x[h+1, w] - x[h, w]
    f2 = np.zeros((2, 1, 1, 1)); f2[0, 0, 0, 0] = -1; f2[1, 0, 0, 0] = 1
    y2 = M.vl_nnconv(x, f2)
    assert np.array_equal(y2[:, :, 0, 0], x[1:, :, 0, 0] - x[:-1, :, 0, 0])


@pytest.mark.parametrize("pad,stride", [(0, 1), (1, 2), ((0, 1, 2, 0), (2, 1)), ((2, 1, 0, 1), (3, 2))])
def test_conv_matches_torch_and_numgrad(pad, stride):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((7, 6, 3, 2))
    f = rng.standard_normal((3, 2, 3, 4))
    b = rng.standard_normal(4)
    y = M.vl_nnconv(x, f, b, pad=pad, stride=stride)
    pt, pb, pl, pr = M._pad4(pad)
    sy, sx = M._stride2(stride)
    xt = F.pad(torch.from_numpy(x.transpose(3, 2, 0, 1)), (pl, pr, pt, pb))
    yt = F.conv2d(xt, torch.from_numpy(f.transpose(3, 2, 0, 1).copy()), torch.from_numpy(b), stride=(sy, sx))
    assert np.allclose(y, yt.numpy().transpose(2, 3, 1, 0), atol=1e-12)
    dy = rng.standard_normal(y.shape)
    dx, df, db = M.vl_nnconv(x, f, b, dy, pad=pad, stride=stride)
    assert np.allclose(dx, numgrad(lambda v: M.vl_nnconv(v, f, b, pad=pad, stride=stride), x.copy(), dy), atol=1e-6)
    assert np.allclose(df, numgrad(lambda v: M.vl_nnconv(x, v, b, pad=pad, stride=stride), f.copy(), dy), atol=1e-6)
    assert np.allclose(db, dy.sum(axis=(0, 1, 3)))


def test_conv_groups():
    rng = np.random.default_rng(1)
    x = rng.standard_normal((5, 5, 4, 2))
    f = rng.standard_normal((3, 3, 2, 6))
    y = M.vl_nnconv(x, f, None, pad=1)
    yt = F.conv2d(torch.from_numpy(x.transpose(3, 2, 0, 1).copy()), torch.from_numpy(f.transpose(3, 2, 0, 1).copy()), padding=1, groups=2)
    assert np.allclose(y, yt.numpy().transpose(2, 3, 1, 0), atol=1e-12)


def test_maxpool_tie_break_first_in_memory_order():
    # all-equal window: the first element of the column-major scan (dh = 0, dw = 0) wins -> index 0;
    # a later strictly greater value wins; equal later values do not
    x = np.zeros((2, 2, 1, 1)); y, idx = M.vl_nnpool(x, 2, return_index=True)
    assert idx[0, 0, 0, 0] == 0
    x[1, 0] = 1.0; x[0, 1] = 1.0   # (h=1,w=0) comes before (h=0,w=1) in memory order
    _, idx = M.vl_nnpool(x, 2, return_index=True)
    assert idx[0, 0, 0, 0] == 1     # dw*PH + dh = 0*2 + 1
    dx = M.vl_nnpool(x, 2, np.full((1, 1, 1, 1), 5.0))
    assert dx[1, 0, 0, 0] == 5.0 and dx.sum() == 5.0


def test_maxpool_padding_is_minus_inf_and_avg_divides_by_inbounds():
    x = -np.ones((3, 3, 1, 1))
    y = M.vl_nnpool(x, 3, pad=1, stride=1, method="max")
    assert np.all(y == -1)          # zero padding would have produced 0
    ya = M.vl_nnpool(np.ones((3, 3, 1, 1)), 3, pad=1, stride=1, method="avg")
    assert np.allclose(ya, 1.0)     # corners average 4 in-bounds ones, not 9 cells


@pytest.mark.parametrize("method", ["max", "avg"])
def test_pool_backward_numgrad(method):
    rng = np.random.default_rng(2)
    x = rng.standard_normal((7, 6, 2, 2))
    args = dict(pad=(0, 1, 1, 0), stride=(2, 1), method=method)
    y = M.vl_nnpool(x, (3, 2), **args)
    dy = rng.standard_normal(y.shape)
    dx = M.vl_nnpool(x, (3, 2), dy, **args)
    assert np.allclose(dx, numgrad(lambda v: M.vl_nnpool(v, (3, 2), **args), x.copy(), dy), atol=1e-6)


def test_bnorm_known_answer_and_moments_are_mu_sigma():
    x = np.array([1.0, 2.0, 3.0, 4.0]).reshape(1, 1, 1, 4)
    y, mom = M.vl_nnbnorm(x, [2.0], [0.5], epsilon=0.0)
    mu, sigma = 2.5, np.sqrt(1.25)  # biased variance
    assert np.allclose(mom, [[mu, sigma]])
    assert np.allclose(y.ravel(), 2.0 * (np.array([1, 2, 3, 4]) - mu) / sigma + 0.5)
    y2, _ = M.vl_nnbnorm(x, [1.0], [0.0], moments=np.array([[1.0, 2.0]]))
    assert np.allclose(y2.ravel(), (np.array([1, 2, 3, 4]) - 1.0) / 2.0)  # sigma, not variance


@pytest.mark.parametrize("test_mode", [False, True])
def test_bnorm_backward_numgrad(test_mode):
    rng = np.random.default_rng(3)
    x = rng.standard_normal((3, 4, 5, 2))
    g, b = rng.uniform(0.5, 1.5, 5), rng.standard_normal(5)
    mom = np.stack([rng.standard_normal(5), rng.uniform(0.5, 2, 5)], 1) if test_mode else None
    dy = rng.standard_normal(x.shape)
    dx, dg, db, _ = M.vl_nnbnorm(x, g, b, dy, epsilon=1e-5, moments=mom)
    assert np.allclose(dx, numgrad(lambda v: M.vl_nnbnorm(v, g, b, epsilon=1e-5, moments=mom)[0], x.copy(), dy), atol=1e-6)
    assert np.allclose(dg, numgrad(lambda v: M.vl_nnbnorm(x, v, b, epsilon=1e-5, moments=mom)[0], g.copy(), dy), atol=1e-6)
    assert np.allclose(db, dy.sum(axis=(0, 1, 3)))
    if not test_mode:
        xt = torch.from_numpy(x.transpose(3, 2, 0, 1).copy())
        yt = F.batch_norm(xt, None, None, torch.from_numpy(g), torch.from_numpy(b), training=True, eps=1e-5)
        assert np.allclose(M.vl_nnbnorm(x, g, b, epsilon=1e-5)[0], yt.numpy().transpose(2, 3, 1, 0), atol=1e-10)


def test_softmaxceloss_known_answer():
    # T = 1, one-hot target: loss = -log softmax(x)[c]
    x = np.array([1.0, 2.0, 3.0]).reshape(1, 1, 3, 1)
    p = np.array([0.0, 0.0, 1.0]).reshape(1, 1, 3, 1)
    assert np.isclose(M.vl_nnsoftmaxceloss(x, p), -np.log(np.exp(3) / np.exp([1, 2, 3]).sum()))
    # logit targets equal to x: loss = entropy of softmax(x/T), gradient zero
    t = 2.0
    q = np.exp(x / t) / np.exp(x / t).sum()
    assert np.isclose(M.vl_nnsoftmaxceloss(x, x, temperature=t, logitTargets=True), -(q * np.log(q)).sum())
    assert np.allclose(M.vl_nnsoftmaxceloss(x, x, 1.0, temperature=t, logitTargets=True), 0)
    # summed over the batch (not averaged), instance weights multiply per sample
    x2 = np.concatenate([x, x], axis=3); p2 = np.concatenate([p, p], axis=3)
    assert np.isclose(M.vl_nnsoftmaxceloss(x2, p2), 2 * M.vl_nnsoftmaxceloss(x, p))
    assert np.isclose(M.vl_nnsoftmaxceloss(x2, p2, instanceWeights=[1.0, 3.0]), 4 * M.vl_nnsoftmaxceloss(x, p))


def test_softmaxceloss_gradient_is_q_minus_p_over_T():
    rng = np.random.default_rng(4)
    x, tl = rng.standard_normal((1, 1, 8, 5)), rng.standard_normal((1, 1, 8, 5))
    dx = M.vl_nnsoftmaxceloss(x, tl, 1.0, temperature=2.0, logitTargets=True)
    num = numgrad(lambda v: np.array(M.vl_nnsoftmaxceloss(v, tl, temperature=2.0, logitTargets=True)), x.copy(), 1.0)
    assert np.allclose(dx, num, atol=1e-6)
    q, p = M.vl_nnsoftmaxt(x, temperature=2.0), M.vl_nnsoftmaxt(tl, temperature=2.0)
    assert np.allclose(dx, (q - p) / 2.0)  # no T^2 factor
    with pytest.raises(AssertionError):
        M.vl_nnsoftmaxceloss(x, tl)  # targets that are not distributions are rejected


def test_euclidean_and_huber_losses_known_answers_and_gradients():
    """emoVoxZoo.m:138-147: dagnn.EuclideanLoss / dagnn.HuberLoss('sigma', 1) on {prediction, logitTarget, instanceWeights}."""
    x = np.array([0.0, 3.0, -2.0, 0.5]).reshape(1, 1, 4, 1)
    t = np.array([0.0, 1.0, 0.0, 0.0]).reshape(1, 1, 4, 1)
    assert np.isclose(M.vl_nneuclideanloss(x, t), 0.5 * (4 + 4 + 0.25))
    # smooth-L1, sigma = 1: |d| - 0.5 beyond 1, d^2 / 2 inside
    assert np.isclose(M.vl_nnhuberloss(x, t), (2 - 0.5) + (2 - 0.5) + 0.125)
    assert np.allclose(M.vl_nnhuberloss(x, t, 1.0).reshape(-1), [0, 1, -1, 0.5])
    # sigma = 2: knee at 1/4, slope 4 d inside, |d| - 1/8 outside
    assert np.isclose(M.vl_nnhuberloss(x, t, sigma=2.0), (2 - 0.125) * 2 + (0.5 - 0.125))
    rng = np.random.default_rng(6)
    x, t = rng.standard_normal((1, 1, 8, 5)) * 2, rng.standard_normal((1, 1, 8, 5))
    w = rng.uniform(0.5, 2, (1, 1, 1, 5))
    for f in (M.vl_nneuclideanloss, M.vl_nnhuberloss):
        num = numgrad(lambda v: np.array(f(v, t, instanceWeights=w)), x.copy(), 1.0)
        assert np.allclose(f(x, t, 1.0, instanceWeights=w), num, atol=1e-5)
        # per-sample weights broadcast over the classes; unit weights are the default
        assert np.isclose(f(x, t, instanceWeights=np.ones(5)), f(x, t))
        assert np.isclose(f(x, t, instanceWeights=3 * np.ones(5)), 3 * f(x, t))


def test_classerror_and_error_stats():
    x = np.zeros((1, 1, 3, 4)); x[0, 0, 2, 0] = 1; x[0, 0, 0, 1] = 1; x[0, 0, 1, 2] = 1  # sample 3: all-equal -> class 1
    c = np.array([3, 1, 1, 1])
    assert M.vl_nnloss(x, c, loss="classerror") == 1.0
    correct, count = M.error_stats(x, c, 3)
    assert list(count) == [3, 0, 1] and list(correct) == [2, 0, 1]


def test_elementwise_se_ops():
    rng = np.random.default_rng(5)
    x = rng.standard_normal((3, 2, 4, 2)); dy = rng.standard_normal(x.shape)
    assert np.allclose(M.vl_nnrelu(x, dy), dy * (x > 0))
    assert np.allclose(M.vl_nnsigmoid(x, dy), numgrad(M.vl_nnsigmoid, x.copy(), dy), atol=1e-6)
    assert np.allclose(M.vl_nnglobalpool(x)[0, 0], x.mean(axis=(0, 1)))
    a = rng.uniform(0, 1, (1, 1, 4, 2)); y = rng.standard_normal(x.shape)
    da, dxx, dyy = M.vl_nnaxpy(a, x, y, dy)
    assert np.allclose(da, numgrad(lambda v: M.vl_nnaxpy(v, x, y), a.copy(), dy), atol=1e-6)
    assert np.allclose(dxx, dy * a) and np.allclose(dyy, dy)
