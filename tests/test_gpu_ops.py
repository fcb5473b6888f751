"""Device-native building blocks (xemo_op_*) against the oracle on exactly-representable data: the fp16
fast-path pooling kernels (BN+ReLU folded into the read, packed compares) must reproduce MatConvNet's arg-max
rule bit-exactly, and the optional pool-gather BN backward must agree with the two-pass form."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import torch

    from mcncrossmodalemotions_b200 import _lib

    stream = torch.cuda.Stream()
    ctx = _lib.Context(0, stream.cuda_stream)
    return torch, ctx, stream


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def nhwc(x):  # H x W x C x N  ->  N H W C
    return np.ascontiguousarray(np.transpose(x, (3, 0, 1, 2)))


def hwcn(x):
    return np.transpose(x, (1, 2, 3, 0))


@pytest.mark.parametrize("geom", [((3, 3), (2, 2), 21, 15), ((5, 3), (3, 2), 30, 17), ((2, 2), (2, 2), 10, 12)])
@pytest.mark.parametrize("affine", [False, True])
def test_fused_maxpool_forward_backward_bit_exact(env, geom, affine):
    torch, ctx, stream = env
    from oracle import mcn_ops as M

    (ph, pw), (sh, sw), H, W = geom
    Cc, N = 24, 3
    rng = np.random.default_rng(ph * 100 + H + affine)
    x = (np.round(rng.standard_normal((H, W, Cc, N)) * 4) / 2).astype(np.float32)        # multiples of 0.5: many exact ties
    a = rng.choice([-2.0, -1.0, 0.5, 1.0, 2.0], Cc).astype(np.float32)                    # both signs
    b = (np.round(rng.standard_normal(Cc) * 4) / 4).astype(np.float32)
    z = np.maximum(a.reshape(1, 1, Cc, 1) * x + b.reshape(1, 1, Cc, 1), 0) if affine else x
    yr, ir = M.vl_nnpool(z, (ph, pw), stride=(sh, sw), method="max", return_index=True)
    OH, OW = yr.shape[:2]
    with torch.cuda.stream(stream):
        xd = torch.from_numpy(nhwc(x)).cuda().half()
        ad, bd = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
        yd = torch.zeros((N, OH, OW, Cc), dtype=torch.float16, device="cuda")
        idd = torch.zeros((N, OH, OW, Cc), dtype=torch.uint8, device="cuda")
        ctx.op_maxpool_fwd(_p(xd), N, H, W, Cc, ph, pw, sh, sw, 0, 0, 0, 0, _p(ad) if affine else None, _p(bd) if affine else None,
                           _p(yd), _p(idd))
        ctx.sync()
        assert np.array_equal(hwcn(yd.float().cpu().numpy()), yr)
        assert np.array_equal(hwcn(idd.cpu().numpy()), ir), "arg-max (first maximum in column-major scan order) must be bit-exact"
        dy = (np.round(rng.standard_normal(yr.shape) * 8) / 8).astype(np.float32)
        dxr = M.vl_nnpool(z, (ph, pw), dy, stride=(sh, sw), method="max")
        dyd = torch.from_numpy(nhwc(dy)).cuda().half()
        dxd = torch.zeros((N, H, W, Cc), dtype=torch.float16, device="cuda")
        ctx.op_maxpool_bwd(_p(dyd), _p(idd), N, H, W, Cc, ph, pw, sh, sw, 0, 0, 0, 0, _p(dxd))
        ctx.sync()
        assert np.array_equal(hwcn(dxd.float().cpu().numpy()), dxr)


def test_pool_gather_bn_backward_matches_two_pass(env):
    torch, ctx, stream = env
    N, H, W, Cc, ph, pw, sh, sw = 3, 21, 15, 32, 3, 3, 2, 2
    OH, OW = (H - ph) // sh + 1, (W - pw) // sw + 1
    rng = np.random.default_rng(0)
    with torch.cuda.stream(stream):
        x = torch.from_numpy(rng.standard_normal((N, H, W, Cc)).astype(np.float32)).cuda().half()
        g = torch.from_numpy(rng.uniform(0.5, 1.5, Cc).astype(np.float32)).cuda()
        beta = torch.from_numpy((0.1 * rng.standard_normal(Cc)).astype(np.float32)).cuda()
        ws = torch.zeros(2 * Cc, dtype=torch.float64, device="cuda")
        mom, a, b = (torch.zeros(n, device="cuda") for n in (2 * Cc, Cc, Cc))
        ctx.op_bn_train(_p(x), N * H * W, Cc, _p(g), _p(beta), 1e-5, _p(ws), _p(mom), _p(a), _p(b))
        y = torch.zeros((N, OH, OW, Cc), dtype=torch.float16, device="cuda")
        arg = torch.zeros((N, OH, OW, Cc), dtype=torch.uint8, device="cuda")
        ctx.op_maxpool_fwd(_p(x), N, H, W, Cc, ph, pw, sh, sw, 0, 0, 0, 0, _p(a), _p(b), _p(y), _p(arg))
        dpool = torch.from_numpy(rng.standard_normal((N, OH, OW, Cc)).astype(np.float32)).cuda().half()
        outs = []
        for fused in (False, True):
            dx = torch.zeros((N, H, W, Cc), dtype=torch.float16, device="cuda")
            dg, db, dbias = (torch.zeros(Cc, device="cuda") for _ in range(3))
            if fused:
                ctx.op_bn_bwd_pool(_p(x), _p(dpool), _p(arg), N, H, W, Cc, ph, pw, sh, sw, 0, 0, 0, 0, _p(mom), _p(a), _p(b), _p(ws),
                                   _p(dx), _p(dg), _p(db), _p(dbias), 1.0)
            else:
                full = torch.zeros((N, H, W, Cc), dtype=torch.float16, device="cuda")
                ctx.op_maxpool_bwd(_p(dpool), _p(arg), N, H, W, Cc, ph, pw, sh, sw, 0, 0, 0, 0, _p(full))
                ctx.op_bn_bwd(_p(x), _p(full), N * H * W, Cc, _p(mom), _p(a), _p(b), 1, 0, _p(ws), _p(dx), _p(dg), _p(db), _p(dbias), 1.0)
            ctx.sync()
            outs.append([t.float().cpu().numpy() for t in (dx, dg, db, dbias)])
    for u, v in list(zip(*outs[:2]))[:3]:   # dx, dg, db
        assert np.abs(u - v).max() <= 2e-3 * max(np.abs(v).max(), 1e-3)
    # sum_rows dx (the bias gradient of a conv feeding train-mode BN) is identically zero in exact arithmetic:
    # both forms must return only rounding noise
    for o in outs:
        assert np.abs(o[3]).max() <= 0.05 * np.abs(o[2]).max()


def test_spectrogram_front_end_matches_runspec_oracle(env):
    """runSpec + row normalisation on the device vs the numpy restatement."""
    torch, ctx, stream = env
    from oracle import nets

    N, W = 3, 300
    L = int(round((0.01 * W + 0.024) * 16000))
    rng = np.random.default_rng(0)
    t = np.arange(L) / 16000.0
    wav = np.stack([0.3 * np.sin(2 * np.pi * (300 + 500 * i) * t) + 0.05 * rng.standard_normal(L) for i in range(N)]).astype(np.float32)
    ref = np.stack([nets.run_spec(w) for w in wav], axis=2)[:, :, None, :]        # 512 x W x 1 x N
    assert ref.shape == (512, W, 1, N)
    with torch.cuda.stream(stream):
        wd = torch.from_numpy(wav).cuda()
        sd = torch.zeros(N * W * 512, device="cuda")
        ctx.op_spectrogram(_p(wd), N, L, 400, 160, 512, 0.97, 32768.0, W, _p(sd))
        ctx.sync()
        got = sd.cpu().numpy().reshape(N, W, 512).transpose(2, 1, 0)[:, :, None, :]
        assert np.abs(got - ref).max() <= 1e-4 * np.abs(ref).max()
        ctx.op_spec_rownorm(_p(sd), 512, W, N)
        ctx.sync()
        gotn = sd.cpu().numpy().reshape(N, W, 512).transpose(2, 1, 0)
        refn = np.stack([nets.normalize_spectrogram(ref[:, :, 0, i]) for i in range(N)], axis=2)
        assert np.abs(gotn - refn).max() <= 1e-3 * np.abs(refn).max()


def test_student_from_waveforms(env):
    from mcncrossmodalemotions_b200 import zoo
    from mcncrossmodalemotions_b200.programs import StudentProgram
    from oracle import nets

    n, W = 4, 100
    p = nets.student_randomize_bn(zoo.student_init())
    prog = StudentProgram(p, n, W, audio_input="wav")
    rng = np.random.default_rng(1)
    wav = (0.1 * rng.standard_normal((n, prog.wav_len))).astype(np.float32)
    spec = np.stack([nets.normalize_spectrogram(nets.run_spec(w)) for w in wav], axis=2)[:, :, None, :]
    ref, _ = nets.student_forward({k: v.astype(np.float64) for k, v in p.items()}, spec, "test", nets.TorchOps)
    got = prog.forward(wav, "test")
    assert np.abs(got - ref.reshape(8, n).T).max() <= 1e-3 * np.abs(ref).max()


# ---- student stem by linearity (csrc/stem_kernels.cuh): the kernels against the numpy restatement of test_stem_algebra.py
@pytest.mark.parametrize("shape", [(40, 38, 3), (70, 400, 2), (512, 100, 2)])
def test_stem_autocorrelation_and_bn_statistics(env, shape):
    torch, ctx, stream = env
    import test_stem_algebra as T
    from oracle import mcn_ops as M
    from mcncrossmodalemotions_b200.programs import student_conv1_to_s2d

    H, W, N = shape
    K = 32
    rng = np.random.default_rng(H + W)
    spec = rng.standard_normal((H, W, 1, N)).astype(np.float16).astype(np.float64)       # fp16-exact inputs
    f = (rng.standard_normal((7, 7, 1, K)) * 0.2).astype(np.float16).astype(np.float64)
    bias = (rng.standard_normal(K) * 0.3).astype(np.float32)
    g = rng.uniform(0.5, 1.5, K).astype(np.float32)
    beta = (rng.standard_normal(K) * 0.2).astype(np.float32)
    X, OH, OW = T.s2d(spec)
    HP = OH + 3
    with torch.cuda.stream(stream):
        sd = torch.from_numpy(np.ascontiguousarray(spec.astype(np.float32).transpose(3, 2, 1, 0)).reshape(-1)).cuda()
        s2d = torch.zeros((N, HP, OW, 16), dtype=torch.float16, device="cuda")
        ws = torch.zeros(int(ctx.lib.xemo_stem_ws_doubles()), dtype=torch.float64, device="cuda")
        w16 = torch.from_numpy(student_conv1_to_s2d(f).reshape(K, 64)).cuda().half()
        bd, gd, betad = (torch.from_numpy(v).cuda() for v in (bias, g, beta))
        mom, a, b = (torch.zeros(n, dtype=torch.float32, device="cuda") for n in (2 * K, K, K))
        ctx.op_spec_s2d(_p(sd), H, W, N, 1, 1, HP, OW, _p(s2d))
        ctx.op_stem_autocorr(_p(s2d), N, HP, OW, OH, _p(ws))
        ctx.op_stem_bn_train(_p(ws), _p(w16), _p(bd), N * OH * OW, K, _p(gd), _p(betad), 1e-5, _p(mom), _p(a), _p(b))
    ctx.sync()
    assert np.array_equal(s2d.cpu().numpy().astype(np.float64), X)
    xs = T.patches(X, OH)
    rs = ws.cpu().numpy()[-(64 * 64 + 64):]
    R, S = rs[: 64 * 64].reshape(64, 64), rs[64 * 64 :]
    Rref, Sref = xs.T @ xs, xs.sum(0)
    assert np.abs(R - Rref).max() <= 5e-6 * np.abs(Rref).max()      # fp32 partial sums, fp64 across warps / CTAs
    assert np.abs(S - Sref).max() <= 5e-6 * np.abs(xs).sum(0).max()
    x = M.vl_nnconv(spec, f, bias.astype(np.float64), pad=1, stride=2)
    _, moments = M.vl_nnbnorm(x, g.astype(np.float64), beta.astype(np.float64), epsilon=1e-5)
    got = mom.cpu().numpy().reshape(2, K).T
    assert np.abs(got - moments).max() <= 1e-5 * np.abs(moments).max()
    assert np.allclose(a.cpu().numpy(), g / moments[:, 1], rtol=1e-5)


def test_stem_pooled_bn_reduce_masks_and_sums(env):
    torch, ctx, stream = env
    P, Cc = 1000, 96
    rng = np.random.default_rng(5)
    xw = rng.standard_normal((P, Cc)).astype(np.float16)
    g = rng.standard_normal((P, Cc)).astype(np.float16)
    mu, sg = rng.standard_normal(Cc).astype(np.float32) * 0.1, rng.uniform(0.5, 1.5, Cc).astype(np.float32)
    a = (rng.choice([-1.0, 1.0], Cc) * rng.uniform(0.5, 1.5, Cc)).astype(np.float32)
    b = (rng.standard_normal(Cc) * 0.3).astype(np.float32)
    alive = (a[None] * xw.astype(np.float32) + b[None]) > 0
    gm = np.where(alive, g, np.float16(0))
    ref1 = gm.astype(np.float64).sum(0)
    ref2 = (gm.astype(np.float64) * (xw.astype(np.float64) - mu) / sg).sum(0)
    with torch.cuda.stream(stream):
        xd, gd = torch.from_numpy(xw).cuda(), torch.from_numpy(g).cuda()
        mom = torch.from_numpy(np.concatenate([mu, sg])).cuda()
        ad, bd = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
        acc = torch.zeros(2 * Cc, dtype=torch.float64, device="cuda")
        ctx.op_stem_pool_bn_reduce(_p(xd), _p(gd), P, Cc, 0, _p(mom), _p(ad), _p(bd), _p(acc))
    ctx.sync()
    assert np.array_equal(gd.cpu().numpy(), gm)
    got = acc.cpu().numpy()
    assert np.abs(got[:Cc] - ref1).max() <= 1e-5 * np.abs(ref1).max()
    assert np.abs(got[Cc:] - ref2).max() <= 1e-5 * np.abs(ref2).max()


def test_maxpool_forward_records_raw_winner(env):
    torch, ctx, stream = env
    N, H, W, Cc = 2, 21, 15, 24
    rng = np.random.default_rng(9)
    x = rng.standard_normal((N, H, W, Cc)).astype(np.float16)
    a = (rng.choice([-1.0, 1.0], Cc) * rng.uniform(0.5, 1.5, Cc)).astype(np.float32)
    b = (rng.standard_normal(Cc) * 0.3).astype(np.float32)
    OH, OW = (H - 3) // 2 + 1, (W - 3) // 2 + 1
    z = a * x.astype(np.float32) + b
    win = np.stack([z[:, dh : dh + 2 * OH - 1 : 2, dw : dw + 2 * OW - 1 : 2] for dw in range(3) for dh in range(3)], 0)
    xr = np.stack([x[:, dh : dh + 2 * OH - 1 : 2, dw : dw + 2 * OW - 1 : 2] for dw in range(3) for dh in range(3)], 0)
    ref = np.take_along_axis(xr, win.argmax(0)[None], 0)[0]
    with torch.cuda.stream(stream):
        xd, ad, bd = torch.from_numpy(x).cuda(), torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
        y = torch.zeros((N, OH, OW, Cc), dtype=torch.float16, device="cuda")
        xw = torch.zeros_like(y)
        arg = torch.zeros((N, OH, OW, Cc), dtype=torch.uint8, device="cuda")
        ctx.op_maxpool_fwd_win(_p(xd), N, H, W, Cc, 3, 3, 2, 2, 0, 0, 0, 0, _p(ad), _p(bd), _p(y), _p(arg), _p(xw), 0)
    ctx.sync()
    got, yy = xw.cpu().numpy(), y.cpu().numpy().astype(np.float32)
    live = yy > 0        # an all-non-positive window has no unique winner (and its gradient is masked)
    assert np.array_equal(got[live], ref[live])
    assert np.all((a * got.astype(np.float32) + b)[~live] <= 0)



@pytest.mark.parametrize("n,width,cin,kout", [(5, 8, 256, 512), (3, 11, 64, 160)])
def test_full_height_data_gradient_matches_the_general_form_and_the_oracle(env, n, width, cin, kout):
    """fc6-type layers (a 9 x 1 filter over a 9 x W map, one output row): xemo_op_conv_dgrad_fullheight (a GEMM over (n, w)
    rows whose N tiles land on the 9 rows of dX) against the parity-decomposed general form and the oracle's vl_nnconv DX."""
    torch, ctx, stream = env
    from oracle import mcn_ops as M

    R = 9
    rng = np.random.default_rng(n + width)
    f = (rng.standard_normal((R, 1, cin, kout)) / np.sqrt(R * cin)).astype(np.float16).astype(np.float32)
    dy = rng.standard_normal((1, width, kout, n)).astype(np.float16).astype(np.float32)
    x = np.zeros((R, width, cin, n), np.float32)
    dx_ref, _, _ = M.vl_nnconv(x.astype(np.float64), f.astype(np.float64), None, dy.astype(np.float64))
    with torch.cuda.stream(stream):
        w16 = torch.from_numpy(np.ascontiguousarray(np.transpose(f, (3, 0, 1, 2)))).cuda().half()      # [K][R][1][C]
        dyd = torch.from_numpy(nhwc(dy)).cuda().half()                                               # [N][1][W][K]
        packed = torch.zeros(int(ctx.lib.xemo_dgrad_pack_elems(cin, kout, R, 1, 1, 1)), dtype=torch.float16, device="cuda")
        out = {}
        for which in ("general", "fullheight"):
            dxd = torch.full((n, R, width, cin), float("nan"), dtype=torch.float16, device="cuda")
            if which == "general":
                ctx.op_pack_dgrad_filters(_p(w16), kout, R, 1, cin, 1, 1, 0, 0, _p(packed))
                ctx.op_conv_dgrad(_p(dyd), n, R, width, cin, _p(packed), kout, R, 1, 1, 1, 0, 0, 0, 0, _p(dxd))
            else:
                ctx.op_pack_dgrad_filters_fullheight(_p(w16), kout, R, cin, _p(packed))
                ctx.op_conv_dgrad_fullheight(_p(dyd), n, R, width, cin, _p(packed), kout, _p(dxd))
            ctx.sync()
            out[which] = hwcn(dxd.float().cpu().numpy())
    scale = np.abs(dx_ref).max()
    assert np.isfinite(out["fullheight"]).all()
    assert np.abs(out["fullheight"] - dx_ref).max() / scale < 1e-3
    assert np.abs(out["fullheight"] - out["general"]).max() / scale < 1e-3


@pytest.mark.parametrize("n", [1, 3, 32, 64, 100, 256])                   # cluster sizes 4, 4, 4, 4, 2, 1
@pytest.mark.parametrize("dims", [(256, 64), (512, 128), (1024, 256), (2048, 512)])   # (C, Cm) of SENet50's four stages
def test_se_gate_kernels_match_fp64_at_every_cluster_size(env, n, dims):
    """mcnExtraLayers SE gate (GlobalPooling -> Conv -> ReLU -> Conv -> Sigmoid): the plain form from s and the form
    by linearity from mean_hw(t2), against numpy fp64.  The cluster of CTAs that shares a group of two samples
    changes with N; an odd N leaves the last group half empty."""
    torch, ctx, stream = env
    Cc, Cm = dims
    Cr = Cc // 16
    rng = np.random.default_rng(n * 7 + Cc)
    s = rng.standard_normal((n, Cc)).astype(np.float32)
    w1 = (rng.standard_normal((Cr, Cc)) / np.sqrt(Cc)).astype(np.float32)
    b1 = (0.1 * rng.standard_normal(Cr)).astype(np.float32)
    w2t = (rng.standard_normal((Cr, Cc)) / np.sqrt(Cr)).astype(np.float32)
    b2 = (0.1 * rng.standard_normal(Cc)).astype(np.float32)

    def gate_of(sv):
        hid = np.maximum(sv.astype(np.float64) @ w1.astype(np.float64).T + b1, 0)
        return 1 / (1 + np.exp(-(hid @ w2t.astype(np.float64) + b2)))

    with torch.cuda.stream(stream):
        d = {k: torch.from_numpy(v).cuda() for k, v in dict(s=s, w1=w1, b1=b1, w2t=w2t, b2=b2).items()}
        g = torch.full((n, Cc), -1.0, device="cuda")
        ctx.op_se_gate(_p(d["s"]), n, Cc, Cr, _p(d["w1"]), _p(d["b1"]), _p(d["w2t"]), _p(d["b2"]), _p(g))
        ctx.sync()
        assert np.abs(g.cpu().numpy() - gate_of(s)).max() < 2e-6

        m2 = rng.standard_normal((n, Cm)).astype(np.float32)
        w3 = (rng.standard_normal((Cc, Cm)) / np.sqrt(Cm)).astype(np.float16)
        a3 = rng.uniform(0.5, 1.5, Cc).astype(np.float32)
        b3 = (0.2 * rng.standard_normal(Cc)).astype(np.float32)
        dl = {k: torch.from_numpy(v).cuda() for k, v in dict(m2=m2, w3=w3, a3=a3, b3=b3).items()}
        sc = torch.full((n, Cc), -1.0, device="cuda")
        sh = torch.full((n, Cc), -1.0, device="cuda")
        ctx.op_se_gate_lin(_p(dl["m2"]), n, Cc, Cm, Cr, _p(dl["w3"]), _p(dl["a3"]), _p(dl["b3"]), _p(d["w1"]), _p(d["b1"]), _p(d["w2t"]),
                           _p(d["b2"]), _p(sc), _p(sh))
        ctx.sync()
        s_lin = a3 * (m2.astype(np.float64) @ w3.astype(np.float64).T) + b3
        gl = gate_of(s_lin)
        assert np.abs(sc.cpu().numpy() - gl * a3).max() < 5e-6
        assert np.abs(sh.cpu().numpy() - gl * b3).max() < 5e-6


@pytest.mark.parametrize("dims", [(256, 64, 56 * 56), (512, 128, 28 * 28), (1024, 256, 14 * 14), (2048, 512, 49)])
def test_se_squeeze_and_gate_do_not_depend_on_the_batch(env, dims):
    """Block sizes (squeeze) and cluster sizes (gate) follow the batch; the summation orders do not: the first two faces of
    a batch of 256 get bit-identical means and gates when they are run as a batch of 2, 32 or 100."""
    torch, ctx, stream = env
    Cc, Cm, HW = dims
    Cr = Cc // 16
    rng = np.random.default_rng(Cc)
    with torch.cuda.stream(stream):
        w1 = torch.from_numpy((rng.standard_normal((Cr, Cc)) / np.sqrt(Cc)).astype(np.float32)).cuda()
        b1 = torch.from_numpy((0.1 * rng.standard_normal(Cr)).astype(np.float32)).cuda()
        w2t = torch.from_numpy((rng.standard_normal((Cr, Cc)) / np.sqrt(Cr)).astype(np.float32)).cuda()
        b2 = torch.from_numpy((0.1 * rng.standard_normal(Cc)).astype(np.float32)).cuda()
        w3 = torch.from_numpy((rng.standard_normal((Cc, Cm)) / np.sqrt(Cm)).astype(np.float16)).cuda()
        a3 = torch.from_numpy(rng.uniform(0.5, 1.5, Cc).astype(np.float32)).cuda()
        b3 = torch.from_numpy((0.2 * rng.standard_normal(Cc)).astype(np.float32)).cuda()
        u = torch.from_numpy(rng.standard_normal((256, HW, Cm)).astype(np.float16)).cuda()

        def run(n):
            m2 = torch.empty((n, Cm), device="cuda")
            ctx.op_se_squeeze(_p(u), n, HW, Cm, _p(m2))
            sc, sh = torch.empty((n, Cc), device="cuda"), torch.empty((n, Cc), device="cuda")
            ctx.op_se_gate_lin(_p(m2), n, Cc, Cm, Cr, _p(w3), _p(a3), _p(b3), _p(w1), _p(b1), _p(w2t), _p(b2), _p(sc), _p(sh))
            s = (sc + sh).contiguous()                   # any [n][C] fp32 vector serves as the plain gate's input
            g = torch.empty((n, Cc), device="cuda")
            ctx.op_se_gate(_p(s), n, Cc, Cr, _p(w1), _p(b1), _p(w2t), _p(b2), _p(g))
            ctx.sync()
            return [t[:2].cpu().numpy() for t in (m2, sc, sh, g)]

        ref = run(256)
        assert np.abs(ref[0] - u[:2].float().mean(1).cpu().numpy()).max() < 1e-5
        for n in (2, 32, 100):
            for a, b, what in zip(ref, run(n), ("squeeze", "nc_scale", "nc_shift", "gate")):
                assert np.array_equal(a, b), (what, n)


@pytest.mark.parametrize("shape", [(3, 49, 8), (2, 49, 16), (5, 200, 24), (4, 784, 40), (2, 3136, 64), (3, 130, 2048)])
def test_se_squeeze_on_narrow_and_odd_tensors(env, shape):
    """Global average pooling (mcnExtraLayers vl_nnglobalpool) on shapes outside SENet50's: fewer than 32 channels (one
    partial warp per block), channel counts that are not powers of two, maps just above the 128-pixel switch."""
    torch, ctx, stream = env
    n, HW, Cc = shape
    rng = np.random.default_rng(HW + Cc)
    u = rng.standard_normal((n, HW, Cc)).astype(np.float16)
    with torch.cuda.stream(stream):
        ud = torch.from_numpy(u).cuda()
        m = torch.full((n, Cc), -7.0, device="cuda")
        ctx.op_se_squeeze(_p(ud), n, HW, Cc, _p(m))
        ctx.sync()
        assert np.abs(m.cpu().numpy() - u.astype(np.float64).mean(1)).max() < 2e-6
