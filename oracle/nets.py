"""CPU restatement of the three graphs on the distillation hot path and of cnn_train_dag's step.

TEST INFRASTRUCTURE ONLY (same rule as oracle/mcn_ops.py): imported by tests/, by
__graft_entry__.smoke() and by bench.py's cpu_baseline / `--impl reference` legs, never by the
product package.  PARITY UNPINNED -- see the header of oracle/mcn_ops.py and SURVEY.md section 8c:
the reference ships neither the graphs (they are downloaded .mat files,
emoVoxCeleb/emoVoxZoo.m:95-97) nor any golden activations, so the architectures below are restated
from SURVEY.md Appendix A and pinned by the shape constraints the reference itself encodes (the
width -> pool6 table at emoVoxCeleb/emoVoxZoo.m:258-259, input 512 x W at
emoVoxCeleb/run_distillation.m:215, 8 outputs at emoVoxCeleb/emoVoxZoo.m:19).

Two operator back-ends walk the same graph description:
  * `NumpyOps`  -- oracle/mcn_ops.py, the literal restatement (any dtype, fp64 for ground truth);
  * `TorchOps`  -- the same operators on torch CPU kernels (multi-threaded; used for the timed CPU
                   baseline and for full-size parity).  tests/test_oracle_nets.py checks that the two
                   agree, so that either can serve as the checker.
All tensors at this interface are numpy arrays of logical shape H x W x C x N (MatConvNet).
"""
from __future__ import annotations

import numpy as np

from . import mcn_ops as M

BN_EPS = 1e-5  # dagnn.BatchNorm default epsilon

# emoVoxCeleb/emoVoxZoo.m:258-259 / external/compute_audio_feats.m:45-46
POOL6_TABLE = {100: 2, 200: 5, 300: 8, 400: 11, 500: 14, 600: 17, 700: 20, 800: 23, 900: 27, 1000: 30}


# ------------------------------------------------------------------------------------------------
# operator back-ends


class NumpyOps:
    name = "numpy"
    conv = staticmethod(M.vl_nnconv)
    pool = staticmethod(M.vl_nnpool)
    bnorm = staticmethod(M.vl_nnbnorm)
    relu = staticmethod(M.vl_nnrelu)


class TorchOps:
    """MatConvNet operator semantics on torch CPU kernels (H x W x C x N numpy in / out)."""

    name = "torch"

    @staticmethod
    def _t(x):
        import torch

        return torch.from_numpy(np.ascontiguousarray(np.transpose(x, (3, 2, 0, 1))))  # N C H W

    @staticmethod
    def _n(t):
        return np.transpose(t.numpy(), (2, 3, 1, 0))

    @classmethod
    def conv(cls, x, f, b=None, dzdy=None, pad=0, stride=1):
        import torch
        import torch.nn.functional as F

        pt, pb, pl, pr = M._pad4(pad)
        sy, sx = M._stride2(stride)
        H, W, C, N = x.shape
        FH, FW, FC, K = f.shape
        OH, OW = M.out_size(H, W, FH, FW, pad, (sy, sx))
        xt = cls._t(x)
        wt = torch.from_numpy(np.ascontiguousarray(np.transpose(f, (3, 2, 0, 1))))  # K C FH FW
        # crop the padded input to exactly what the OH x OW outputs read (vl_nnconv floors)
        need_h, need_w = (OH - 1) * sy + FH, (OW - 1) * sx + FW
        xp = F.pad(xt, (pl, pr, pt, pb))[:, :, :need_h, :need_w]
        bt = None if b is None or not np.size(b) else torch.from_numpy(np.asarray(b, dtype=x.dtype).reshape(-1))
        if dzdy is None:
            return cls._n(F.conv2d(xp, wt, bt, stride=(sy, sx)))
        dyt = cls._t(dzdy)
        dxp = torch.nn.grad.conv2d_input(xp.shape, wt, dyt, stride=(sy, sx))
        dw = torch.nn.grad.conv2d_weight(xp, wt.shape, dyt, stride=(sy, sx))
        full = torch.zeros((N, C, H + pt + pb, W + pl + pr), dtype=xt.dtype)
        full[:, :, :need_h, :need_w] = dxp
        dx = full[:, :, pt : pt + H, pl : pl + W]
        db = None if bt is None else dzdy.sum(axis=(0, 1, 3))
        return cls._n(dx), np.transpose(dw.numpy(), (2, 3, 1, 0)), db

    pool = staticmethod(M.vl_nnpool)  # pooling is cheap next to conv; keep the literal restatement
    bnorm = staticmethod(M.vl_nnbnorm)
    relu = staticmethod(M.vl_nnrelu)


def _r16(a):
    return a.astype(np.float16).astype(a.dtype)


class Fp16ModelOps(TorchOps):
    """The device number-format model: convolution operands (activations, filters, output gradients)
    rounded to fp16 before a wide-accumulate contraction; activations and data gradients stored in
    fp16; everything else as in TorchOps.  Used to separate "the kernels implement the fp16 model
    exactly" (tight tolerance against this back-end) from "the fp16 model is close to fp32" (the
    documented tolerance against TorchOps / NumpyOps)."""

    name = "fp16-model"

    @classmethod
    def conv(cls, x, f, b=None, dzdy=None, pad=0, stride=1):
        if dzdy is None:
            return _r16(TorchOps.conv(_r16(x), _r16(f), b, None, pad, stride))
        dx, df, db = TorchOps.conv(_r16(x), _r16(f), b, _r16(dzdy), pad, stride)
        return _r16(dx), df, db


# ------------------------------------------------------------------------------------------------
# VGGVox student.  Layer tuples: (name, type, args)

STUDENT_CONVS = [
    # name, FH, FW, Cin, Cout, stride, pad, has_bn
    ("conv1", 7, 7, 1, 96, (2, 2), (1, 1, 1, 1), True),
    ("conv2", 5, 5, 96, 256, (2, 2), (1, 1, 1, 1), True),
    ("conv3", 3, 3, 256, 384, (1, 1), (1, 1, 1, 1), True),
    ("conv4", 3, 3, 384, 256, (1, 1), (1, 1, 1, 1), True),
    ("conv5", 3, 3, 256, 256, (1, 1), (1, 1, 1, 1), True),
    ("fc6", 9, 1, 256, 4096, (1, 1), (0, 0, 0, 0), True),
    ("fc7", 1, 1, 4096, 1024, (1, 1), (0, 0, 0, 0), True),
    ("fc8", 1, 1, 1024, 8, (1, 1), (0, 0, 0, 0), False),
]
# pooling that follows the (conv, bn, relu) trio of that name
STUDENT_POOLS = {
    "conv1": ("pool1", "max", (3, 3), (2, 2)),
    "conv2": ("pool2", "max", (3, 3), (2, 2)),
    "conv5": ("pool5", "max", (5, 3), (3, 2)),
    "fc6": ("pool6", "avg", None, (1, 1)),  # window [1 p] from the width bucket
}


def student_pool6_width(spec_width):
    """Width of fc6's output (= the pool6 window that averages all of it) for a 512 x W input."""
    w = spec_width
    w = (w + 2 - 7) // 2 + 1
    w = (w - 3) // 2 + 1
    w = (w + 2 - 5) // 2 + 1
    w = (w - 3) // 2 + 1
    w = (w - 3) // 2 + 1
    return w


def student_init(seed=3, num_outputs=8, dtype=np.float32):
    """dag.initParams() as called at emoVoxCeleb/emoVoxZoo.m:54: conv W ~ N(0, 2/(FH*FW*FC)),
    bias 0; BatchNorm mult 1, bias 0, moments 0."""
    rng = np.random.default_rng(seed)
    p = {}
    for name, fh, fw, cin, cout, _, _, has_bn in STUDENT_CONVS:
        if name == "fc8":
            cout = num_outputs
        p[name + "f"] = (rng.standard_normal((fh, fw, cin, cout)) * np.sqrt(2.0 / (fh * fw * cin))).astype(dtype)
        p[name + "b"] = np.zeros((cout,), dtype)
        if has_bn:
            bn = "bn" + name[-1]
            p[bn + "m"] = np.ones((cout,), dtype)
            p[bn + "b"] = np.zeros((cout,), dtype)
            p[bn + "x"] = np.zeros((cout, 2), dtype)  # moments [mu sigma]
    return p


def student_randomize_bn(p, seed=5):
    """Non-trivial BN scale/shift/moments so that test-mode parity exercises them (tests only)."""
    rng = np.random.default_rng(seed)
    for k in list(p):
        if k.startswith("bn") and k.endswith("m"):
            c = p[k].shape[0]
            dt = p[k].dtype
            p[k] = rng.uniform(0.5, 1.5, c).astype(dt)
            p[k[:-1] + "b"] = (0.1 * rng.standard_normal(c)).astype(dt)
            p[k[:-1] + "x"] = np.stack([0.1 * rng.standard_normal(c), rng.uniform(0.5, 1.5, c)], axis=1).astype(dt)
    return p


def student_forward(p, x, mode="train", ops=NumpyOps, keep=False):
    """Forward of the student DagNN.  x: 512 x W x 1 x N.  Returns (prediction 1x1xKxN, tape).
    mode 'train': batch statistics (cnn_train_dag); 'test': stored moments
    (external/compute_audio_feats.m:106 sets dag.mode = 'test')."""
    tape = {}
    cur = x
    for name, fh, fw, cin, cout, stride, pad, has_bn in STUDENT_CONVS:
        if keep:
            tape[name + ":x"] = cur
        cur = ops.conv(cur, p[name + "f"], p[name + "b"], pad=pad, stride=stride)
        if has_bn:
            bn = "bn" + name[-1]
            if keep:
                tape[bn + ":x"] = cur
            cur, mom = ops.bnorm(cur, p[bn + "m"], p[bn + "b"], epsilon=BN_EPS, moments=p[bn + "x"] if mode == "test" else None)
            if keep:
                tape[bn + ":moments"] = mom
                tape["relu" + name[-1] + ":x"] = cur
            cur = ops.relu(cur)
        if name in STUDENT_POOLS:
            pname, method, win, pstride = STUDENT_POOLS[name]
            if win is None:
                win = (1, cur.shape[1])
            if keep:
                tape[pname + ":x"] = cur
                tape[pname + ":win"] = win
            if method == "max" and keep:
                cur, arg = ops.pool(cur, win, pad=0, stride=pstride, method="max", return_index=True)
                tape[pname + ":argmax"] = arg
            else:
                cur = ops.pool(cur, win, pad=0, stride=pstride, method=method)
    return cur, tape


def student_backward(p, tape, dzdy, ops=NumpyOps, relu_masks=None, pool_index=None):
    """Reverse sweep (derOutputs = {'objective', 1}).  Returns the gradient dict keyed like `p`
    (BN moments entries hold the batch moments, which cnn_train_dag averages in).
    Test instrumentation: `relu_masks` {'relu<i>': bool H x W x C x N} and `pool_index` {'pool<i>': uint8 window-local
    indices} replace the oracle's own discrete decisions (x > 0, first arg-max) by another implementation's, so that
    what remains of a gradient difference is arithmetic, not decision flips."""
    g = {}
    cur = dzdy
    for name, fh, fw, cin, cout, stride, pad, has_bn in reversed(STUDENT_CONVS):
        if name in STUDENT_POOLS:
            pname, method, _, pstride = STUDENT_POOLS[name]
            if pool_index is not None and pname in pool_index:
                cur = M.vl_nnpool(tape[pname + ":x"], tape[pname + ":win"], cur, pad=0, stride=pstride, method=method, index=pool_index[pname])
            else:
                cur = ops.pool(tape[pname + ":x"], tape[pname + ":win"], cur, pad=0, stride=pstride, method=method)
        if has_bn:
            bn = "bn" + name[-1]
            if relu_masks is not None and "relu" + name[-1] in relu_masks:
                cur = cur * relu_masks["relu" + name[-1]]
            else:
                cur = ops.relu(tape["relu" + name[-1] + ":x"], cur)
            cur, dg, db, mom = ops.bnorm(tape[bn + ":x"], p[bn + "m"], p[bn + "b"], cur, epsilon=BN_EPS)
            g[bn + "m"], g[bn + "b"], g[bn + "x"] = dg, db, mom
        dx, df, dbias = ops.conv(tape[name + ":x"], p[name + "f"], p[name + "b"], cur, pad=pad, stride=stride)
        g[name + "f"], g[name + "b"] = df, dbias
        cur = dx
    return g


def aggregate_logits(frame_logits, method="max", num_pred=8):
    """emoVoxCeleb/getBatchEmoVoxCeleb.m:179-188,30: aggregate the F x 8 frame logits of one clip."""
    lg = np.asarray(frame_logits)
    out = lg.max(axis=0) if method == "max" else lg.mean(axis=0)
    return out[:num_pred]


def time2idx(t, fps=25, stride=6):
    """emoVoxCeleb/getBatchEmoVoxCeleb.m:210-214: 1-based teacher-frame index of time t (seconds)."""
    return int(np.floor(max(fps * t - 1, 0) / stride) + 1)


def sgd_momentum(p, state, g, lr, batch_size, momentum=0.9, weight_decay=5e-4, bn_rate=0.1):
    """cnn_train_dag accumulateGradients, in place:
       m <- mu*m - (lambda*w + g/B) ; w <- w + lr*m      for 'gradient' parameters
       moments <- (1-rate)*moments + rate*batch_moments    for BatchNorm moments ('average', lr 0.1)."""
    for k in p:
        if k.startswith("bn") and k.endswith("x"):
            p[k] = ((1 - bn_rate) * p[k] + bn_rate * g[k]).astype(p[k].dtype)
            continue
        m = state.setdefault(k, np.zeros_like(p[k]))
        m[...] = momentum * m - (weight_decay * p[k] + g[k].reshape(p[k].shape) / batch_size)
        p[k] = (p[k] + lr * m).astype(p[k].dtype)
    return p, state


def distillation_student_step(p, state, spec, logit_target, lr=1e-4, T=2.0, ops=NumpyOps, update=True, loss_type="hot-cross-ent",
                              instance_weights=None):
    """One cnn_train_dag iteration on the student (emoVoxCeleb/run_distillation.m:170-182) with the
    loss wired at emoVoxCeleb/emoVoxZoo.m:137-157 (`loss_type`: hot-cross-ent [default, :151-152], softmaxlog, euclidean,
    huber).  logit_target: 1 x 1 x 8 x N teacher logits.  Returns dict(prediction, objective, classerror, grads)."""
    N = spec.shape[3]
    pred, tape = student_forward(p, spec, "train", ops, keep=True)
    max_label = logit_target.argmax(axis=2).reshape(1, 1, 1, N) + 1  # getBatchEmoVoxCeleb.m:32
    one = np.array(1.0, pred.dtype)
    if loss_type == "hot-cross-ent":
        loss = lambda *dz: M.vl_nnsoftmaxceloss(pred, logit_target, *dz, temperature=T, logitTargets=True)
    elif loss_type == "softmaxlog":
        loss = lambda *dz: M.vl_nnloss(pred, max_label, *dz, loss="softmaxlog")
    elif loss_type == "euclidean":
        loss = lambda *dz: M.vl_nneuclideanloss(pred, logit_target, *dz, instanceWeights=instance_weights)
    elif loss_type == "huber":
        loss = lambda *dz: M.vl_nnhuberloss(pred, logit_target, *dz, sigma=1.0, instanceWeights=instance_weights)
    else:
        raise ValueError("unrecognised regression loss: %s" % loss_type)
    objective = loss()
    classerror = M.vl_nnloss(pred, max_label, loss="classerror")
    dpred = loss(one)
    grads = student_backward(p, tape, dpred.astype(pred.dtype), ops)
    if update:
        sgd_momentum(p, state, grads, lr, N)
    return dict(prediction=pred, objective=float(objective), classerror=classerror, grads=grads,
                max_label=max_label, tape=tape)


# ------------------------------------------------------------------------------------------------
# ResNet50 / SENet50 -ferplus teachers

TEACHER_STAGES = [(3, 64, 256, 1), (4, 128, 512, 2), (6, 256, 1024, 2), (3, 512, 2048, 2)]  # blocks, mid, out, stride
PIXEL_SCALE = 100.0  # ~ std of a He-initialised conv1 response to mean-subtracted 0..255 pixels
AVERAGE_IMAGE = np.array([131.0912, 103.8827, 91.4953], np.float32)  # VGGFace2 RGB means


def teacher_init(arch="senet50", seed=4, num_outputs=8, dtype=np.float32):
    """Synthetic, seeded teacher weights: conv He-normal, BN mult ~ U(0.5,1.5),
    bias ~ N(0,0.1), mu ~ N(0,0.1), sigma ~ U(0.5,1.5), SE FC biases 0.  The last BN of every
    bottleneck gets mult ~ U(0.1,0.3) so that activations stay O(1) through 16 residual sums."""
    assert arch in ("resnet50", "senet50")
    rng = np.random.default_rng(seed)
    p = {"arch": arch}

    def conv(name, fh, fw, cin, cout):
        p[name + "f"] = (rng.standard_normal((fh, fw, cin, cout)) * np.sqrt(2.0 / (fh * fw * cin))).astype(dtype)

    def bn(name, c, small=False, scale=1.0):
        lo, hi = (0.1, 0.3) if small else (0.5, 1.5)
        p[name + "m"] = rng.uniform(lo, hi, c).astype(dtype)
        p[name + "b"] = (0.1 * rng.standard_normal(c)).astype(dtype)
        p[name + "x"] = (scale * np.stack([0.1 * rng.standard_normal(c), rng.uniform(0.5, 1.5, c)], axis=1)).astype(dtype)

    conv("conv1", 7, 7, 3, 64)
    bn("bn1", 64, scale=PIXEL_SCALE)  # a trained bn1 absorbs the 0..255 pixel range
    cin = 64
    for si, (blocks, mid, cout, stride) in enumerate(TEACHER_STAGES):
        for bi in range(blocks):
            pre = "s%db%d_" % (si + 2, bi + 1)
            conv(pre + "c1", 1, 1, cin, mid); bn(pre + "bn1", mid)
            conv(pre + "c2", 3, 3, mid, mid); bn(pre + "bn2", mid)
            conv(pre + "c3", 1, 1, mid, cout); bn(pre + "bn3", cout, small=True)
            if bi == 0:
                conv(pre + "proj", 1, 1, cin, cout); bn(pre + "bnp", cout)
            if arch == "senet50":
                r = cout // 16
                conv(pre + "se1", 1, 1, cout, r); p[pre + "se1b"] = np.zeros(r, dtype)
                conv(pre + "se2", 1, 1, r, cout); p[pre + "se2b"] = np.zeros(cout, dtype)
            cin = cout
    conv("classifier", 1, 1, 2048, num_outputs)
    p["classifierb"] = (0.01 * rng.standard_normal(num_outputs)).astype(dtype)
    return p


def teacher_forward(p, x, ops=NumpyOps, taps=None):
    """Teacher forward with dag.mode = 'test' (emoVoxCeleb/fetch_emovoxceleb_imdb.m:107,129).
    x: 224 x 224 x 3 x N mean-subtracted faces.  Returns logits 1 x 1 x 8 x N.  `taps` (dict) collects
    the block outputs for per-stage parity."""
    se = p["arch"] == "senet50"

    def cbr(name, bnname, t, stride=1, pad=0, relu=True):
        t = ops.conv(t, p[name + "f"], None, pad=pad, stride=stride)
        t, _ = ops.bnorm(t, p[bnname + "m"], p[bnname + "b"], epsilon=BN_EPS, moments=p[bnname + "x"])
        return ops.relu(t) if relu else t

    cur = cbr("conv1", "bn1", x, stride=2, pad=3)
    cur = ops.pool(cur, (3, 3), pad=(0, 1, 0, 1), stride=2, method="max")
    if taps is not None:
        taps["pool1"] = cur
    for si, (blocks, mid, cout, stride) in enumerate(TEACHER_STAGES):
        for bi in range(blocks):
            pre = "s%db%d_" % (si + 2, bi + 1)
            s = stride if bi == 0 else 1
            u = cbr(pre + "c1", pre + "bn1", cur, stride=s)
            u = cbr(pre + "c2", pre + "bn2", u, pad=1)
            u = cbr(pre + "c3", pre + "bn3", u, relu=False)
            sc = cbr(pre + "proj", pre + "bnp", cur, stride=s, relu=False) if bi == 0 else cur
            if se:
                z = M.vl_nnglobalpool(u)
                z = ops.relu(ops.conv(z, p[pre + "se1f"], p[pre + "se1b"]))
                a = M.vl_nnsigmoid(ops.conv(z, p[pre + "se2f"], p[pre + "se2b"]))
                cur = ops.relu(M.vl_nnaxpy(a, u, sc))
            else:
                cur = ops.relu(u + sc)
            if taps is not None:
                taps[pre[:-1]] = cur
    cur = ops.pool(cur, (7, 7), pad=0, stride=1, method="avg")
    return ops.conv(cur, p["classifierf"], p["classifierb"])


# ------------------------------------------------------------------------------------------------
# synthetic inputs


def synth_faces(n, seed=0, size=224, dtype=np.float32):
    """224 x 224 x 3 x N: uniform[0,255) minus the per-channel average image."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, 255, (size, size, 3, n)).astype(dtype)
    return x - AVERAGE_IMAGE.reshape(1, 1, 3, 1).astype(dtype)


def synth_faces48(n, seed=0):
    """48 x 48 x N uint8 grey faces (FER layout, teacher/ferplus_baselines.m:181)."""
    return np.random.default_rng(seed).integers(0, 256, (48, 48, n), dtype=np.uint8)


def normalize_spectrogram(s):
    """emoVoxCeleb/getBatchEmoVoxCeleb.m:164-169: per frequency row (x - mean)/std over time, std with
    the N-1 normalisation (MATLAB default)."""
    mu = s.mean(axis=1, keepdims=True)
    sd = s.std(axis=1, ddof=1, keepdims=True)
    return (s - mu) / sd


def synth_spectrograms(n, width=300, seed=1, dtype=np.float32):
    rng = np.random.default_rng(seed)
    s = rng.standard_normal((512, width, 1, n))
    return normalize_spectrogram(s).astype(dtype)


def synth_teacher_logits(n, seed=2, dtype=np.float32):
    return (3.0 * np.random.default_rng(seed).standard_normal((1, 1, 8, n))).astype(dtype)


def bilinear_resize_hw(x, out_h, out_w):
    """Identity-affine vl_nnaffinegrid + vl_nnbilinearsampler (teacher/ferplus_baselines.m:203-213):
    the output grid spans [-1, 1] inclusive over the input (corner-aligned), bilinear taps.
    x: H x W x C x N."""
    H, W = x.shape[:2]
    ys = np.linspace(0, H - 1, out_h)
    xs = np.linspace(0, W - 1, out_w)
    y0 = np.clip(np.floor(ys).astype(int), 0, H - 1); y1 = np.clip(y0 + 1, 0, H - 1)
    x0 = np.clip(np.floor(xs).astype(int), 0, W - 1); x1 = np.clip(x0 + 1, 0, W - 1)
    wy = (ys - y0).reshape(-1, 1, 1, 1).astype(x.dtype)
    wx = (xs - x0).reshape(1, -1, 1, 1).astype(x.dtype)
    top = x[y0][:, x0] * (1 - wx) + x[y0][:, x1] * wx
    bot = x[y1][:, x0] * (1 - wx) + x[y1][:, x1] * wx
    return top * (1 - wy) + bot * wy


def faces48_to_input(u8, dtype=np.float32):
    """48 x 48 x N uint8 grey -> 224 x 224 x 3 x N teacher input: replicate to 3 channels, single,
    subtract averageImage (emoVoxCeleb/fetch_emovoxceleb_imdb.m:175-193), bilinear 48 -> 224."""
    g = u8.astype(dtype)[:, :, None, :]
    x = np.repeat(g, 3, axis=2) - AVERAGE_IMAGE.reshape(1, 1, 3, 1).astype(dtype)
    return bilinear_resize_hw(x, 224, 224)


def run_spec(speech, fs=16000, Tw=25, Ts=10, alpha=0.97):
    """VGGVox `runSpec` [UPSTREAM, restated from the public VGGVox / HTK-MFCC code it derives from; not in
    /root/reference]: called at emoVoxCeleb/getBatchEmoVoxCeleb.m:162 with the constants of
    emoVoxCeleb/run_distillation.m:109-117.  speech * 2^15 when max|speech| <= 1; pre-emphasis filter([1 -alpha], 1, .);
    vec2frames(Nw, Ns, hamming, no padding); abs(fft(frames, 2^nextpow2(Nw))) -> nfft x M (full spectrum: 512 rows)."""
    z = np.asarray(speech, np.float64).reshape(-1)
    if np.abs(z).max() <= 1:
        z = z * 2.0 ** 15
    Nw, Ns = int(round(1e-3 * Tw * fs)), int(round(1e-3 * Ts * fs))
    nfft = 1 << int(np.ceil(np.log2(Nw)))
    y = np.concatenate([z[:1], z[1:] - alpha * z[:-1]])
    M_ = (len(y) - Nw) // Ns + 1
    idx = np.arange(Nw)[:, None] + Ns * np.arange(M_)[None, :]
    win = 0.54 - 0.46 * np.cos(2 * np.pi * np.arange(Nw) / (Nw - 1))  # MATLAB hamming(Nw), symmetric
    return np.abs(np.fft.fft(y[idx] * win[:, None], nfft, axis=0))
