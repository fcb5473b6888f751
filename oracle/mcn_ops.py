"""CPU restatement of the MatConvNet / mcnExtraLayers operator semantics used on the hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this module; it is imported by
`tests/`, by `__graft_entry__.smoke()` and by `bench.py`'s cpu_baseline / `--impl reference` legs.

PARITY UNPINNED: the reference repository (albanie/mcnCrossModalEmotions) contains no arithmetic, no
tests and no golden vectors for this path -- every operator below lives in un-vendored, un-pinned
third-party MATLAB/MEX code (MatConvNet `vl_nn*`, mcnExtraLayers `vl_nnsoftmaxceloss`, ...; see
SURVEY.md section 8c).  Each function therefore restates the *published* semantics of the upstream
operator, cites the reference call site that reaches it, and is pinned by
hand-computed known-answer tests, numerical-gradient checks and an independent torch-CPU
cross-check in tests/test_oracle_*.py.

Conventions (MatConvNet): tensors are H x W x C x N (numpy arrays of that *logical* shape; memory
order is irrelevant here, the C ABI boundary uses column-major buffers), filters FH x FW x FC x K,
pad = [top bottom left right], stride = [sy sx].  Forward when `dzdy` is None, backward otherwise --
the same calling convention as the MATLAB functions.
"""
from __future__ import annotations

import numpy as np

# ----------------------------------------------------------------------------------------------
# helpers


def _pad4(pad):
    pad = np.atleast_1d(np.asarray(pad, dtype=np.int64))
    if pad.size == 1:
        return [int(pad[0])] * 4
    assert pad.size == 4, "pad must be scalar or [top bottom left right]"
    return [int(p) for p in pad]


def _stride2(stride):
    stride = np.atleast_1d(np.asarray(stride, dtype=np.int64))
    if stride.size == 1:
        return int(stride[0]), int(stride[0])
    return int(stride[0]), int(stride[1])


def out_size(h, w, fh, fw, pad, stride):
    """vl_nnconv / vl_nnpool output size: floor((H + pt + pb - FH)/sy) + 1."""
    pt, pb, pl, pr = _pad4(pad)
    sy, sx = _stride2(stride)
    return (h + pt + pb - fh) // sy + 1, (w + pl + pr - fw) // sx + 1


def _padded(x, pad, value=0.0):
    pt, pb, pl, pr = pad
    if pt == pb == pl == pr == 0:
        return x
    return np.pad(x, ((pt, pb), (pl, pr), (0, 0), (0, 0)), mode="constant", constant_values=value)


# ----------------------------------------------------------------------------------------------
# vl_nnconv -- reached through dagnn.Conv from dag.eval at
#   emoVoxCeleb/fetch_emovoxceleb_imdb.m:129, external/compute_visual_feats.m:90,
#   external/compute_audio_feats.m:126 and cnn_train_dag (emoVoxCeleb/run_distillation.m:170);
#   the block type is named at emoVoxCeleb/emoVoxZoo.m:118.


def vl_nnconv(x, f, b=None, dzdy=None, pad=0, stride=1):
    """Cross-correlation (no filter flip) with optional bias; groups = C / FC.

    forward : y[oh,ow,k,n] = b[k] + sum_{r,s,c} x[oh*sy+r-pt, ow*sx+s-pl, c, n] * f[r,s,c,k]
    backward: returns (dx, df, db); db = sum_{h,w,n} dzdy.
    Implemented the way the upstream CPU path works (im2row + GEMM), one filter tap at a time.
    """
    x = np.asarray(x)
    f = np.asarray(f)
    H, W, C, N = x.shape
    FH, FW, FC, K = f.shape
    pad = _pad4(pad)
    sy, sx = _stride2(stride)
    OH, OW = out_size(H, W, FH, FW, pad, (sy, sx))
    assert C % FC == 0, "filter depth must divide the input depth"
    groups = C // FC
    assert K % groups == 0
    kg = K // groups
    xp = _padded(x, pad)

    def tap(r, s):  # (OH, OW, C, N) view of the inputs seen by filter tap (r, s)
        return xp[r : r + (OH - 1) * sy + 1 : sy, s : s + (OW - 1) * sx + 1 : sx]

    if dzdy is None:
        y = np.zeros((OH, OW, K, N), dtype=x.dtype)
        for g in range(groups):
            cs, ks = slice(g * FC, (g + 1) * FC), slice(g * kg, (g + 1) * kg)
            for r in range(FH):
                for s in range(FW):
                    # (OH,OW,FC,N) x (FC,kg) -> (OH,OW,N,kg)
                    y[:, :, ks, :] += np.einsum("hwcn,ck->hwkn", tap(r, s)[:, :, cs, :], f[r, s, :, ks], optimize=True)
        if b is not None and np.size(b):
            y += np.asarray(b, dtype=x.dtype).reshape(1, 1, K, 1)
        return y

    dzdy = np.asarray(dzdy)
    assert dzdy.shape == (OH, OW, K, N)
    dxp = np.zeros_like(xp)
    df = np.zeros_like(f)
    for g in range(groups):
        cs, ks = slice(g * FC, (g + 1) * FC), slice(g * kg, (g + 1) * kg)
        for r in range(FH):
            for s in range(FW):
                df[r, s, :, ks] = np.einsum("hwcn,hwkn->ck", tap(r, s)[:, :, cs, :], dzdy[:, :, ks, :], optimize=True)
                dxp[r : r + (OH - 1) * sy + 1 : sy, s : s + (OW - 1) * sx + 1 : sx, cs, :] += np.einsum(
                    "hwkn,ck->hwcn", dzdy[:, :, ks, :], f[r, s, :, ks], optimize=True
                )
    pt, pb, pl, pr = pad
    dx = dxp[pt : pt + H, pl : pl + W]
    db = dzdy.sum(axis=(0, 1, 3)) if (b is not None and np.size(b)) else None
    return dx, df, db


# ----------------------------------------------------------------------------------------------
# vl_nnpool -- dagnn.Pooling; `pool6` is re-sized at emoVoxCeleb/emoVoxZoo.m:263-269 and
#   external/compute_audio_feats.m:121-125.


def vl_nnpool(x, pool, dzdy=None, pad=0, stride=1, method="max", return_index=False, index=None):
    """Max / average pooling.

    max: padding behaves as -inf.  The backward pass routes dzdy to the FIRST element attaining the
         maximum when the window is scanned in memory order (column-major: w outer, h inner, strict
         `>` update) -- the rule behind "bit-exact pooling indices" in BASELINE.json's north_star.
    avg: divides by the number of in-bounds elements of the window.
    `return_index` (forward, max only) additionally returns the window-local arg-max index
    idx = dw * PH + dh as uint8 -- the quantity the CUDA kernels emit and the tests compare exactly.
    `index` (backward, max only; test instrumentation): route dzdy through THESE window-local indices instead of the
    arg-max of x -- the oracle's arithmetic under another implementation's discrete choices.
    """
    x = np.asarray(x)
    H, W, C, N = x.shape
    pool = np.atleast_1d(np.asarray(pool, dtype=np.int64))
    PH, PW = (int(pool[0]), int(pool[0])) if pool.size == 1 else (int(pool[0]), int(pool[1]))
    pad = _pad4(pad)
    pt, pb, pl, pr = pad
    sy, sx = _stride2(stride)
    OH, OW = out_size(H, W, PH, PW, pad, (sy, sx))
    if method == "max":
        xp = _padded(x, pad, value=-np.inf)
        best = np.full((OH, OW, C, N), -np.inf, dtype=x.dtype)
        arg = np.zeros((OH, OW, C, N), dtype=np.uint8)
        for dw in range(PW):  # memory-order scan: w outer, h inner
            for dh in range(PH):
                v = xp[dh : dh + (OH - 1) * sy + 1 : sy, dw : dw + (OW - 1) * sx + 1 : sx]
                upd = v > best
                best = np.where(upd, v, best)
                arg = np.where(upd, np.uint8(dw * PH + dh), arg)
        if dzdy is None:
            return (best, arg) if return_index else best
        dzdy = np.asarray(dzdy)
        if index is not None:
            arg = np.asarray(index).reshape(arg.shape)
        dxp = np.zeros(xp.shape, dtype=dzdy.dtype)
        for dw in range(PW):
            for dh in range(PH):
                sel = arg == (dw * PH + dh)
                dxp[dh : dh + (OH - 1) * sy + 1 : sy, dw : dw + (OW - 1) * sx + 1 : sx] += np.where(sel, dzdy, 0)
        return dxp[pt : pt + H, pl : pl + W]
    if method == "avg":
        xp = _padded(x, pad, value=0.0)
        ones = _padded(np.ones((H, W, 1, 1), dtype=x.dtype), pad, value=0.0)
        acc = np.zeros((OH, OW, C, N), dtype=x.dtype)
        cnt = np.zeros((OH, OW, 1, 1), dtype=x.dtype)
        for dw in range(PW):
            for dh in range(PH):
                acc += xp[dh : dh + (OH - 1) * sy + 1 : sy, dw : dw + (OW - 1) * sx + 1 : sx]
                cnt += ones[dh : dh + (OH - 1) * sy + 1 : sy, dw : dw + (OW - 1) * sx + 1 : sx]
        if dzdy is None:
            return acc / cnt
        dzdy = np.asarray(dzdy)
        dxp = np.zeros(xp.shape, dtype=dzdy.dtype)
        g = dzdy / cnt
        for dw in range(PW):
            for dh in range(PH):
                dxp[dh : dh + (OH - 1) * sy + 1 : sy, dw : dw + (OW - 1) * sx + 1 : sx] += g
        return dxp[pt : pt + H, pl : pl + W]
    raise ValueError("unknown pooling method %r" % (method,))


# ----------------------------------------------------------------------------------------------
# vl_nnbnorm -- dagnn.BatchNorm (student graph: train mode under cnn_train_dag,
#   emoVoxCeleb/run_distillation.m:170; test mode after `dag.mode = 'test'` at
#   emoVoxCeleb/fetch_emovoxceleb_imdb.m:107, external/compute_visual_feats.m:58).


def vl_nnbnorm(x, g, b, dzdy=None, epsilon=1e-4, moments=None):
    """Batch normalisation over (H, W, N) per channel.

    mu = sum(x)/M, v = sum((x-mu)^2)/M (biased), sigma = sqrt(v + eps), y = g*(x-mu)/sigma + b.
    `moments` is C x 2 = [mu sigma] (sigma, NOT variance); when given (test mode) it is used as is.
    forward returns (y, moments); backward returns (dx, dg, db, moments).  In test mode (`moments`
    given) the backward treats mu/sigma as constants, as upstream does.
    """
    x = np.asarray(x)
    H, W, C, N = x.shape
    M = H * W * N
    g = np.asarray(g, dtype=x.dtype).reshape(1, 1, C, 1)
    b = np.asarray(b, dtype=x.dtype).reshape(1, 1, C, 1)
    given = moments is not None
    if given:
        moments = np.asarray(moments, dtype=x.dtype).reshape(C, 2)
        mu = moments[:, 0].reshape(1, 1, C, 1)
        sigma = moments[:, 1].reshape(1, 1, C, 1)
    else:
        mu = x.sum(axis=(0, 1, 3), keepdims=True) / M
        var = ((x - mu) ** 2).sum(axis=(0, 1, 3), keepdims=True) / M
        sigma = np.sqrt(var + epsilon)
        moments = np.concatenate([mu.reshape(C, 1), sigma.reshape(C, 1)], axis=1)
    xhat = (x - mu) / sigma
    if dzdy is None:
        return g * xhat + b, moments
    dzdy = np.asarray(dzdy)
    db = dzdy.sum(axis=(0, 1, 3))
    dg = (dzdy * xhat).sum(axis=(0, 1, 3))
    if given:
        dx = dzdy * (g / sigma)
    else:
        dx = (g / sigma) * (dzdy - db.reshape(1, 1, C, 1) / M - xhat * dg.reshape(1, 1, C, 1) / M)
    return dx, dg, db, moments


# ----------------------------------------------------------------------------------------------
# element-wise blocks living inside the .mat graphs loaded at emoVoxCeleb/emoVoxZoo.m:44


def vl_nnrelu(x, dzdy=None, leak=0.0):
    x = np.asarray(x)
    if dzdy is None:
        return np.where(x > 0, x, leak * x) if leak else np.maximum(x, 0)
    return np.asarray(dzdy) * np.where(x > 0, 1.0, leak).astype(x.dtype)


def vl_nnsigmoid(x, dzdy=None):
    y = 1.0 / (1.0 + np.exp(-np.asarray(x)))
    if dzdy is None:
        return y
    return np.asarray(dzdy) * y * (1.0 - y)


def vl_nnsum(inputs, dzdy=None):
    """dagnn.Sum: element-wise sum of its inputs; the backward copies dzdy to every input."""
    if dzdy is None:
        out = np.array(inputs[0], copy=True)
        for t in inputs[1:]:
            out = out + t
        return out
    return [np.asarray(dzdy) for _ in inputs]


def vl_nnglobalpool(x, dzdy=None, method="avg"):
    """mcnExtraLayers global pooling (SE squeeze): mean (or max) over H x W -> 1 x 1 x C x N."""
    x = np.asarray(x)
    H, W, C, N = x.shape
    if method != "avg":
        raise ValueError("only the 'avg' squeeze is on the hot path")
    if dzdy is None:
        return x.mean(axis=(0, 1), keepdims=True)
    return np.broadcast_to(np.asarray(dzdy) / (H * W), x.shape).copy()


def vl_nnaxpy(a, x, y, dzdy=None):
    """mcnExtraLayers Axpy (SE excite + shortcut): out = a (.) x + y with a (1x1xCxN) broadcast over HxW."""
    a, x, y = np.asarray(a), np.asarray(x), np.asarray(y)
    if dzdy is None:
        return a * x + y
    dzdy = np.asarray(dzdy)
    return (dzdy * x).sum(axis=(0, 1), keepdims=True), dzdy * a, dzdy


def vl_nnscale(x, a, dzdy=None):
    """mcnExtraLayers Scale as used by SE graphs: out = a (.) x, a broadcast over H x W."""
    a, x = np.asarray(a), np.asarray(x)
    if dzdy is None:
        return a * x
    dzdy = np.asarray(dzdy)
    return dzdy * a, (dzdy * x).sum(axis=(0, 1), keepdims=True)


# ----------------------------------------------------------------------------------------------
# softmax family -- vl_nnsoftmaxt is called at emoVoxCeleb/student_stats.m:95; the distillation loss
#   dagnn.SoftmaxCELoss('temperature', 2, 'logitTargets', true) is attached at
#   emoVoxCeleb/emoVoxZoo.m:151-157.


def vl_nnsoftmaxt(x, dim=3, temperature=1.0, dzdy=None):
    """Numerically stable softmax along MATLAB dimension `dim` (1-based; 3 = channels)."""
    x = np.asarray(x)
    ax = dim - 1
    z = x / temperature
    z = z - z.max(axis=ax, keepdims=True)
    e = np.exp(z)
    y = e / e.sum(axis=ax, keepdims=True)
    if dzdy is None:
        return y
    dzdy = np.asarray(dzdy)
    return y * (dzdy - (dzdy * y).sum(axis=ax, keepdims=True)) / temperature


def _log_softmax(z, ax):
    z = z - z.max(axis=ax, keepdims=True)
    return z - np.log(np.exp(z).sum(axis=ax, keepdims=True))


def vl_nnsoftmaxceloss(x, p, dzdy=None, temperature=1.0, logitTargets=False, instanceWeights=None, tol=1e-5):
    """Temperature-softmax cross-entropy against (soft) targets.

    p <- softmax_3(p/T) when `logitTargets`; q = softmax_3(x/T);
    forward : y = sum_n w_n * ( - sum_c p * log q )          (summed over the batch, not averaged --
              cnn_train_dag divides by the batch size in its update)
    backward: dx = dzdy * w_n * (q - p) / T                   (no T^2 rescaling)
    """
    x = np.asarray(x)
    p = np.asarray(p, dtype=x.dtype)
    T = float(temperature)
    if logitTargets:
        p = vl_nnsoftmaxt(p, dim=3, temperature=T)
    s = p.sum(axis=2)
    assert np.all(np.abs(s - 1) < tol), "targets must be distributions over dim 3"
    w = 1.0 if instanceWeights is None else np.asarray(instanceWeights, dtype=x.dtype).reshape(1, 1, 1, -1)
    logq = _log_softmax(x / T, 2)
    if dzdy is None:
        return (w * (-(p * logq).sum(axis=2, keepdims=True))).sum()
    q = np.exp(logq)
    return np.asarray(dzdy) * w * (q - p) / T


def vl_nneuclideanloss(x, t, dzdy=None, instanceWeights=None):
    """dagnn.EuclideanLoss -> vl_nneuclideanloss [UPSTREAM mcnExtraLayers, restated from recollection; wired at
    /root/reference/emoVoxCeleb/emoVoxZoo.m:138-144 on {prediction, logitTarget, instanceWeights}].
    forward : y = 1/2 * sum(w .* (x - t).^2)      (w broadcast over dims 1-3: getBatchEmoVoxCeleb.m:37 passes 1 x 1 x 1 x N)
    backward: dx = dzdy * w .* (x - t)"""
    x = np.asarray(x)
    d = x - np.asarray(t, dtype=x.dtype)
    w = 1.0 if instanceWeights is None else np.asarray(instanceWeights, dtype=x.dtype).reshape(1, 1, 1, -1)
    if dzdy is None:
        return 0.5 * (w * d * d).sum()
    return np.asarray(dzdy) * w * d


def vl_nnhuberloss(x, t, dzdy=None, sigma=1.0, instanceWeights=None):
    """dagnn.HuberLoss('sigma', s) -> vl_nnhuberloss [UPSTREAM mcnExtraLayers, restated from recollection: the smooth-L1
    of Fast R-CNN; wired at /root/reference/emoVoxCeleb/emoVoxZoo.m:145-147].  With s2 = sigma^2 and d = x - t:
    forward : y = sum w .* ( |d| - 0.5/s2  where |d| > 1/s2,  0.5*s2*d^2 elsewhere )
    backward: dx = dzdy * w .* ( sign(d)   where |d| > 1/s2,  s2*d       elsewhere )"""
    x = np.asarray(x)
    d = x - np.asarray(t, dtype=x.dtype)
    w = 1.0 if instanceWeights is None else np.asarray(instanceWeights, dtype=x.dtype).reshape(1, 1, 1, -1)
    s2 = float(sigma) ** 2
    lin = np.abs(d) > 1.0 / s2
    if dzdy is None:
        return (w * np.where(lin, np.abs(d) - 0.5 / s2, 0.5 * s2 * d * d)).sum()
    return np.asarray(dzdy) * w * np.where(lin, np.sign(d), s2 * d)


def vl_nnloss(x, c, dzdy=None, loss="classerror"):
    """The two vl_nnloss modes the reference can attach (emoVoxCeleb/emoVoxZoo.m:147-149,160-163).

    classerror: sum_n [argmax_c x != c_n] (labels are 1-based, first maximum wins); no gradient.
    softmaxlog: sum_n ( logsumexp(x) - x[c_n] ); gradient softmax(x) - onehot(c).
    """
    x = np.asarray(x)
    H, W, C, N = x.shape
    c = np.asarray(c).reshape(H, W, 1, N).astype(np.int64)
    if loss == "classerror":
        pred = x.argmax(axis=2)[:, :, None, :] + 1
        if dzdy is None:
            return float((pred != c).sum())
        return np.zeros_like(x)
    if loss == "softmaxlog":
        logq = _log_softmax(x, 2)
        onehot = np.zeros_like(x)
        np.put_along_axis(onehot, c - 1, 1.0, axis=2)
        if dzdy is None:
            return float(-(onehot * logq).sum())
        return np.asarray(dzdy) * (np.exp(logq) - onehot)
    raise ValueError("unsupported loss %r" % (loss,))


def error_stats(x, c, num_classes):
    """Per-class accuracy counters of the reference's dagnn.ErrorStats metric layer
    (emoVoxCeleb/emoVoxZoo.m:166-169; fields read at emoVoxCeleb/run_distillation.m:190-192).
    Returns (correct_per_class, count_per_class)."""
    x = np.asarray(x)
    pred = x.argmax(axis=2).reshape(-1) + 1
    c = np.asarray(c).reshape(-1).astype(np.int64)
    correct = np.zeros(num_classes, dtype=np.int64)
    count = np.zeros(num_classes, dtype=np.int64)
    for k in range(1, num_classes + 1):
        sel = c == k
        count[k - 1] = sel.sum()
        correct[k - 1] = (pred[sel] == k).sum()
    return correct, count
