"""The MatConvNet / mcnExtraLayers operator surface (`vl_nn*`) on top of the C ABI.

Same names, argument order, option names and forward/backward convention as the upstream MATLAB
functions that dagnn blocks call from `dag.eval` in the reference
(emoVoxCeleb/fetch_emovoxceleb_imdb.m:129, external/compute_visual_feats.m:90,
external/compute_audio_feats.m:126, cnn_train_dag at emoVoxCeleb/run_distillation.m:170;
vl_nnsoftmaxt at emoVoxCeleb/student_stats.m:95; the loss at emoVoxCeleb/emoVoxZoo.m:151-157):
forward when `dzdy` is None, backward otherwise; tensors are `single`, H x W x C x N.

Arrays may be numpy arrays (the CPU-array case of MATLAB: staged through the device by the library)
or `GpuArray`s (MATLAB's gpuArray: the data stays on the device, outputs are GpuArrays).  All
arithmetic happens in libxemo.so on the GPU; nothing here computes on the CPU.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import I2, I4, XemoArray

_ctx = None


def default_context():
    """One context on cuda:0 for the boundary operators (created on first use)."""
    global _ctx
    if _ctx is None:
        _ctx = _lib.Context(0)
    return _ctx


class GpuArray:
    """Minimal gpuArray: a device buffer holding a column-major H x W x C x N single array."""

    def __init__(self, shape, tensor):
        self.shape = tuple(int(s) for s in shape)
        self.tensor = tensor  # flat torch.float32 CUDA tensor, column-major order

    @property
    def ptr(self):
        return self.tensor.data_ptr()


def gpuArray(a):
    import torch

    a = np.asarray(a, dtype=np.float32)
    shape = _shape4(a.shape)
    flat = np.ascontiguousarray(a.reshape(shape).transpose(3, 2, 1, 0)).reshape(-1)  # column-major order
    t = torch.from_numpy(flat).cuda()
    torch.cuda.current_stream().synchronize()
    return GpuArray(shape, t)


def gather(a):
    if isinstance(a, GpuArray):
        default_context().sync()
        flat = a.tensor.cpu().numpy()
        h, w, c, n = a.shape
        return flat.reshape(n, c, w, h).transpose(3, 2, 1, 0)
    return a


def _shape4(shape):
    shape = tuple(int(s) for s in shape)
    return shape + (1,) * (4 - len(shape))


class _Marshal:
    """Keeps the column-major staging buffers of one call alive and builds xemo_array structs."""

    def __init__(self):
        self.keep = []
        self.outputs = []
        self.on_gpu = False

    def arr(self, a):
        if a is None:
            return None
        if isinstance(a, GpuArray):
            self.on_gpu = True
            h, w, c, n = a.shape
            s = XemoArray(C.c_void_p(a.ptr), h, w, c, n)
            self.keep.append((a, s))
            return C.byref(s)
        a = np.asarray(a, dtype=np.float32)
        shape = _shape4(a.shape)
        buf = np.ascontiguousarray(a.reshape(shape).transpose(3, 2, 1, 0))  # memory order = column-major HWCN
        s = XemoArray(buf.ctypes.data_as(C.c_void_p), *shape)
        self.keep.append((buf, s))
        return C.byref(s)

    def vec(self, v):
        """float vector argument (host or device)."""
        if v is None:
            return None
        if isinstance(v, GpuArray):
            self.keep.append(v)
            return C.c_void_p(v.ptr)
        buf = np.ascontiguousarray(np.asarray(v, dtype=np.float32).reshape(-1, order="F"))
        self.keep.append(buf)
        return buf.ctypes.data_as(C.c_void_p)

    def out(self, shape, dtype=np.float32):
        shape = _shape4(shape)
        if self.on_gpu:
            import torch

            t = torch.empty(int(np.prod(shape)), dtype=torch.float32 if dtype == np.float32 else torch.uint8, device="cuda")
            torch.cuda.current_stream().synchronize()
            g = GpuArray(shape, t)
            s = XemoArray(C.c_void_p(g.ptr), *shape)
            self.keep.append((g, s))
            self.outputs.append(g)
            return C.byref(s), len(self.outputs) - 1
        buf = np.empty(shape[::-1], dtype=dtype)  # (N, C, W, H) C-order == column-major HWCN
        s = XemoArray(buf.ctypes.data_as(C.c_void_p), *shape)
        self.keep.append((buf, s))
        self.outputs.append(buf)
        return C.byref(s), len(self.outputs) - 1

    def result(self, i):
        o = self.outputs[i]
        if isinstance(o, GpuArray):
            return o
        return o.transpose(3, 2, 1, 0)  # logical H x W x C x N view


def _pad4(pad):
    pad = np.atleast_1d(np.asarray(pad, dtype=np.int64))
    if pad.size == 1:
        pad = np.repeat(pad, 4)
    if pad.size != 4:
        raise ValueError("pad must be a scalar or [top bottom left right]")
    return I4(*[int(p) for p in pad])


def _pair(v, what):
    v = np.atleast_1d(np.asarray(v, dtype=np.int64))
    if v.size == 1:
        v = np.repeat(v, 2)
    if v.size != 2:
        raise ValueError("%s must be a scalar or a pair" % what)
    return I2(int(v[0]), int(v[1]))


def _shape_of(a):
    return a.shape if isinstance(a, GpuArray) else _shape4(np.shape(a))


def _out_size(h, w, fh, fw, pad, stride):
    return (h + pad[0] + pad[1] - fh) // stride[0] + 1, (w + pad[2] + pad[3] - fw) // stride[1] + 1


# ------------------------------------------------------------------------------------------------


def vl_nnconv(x, f, b=None, dzdy=None, pad=0, stride=1, dilate=1):
    """Y = vl_nnconv(X, F, B, 'pad', P, 'stride', S)  /  [DX, DF, DB] = vl_nnconv(X, F, B, DZDY, ...)."""
    if np.any(np.asarray(dilate) != 1):
        raise ValueError("vl_nnconv: dilation is not used on this path and is not supported")
    ctx = default_context()
    m = _Marshal()
    pad, stride = _pad4(pad), _pair(stride, "stride")
    H, W, Cc, N = _shape_of(x)
    FH, FW, FC, K = _shape_of(f)
    has_b = b is not None and (isinstance(b, GpuArray) or np.size(b) > 0)
    xa, fa = m.arr(x), m.arr(f)
    ba = m.arr(b if isinstance(b, GpuArray) else np.asarray(b, np.float32).reshape(K, 1, 1, 1)) if has_b else None
    OH, OW = _out_size(H, W, FH, FW, pad, stride)
    if dzdy is None:
        ya, yi = m.out((OH, OW, K, N))
        ctx.vl_nnconv(xa, fa, ba, None, pad, stride, ya, None, None, None)
        return m.result(yi)
    dya = m.arr(dzdy)
    dxa, dxi = m.out((H, W, Cc, N))
    dfa, dfi = m.out((FH, FW, FC, K))
    dba, dbi = m.out((K, 1, 1, 1)) if has_b else (None, None)
    ctx.vl_nnconv(xa, fa, ba, dya, pad, stride, None, dxa, dfa, dba)
    db = None
    if has_b:
        db = m.result(dbi)
        db = db if isinstance(db, GpuArray) else db.reshape(K)
    return m.result(dxi), m.result(dfi), db


def vl_nnpool(x, pool, dzdy=None, pad=0, stride=1, method="max", return_index=False):
    """Y = vl_nnpool(X, POOL, 'pad', P, 'stride', S, 'method', M)  /  DX = vl_nnpool(X, POOL, DZDY, ...).
    `return_index` (forward, max) also returns the uint8 window-local arg-max dw*PH + dh."""
    if method not in ("max", "avg"):
        raise ValueError("vl_nnpool: unknown method %r" % (method,))
    ctx = default_context()
    m = _Marshal()
    pad, stride, pool = _pad4(pad), _pair(stride, "stride"), _pair(pool, "pool")
    H, W, Cc, N = _shape_of(x)
    xa = m.arr(x)
    OH, OW = _out_size(H, W, pool[0], pool[1], pad, stride)
    meth = 0 if method == "max" else 1
    if dzdy is None:
        ya, yi = m.out((OH, OW, Cc, N))
        if not (return_index and meth == 0):
            ctx.vl_nnpool(xa, pool, None, pad, stride, meth, ya, None)
            return m.result(yi)
        if m.on_gpu:       # gpuArray in: the uint8 indices are produced on the device and gathered
            import torch

            idx_t = torch.empty(N * Cc * OW * OH, dtype=torch.uint8, device="cuda")
            torch.cuda.current_stream().synchronize()
            ctx.vl_nnpool(xa, pool, None, pad, stride, meth, ya, C.c_void_p(idx_t.data_ptr()))
            ctx.sync()
            return m.result(yi), idx_t.cpu().numpy().reshape(N, Cc, OW, OH).transpose(3, 2, 1, 0)
        idx = np.empty((N, Cc, OW, OH), dtype=np.uint8)
        ctx.vl_nnpool(xa, pool, None, pad, stride, meth, ya, idx.ctypes.data_as(C.c_void_p))
        return m.result(yi), idx.transpose(3, 2, 1, 0)
    dya = m.arr(dzdy)
    dxa, dxi = m.out((H, W, Cc, N))
    ctx.vl_nnpool(xa, pool, dya, pad, stride, meth, dxa, None)
    return m.result(dxi)


def vl_nnbnorm(x, g, b, dzdy=None, epsilon=1e-4, moments=None):
    """[Y, MOMENTS] = vl_nnbnorm(X, G, B, 'epsilon', E [, 'moments', M])  /
    [DX, DG, DB, MOMENTS] = vl_nnbnorm(X, G, B, DZDY, ...).  MOMENTS is C x 2 = [mu sigma]."""
    ctx = default_context()
    m = _Marshal()
    H, W, Cc, N = _shape_of(x)
    xa = m.arr(x)
    gv, bv = m.vec(g), m.vec(b)
    mi = None
    if moments is not None:
        mom = np.asarray(moments, np.float32).reshape(Cc, 2)
        mi = m.vec(np.concatenate([mom[:, 0], mom[:, 1]]))  # column-major C x 2
    mo = np.empty(2 * Cc, np.float32)
    oa, oi = m.out((H, W, Cc, N))
    if dzdy is None:
        ctx.vl_nnbnorm(xa, gv, bv, None, float(epsilon), mi, oa, None, None, mo.ctypes.data_as(C.c_void_p))
        return m.result(oi), mo.reshape(2, Cc).T.copy()
    dya = m.arr(dzdy)
    dg, db = np.empty(Cc, np.float32), np.empty(Cc, np.float32)
    ctx.vl_nnbnorm(xa, gv, bv, dya, float(epsilon), mi, oa, dg.ctypes.data_as(C.c_void_p), db.ctypes.data_as(C.c_void_p),
                   mo.ctypes.data_as(C.c_void_p))
    return m.result(oi), dg, db, mo.reshape(2, Cc).T.copy()


def vl_nnrelu(x, dzdy=None, leak=0.0):
    ctx = default_context()
    m = _Marshal()
    xa = m.arr(x)
    dya = m.arr(dzdy)
    oa, oi = m.out(_shape_of(x))
    ctx.vl_nnrelu(xa, dya, float(leak), oa)
    return m.result(oi)


def vl_nnsigmoid(x, dzdy=None):
    ctx = default_context()
    m = _Marshal()
    xa = m.arr(x)
    dya = m.arr(dzdy)
    oa, oi = m.out(_shape_of(x))
    ctx.vl_nnsigmoid(xa, dya, oa)
    return m.result(oi)


def vl_nnsoftmaxt(x, dim=3, temperature=1.0):
    """Y = vl_nnsoftmaxt(X, 'dim', D): softmax along MATLAB dimension D (only D = 3, channels, is used by
    the reference: emoVoxCeleb/student_stats.m:95)."""
    if dim != 3:
        raise ValueError("vl_nnsoftmaxt: only 'dim', 3 is supported")
    ctx = default_context()
    m = _Marshal()
    xa = m.arr(x)
    oa, oi = m.out(_shape_of(x))
    ctx.vl_nnsoftmaxt(xa, float(temperature), oa)
    return m.result(oi)


def vl_nnsoftmaxceloss(x, p, dzdy=None, temperature=1.0, logitTargets=False, instanceWeights=None, tol=1e-5):
    """Y = vl_nnsoftmaxceloss(X, P, 'temperature', T, 'logitTargets', tf, 'instanceWeights', W) and its
    derivative DX = vl_nnsoftmaxceloss(X, P, DZDY, ...) (mcnExtraLayers; wired at emoVoxCeleb/emoVoxZoo.m:151-157)."""
    ctx = default_context()
    m = _Marshal()
    shape = _shape_of(x)
    if _shape_of(p) != shape:
        raise ValueError("vl_nnsoftmaxceloss: X and P must have the same size")
    if not logitTargets and not isinstance(p, GpuArray):
        s = np.asarray(p, np.float64).reshape(shape).sum(axis=2)
        if np.any(np.abs(s - 1) > tol):
            raise ValueError("vl_nnsoftmaxceloss: targets must sum to one along dimension 3")
    xa, pa = m.arr(x), m.arr(p)
    w = m.vec(instanceWeights)
    if dzdy is None:
        loss = np.zeros(1, np.float32)
        ctx.vl_nnsoftmaxceloss(xa, pa, None, float(temperature), int(bool(logitTargets)), w, loss.ctypes.data_as(C.c_void_p), None)
        return float(loss[0])
    dz = np.asarray(dzdy, np.float32).reshape(-1)[:1].copy()
    oa, oi = m.out(shape)
    ctx.vl_nnsoftmaxceloss(xa, pa, dz.ctypes.data_as(C.c_void_p), float(temperature), int(bool(logitTargets)), w, None, oa)
    return m.result(oi)


def vl_nnloss(x, c, dzdy=None, loss="classerror"):
    """vl_nnloss(X, c, [], 'loss', 'classerror'): number of samples whose arg-max differs from the
    1-based label (the metric layer at emoVoxCeleb/emoVoxZoo.m:160-163)."""
    if loss != "classerror":
        raise ValueError("vl_nnloss: only 'classerror' is on this path")
    if dzdy is not None:
        return np.zeros(_shape_of(x), np.float32)
    ctx = default_context()
    m = _Marshal()
    xa = m.arr(x)
    lab = m.vec(np.asarray(c, np.float32))
    nerr = np.zeros(1, np.float32)
    ctx.vl_nnloss_classerror(xa, lab, nerr.ctypes.data_as(C.c_void_p))
    return float(nerr[0])


def vl_nnglobalpool(x, dzdy=None):
    ctx = default_context()
    m = _Marshal()
    H, W, Cc, N = _shape_of(x)
    xa = m.arr(x)
    dya = m.arr(dzdy)
    oa, oi = m.out((1, 1, Cc, N) if dzdy is None else (H, W, Cc, N))
    ctx.vl_nnglobalpool(xa, dya, oa)
    return m.result(oi)


def vl_nnaxpy(a, x, y):
    ctx = default_context()
    m = _Marshal()
    aa, xa, ya = m.arr(a), m.arr(x), m.arr(y)
    oa, oi = m.out(_shape_of(x))
    ctx.vl_nnaxpy(aa, xa, ya, oa)
    return m.result(oi)
