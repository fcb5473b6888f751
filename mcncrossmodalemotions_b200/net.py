"""Thin ctypes caller of the graph-level C ABI (include/xemo.h section C; csrc/xemo_net.cu).

Everything that matters happens inside libxemo.so: the network lives on the device, the kernels are sequenced and
captured in CUDA graphs by the library, the gradient exchange is ncclAllReduce issued by the library inside the
captured step.  This module only marshals: zoo parameter dictionaries -> MatConvNet-layout host arrays -> xemo_net_set_param,
inputs in, logits / metrics out -- exactly what the MEX shim mex/xemo_dagnn_mex.c does for a MATLAB host.  It does not
import programs.py (the Python-side assembly of the same kernels, kept for per-operator profiling and experiments).

Reference call sites replaced (all `file:line` under /root/reference):
  TeacherNet.forward        dag.eval at emoVoxCeleb/fetch_emovoxceleb_imdb.m:129, external/compute_visual_feats.m:90
  StudentNet.forward        dag.eval at external/compute_audio_feats.m:126
  StudentNet.train_step     one cnn_train_dag iteration, emoVoxCeleb/run_distillation.m:170-182 (loss emoVoxZoo.m:137-157)
  DistillStep.step          teacher forward + coupling (getBatchEmoVoxCeleb.m:133-188) + student step, one graph replay
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

KINDS = {"resnet50": 0, "senet50": 1, "vggvox": 2}
LOSSES = {"hot-cross-ent": 0, "softmaxlog": 1, "euclidean": 2, "huber": 3}
VP = C.c_void_p


def _colmajor(name, v):
    """zoo value -> the column-major MatConvNet array the C ABI takes (filters FH x FW x FC x K; moments C x 2 = [mu sigma])."""
    v = np.asarray(v, np.float32)
    if v.ndim <= 1:
        return np.ascontiguousarray(v.reshape(-1))
    return np.ascontiguousarray(np.transpose(v, tuple(range(v.ndim))[::-1])).reshape(-1)


def _from_colmajor(flat, dims):
    dims = [int(d) for d in dims]
    while len(dims) > 1 and dims[-1] == 1:
        dims.pop()
    if len(dims) == 1:
        return flat.copy()
    return np.ascontiguousarray(np.transpose(flat.reshape(dims[::-1]), tuple(range(len(dims)))[::-1]))


def _ptr(x):
    """host numpy array / torch tensor (host or device) / int address -> void*"""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x.ctypes.data_as(VP)
    if hasattr(x, "data_ptr"):
        return VP(x.data_ptr())
    return VP(int(x))


class Comm:
    """Data-parallel communicator of the library (NCCL resolved inside libxemo at run time).  `broadcast(bytes_or_None)`
    distributes rank 0's 128-byte unique id -- any host mechanism will do; `Comm.from_torch()` uses torch.distributed."""

    def __init__(self, ctx, rank, world, broadcast):
        self.ctx, self.rank, self.world = ctx, rank, world
        ident = (C.c_char * 128)()
        if rank == 0:
            rc = ctx.lib.xemo_comm_unique_id(ident)
            if rc:
                raise _lib.XemoError(rc, "NCCL is not available to libxemo (libnccl.so.2 could not be resolved)")
        raw = broadcast(bytes(ident.raw) if rank == 0 else None)
        h = VP()
        rc = ctx.lib.xemo_comm_create(ctx.handle, C.c_char_p(raw), rank, world, C.byref(h))
        if rc:
            raise _lib.XemoError(rc, ctx.lib.xemo_last_error(ctx.handle).decode())
        self.handle = h
        ctx.adopt(self)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @classmethod
    def from_torch(cls, ctx):
        import torch
        import torch.distributed as dist

        def bcast(raw):
            t = torch.zeros(128, dtype=torch.uint8)
            if raw is not None:
                t = torch.frombuffer(bytearray(raw), dtype=torch.uint8).clone()
            dev = t.cuda() if dist.get_backend() == "nccl" else t
            dist.broadcast(dev, src=0)
            return bytes(dev.cpu().numpy().tobytes())

        return cls(ctx, dist.get_rank(), dist.get_world_size(), bcast)

    def close(self):
        """(close the networks whose captured steps used this communicator first)"""
        if getattr(self, "handle", None) and getattr(self.ctx, "handle", None):
            self.ctx.lib.xemo_comm_destroy(self.handle)
        self.handle = None


class Net:
    def __init__(self, kind, params, batch, size, input_mode=0, num_outputs=8, ctx=None, device=0, stream=None):
        self.ctx = ctx or _lib.Context(device, stream)
        self.lib = self.ctx.lib
        self.kind, self.N, self.K = kind, int(batch), int(num_outputs)
        h = VP()
        self._check(self.lib.xemo_net_create(self.ctx.handle, KINDS[kind], self.N, int(size), int(input_mode), self.K, C.byref(h)))
        self.handle = h
        self.ctx.adopt(self)
        self.names = [self.lib.xemo_net_param_name(h, i).decode() for i in range(self.lib.xemo_net_num_params(h))]
        self.dims = {}
        for name in self.names:
            d = (C.c_int64 * 4)()
            self._check(self.lib.xemo_net_param_dims(h, name.encode(), d))
            self.dims[name] = tuple(d)
        missing = [k for k in self.names if k not in params]
        if missing:
            raise KeyError("parameters missing from the dictionary: %s" % missing[:5])
        for name in self.names:
            self.set_param(name, params[name])
        self._check(self.lib.xemo_net_finalize(h))
        nb = C.c_size_t()
        self.lib.xemo_net_input_bytes(h, C.byref(nb))
        self.input_bytes = nb.value

    def _check(self, rc):
        if rc:
            raise _lib.XemoError(rc, self.lib.xemo_last_error(self.ctx.handle).decode())

    def set_param(self, name, value):
        flat = _colmajor(name, value)
        want = int(np.prod(self.dims[name]))
        if flat.size != want:
            raise ValueError("%s: expected %s (%d elements), got shape %s" % (name, self.dims[name], want, np.shape(value)))
        self._check(self.lib.xemo_net_set_param(self.handle, name.encode(), flat.ctypes.data_as(VP), flat.size))

    def set_input(self, x):
        """numpy in the logical MatConvNet shape (H x W x C x N / H x W x N uint8), or a flat host / device tensor already in
        column-major order."""
        if isinstance(x, np.ndarray) and x.ndim > 1:
            x = np.ascontiguousarray(np.transpose(x, tuple(range(x.ndim))[::-1]))
        nbytes = x.nbytes if isinstance(x, np.ndarray) else x.numel() * x.element_size()
        self._keep = x
        self._check(self.lib.xemo_net_set_input(self.handle, _ptr(x), nbytes))

    def buffer(self, name):
        return self.lib.xemo_net_buffer(self.handle, name.encode())

    def num_kernels(self):
        return self.lib.xemo_net_num_kernels(self.handle)

    def sync(self):
        self.ctx.sync()

    def close(self):
        if getattr(self, "handle", None) and getattr(self.ctx, "handle", None):
            self.lib.xemo_net_destroy(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class TeacherNet(Net):
    """ResNet50 / SENet50 -ferplus behind xemo_teacher_forward."""

    def __init__(self, params, batch, input_mode="hwcn224", face_size=48, ctx=None, device=0, stream=None):
        k = int(np.asarray(params["classifierf"]).shape[-1])
        super().__init__(params["arch"], params, batch, face_size if input_mode == "u8" else 224, 1 if input_mode == "u8" else 0, k, ctx, device, stream)

    def run(self):
        self._check(self.lib.xemo_teacher_forward(self.handle, None))

    def forward(self, faces):
        """dag.eval({'data', faces}); gather(squeeze(dag.vars(end).value))' -> N x K."""
        self.set_input(faces)
        out = np.empty((self.N, self.K), np.float32)
        self._check(self.lib.xemo_teacher_forward(self.handle, out.ctypes.data_as(VP)))
        return out


class StudentNet(Net):
    """VGGVox student behind xemo_student_forward / xemo_student_train_step / xemo_sgd_step."""

    def __init__(self, params, batch, width=300, num_classes=8, temperature=2.0, loss_type="hot-cross-ent", grad_scale=1024.0, ctx=None,
                 device=0, stream=None):
        if loss_type not in LOSSES:
            raise ValueError("unrecognised regression loss: %s" % (loss_type,))
        super().__init__("vggvox", params, batch, width, 0, num_classes, ctx, device, stream)
        self.W, self.loss_type = width, loss_type
        self.temperature, self.grad_scale = float(temperature), float(grad_scale)
        self._check(self.lib.xemo_net_set_loss(self.handle, LOSSES[loss_type], self.temperature, self.grad_scale))
        self.hyper = dict(lr=1e-4, momentum=0.9, weight_decay=5e-4, batch_size=self.N)

    def set_grad_scale(self, grad_scale):
        """Loss scale of the fp16 activation-gradient chain (the filter / bias / BN reductions undo it in fp32).  It is baked
        into the captured step: the next step re-captures (and restarts the epoch's per-class counters)."""
        self.grad_scale = float(grad_scale)
        self._check(self.lib.xemo_net_set_loss(self.handle, LOSSES[self.loss_type], self.temperature, self.grad_scale))

    def set_overlap(self, mode):
        """-1 auto (on for batch <= 64), 0 one stream, 1 forked branches inside the captured step"""
        self._check(self.lib.xemo_net_set_overlap(self.handle, int(mode)))

    def set_hyper(self, **kw):
        self.hyper.update({k: v for k, v in kw.items() if v is not None})

    def set_target(self, target=None, weights=None):
        t = w = None
        if target is not None:
            if self.loss_type == "softmaxlog":   # maxLabel (1-based) -> one-hot rows
                lab = np.asarray(target).reshape(-1).astype(np.int64)
                t = np.zeros((self.N, self.K), np.float32)
                t[np.arange(self.N), lab - 1] = 1.0
            else:
                t = np.ascontiguousarray(np.asarray(target, np.float32).reshape(self.K, self.N).T)
        if weights is not None:
            w = np.ascontiguousarray(np.asarray(weights, np.float32).reshape(self.N))
        self._check(self.lib.xemo_net_set_target(self.handle, _ptr(t), _ptr(w)))
        self.ctx.sync()     # (t / w are pageable temporaries)

    def forward(self, spec, mode="test"):
        """dag.eval({'data', spec}) -> N x K."""
        self.set_input(spec)
        out = np.empty((self.N, self.K), np.float32)
        self._check(self.lib.xemo_student_forward(self.handle, 0 if mode == "test" else 1, out.ctypes.data_as(VP)))
        return out

    def grad_step(self, comm=None):
        self._check(self.lib.xemo_student_train_step(self.handle, comm.handle if comm else None))

    def update(self):
        h = self.hyper
        self._check(self.lib.xemo_sgd_step(self.handle, float(h["lr"]), float(h["momentum"]), float(h["weight_decay"]), int(h["batch_size"])))

    def train_step(self, spec, target, comm=None, weights=None):
        """One cnn_train_dag iteration."""
        self.set_input(spec)
        self.set_target(target, weights)
        self.grad_step(comm)
        self.update()

    def reset_metrics(self):
        self._check(self.lib.xemo_net_reset_metrics(self.handle))

    def metrics(self):
        out = np.zeros(4 + 2 * self.K, np.float32)
        self._check(self.lib.xemo_net_metrics(self.handle, out.ctypes.data_as(VP), out.size))
        K = self.K
        return dict(objective=float(out[0]), classerror=float(out[1]), correct=out[2 : 2 + K].copy(), count=out[2 + K : 2 + 2 * K].copy(),
                    nonfinite_grad=bool(out[2 + 2 * K]), skipped_steps=int(out[3 + 2 * K]))

    def prediction(self):
        out = np.empty((self.N, 16), np.float32)
        self.ctx.sync()
        self.ctx.d2h(out.ctypes.data_as(VP), VP(self.buffer("pred32")), out.nbytes)
        self.ctx.sync()
        return out[:, : self.K].copy()

    def _tensors(self, which):
        out = {}
        for name in self.names:
            flat = np.empty(int(np.prod(self.dims[name])), np.float32)
            self._check(self.lib.xemo_net_get_tensor(self.handle, which, name.encode(), flat.ctypes.data_as(VP), flat.size))
            out[name] = _from_colmajor(flat, self.dims[name])
        return out

    def export_params(self):
        return self._tensors(0)

    def export_grads(self):
        return self._tensors(1)

    def export_momentum(self):
        return {k: v for k, v in self._tensors(2).items() if not k.endswith("x")}

    def load_momentum(self, momentum):
        for name, v in momentum.items():
            flat = _colmajor(name, v)
            self._check(self.lib.xemo_net_set_momentum(self.handle, name.encode(), flat.ctypes.data_as(VP), flat.size))


class DistillStep:
    """The full distillation step through xemo_distill_step: one context, two networks, one captured graph."""

    def __init__(self, teacher_params, student_params, batch, width=300, frames_per_clip=1, aggregator="max", device=0, stream=None,
                 face_input="u8", face_size=48, temperature=2.0, loss_type="hot-cross-ent", comm=None):
        self.ctx = _lib.Context(device, stream)
        self.N, self.F = batch, frames_per_clip
        self.teacher = TeacherNet(teacher_params, batch * frames_per_clip, face_input, face_size, ctx=self.ctx)
        self.student = StudentNet(student_params, batch, width, temperature=temperature, loss_type=loss_type, ctx=self.ctx)
        self.use_mean = 1 if aggregator == "mean" else 0
        self.comm = comm
        # frame windows of the coupling: clip i owns teacher rows [i F, (i + 1) F)
        self.set_windows(np.arange(batch) * frames_per_clip, (np.arange(batch) + 1) * frames_per_clip)

    def set_windows(self, start, end):
        """half-open teacher-row ranges per clip (batch.frame_window gives them for cached logits)"""
        s, e = np.ascontiguousarray(start, np.int32), np.ascontiguousarray(end, np.int32)
        self.student._check(self.ctx.lib.xemo_distill_set_windows(self.student.handle, s.ctypes.data_as(VP), e.ctypes.data_as(VP)))

    def step(self):
        h = self.student.hyper
        rc = self.ctx.lib.xemo_distill_step(self.teacher.handle, self.student.handle, self.comm.handle if self.comm else None,
                                            None, None, self.use_mean, float(h["lr"]),
                                            float(h["momentum"]), float(h["weight_decay"]), int(h["batch_size"]))
        self.student._check(rc)

    def num_kernels(self):
        return self.student.num_kernels()

    def sync(self):
        self.ctx.sync()

    def close(self):
        self.student.close()        # (networks before the communicator their captured step used)
        self.teacher.close()
        if self.comm:
            self.comm.close()
