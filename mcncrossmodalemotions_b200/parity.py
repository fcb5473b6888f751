"""The student in fp32-equivalent arithmetic: one cnn_train_dag iteration as dagnn would run it -- a chain of the
MatConvNet-boundary operators (`vl_nn*` on gpuArrays, single, H x W x C x N) -- with the convolutions in the
split-operand mode of libxemo (XEMO_CONV_F32X3: three tcgen05 products per single-precision product, fp32 accumulation)
and every other operator in fp32.

This is the parity mode of the hot path: the reference trains in `single` end to end (gpuArray(single) batches,
emoVoxCeleb/getBatchEmoVoxCeleb.m:197; cnn_train_dag, emoVoxCeleb/run_distillation.m:170-182; loss
emoVoxCeleb/emoVoxZoo.m:137-157), and the fp16-operand fast programs (programs.StudentProgram) cannot hold a 1e-3
tolerance on train-mode logits and gradients (ReLU masks / pooling winners flip under 2^-11 perturbations).  Same
interface as StudentProgram (train_step / grad_step / update / forward / metrics / export_*), no CUDA graphs, activations
kept as fp32 gpuArrays on the tape; throughput is not the point (bench.py --precision f32x3 reports it beside the fast
mode).  All arithmetic runs in libxemo.so on the GPU; nothing here computes on the CPU."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import vl_nn
from .arch import TEACHER_STAGES
from .programs import BN_EPS, LOSS_TYPES, STUDENT_CONVS, STUDENT_POOLS, _out
from .vl_nn import GpuArray, gather, gpuArray

VP = C.c_void_p


def _ptr(t):
    return VP(t.data_ptr()) if t is not None else None


class StudentProgramF32:
    def __init__(self, params, batch, width=300, num_classes=8, temperature=2.0, loss_type="hot-cross-ent"):
        if loss_type not in LOSS_TYPES:
            raise ValueError("unrecognised regression loss: %s" % (loss_type,))
        self.ctx = vl_nn.default_context()
        self.N, self.W, self.K, self.T, self.loss_type = batch, width, num_classes, float(temperature), loss_type
        self.layers = []
        h, w = 512, width
        for name, fh, fw, cin, cout, stride, pad, has_bn in STUDENT_CONVS:
            cout = num_classes if name == "fc8" else cout
            oh, ow = _out(h, w, fh, fw, stride, pad)
            L = dict(name=name, stride=stride, pad=pad, bn="bn" + name[-1] if has_bn else None, pool=None)
            h, w = oh, ow
            if name in STUDENT_POOLS:
                method, win, ps = STUDENT_POOLS[name]
                win = win or (1, w)      # pool6 averages the whole remaining width (emoVoxZoo.m:258-269)
                L["pool"] = (method, win, ps)
                h, w = _out(h, w, win[0], win[1], ps, (0, 0, 0, 0))
            self.layers.append(L)
        assert (h, w) == (1, 1), "student graph must reduce to 1 x 1 (got %d x %d)" % (h, w)
        self.p, self.m, self.moments = {}, {}, {}
        for k, v in params.items():
            if k.startswith("bn") and k.endswith("x"):
                self.moments[k] = np.asarray(v, np.float32).copy()      # C x 2 = [mu sigma]
            elif isinstance(v, np.ndarray):
                self.p[k] = gpuArray(v.reshape(-1, 1) if v.ndim == 1 else v)
                self.m[k] = torch.zeros_like(self.p[k].tensor)
        dev = self.p["conv1f"].tensor.device
        self.hyper = torch.tensor([1e-4, 0.9, 5e-4, 1.0 / batch], dtype=torch.float32, device=dev)
        self.scalars = torch.zeros(2, dtype=torch.float32, device=dev)
        self.class_stats = torch.zeros(2 * num_classes, dtype=torch.float32, device=dev)
        self.max_label = torch.zeros(batch, dtype=torch.int32, device=dev)
        self.weights = torch.ones(batch, dtype=torch.float32, device=dev)
        self.target = torch.zeros(batch * num_classes, dtype=torch.float32, device=dev)
        self.grads, self.batch_moments, self.tape = {}, {}, {}
        self.spec = self.pred = None
        torch.cuda.synchronize()

    # ---- inputs
    def set_hyper(self, lr=None, momentum=None, weight_decay=None, batch_size=None):
        h = self.hyper.cpu().numpy()
        for i, v in enumerate((lr, momentum, weight_decay, None if batch_size is None else 1.0 / batch_size)):
            if v is not None:
                h[i] = v
        self.hyper.copy_(torch.from_numpy(h))
        torch.cuda.synchronize()

    def set_input(self, spec, target=None, weights=None):
        self.spec = spec if isinstance(spec, GpuArray) else gpuArray(spec)
        if target is not None:
            if self.loss_type == "softmaxlog":   # maxLabel (1-based) -> one-hot rows
                lab = np.asarray(target).reshape(-1).astype(np.int64)
                t = np.zeros((self.N, self.K), np.float32)
                t[np.arange(self.N), lab - 1] = 1.0
            else:
                t = np.ascontiguousarray(np.asarray(target, np.float32).reshape(self.K, self.N).T)
            self.target.copy_(torch.from_numpy(t.reshape(-1)))
        if weights is not None:
            self.weights.copy_(torch.from_numpy(np.ascontiguousarray(np.asarray(weights, np.float32).reshape(self.N))))
        torch.cuda.synchronize()

    # ---- forward / backward as dagnn.DagNN.eval sequences the blocks
    def _forward(self, train, keep):
        p, tape = self.p, {}
        cur = self.spec
        prev = self.ctx.lib.xemo_get_conv_precision(self.ctx.handle)
        self.ctx.set_conv_precision(1)
        try:
            for L in self.layers:
                n = L["name"]
                if keep:
                    tape[n + ":x"] = cur
                cur = vl_nn.vl_nnconv(cur, p[n + "f"], p[n + "b"], pad=L["pad"], stride=L["stride"])
                if L["bn"]:
                    bn = L["bn"]
                    if keep:
                        tape[bn + ":x"] = cur
                    cur, mom = vl_nn.vl_nnbnorm(cur, p[bn + "m"], p[bn + "b"], epsilon=BN_EPS, moments=None if train else self.moments[bn + "x"])
                    if train:
                        self.batch_moments[bn + "x"] = mom
                    if keep:
                        tape[bn + ":relu"] = cur
                    cur = vl_nn.vl_nnrelu(cur)
                if L["pool"]:
                    method, win, ps = L["pool"]
                    if keep:
                        tape[n + ":pool"] = cur
                    cur = vl_nn.vl_nnpool(cur, win, stride=ps, method=method)
        finally:
            self.ctx.set_conv_precision(prev)
        self.tape = tape
        self.pred = cur      # 1 x 1 x K x N == [N][K] row-major
        return cur

    def _loss(self, backward):
        lt, soft = LOSS_TYPES[self.loss_type], self.loss_type == "hot-cross-ent"
        dpred = GpuArray(self.pred.shape, torch.empty_like(self.pred.tensor)) if backward else None
        torch.cuda.synchronize()
        self.ctx.memset(_ptr(self.scalars), 0, 8)
        self.ctx.op_loss(VP(self.pred.ptr), 1, self.K, _ptr(self.target), self.K, _ptr(self.weights) if lt else None, self.N, self.K, lt,
                         self.T if soft else 1.0, 1 if soft else 0, 1.0, 1.0, VP(dpred.ptr) if backward else None, 1, self.K,
                         _ptr(self.scalars), _ptr(self.class_stats), _ptr(self.max_label))
        return dpred

    def _backward(self, dzdy):
        p, tape, g = self.p, self.tape, {}
        cur = dzdy
        prev = self.ctx.lib.xemo_get_conv_precision(self.ctx.handle)
        self.ctx.set_conv_precision(1)
        try:
            for L in reversed(self.layers):
                n = L["name"]
                if L["pool"]:
                    method, win, ps = L["pool"]
                    cur = vl_nn.vl_nnpool(tape[n + ":pool"], win, cur, stride=ps, method=method)
                if L["bn"]:
                    bn = L["bn"]
                    cur = vl_nn.vl_nnrelu(tape[bn + ":relu"], cur)
                    cur, dg, db, _ = vl_nn.vl_nnbnorm(tape[bn + ":x"], p[bn + "m"], p[bn + "b"], cur, epsilon=BN_EPS)
                    g[bn + "m"], g[bn + "b"] = gpuArray(dg.reshape(-1, 1)), gpuArray(db.reshape(-1, 1))
                dx, df, dbias = vl_nn.vl_nnconv(tape[n + ":x"], p[n + "f"], p[n + "b"], cur, pad=L["pad"], stride=L["stride"])
                g[n + "f"], g[n + "b"] = df, dbias
                cur = dx
        finally:
            self.ctx.set_conv_precision(prev)
        self.grads = g      # (the tape stays until the next forward: export_decisions reads it)

    # ---- StudentProgram interface
    def forward(self, spec, mode="test"):
        """dag.eval({'data', spec}) -> N x K numpy predictions."""
        self.set_input(spec)
        pred = self._forward(mode != "test", keep=False)
        return gather(pred).reshape(self.K, self.N).T.copy()

    def grad_step(self):
        self._forward(True, keep=True)
        self._backward(self._loss(True))

    def update(self):
        """cnn_train_dag accumulateGradients: m <- mu m - (wd w + g / B); w <- w + lr m; BN moments moving average."""
        for k, w in self.p.items():
            gk = self.grads[k]
            self.ctx.op_sgd_momentum(VP(w.ptr), _ptr(self.m[k]), VP(gk.ptr), w.tensor.numel(), _ptr(self.hyper), 1.0, 1.0, 1.0, None)
        self.ctx.sync()
        for k, mom in self.batch_moments.items():
            # (2 x C numbers per layer; the fast program's moments_average kernel computes the same expression)
            dev = gpuArray(self.moments[k]), gpuArray(mom)
            self.ctx.op_moments_average(VP(dev[0].ptr), VP(dev[1].ptr), dev[0].tensor.numel(), 0.1)
            self.moments[k] = gather(dev[0]).reshape(self.moments[k].shape)

    def train_step(self, spec, target, allreduce=None, weights=None):
        self.set_input(spec, target, weights)
        self.grad_step()
        if allreduce is not None:
            for k in self.grads:
                allreduce(self.grads[k].tensor)
        self.update()

    def reset_metrics(self):
        self.ctx.memset(_ptr(self.scalars), 0, 8)
        self.ctx.memset(_ptr(self.class_stats), 0, 8 * self.K)

    def metrics(self):
        self.ctx.sync()
        s, cs = self.scalars.cpu(), self.class_stats.cpu()
        return dict(objective=float(s[0]), classerror=float(s[1]), correct=cs[: self.K].numpy(), count=cs[self.K :].numpy())

    def export_decisions(self):
        """The discrete decisions of the last train-mode forward (see StudentProgram.export_decisions): ReLU masks x > 0
        ('relu<i>', H x W x C x N bool) and the uint8 window-local arg-max of every max pool ('pool<i>')."""
        out = {}
        for L in self.layers:
            n, i = L["name"], L["name"][-1]
            if L["bn"]:
                out["relu" + i] = gather(self.tape[L["bn"] + ":relu"]) > 0
            if L["pool"] and L["pool"][0] == "max":
                _, idx = vl_nn.vl_nnpool(self.tape[n + ":pool"], L["pool"][1], stride=L["pool"][2], method="max", return_index=True)
                out["pool" + i] = idx
        return out

    def prediction(self):
        return gather(self.pred).reshape(self.K, self.N).T.copy()

    def _export(self, d):
        out = {}
        for k, v in d.items():
            a = gather(v)
            out[k] = a.copy() if k.endswith("f") else a.reshape(-1).copy()   # filters FH x FW x FC x K; vectors flat
        return out

    def export_params(self):
        out = self._export(self.p)
        out.update({k: v.copy() for k, v in self.moments.items()})
        return out

    def export_grads(self):
        out = self._export(self.grads)
        out.update({k: v.copy() for k, v in self.batch_moments.items()})
        return out


class TeacherProgramF32:
    """ResNet50 / SENet50 -ferplus forward (dag.mode = 'test', emoVoxCeleb/fetch_emovoxceleb_imdb.m:107,129) in
    fp32-equivalent arithmetic: the dagnn layer sequence on gpuArrays through the boundary operators, convolutions in the
    split-operand mode.  The fast program's worst logit over a 256-face batch sits at 1.06e-3 of the range (fp16 rounding of
    the residual stream at each of the 16 bottlenecks); this mode is the one that holds 1e-3 on every logit."""

    def __init__(self, params):
        self.ctx = vl_nn.default_context()
        self.arch = params["arch"]
        self.p = {}
        for k, v in params.items():
            if isinstance(v, np.ndarray):
                self.p[k] = v if k.endswith("x") else gpuArray(v.reshape(-1, 1) if v.ndim == 1 else v)

    def _cbr(self, name, bn, t, stride=1, pad=0, relu=True):
        t = vl_nn.vl_nnconv(t, self.p[name + "f"], None, pad=pad, stride=stride)
        t, _ = vl_nn.vl_nnbnorm(t, self.p[bn + "m"], self.p[bn + "b"], epsilon=BN_EPS, moments=self.p[bn + "x"])
        return vl_nn.vl_nnrelu(t) if relu else t

    def forward(self, faces):
        """faces: 224 x 224 x 3 x N normalised singles -> N x K logits."""
        p, se = self.p, self.arch == "senet50"
        prev = self.ctx.lib.xemo_get_conv_precision(self.ctx.handle)
        self.ctx.set_conv_precision(1)
        try:
            cur = self._cbr("conv1", "bn1", gpuArray(faces), stride=2, pad=3)
            cur = vl_nn.vl_nnpool(cur, (3, 3), pad=(0, 1, 0, 1), stride=2, method="max")
            for si, (blocks, mid, cout, stride) in enumerate(TEACHER_STAGES):
                for bi in range(blocks):
                    pre = "s%db%d_" % (si + 2, bi + 1)
                    s = stride if bi == 0 else 1
                    u = self._cbr(pre + "c1", pre + "bn1", cur, stride=s)
                    u = self._cbr(pre + "c2", pre + "bn2", u, pad=1)
                    u = self._cbr(pre + "c3", pre + "bn3", u, relu=False)
                    sc = self._cbr(pre + "proj", pre + "bnp", cur, stride=s, relu=False) if bi == 0 else cur
                    if se:
                        z = vl_nn.vl_nnglobalpool(u)
                        z = vl_nn.vl_nnrelu(vl_nn.vl_nnconv(z, p[pre + "se1f"], p[pre + "se1b"]))
                        a = vl_nn.vl_nnsigmoid(vl_nn.vl_nnconv(z, p[pre + "se2f"], p[pre + "se2b"]))
                        cur = vl_nn.vl_nnrelu(vl_nn.vl_nnaxpy(a, u, sc))
                    else:
                        cur = vl_nn.vl_nnrelu(vl_nn.vl_nnaxpy(GpuArray((1, 1, cout, faces.shape[3]), self._ones(cout * faces.shape[3])), u, sc))
            cur = vl_nn.vl_nnpool(cur, (7, 7), stride=1, method="avg")
            out = vl_nn.vl_nnconv(cur, p["classifierf"], p["classifierb"])
        finally:
            self.ctx.set_conv_precision(prev)
        k = out.shape[2]
        return gather(out).reshape(k, -1).T.copy()

    def _ones(self, n):
        t = torch.ones(n, dtype=torch.float32, device="cuda")
        torch.cuda.synchronize()
        return t
