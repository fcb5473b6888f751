"""Fused device programs for the three graphs on the distillation hot path.

A program owns the device-resident state of one network on one GPU (fp16 KRSC filters, fp32 master
copies, NHWC fp16 activations, gradients, optimiser state -- torch tensors are used purely as device
memory) and strings the `xemo_op_*` building blocks of libxemo.so into one CUDA graph per phase.
Reference call sites replaced:
  TeacherProgram.forward        dag.eval at emoVoxCeleb/fetch_emovoxceleb_imdb.m:129 and
                                external/compute_visual_feats.m:90 (dag.mode = 'test', losses removed)
  StudentProgram.forward        dag.eval at external/compute_audio_feats.m:126 (test mode)
  StudentProgram.train_step     one cnn_train_dag iteration (emoVoxCeleb/run_distillation.m:170-182)
                                with the loss of emoVoxCeleb/emoVoxZoo.m:151-157 and the metric layers
                                of emoVoxZoo.m:160-169
Parameter dictionaries use MatConvNet layouts (filters FH x FW x FC x K, BN moments C x 2) and the
key names of the zoo (`<layer>f`, `<layer>b`, `bnNm`, `bnNb`, `bnNx`).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib

BN_EPS = 1e-5  # dagnn.BatchNorm default
LOSS_TYPES = {"hot-cross-ent": 0, "softmaxlog": 0, "euclidean": 1, "huber": 2}   # -> loss_type of xemo_op_loss
VP = C.c_void_p

from .arch import STUDENT_CONVS, STUDENT_POOLS, TEACHER_STAGES  # noqa: E402,F401


def _pad16(v):
    return (v + 15) // 16 * 16


def _p(t):
    return VP(t.data_ptr()) if t is not None else None


def krsc(f, kp=None, cp=None):
    """FH x FW x FC x K (MatConvNet) -> [Kp][FH][FW][Cp] (device layout), zero padded."""
    fh, fw, fc, k = f.shape
    kp, cp = kp or _pad16(k), cp or _pad16(fc)
    out = np.zeros((kp, fh, fw, cp), np.float32)
    out[:k, :, :, :fc] = np.transpose(f, (3, 0, 1, 2))
    return out


def unkrsc(w, fh, fw, fc, k):
    return np.ascontiguousarray(np.transpose(w[:k, :, :, :fc], (1, 2, 3, 0)))


def student_conv1_to_s2d(f):
    """7 x 7 x 1 x K stride-2 filter -> [K][4][1][16] filter of the space-to-depth formulation:
    G[k][j][0][dr*8 + s] = F[2j + dr, s, 0, k]  (zero where 2j+dr > 6 or s > 6)."""
    k = f.shape[3]
    g = np.zeros((k, 4, 1, 16), np.float32)
    for j in range(4):
        for dr in range(2):
            r = 2 * j + dr
            if r < 7:
                g[:, j, 0, dr * 8 : dr * 8 + 7] = f[r, :, 0, :].T
    return g


def student_conv1_from_s2d(g):
    k = g.shape[0]
    f = np.zeros((7, 7, 1, k), np.float32)
    for j in range(4):
        for dr in range(2):
            r = 2 * j + dr
            if r < 7:
                f[r, :, 0, :] = g[:, j, 0, dr * 8 : dr * 8 + 7].T
    return f


def pair_filter(g):
    """[K][R][1][C] filter -> block-diagonal [2K][R][1][2C] filter of the pixel-pair form (two horizontally adjacent
    output pixels as one GEMM row): G2[e*K + k][r][0][e'*C + c] = [e == e'] G[k][r][0][c]."""
    k, r, _, c = g.shape
    g2 = np.zeros((2 * k, r, 1, 2 * c), g.dtype)
    g2[:k, :, :, :c] = g
    g2[k:, :, :, c:] = g
    return g2


def teacher_conv1_to_rows(f):
    """7 x 7 x 3 x K stride-2 filter -> [K][7][1][32] filter over the row-im2col input:
    G[k][r][0][s*4 + c] = F[r, s, c, k]."""
    k = f.shape[3]
    g = np.zeros((k, 7, 1, 32), np.float32)
    for s in range(7):
        for c in range(3):
            g[:, :, 0, s * 4 + c] = f[:, s, c, :].T
    return g


class _Base:
    def __init__(self, device=0, stream=None, ctx=None):
        """`stream` (torch.cuda.Stream) and `ctx` may be shared between programs that are captured into
        one CUDA graph (the distillation step)."""
        self.device = torch.device("cuda", device)
        torch.cuda.set_device(self.device)
        self.stream = stream or torch.cuda.Stream(self.device)
        self.ctx = ctx or _lib.Context(device, self.stream.cuda_stream)
        self._keep = []

    # ---- device memory (torch tensors as raw buffers)
    def f16(self, *shape):
        with torch.cuda.stream(self.stream):
            return torch.zeros(shape, dtype=torch.float16, device=self.device)

    def f32(self, *shape):
        with torch.cuda.stream(self.stream):
            return torch.zeros(shape, dtype=torch.float32, device=self.device)

    def upload(self, a, dtype=torch.float32):
        with torch.cuda.stream(self.stream):
            return torch.from_numpy(np.ascontiguousarray(a)).to(self.device).to(dtype)

    def sync(self):
        self.ctx.sync()

    # ---- op helpers
    def conv(self, x, n, h, w, cin, wt, kout, r, s, stride, pad, scale=None, shift=None, residual=None, relu=0, out16=None,
             out32=None, ldc=0):
        self.ctx.op_conv_fwd(_p(x), n, h, w, cin, _p(wt), kout, r, s, stride[0], stride[1], pad[0], pad[1], pad[2], pad[3],
                             _p(scale), _p(shift), _p(residual), relu, _p(out16), _p(out32), ldc)


def _out(h, w, fh, fw, stride, pad):
    return (h + pad[0] + pad[1] - fh) // stride[0] + 1, (w + pad[2] + pad[3] - fw) // stride[1] + 1


# ================================================================================================
class TeacherProgram(_Base):
    """ResNet50 / SENet50 -ferplus forward (test-mode BN folded into the conv epilogues)."""

    def __init__(self, params, batch, device=0, stream=None, use_graph=True, ctx=None, input_mode="hwcn224", face_size=48,
                 average_image=(131.0912, 103.8827, 91.4953)):
        """input_mode 'hwcn224': 224 x 224 x 3 x N single, already normalised (what dag.eval receives at
        emoVoxCeleb/fetch_emovoxceleb_imdb.m:129).  'u8': face_size x face_size x N uint8 grey faces; the
        reference's normalizeFace + resize (fetch_emovoxceleb_imdb.m:175-193) run fused on the device."""
        super().__init__(device, stream, ctx)
        self.arch = params["arch"]
        self.N = batch
        self.use_graph = use_graph
        self.input_mode = input_mode
        self.face_size = face_size
        self.mean3 = self.upload(np.asarray(average_image, np.float32))
        self.graph = None
        # SE blocks by linearity -- squeeze of the C/4-channel 3x3 output, gate, and the excite folded into the expand
        # convolution's epilogue; the expand output u is never materialised (2.5 C instead of 5.25 C bytes per pixel of SE
        # traffic) -- on the stages whose feature map is at least se_lin_min_hw wide: measured on B200, a gain at 56 x 56 and
        # 28 x 28, a loss at 14 x 14 and 7 x 7 (profiles/r02_ab_experimental_options.json).  XEMO_SE_LIN_MIN_HW overrides
        # (0 disables; 7 = every stage).  The same rule as csrc/xemo_net.cu.
        self.se_lin_min_hw = int(os.environ.get("XEMO_SE_LIN_MIN_HW", "28"))
        self.se_lin = lambda hw: self.se_lin_min_hw > 0 and hw >= self.se_lin_min_hw
        self._load(params)
        self._alloc()

    def _fold(self, p, bn):
        """test-mode BN as y = a*x + b: a = g/sigma, b = beta - a*mu (moments = [mu sigma])."""
        g, beta, mom = p[bn + "m"].astype(np.float64), p[bn + "b"].astype(np.float64), p[bn + "x"].astype(np.float64)
        a = g / mom[:, 1]
        return self.upload(a.astype(np.float32)), self.upload((beta - a * mom[:, 0]).astype(np.float32))

    def _load(self, p):
        W = self.w = {}
        # the stem in pixel-pair form: 128-byte TMA rows over the [N][224][56][64] view of the row-im2col tensor,
        # block-diagonal [128][7][1][64] filter (resident in smem), output = the [N][112][56][128] view of c1
        self.stem_pairs = os.environ.get("XEMO_TEACHER_STEM_PAIRS", "1") != "0"
        g = teacher_conv1_to_rows(p["conv1f"])
        W["conv1"] = self.upload(pair_filter(g) if self.stem_pairs else g, torch.float16)
        a1, b1 = self._fold(p, "bn1")
        if self.stem_pairs:
            with torch.cuda.stream(self.stream):
                a1, b1 = torch.cat([a1, a1]), torch.cat([b1, b1])
        W["bn1"] = (a1, b1)
        cin = 64
        self.blocks = []
        for si, (blocks, mid, cout, stride) in enumerate(TEACHER_STAGES):
            for bi in range(blocks):
                pre = "s%db%d_" % (si + 2, bi + 1)
                for c, bn in (("c1", "bn1"), ("c2", "bn2"), ("c3", "bn3")) + ((("proj", "bnp"),) if bi == 0 else ()):
                    W[pre + c] = self.upload(krsc(p[pre + c + "f"]), torch.float16)
                    W[pre + bn] = self._fold(p, pre + bn)
                if self.arch == "senet50":
                    W[pre + "se1"] = self.upload(p[pre + "se1f"][0, 0].T)  # [Cr][C]
                    W[pre + "se1b"] = self.upload(p[pre + "se1b"])
                    W[pre + "se2"] = self.upload(p[pre + "se2f"][0, 0])  # [Cr][C]: W2 transposed (coalesced reads)
                    W[pre + "se2b"] = self.upload(p[pre + "se2b"])
                self.blocks.append((pre, cin, mid, cout, stride if bi == 0 else 1, bi == 0))
                cin = cout
        k = p["classifierf"].shape[3]
        self.num_outputs = k
        W["classifier"] = self.upload(krsc(p["classifierf"]), torch.float16)
        cb = np.zeros(_pad16(k), np.float32)
        cb[:k] = p["classifierb"]
        W["classifierb"] = self.upload(cb)

    def _alloc(self):
        N = self.N
        A = self.a = {}
        if self.input_mode == "u8":
            A["faces"] = torch.zeros(N * self.face_size * self.face_size, dtype=torch.uint8, device=self.device)
        else:
            A["faces"] = self.f32(N * 3 * 224 * 224)      # H x W x C x N column-major fp32 (MatConvNet layout)
        A["rows"] = self.f16(N, 224, 112, 32)
        A["c1"] = self.f16(N, 112, 112, 64)
        A["p1"] = self.f16(N, 56, 56, 64)
        hw = 56
        for pre, cin, mid, cout, stride, proj in self.blocks:
            ohw = hw // stride
            A[pre + "t1"] = self.f16(N, ohw, ohw, mid)
            A[pre + "t2"] = self.f16(N, ohw, ohw, mid)
            if proj:
                A[pre + "sc"] = self.f16(N, ohw, ohw, cout)
            if self.arch == "senet50" and self.se_lin(ohw):
                A[pre + "m2"], A[pre + "gs"], A[pre + "gh"] = self.f32(N, mid), self.f32(N, cout), self.f32(N, cout)
            elif self.arch == "senet50":
                A[pre + "u"] = self.f16(N, ohw, ohw, cout)
                A[pre + "s"] = self.f32(N, cout)
                A[pre + "g"] = self.f32(N, cout)
            A[pre + "y"] = self.f16(N, ohw, ohw, cout)
            hw = ohw
        A["pool5"] = self.f16(N, 2048)
        A["logits"] = self.f32(N, _pad16(self.num_outputs))

    def _record(self):
        N, A, W, ctx = self.N, self.a, self.w, self.ctx
        if self.input_mode == "u8":
            fs = self.face_size
            ctx.op_face_u8_rows_im2col(_p(A["faces"]), fs, fs, N, 224, 224, _p(self.mean3), 7, 2, 3, 112, _p(A["rows"]))
        else:
            ctx.op_face_rows_im2col(_p(A["faces"]), 224, 224, 3, N, 7, 2, 3, 112, _p(A["rows"]))
        a, b = W["bn1"]
        if self.stem_pairs:
            self.conv(A["rows"], N, 224, 56, 64, W["conv1"], 128, 7, 1, (2, 1), (3, 3, 0, 0), a, b, None, 1, A["c1"])
        else:
            self.conv(A["rows"], N, 224, 112, 32, W["conv1"], 64, 7, 1, (2, 1), (3, 3, 0, 0), a, b, None, 1, A["c1"])
        ctx.op_maxpool_fwd(_p(A["c1"]), N, 112, 112, 64, 3, 3, 2, 2, 0, 1, 0, 1, None, None, _p(A["p1"]), None)
        cur, hw = A["p1"], 56
        se = self.arch == "senet50"
        for pre, cin, mid, cout, stride, proj in self.blocks:
            ohw = hw // stride
            a, b = W[pre + "bn1"]
            self.conv(cur, N, hw, hw, cin, W[pre + "c1"], mid, 1, 1, (stride, stride), (0, 0, 0, 0), a, b, None, 1, A[pre + "t1"])
            a, b = W[pre + "bn2"]
            self.conv(A[pre + "t1"], N, ohw, ohw, mid, W[pre + "c2"], mid, 3, 3, (1, 1), (1, 1, 1, 1), a, b, None, 1, A[pre + "t2"])
            if proj:
                a, b = W[pre + "bnp"]
                self.conv(cur, N, hw, hw, cin, W[pre + "proj"], cout, 1, 1, (stride, stride), (0, 0, 0, 0), a, b, None, 0,
                          A[pre + "sc"])
                sc = A[pre + "sc"]
            else:
                sc = cur
            a, b = W[pre + "bn3"]
            if se and self.se_lin(ohw):
                ctx.op_se_squeeze(_p(A[pre + "t2"]), N, ohw * ohw, mid, _p(A[pre + "m2"]))
                ctx.op_se_gate_lin(_p(A[pre + "m2"]), N, cout, mid, cout // 16, _p(W[pre + "c3"]), _p(a), _p(b), _p(W[pre + "se1"]),
                                   _p(W[pre + "se1b"]), _p(W[pre + "se2"]), _p(W[pre + "se2b"]), _p(A[pre + "gs"]), _p(A[pre + "gh"]))
                ctx.op_conv_fwd_nc(_p(A[pre + "t2"]), N, ohw, ohw, mid, _p(W[pre + "c3"]), cout, 1, 1, 1, 1, 0, 0, 0, 0,
                                   _p(A[pre + "gs"]), _p(A[pre + "gh"]), _p(sc), 1, _p(A[pre + "y"]))
            elif se:
                self.conv(A[pre + "t2"], N, ohw, ohw, mid, W[pre + "c3"], cout, 1, 1, (1, 1), (0, 0, 0, 0), a, b, None, 0, A[pre + "u"])
                # (computing the squeeze by linearity from mean_hw(t2) -- 4x fewer bytes -- was measured: the extra tiny
                # launches cost what the narrower read saves, so the direct form stays)
                ctx.op_se_squeeze(_p(A[pre + "u"]), N, ohw * ohw, cout, _p(A[pre + "s"]))
                ctx.op_se_gate(_p(A[pre + "s"]), N, cout, cout // 16, _p(W[pre + "se1"]), _p(W[pre + "se1b"]), _p(W[pre + "se2"]),
                               _p(W[pre + "se2b"]), _p(A[pre + "g"]))
                ctx.op_se_excite(_p(A[pre + "u"]), _p(A[pre + "g"]), _p(sc), N, ohw * ohw, cout, 1, _p(A[pre + "y"]))
            else:
                self.conv(A[pre + "t2"], N, ohw, ohw, mid, W[pre + "c3"], cout, 1, 1, (1, 1), (0, 0, 0, 0), a, b, sc, 1, A[pre + "y"])
            cur, hw = A[pre + "y"], ohw
        ctx.op_avgpool_fwd(_p(cur), N, 7, 7, 2048, 7, 7, 1, 1, 0, 0, 0, 0, _p(A["pool5"]))
        kp = _pad16(self.num_outputs)
        self.conv(A["pool5"], N, 1, 1, 2048, W["classifier"], kp, 1, 1, (1, 1), (0, 0, 0, 0), None, W["classifierb"], None, 0,
                  None, A["logits"], kp)

    def run(self):
        """Forward on whatever a['faces'] currently holds; logits land in a['logits'] ([N][16] fp32)."""
        if not self.use_graph:
            self._record()
            return
        if self.graph is None:
            self._record()  # eager warm-up (sets kernel attributes outside the capture)
            self.ctx.capture_begin()
            self._record()
            self.graph = self.ctx.capture_end()
        self.graph.launch()

    def set_input(self, faces):
        """faces: 224 x 224 x 3 x N numpy (MatConvNet layout) or a pinned/device flat torch tensor in
        column-major order."""
        if isinstance(faces, np.ndarray):
            if self.input_mode == "u8":
                faces = torch.from_numpy(np.ascontiguousarray(faces.astype(np.uint8).transpose(2, 1, 0)).reshape(-1))
            else:
                faces = torch.from_numpy(np.ascontiguousarray(faces.astype(np.float32).transpose(3, 2, 1, 0)).reshape(-1))
        with torch.cuda.stream(self.stream):
            self.a["faces"].copy_(faces.reshape(-1), non_blocking=True)

    def forward(self, faces):
        """dag.eval({'data', faces}); gather(squeeze(dag.vars(end).value))' -> N x 8 numpy."""
        self.set_input(faces)
        self.run()
        with torch.cuda.stream(self.stream):
            out = self.a["logits"][:, : self.num_outputs].cpu()
        self.sync()
        return out.numpy()


# ================================================================================================
class StudentProgram(_Base):
    """VGGVox student: forward (train / test mode BN), backward, loss + metrics, SGD-momentum."""

    AUDIO = dict(fs=16000, Tw=25, Ts=10, alpha=0.97)   # emoVoxCeleb/run_distillation.m:109-117

    def __init__(self, params, batch, width=300, device=0, stream=None, use_graph=True, grad_scale=1024.0, num_classes=8,
                 temperature=2.0, ctx=None, audio_input="spectrogram", stem_algebra=None, stem_pairs=None,
                 loss_type="hot-cross-ent"):
        """audio_input 'spectrogram': 512 x W x 1 x N row-normalised spectrograms (what getBatchEmoVoxCeleb hands to
        dag.eval); 'wav': N x L waveform crops of L = (0.01 W + 0.024) * fs samples -- runSpec + the row normalisation
        (getBatchEmoVoxCeleb.m:162-169) then run on the device ahead of the graph."""
        # emoVoxZoo.m:137-157: 'hot-cross-ent' = SoftmaxCELoss(temperature, logitTargets) on {prediction, logitTarget};
        # 'softmaxlog' = dagnn.Loss('softmaxlog') on {prediction, maxLabel}, i.e. the same cross-entropy against a one-hot
        # distribution at T = 1 (the fused kernel's logit_targets = 0 mode); 'euclidean' = dagnn.EuclideanLoss and
        # 'huber' = dagnn.HuberLoss('sigma', 1) on {prediction, logitTarget, instanceWeights} -- all one fused kernel.
        if loss_type not in LOSS_TYPES:
            raise ValueError("unrecognised regression loss: %s" % (loss_type,))   # emoVoxZoo.m:154
        super().__init__(device, stream, ctx)
        self.audio_input = audio_input
        self.loss_type = loss_type
        a = self.AUDIO
        self.Nw, self.Ns = int(round(1e-3 * a["Tw"] * a["fs"])), int(round(1e-3 * a["Ts"] * a["fs"]))
        self.wav_len = (width - 1) * self.Ns + self.Nw + (a["fs"] * 24 // 1000 - self.Nw + self.Ns)  # = (0.01 W + 0.024) fs
        self.N, self.W = batch, width
        self.use_graph = use_graph
        self.grad_scale = float(grad_scale)
        self.K = num_classes
        self.T = float(temperature)
        self.graphs = {}
        self.fuse_pool_bwd = False
        # conv1 + bn1 + pool1 by linearity in the one-channel input (csrc/stem_kernels.cuh): BN statistics from the patch
        # autocorrelation, BN reductions at the pooled resolution, filter gradient without materialising dY.
        # XEMO_STEM_ALGEBRA=0 selects the generic per-layer path (A/B measurements, parity tests of both).
        self.stem_algebra = os.environ.get("XEMO_STEM_ALGEBRA", "1") != "0" if stem_algebra is None else bool(stem_algebra)
        # conv1 forward in pixel-pair form (32-channel view of the s2d tensor, block-diagonal filter): half the TMA row
        # requests per output pixel.  Needs an even conv1 output width (true for every width bucket 100..1000).
        self.stem_pairs = os.environ.get("XEMO_STEM_PAIRS", "1") != "0" if stem_pairs is None else bool(stem_pairs)
        # conv2 reads its input (pool1's output) with a 128-channel pitch (96 real + 32 zero channels): 128-byte TMA rows for
        # the forward / filter-gradient operand tiles instead of 64-byte ones (the kernels are bound by TMA rows, not bytes).
        # Rides on the stem path (its pooling kernels take a pooled-side pitch).  XEMO_CONV2_PAD=0 disables.
        self.conv2_pad = self.stem_algebra and os.environ.get("XEMO_CONV2_PAD", "1") != "0"
        self.stem_wgrad_pairs = os.environ.get("XEMO_STEM_WGRAD_PAIRS", "0") != "0"
        self.side_stream = None   # torch.cuda.Stream: filter gradients run there, off the dgrad critical path
        self._geometry()
        self._load(params)
        self._alloc()

    # ---- geometry walk (matches SURVEY.md Appendix A.1; pool6 averages the whole remaining width)
    def _geometry(self):
        self.layers = []
        h, w, c = 512, self.W, 1
        for name, fh, fw, cin, cout, stride, pad, has_bn in STUDENT_CONVS:
            if name == "fc8":
                cout = self.K
            oh, ow = _out(h, w, fh, fw, stride, pad)
            L = dict(name=name, fh=fh, fw=fw, cin=cin, cout=cout, kp=_pad16(cout), stride=stride, pad=pad, bn=has_bn, h=h, w=w, oh=oh,
                     ow=ow, pool=None)
            h, w, c = oh, ow, cout
            if name in STUDENT_POOLS:
                method, win, ps = STUDENT_POOLS[name]
                if win is None:
                    win = (1, w)
                ph, pw = _out(h, w, win[0], win[1], ps, (0, 0, 0, 0))
                L["pool"] = dict(method=method, win=win, stride=ps, h=h, w=w, oh=ph, ow=pw)
                h, w = ph, pw
            self.layers.append(L)
        assert (h, w) == (1, 1), "student graph must reduce to 1 x 1 (got %d x %d)" % (h, w)
        L1 = self.layers[0]
        self.s2d_hp, self.s2d_ow = L1["oh"] + 3, L1["ow"]
        self.stem_pairs = self.stem_pairs and L1["ow"] % 2 == 0
        for L in self.layers:
            L["cp"] = _pad16(L["cin"])
        if self.conv2_pad:
            self.layers[1]["cp"] = (self.layers[1]["cin"] + 63) // 64 * 64
        self.pool1_ld = self.layers[1]["cp"]   # channel pitch of pool1's output / gradient / arg-max / winner tensors

    # ---- parameters: one flat fp32 master / momentum / gradient buffer (single all-reduce payload)
    def _load(self, p):
        segs, off = {}, 0

        def seg(name, arr):
            nonlocal off
            segs[name] = (off, arr.shape)
            off += (arr.size + 63) // 64 * 64
            return arr

        host = {}
        for L in self.layers:
            n = L["name"]
            f = p[n + "f"].astype(np.float32)
            dev = student_conv1_to_s2d(f) if n == "conv1" else krsc(f, cp=L["cp"])
            host[n + "f"] = seg(n + "f", dev)
            b = np.zeros(L["kp"], np.float32)
            b[: L["cout"]] = p[n + "b"]
            host[n + "b"] = seg(n + "b", b)
            if L["bn"]:
                bn = "bn" + n[-1]
                host[bn + "m"] = seg(bn + "m", p[bn + "m"].astype(np.float32))
                host[bn + "b"] = seg(bn + "b", p[bn + "b"].astype(np.float32))
        self.segs, self.nparam = segs, off
        flat = np.zeros(off, np.float32)
        for k, (o, shape) in segs.items():
            flat[o : o + host[k].size] = host[k].reshape(-1)
        self.master = self.upload(flat)
        self.momentum = self.f32(off)
        self.grad = self.f32(off)
        self.w16 = self.upload(flat, torch.float16)  # fp16 mirror at identical offsets (filters read by tcgen05)
        # BN moments parameters (C x 2 = [mu sigma], stored [mu | sigma]) and the batch moments of the step
        self.moments, self.batch_moments = {}, {}
        for L in self.layers:
            if L["bn"]:
                bn = "bn" + L["name"][-1]
                m = p[bn + "x"].astype(np.float32)
                self.moments[bn] = self.upload(np.concatenate([m[:, 0], m[:, 1]]))
                self.batch_moments[bn] = self.f32(2 * L["cout"])
        self.hyper = self.upload(np.array([1e-4, 0.9, 5e-4, 1.0 / self.N], np.float32))
        with torch.cuda.stream(self.stream):
            self.guard = torch.zeros(3, dtype=torch.int32, device=self.device)   # xemo_op_grad_guard state

    def view(self, buf, name):
        o, shape = self.segs[name]
        return buf[o : o + int(np.prod(shape))]

    def _alloc(self):
        N = self.N
        A = self.a = {}
        A["spec"] = self.f32(N * 512 * self.W)                   # 512 x W x 1 x N column-major fp32
        if self.audio_input == "wav":
            A["wav"] = self.f32(N, self.wav_len)
        A["s2d"] = self.f16(N, self.s2d_hp, self.s2d_ow, 16)
        if self.stem_pairs:
            c1 = self.layers[0]["kp"]
            A["stem:w2"] = self.f16(2 * c1 * 4 * 32)
            A["stem:shift2"], A["stem:scale2"] = self.f32(2 * c1), self.f32(2 * c1)
            A["stem:g1pair"] = self.f32(2 * c1 * 4 * 32)
        if self.stem_algebra:
            A["stem:ws"] = torch.zeros(int(self.ctx.lib.xemo_stem_ws_doubles()), dtype=torch.float64, device=self.device)
        A["target"] = self.f32(N, self.K)                        # aggregated teacher logits
        with torch.cuda.stream(self.stream):
            A["weights"] = torch.ones(N, dtype=torch.float32, device=self.device)   # instanceWeights (getBatchEmoVoxCeleb.m:37)
        for L in self.layers:
            n = L["name"]
            A[n + ":raw"] = self.f16(N, L["oh"], L["ow"], L["kp"])
            A[n + ":draw"] = self.f16(N, L["oh"], L["ow"], L["kp"])
            if L["bn"]:
                A[n + ":a"], A[n + ":b"] = self.f32(L["cout"]), self.f32(L["cout"])
                A[n + ":ws"] = torch.zeros(2 * L["cout"], dtype=torch.float64, device=self.device)
            P = L["pool"]
            if P:
                pc = self.pool1_ld if n == "conv1" else L["cout"]   # (padding channels stay zero: never written)
                A[n + ":out"] = self.f16(N, P["oh"], P["ow"], pc)
                A[n + ":dout"] = self.f16(N, P["oh"], P["ow"], pc)
                if P["method"] == "max":
                    A[n + ":arg"] = torch.zeros((N, P["oh"], P["ow"], pc), dtype=torch.uint8, device=self.device)
                    if n == "conv1" and self.stem_algebra:
                        A[n + ":xwin"] = self.f16(N, P["oh"], P["ow"], pc)
                else:
                    A[n + ":act"] = self.f16(N, L["oh"], L["ow"], L["cout"])
                    A[n + ":dact"] = self.f16(N, L["oh"], L["ow"], L["cout"])
            elif L["bn"]:
                A[n + ":out"] = self.f16(N, L["oh"], L["ow"], L["cout"])
                A[n + ":dout"] = self.f16(N, L["oh"], L["ow"], L["cout"])
            if n not in ("conv1",):
                A[n + ":packed"] = self.f16(int(self.ctx.lib.xemo_dgrad_pack_elems(L["cp"], L["kp"], L["fh"], L["fw"],
                                                                                  L["stride"][0], L["stride"][1])))
        A["pred32"] = self.f32(N, self.layers[-1]["kp"])
        A["scalars"] = self.f32(2)          # objective, classerror (accumulated)
        A["class_stats"] = self.f32(2 * self.K)
        A["max_label"] = torch.zeros(N, dtype=torch.int32, device=self.device)

    # ---- forward
    def _record_forward(self, train):
        if not train:
            return self._record_forward_test()
        N, A, ctx = self.N, self.a, self.ctx
        self._record_frontend()
        ctx.op_spec_s2d(_p(A["spec"]), 512, self.W, N, 1, 1, self.s2d_hp, self.s2d_ow, _p(A["s2d"]))
        if self.stem_algebra:
            ctx.op_stem_autocorr(_p(A["s2d"]), N, self.s2d_hp, self.s2d_ow, self.layers[0]["oh"], _p(A["stem:ws"]))
        cur = A["s2d"]
        for L in self.layers:
            n = L["name"]
            wt, bias = self.view(self.w16, n + "f"), self.view(self.master, n + "b")
            last = n == "fc8"
            if n == "conv1":
                self._stem_conv(wt, None, bias, 0, A[n + ":raw"])
            else:
                self.conv(cur, N, L["h"], L["w"], L["cp"], wt, L["kp"], L["fh"], L["fw"], L["stride"], L["pad"], None, bias,
                          None, 0, A[n + ":raw"], A["pred32"] if last else None, L["kp"])
            cur = A[n + ":raw"]
            if not L["bn"]:
                continue
            bn = "bn" + n[-1]
            g, beta = self.view(self.master, bn + "m"), self.view(self.master, bn + "b")
            rows = N * L["oh"] * L["ow"]
            stem = n == "conv1" and self.stem_algebra
            if stem:   # batch statistics of w.patch + b from the patch autocorrelation: no pass over the activation
                ctx.op_stem_bn_train(_p(A["stem:ws"]), _p(wt), _p(bias), rows, L["cout"], _p(g), _p(beta), BN_EPS,
                                     _p(self.batch_moments[bn]), _p(A[n + ":a"]), _p(A[n + ":b"]))
            else:
                ctx.op_bn_train(_p(cur), rows, L["cout"], _p(g), _p(beta), BN_EPS, _p(A[n + ":ws"]), _p(self.batch_moments[bn]),
                                _p(A[n + ":a"]), _p(A[n + ":b"]))
            P = L["pool"]
            if P and P["method"] == "max" and stem:
                ctx.op_maxpool_fwd_win(_p(cur), N, L["oh"], L["ow"], L["cout"], P["win"][0], P["win"][1], P["stride"][0],
                                       P["stride"][1], 0, 0, 0, 0, _p(A[n + ":a"]), _p(A[n + ":b"]), _p(A[n + ":out"]),
                                       _p(A[n + ":arg"]), _p(A[n + ":xwin"]), self.pool1_ld)
            elif P and P["method"] == "max":
                ctx.op_maxpool_fwd(_p(cur), N, L["oh"], L["ow"], L["cout"], P["win"][0], P["win"][1], P["stride"][0], P["stride"][1],
                                   0, 0, 0, 0, _p(A[n + ":a"]), _p(A[n + ":b"]), _p(A[n + ":out"]), _p(A[n + ":arg"]))
            elif P:
                ctx.op_affine_act(_p(cur), rows, L["cout"], _p(A[n + ":a"]), _p(A[n + ":b"]), 1, _p(A[n + ":act"]))
                ctx.op_avgpool_fwd(_p(A[n + ":act"]), N, L["oh"], L["ow"], L["cout"], P["win"][0], P["win"][1], P["stride"][0],
                                   P["stride"][1], 0, 0, 0, 0, _p(A[n + ":out"]))
            else:
                ctx.op_affine_act(_p(cur), rows, L["cout"], _p(A[n + ":a"]), _p(A[n + ":b"]), 1, _p(A[n + ":out"]))
            cur = A[n + ":out"]

    def _stem_conv(self, wt, scale, shift, relu, dst, x=None, n=None, prepare=True):
        """conv1 as a 4 x 1 convolution over the space-to-depth tensor (`x`, `n` clips: default the whole batch); in
        pixel-pair form when enabled (`prepare`: expand the block-diagonal filter / duplicated epilogue vectors first)."""
        A, ctx, L = self.a, self.ctx, self.layers[0]
        x = A["s2d"] if x is None else x
        N = self.N if n is None else n
        if not self.stem_pairs:
            self.conv(x, N, self.s2d_hp, self.s2d_ow, 16, wt, L["kp"], 4, 1, (1, 1), (0, 0, 0, 0), scale, shift, None, relu, dst)
            return
        kp = L["kp"]
        if prepare:
            ctx.op_stem_pair_filter(_p(wt), kp, _p(A["stem:w2"]))
            ctx.op_tile_f32(_p(shift), kp, 2, 0.0, _p(A["stem:shift2"]))
            if scale is not None:
                ctx.op_tile_f32(_p(scale), kp, 2, 1.0, _p(A["stem:scale2"]))
        self.conv(x, N, self.s2d_hp, self.s2d_ow // 2, 32, A["stem:w2"], 2 * kp, 4, 1, (1, 1), (0, 0, 0, 0),
                  A["stem:scale2"] if scale is not None else None, A["stem:shift2"], None, relu, dst)

    def _record_forward_test(self):
        """dag.mode = 'test' (external/compute_audio_feats.m:106): BN uses the stored moments, so it folds -- together
        with the conv bias -- into the convolution's scale/shift epilogue and the activation is rounded to fp16 once."""
        N, A, ctx = self.N, self.a, self.ctx
        self._record_frontend()
        ctx.op_spec_s2d(_p(A["spec"]), 512, self.W, N, 1, 1, self.s2d_hp, self.s2d_ow, _p(A["s2d"]))
        cur = A["s2d"]
        for L in self.layers:
            n = L["name"]
            wt, bias = self.view(self.w16, n + "f"), self.view(self.master, n + "b")
            scale, shift, relu, out32 = None, bias, 0, None
            if L["bn"]:
                bn = "bn" + n[-1]
                ctx.op_bn_test(_p(self.moments[bn]), L["cout"], _p(self.view(self.master, bn + "m")), _p(self.view(self.master, bn + "b")),
                               _p(bias), _p(A[n + ":a"]), _p(A[n + ":b"]))
                scale, shift, relu = A[n + ":a"], A[n + ":b"], 1
            else:
                out32 = A["pred32"]
            P = L["pool"]
            dst = A[n + ":raw"] if (P or not L["bn"]) else A[n + ":out"]
            if n == "conv1":
                self._stem_conv(wt, scale, shift, relu, dst)
            else:
                self.conv(cur, N, L["h"], L["w"], L["cp"], wt, L["kp"], L["fh"], L["fw"], L["stride"], L["pad"], scale, shift,
                          None, relu, dst, out32, L["kp"])
            if P and P["method"] == "max" and n == "conv1" and self.pool1_ld != L["cout"]:
                ctx.op_maxpool_fwd_win(_p(dst), N, L["oh"], L["ow"], L["cout"], P["win"][0], P["win"][1], P["stride"][0], P["stride"][1],
                                       0, 0, 0, 0, None, None, _p(A[n + ":out"]), None, None, self.pool1_ld)
                dst = A[n + ":out"]
            elif P and P["method"] == "max":
                ctx.op_maxpool_fwd(_p(dst), N, L["oh"], L["ow"], L["cout"], P["win"][0], P["win"][1], P["stride"][0], P["stride"][1],
                                   0, 0, 0, 0, None, None, _p(A[n + ":out"]), None)
                dst = A[n + ":out"]
            elif P:
                ctx.op_avgpool_fwd(_p(dst), N, L["oh"], L["ow"], L["cout"], P["win"][0], P["win"][1], P["stride"][0], P["stride"][1],
                                   0, 0, 0, 0, _p(A[n + ":out"]))
                dst = A[n + ":out"]
            cur = dst

    def _record_frontend(self):
        """runSpec + row normalisation on the device (audio_input == 'wav'); samples are scaled by 2^15 as runSpec does
        for [-1, 1] audio (the row normalisation makes the result scale-free)."""
        if self.audio_input != "wav":
            return
        A, ctx = self.a, self.ctx
        ctx.op_spectrogram(_p(A["wav"]), self.N, self.wav_len, self.Nw, self.Ns, 512, self.AUDIO["alpha"], 32768.0, self.W, _p(A["spec"]))
        ctx.op_spec_rownorm(_p(A["spec"]), 512, self.W, self.N)

    # ---- loss + backward
    def _record_backward(self, lo=0, hi=None, loss=True):
        """Loss (when `loss`) and the backward sweep over layers [lo, hi) in reverse order.  The split form lets the
        data-parallel step all-reduce the gradients of the late layers (fc6-fc8: 82 % of the bytes) while the early
        layers are still being differentiated."""
        N, A, ctx, gs = self.N, self.a, self.ctx, self.grad_scale
        inv = 1.0 / gs
        last = self.layers[-1]
        hi = len(self.layers) if hi is None else hi
        if loss:
            ctx.memset(_p(self.grad), 0, self.nparam * 4)
            ctx.memset(_p(A["fc8:draw"]), 0, A["fc8:draw"].numel() * 2)
            ctx.memset(_p(A["scalars"]), 0, 8)  # objective / classerror of THIS batch (class_stats keep accumulating)
            soft, lt = self.loss_type == "hot-cross-ent", LOSS_TYPES[self.loss_type]
            # the loss reads the fp32 logits the fc8 convolution writes beside its fp16 output (T of the non-softmax losses:
            # huber's sigma = 1, emoVoxZoo.m:146)
            ctx.op_loss(_p(A["pred32"]), 1, last["kp"], _p(A["target"]), self.K, _p(A["weights"]) if lt else None, N, self.K, lt,
                        self.T if soft else 1.0, 1 if soft else 0, 1.0, gs, _p(A["fc8:draw"]), 0, last["kp"], _p(A["scalars"]),
                        _p(A["class_stats"]), _p(A["max_label"]))
        for i in range(hi - 1, lo - 1, -1):
            L = self.layers[i]
            n = L["name"]
            rows = N * L["oh"] * L["ow"]
            fused_bias = False
            stem = n == "conv1" and self.stem_algebra
            if stem:
                bn, P = "bn1", L["pool"]
                prow = N * P["oh"] * P["ow"]
                # ReLU mask + the two BN reductions at the pooled resolution, then the (masked) gradient w.r.t. the
                # never-materialised ReLU output at the conv resolution: dz, which the filter gradient consumes directly
                ld = self.pool1_ld
                ctx.op_stem_pool_bn_reduce(_p(A[n + ":xwin"]), _p(A[n + ":dout"]), prow, L["cout"], ld, _p(self.batch_moments[bn]),
                                           _p(A[n + ":a"]), _p(A[n + ":b"]), _p(A[n + ":ws"]))
                ctx.op_maxpool_bwd_ld(_p(A[n + ":dout"]), _p(A[n + ":arg"]), N, L["oh"], L["ow"], L["cout"], P["win"][0], P["win"][1],
                                      P["stride"][0], P["stride"][1], 0, 0, 0, 0, _p(A[n + ":draw"]), ld)
                fused_bias = True
            elif L["bn"]:
                bn = "bn" + n[-1]
                P = L["pool"]
                dcur = A[n + ":dout"]
                gbias = self.view(self.grad, n + "b")
                common = (_p(self.batch_moments[bn]), _p(A[n + ":a"]), _p(A[n + ":b"]))
                outs = (_p(A[n + ":ws"]), _p(A[n + ":draw"]), _p(self.view(self.grad, bn + "m")), _p(self.view(self.grad, bn + "b")),
                        _p(gbias), inv)
                fused_bias = L["kp"] == L["cout"]
                if not fused_bias:
                    outs = outs[:4] + (None, inv)
                if P and P["method"] == "max" and self.fuse_pool_bwd:
                    # BN backward reads the pooled gradient through the arg-max (measured slower than the two-pass
                    # form on B200 -- the gather is instruction-bound -- so off by default)
                    ctx.op_bn_bwd_pool(_p(A[n + ":raw"]), _p(dcur), _p(A[n + ":arg"]), N, L["oh"], L["ow"], L["cout"], P["win"][0],
                                       P["win"][1], P["stride"][0], P["stride"][1], 0, 0, 0, 0, *common, *outs)
                else:
                    if P and P["method"] == "max":
                        # gradient w.r.t. the (never materialised) ReLU output, NHWC at the conv resolution
                        ctx.op_maxpool_bwd(_p(dcur), _p(A[n + ":arg"]), N, L["oh"], L["ow"], L["cout"], P["win"][0], P["win"][1],
                                           P["stride"][0], P["stride"][1], 0, 0, 0, 0, _p(A[n + ":draw"]))
                        dcur = A[n + ":draw"]
                    elif P:
                        ctx.op_avgpool_bwd(_p(dcur), N, L["oh"], L["ow"], L["cout"], P["win"][0], P["win"][1], P["stride"][0],
                                           P["stride"][1], 0, 0, 0, 0, _p(A[n + ":dact"]))
                        dcur = A[n + ":dact"]
                    ctx.op_bn_bwd(_p(A[n + ":raw"]), _p(dcur), rows, L["cout"], *common, 1, 0, *outs)
            dy = A[n + ":draw"]
            x = A["s2d"] if i == 0 else self._input_of(i)
            side = self.side_stream is not None
            if side:  # fork: the filter gradient needs dy (just produced) but nothing downstream needs it before the update
                ctx.stream_wait(VP(self.side_stream.cuda_stream), None)
                ctx.set_stream(VP(self.side_stream.cuda_stream))
            if stem:
                gf = self.view(self.grad, n + "f")
                if self.stem_pairs and self.stem_wgrad_pairs:   # G1 in pixel-pair form (measured slower: 0.80 vs 0.74 ms)
                    g1p = A["stem:g1pair"]
                    ctx.memset(_p(g1p), 0, g1p.numel() * 4)
                    ctx.op_conv_wgrad(_p(x), N, self.s2d_hp, self.s2d_ow // 2, 32, _p(dy), 2 * L["kp"], 2 * L["kp"], 4, 1, 1, 1,
                                      0, 0, 0, 0, _p(g1p), inv)
                else:
                    g1p = None
                    ctx.op_conv_wgrad(_p(x), N, self.s2d_hp, self.s2d_ow, 16, _p(dy), L["kp"], L["kp"], 4, 1, 1, 1, 0, 0, 0, 0, _p(gf), inv)
                ctx.op_stem_wgrad_finalize(_p(A["stem:ws"]), _p(self.view(self.w16, n + "f")), _p(self.view(self.master, n + "b")),
                                           _p(A[n + ":ws"]), rows, L["cout"], _p(self.batch_moments["bn1"]), _p(A[n + ":a"]), inv,
                                           _p(gf), _p(self.view(self.grad, n + "b")), _p(self.view(self.grad, "bn1m")),
                                           _p(self.view(self.grad, "bn1b")), _p(g1p))
            elif n == "conv1":
                ctx.op_conv_wgrad(_p(x), N, self.s2d_hp, self.s2d_ow, 16, _p(dy), L["kp"], L["kp"], 4, 1, 1, 1, 0, 0, 0, 0,
                                  _p(self.view(self.grad, n + "f")), inv)
                # structurally-zero slots of the space-to-depth filter: column 7 / 15 of every tap, and tap 3 rows 8..15
                gf = self.view(self.grad, n + "f")
                ctx.op_fill_strided_f32(_p(gf), L["kp"] * 4, 16, 7, 1, 0.0)
                ctx.op_fill_strided_f32(_p(gf), L["kp"] * 4, 16, 15, 1, 0.0)
                ctx.op_fill_strided_f32(_p(gf), L["kp"], 64, 3 * 16 + 8, 8, 0.0)
            else:
                cp = L["cp"]
                ctx.op_conv_wgrad(_p(x), N, L["h"], L["w"], cp, _p(dy), L["kp"], L["kp"], L["fh"], L["fw"], L["stride"][0],
                                  L["stride"][1], *L["pad"], _p(self.view(self.grad, n + "f")), inv)
            if not fused_bias:
                ctx.op_colsum(_p(dy), rows, L["kp"], L["kp"], inv, _p(self.view(self.grad, n + "b")))
            if side:
                ctx.set_stream(None)
            if i > 0 and self._full_height(L):
                # fc6: a 9 x 1 filter over a 9 x W map -> one output row: the data gradient is a plain GEMM over (n, w) rows
                # (the general form walks 9 taps of which 8 fall outside dY for every pixel)
                cp = L["cp"]
                ctx.op_pack_dgrad_filters_fullheight(_p(self.view(self.w16, n + "f")), L["kp"], L["fh"], cp, _p(A[n + ":packed"]))
                ctx.op_conv_dgrad_fullheight(_p(dy), N, L["h"], L["w"], cp, _p(A[n + ":packed"]), L["kp"],
                                             _p(A[self.layers[i - 1]["name"] + ":dout"]))
            elif i > 0:
                cp = L["cp"]
                ctx.op_pack_dgrad_filters(_p(self.view(self.w16, n + "f")), L["kp"], L["fh"], L["fw"], cp, L["stride"][0],
                                          L["stride"][1], L["pad"][0], L["pad"][2], _p(A[n + ":packed"]))
                ctx.op_conv_dgrad(_p(dy), N, L["h"], L["w"], cp, _p(A[n + ":packed"]), L["kp"], L["fh"], L["fw"], L["stride"][0],
                                  L["stride"][1], *L["pad"], _p(A[self.layers[i - 1]["name"] + ":dout"]))

        if self.side_stream is not None:
            ctx.stream_wait(None, VP(self.side_stream.cuda_stream))  # join before the update / all-reduce

    @staticmethod
    def _full_height(L):
        """a filter as tall as its input, one column wide, unpadded, stride 1: one output row (the same rule as xemo_net.cu)"""
        return (L["fh"] == L["h"] and L["fw"] == 1 and L["fh"] > 1 and L["stride"] == (1, 1) and L["pad"] == (0, 0, 0, 0)
                and L["cp"] <= 256 and os.environ.get("XEMO_DGRAD_FULLHEIGHT", "1") != "0")

    def _input_of(self, i):
        return self.a[self.layers[i - 1]["name"] + ":out"]

    def _record_update(self):
        """cnn_train_dag's accumulateGradients: every 'gradient' parameter has learningRate = weightDecay = 1 in
        this graph (dag.initParams defaults), so the whole flat master buffer is updated by ONE launch that also
        refreshes the fp16 mirror the tensor-core kernels read (alignment padding stays zero: g = w = m = 0)."""
        ctx = self.ctx
        # the activation gradients travel in fp16 under a fixed loss scale: a non-finite element in the (all-reduced) flat
        # gradient sets the guard, the update of this step is skipped and counted (metrics()['skipped_steps'])
        guard = _p(self.guard)
        ctx.op_grad_guard(_p(self.grad), self.nparam, guard)
        ctx.op_sgd_momentum_guarded(_p(self.master), _p(self.momentum), _p(self.grad), self.nparam, _p(self.hyper), 1.0, 1.0, 1.0,
                                    _p(self.w16), guard)
        for bn, m in self.moments.items():
            ctx.op_moments_average_guarded(_p(m), _p(self.batch_moments[bn]), m.numel(), 0.1, 1.0, guard)

    # ---- graph plumbing
    def _run(self, key, record):
        if not self.use_graph:
            record()
            return
        g = self.graphs.get(key)
        if g is None:
            record()  # eager warm-up: kernel attributes must be set outside a capture
            self.ctx.capture_begin()
            record()
            g = self.graphs[key] = self.ctx.capture_end()
            if key != "fwd_test":
                return  # the warm-up already executed this phase once with the same inputs
        g.launch()

    def set_hyper(self, lr=None, momentum=None, weight_decay=None, batch_size=None):
        h = self.hyper.cpu().numpy()
        for i, v in enumerate((lr, momentum, weight_decay, None if batch_size is None else 1.0 / batch_size)):
            if v is not None:
                h[i] = v
        with torch.cuda.stream(self.stream):
            self.hyper.copy_(torch.from_numpy(h))

    def set_input(self, spec, target=None, weights=None):
        """spec: 512 x W x 1 x N spectrograms, or N x L waveforms when audio_input == 'wav'; target: logitTarget
        (1 x 1 x K x N) or maxLabel (softmaxlog); weights: instanceWeights 1 x 1 x 1 x N (euclidean / huber)."""
        key = "wav" if self.audio_input == "wav" else "spec"
        if isinstance(spec, np.ndarray):
            if key == "wav":
                spec = torch.from_numpy(np.ascontiguousarray(spec, dtype=np.float32).reshape(-1))
            else:
                spec = torch.from_numpy(np.ascontiguousarray(spec.astype(np.float32).transpose(3, 2, 1, 0)).reshape(-1))
        with torch.cuda.stream(self.stream):
            self.a[key].view(-1).copy_(spec.reshape(-1), non_blocking=True)
            if target is not None:
                if isinstance(target, np.ndarray) and self.loss_type == "softmaxlog":
                    # maxLabel (1 x 1 x 1 x N, 1-based; getBatchEmoVoxCeleb.m:32) -> one-hot rows
                    lab = np.asarray(target).reshape(-1).astype(np.int64)
                    assert lab.size == self.N and lab.min() >= 1 and lab.max() <= self.K, "maxLabel must hold N labels in 1..K"
                    onehot = np.zeros((self.N, self.K), np.float32)
                    onehot[np.arange(self.N), lab - 1] = 1.0
                    target = torch.from_numpy(onehot)
                elif isinstance(target, np.ndarray):
                    target = torch.from_numpy(np.ascontiguousarray(target.astype(np.float32).reshape(self.K, self.N).T))
                self.a["target"].copy_(target.reshape(self.N, self.K), non_blocking=True)
            if weights is not None:
                self.a["weights"].copy_(torch.from_numpy(np.ascontiguousarray(np.asarray(weights, np.float32).reshape(self.N))),
                                        non_blocking=True)

    def forward(self, spec, mode="test"):
        """dag.eval({'data', spec}) -> N x K numpy predictions."""
        self.set_input(spec)
        if mode == "test":
            self._run("fwd_test", lambda: self._record_forward(False))
        else:
            self._record_forward(True)
        with torch.cuda.stream(self.stream):
            out = self.a["pred32"][:, : self.K].cpu()
        self.sync()
        return out.numpy()

    def grad_step(self):
        """forward (train-mode BN) + loss + backward on the current inputs; gradients land in self.grad."""
        def rec():
            self._record_forward(True)
            self._record_backward()
        if not self.use_graph:
            rec()
            return
        g = self.graphs.get("grad")
        if g is None:
            rec()  # eager warm-up
            self.reset_metrics()
            self.ctx.capture_begin()
            rec()
            g = self.graphs["grad"] = self.ctx.capture_end()
        g.launch()

    def update(self):
        def rec():
            self._record_update()
        if not self.use_graph:
            rec()
            return
        g = self.graphs.get("update")
        if g is None:
            self.ctx.capture_begin()
            rec()
            g = self.graphs["update"] = self.ctx.capture_end()
        g.launch()

    def reset_metrics(self):
        self.ctx.memset(_p(self.a["scalars"]), 0, 8)
        self.ctx.memset(_p(self.a["class_stats"]), 0, 8 * self.K)

    def train_step(self, spec, target, allreduce=None, weights=None):
        """One cnn_train_dag iteration.  `allreduce(flat_grad_tensor)` (optional) sums gradients across
        data-parallel ranks between the backward pass and the update."""
        self.set_input(spec, target, weights)
        self.grad_step()
        if allreduce is not None:
            with torch.cuda.stream(self.stream):
                allreduce(self.grad)
        self.update()

    def metrics(self):
        with torch.cuda.stream(self.stream):
            s = self.a["scalars"].cpu()
            cs = self.a["class_stats"].cpu()
            gd = self.guard.cpu()
        self.sync()
        return dict(objective=float(s[0]), classerror=float(s[1]), correct=cs[: self.K].numpy(), count=cs[self.K :].numpy(),
                    nonfinite_grad=bool(gd[0]), skipped_steps=int(gd[1]))

    # ---- export in MatConvNet layouts (parity checks, checkpoints)
    def _export(self, buf):
        flat = buf.cpu().numpy()
        out = {}
        for L in self.layers:
            n = L["name"]
            o, shape = self.segs[n + "f"]
            w = flat[o : o + int(np.prod(shape))].reshape(shape)
            out[n + "f"] = student_conv1_from_s2d(w) if n == "conv1" else unkrsc(w, L["fh"], L["fw"], L["cin"], L["cout"])
            o, shape = self.segs[n + "b"]
            out[n + "b"] = flat[o : o + L["cout"]].copy()
            if L["bn"]:
                bn = "bn" + n[-1]
                for s in ("m", "b"):
                    o, shape = self.segs[bn + s]
                    out[bn + s] = flat[o : o + L["cout"]].copy()
        return out

    def export_momentum(self):
        self.sync()
        return self._export(self.momentum)

    def load_momentum(self, momentum):
        """Restore the optimiser state exported by `_export(self.momentum)` (checkpoint resume)."""
        flat = np.zeros(self.nparam, np.float32)
        for L in self.layers:
            n = L["name"]
            o, shape = self.segs[n + "f"]
            w = student_conv1_to_s2d(momentum[n + "f"]) if n == "conv1" else krsc(momentum[n + "f"], cp=L["cp"])
            flat[o : o + w.size] = w.reshape(-1)
            o, _ = self.segs[n + "b"]
            flat[o : o + L["cout"]] = momentum[n + "b"]
            if L["bn"]:
                bn = "bn" + n[-1]
                for s_ in ("m", "b"):
                    o, _ = self.segs[bn + s_]
                    flat[o : o + L["cout"]] = momentum[bn + s_]
        with torch.cuda.stream(self.stream):
            self.momentum.copy_(torch.from_numpy(flat))

    def export_params(self):
        self.sync()
        out = self._export(self.master)
        for bn, m in self.moments.items():
            c = m.numel() // 2
            out[bn + "x"] = m.cpu().numpy().reshape(2, c).T.copy()
        return out

    def export_decisions(self):
        """The discrete decisions of the last train-mode forward, in MatConvNet layout (H x W x C x N): the ReLU masks
        `a x + b > 0` the backward kernels recompute from the raw convolution outputs ('relu<i>') and the uint8 window-local
        arg-max dw*PH + dh of every max pool ('pool<i>').  Diagnostics / parity tests: running a reference backward under
        these decisions separates arithmetic error from decision flips."""
        self.sync()
        out = {}
        for L in self.layers:
            n, i = L["name"], L["name"][-1]
            if not L["bn"]:
                continue
            raw = self.a[n + ":raw"].cpu().numpy()[..., : L["cout"]].astype(np.float32)
            a, b = self.a[n + ":a"].cpu().numpy(), self.a[n + ":b"].cpu().numpy()
            out["relu" + i] = np.transpose(a * raw + b > 0, (1, 2, 3, 0))
            if L["pool"] and L["pool"]["method"] == "max":
                out["pool" + i] = np.transpose(self.a[n + ":arg"].cpu().numpy()[..., : L["cout"]], (1, 2, 3, 0))
        return out

    def export_grads(self):
        self.sync()
        out = self._export(self.grad)
        for bn, m in self.batch_moments.items():
            c = m.numel() // 2
            out[bn + "x"] = m.cpu().numpy().reshape(2, c).T.copy()
        return out
