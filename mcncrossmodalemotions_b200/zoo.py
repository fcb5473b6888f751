"""Model zoo: the graphs `emoVoxZoo` / `ferPlusZoo` return in the reference
(emoVoxCeleb/emoVoxZoo.m:1-62, teacher/ferPlusZoo.m:95-114), as parameter dictionaries in MatConvNet
layouts for the fused programs.

The reference downloads `<name>.mat` (emoVoxZoo.m:95-97); with no file at hand the architectures are
instantiated with seeded synthetic parameters: the student exactly as `dag.initParams()` leaves it
(emoVoxZoo.m:54 -- He-normal filters, zero biases, BN mult 1 / bias 0 / moments 0), the teachers with a
trained-like BN state so that activations stay O(1) through the 16 bottlenecks."""
from __future__ import annotations

import numpy as np

from .programs import STUDENT_CONVS, TEACHER_STAGES

TEACHERS = ("resnet50-ferplus", "senet50-ferplus")
STUDENTS = ("emovoxceleb-student",)
# width bucket -> pool6 window (emoVoxCeleb/emoVoxZoo.m:258-259, external/compute_audio_feats.m:45-46)
POOL6_BUCKETS = {100: 2, 200: 5, 300: 8, 400: 11, 500: 14, 600: 17, 700: 20, 800: 23, 900: 27, 1000: 30}
AVERAGE_IMAGE = (131.0912, 103.8827, 91.4953)
PIXEL_SCALE = 100.0


def student_init(seed=3, num_outputs=8):
    """prepareFromDagNN + dag.initParams() (emoVoxZoo.m:187-253,54)."""
    rng = np.random.default_rng(seed)
    p = {}
    for name, fh, fw, cin, cout, _, _, has_bn in STUDENT_CONVS:
        if name == "fc8":
            cout = num_outputs
        p[name + "f"] = (rng.standard_normal((fh, fw, cin, cout)) * np.sqrt(2.0 / (fh * fw * cin))).astype(np.float32)
        p[name + "b"] = np.zeros(cout, np.float32)
        if has_bn:
            bn = "bn" + name[-1]
            p[bn + "m"] = np.ones(cout, np.float32)
            p[bn + "b"] = np.zeros(cout, np.float32)
            p[bn + "x"] = np.zeros((cout, 2), np.float32)
    return p


def teacher_init(arch="senet50", seed=4, num_outputs=8):
    arch = arch.replace("-ferplus", "")
    if arch not in ("resnet50", "senet50"):
        raise ValueError("unknown teacher %r" % (arch,))
    rng = np.random.default_rng(seed)
    p = {"arch": arch}

    def conv(name, fh, fw, cin, cout):
        p[name + "f"] = (rng.standard_normal((fh, fw, cin, cout)) * np.sqrt(2.0 / (fh * fw * cin))).astype(np.float32)

    def bn(name, c, small=False, scale=1.0):
        lo, hi = (0.1, 0.3) if small else (0.5, 1.5)
        p[name + "m"] = rng.uniform(lo, hi, c).astype(np.float32)
        p[name + "b"] = (0.1 * rng.standard_normal(c)).astype(np.float32)
        p[name + "x"] = (scale * np.stack([0.1 * rng.standard_normal(c), rng.uniform(0.5, 1.5, c)], axis=1)).astype(np.float32)

    conv("conv1", 7, 7, 3, 64)
    bn("bn1", 64, scale=PIXEL_SCALE)
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
                conv(pre + "se1", 1, 1, cout, r); p[pre + "se1b"] = np.zeros(r, np.float32)
                conv(pre + "se2", 1, 1, r, cout); p[pre + "se2b"] = np.zeros(cout, np.float32)
            cin = cout
    conv("classifier", 1, 1, 2048, num_outputs)
    p["classifierb"] = (0.01 * rng.standard_normal(num_outputs)).astype(np.float32)
    return p


def pool6_window(width):
    """updatePooling (emoVoxZoo.m:256-269): pool6.poolSize = [1 p] for the clip's width bucket."""
    if width not in POOL6_BUCKETS:
        raise ValueError("spectrogram width %d is not one of the buckets %s" % (width, sorted(POOL6_BUCKETS)))
    return (1, POOL6_BUCKETS[width])


class Model:
    """What `emoVoxZoo` / `ferPlusZoo` hand to the reference scripts, reduced to what they touch: the parameters
    (MatConvNet layouts), `meta.normalization`, the layer names they look up (`pool6`), `mode`, `move` and `eval`."""

    def __init__(self, name, params, kind, width=None):
        self.name, self.params, self.kind = name, params, kind
        self.mode = "test" if kind == "teacher" else "normal"
        self.device = "cpu"
        self.meta = {"normalization": {"imageSize": (224, 224, 3) if kind == "teacher" else (512, width, 1),
                                       "averageImage": AVERAGE_IMAGE if kind == "teacher" else None}}
        self.pool6 = pool6_window(width) if kind == "student" else None
        self._prog = None

    def move(self, device):
        if device not in ("gpu", "cpu"):
            raise ValueError("move: device must be 'gpu' or 'cpu'")
        if device == "cpu":
            raise RuntimeError("this back-end has no CPU path; the model stays on the GPU")
        self.device = device

    def eval(self, inputs):
        """dag.eval({'data', x}) -> N x 8 (teacher: logits; student: predictions in the current mode's BN)."""
        from .programs import StudentProgram, TeacherProgram

        x = inputs["data"]
        n = x.shape[-1]
        if self._prog is None or self._prog.N != n:
            if self.kind == "teacher":
                self._prog = TeacherProgram(self.params, n, input_mode="u8" if x.ndim == 3 else "hwcn224",
                                            face_size=x.shape[0] if x.ndim == 3 else 48)
            else:
                self._prog = StudentProgram(self.params, n, x.shape[1])
        if self.kind == "teacher":
            return self._prog.forward(x)
        return self._prog.forward(x, "test" if self.mode == "test" else "train")


def emoVoxZoo(modelName, scratch=False, lossType="hot-cross-ent", numSeconds=4, numOutputs=8, seed=3):
    """emoVoxZoo(modelName, 'scratch', tf, 'lossType', ..., 'numSeconds', ..., 'numOutputs', ...) (emoVoxZoo.m:1,17-23).
    Student names build the VGGVox graph re-initialised as dag.initParams() leaves it (:50-62); teacher names
    (emoVoxZoo.m:28-31) return the FER+ teachers.  The released .mat weights are remote downloads (:95-97): with no
    file at hand the parameters are seeded synthetic ones (scratch=True is what run_distillation uses anyway)."""
    if modelName in TEACHERS:
        return Model(modelName, teacher_init(modelName), "teacher")
    if modelName not in STUDENTS:
        raise ValueError("unrecognised model: %s" % modelName)
    if lossType not in ("hot-cross-ent", "softmaxlog", "euclidean", "huber"):
        raise ValueError("unrecognised loss type: %s" % lossType)
    width = 100 * int(numSeconds)
    return Model(modelName, student_init(seed, numOutputs), "student", width)


def ferPlusZoo(modelName):
    """ferPlusZoo (teacher/ferPlusZoo.m:95-114): the pretrained branch falls through to emoVoxZoo(modelName)."""
    if modelName not in TEACHERS:
        raise ValueError("unrecognised FER+ model: %s" % modelName)
    return emoVoxZoo(modelName)
