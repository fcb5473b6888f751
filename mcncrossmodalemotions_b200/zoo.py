"""Model zoo: the graphs `emoVoxZoo` / `ferPlusZoo` return in the reference
(emoVoxCeleb/emoVoxZoo.m:1-62, teacher/ferPlusZoo.m:95-114), as parameter dictionaries in MatConvNet
layouts for the fused programs.

The reference downloads `<name>.mat` (emoVoxZoo.m:95-97); with no file at hand the architectures are
instantiated with seeded synthetic parameters: the student exactly as `dag.initParams()` leaves it
(emoVoxZoo.m:54 -- He-normal filters, zero biases, BN mult 1 / bias 0 / moments 0), the teachers with a
trained-like BN state so that activations stay O(1) through the 16 bottlenecks."""
from __future__ import annotations

import numpy as np

from .arch import STUDENT_CONVS, TEACHER_STAGES

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


class _Block:
    """dagnn block stand-in: `type` (e.g. 'dagnn.Pooling') plus the attributes the reference scripts touch
    (`poolSize`, emoVoxCeleb/emoVoxZoo.m:269, external/compute_audio_feats.m:125)."""

    def __init__(self, type_, **attrs):
        self.type = type_
        self.__dict__.update(attrs)

    def isa(self, cls):
        loss_family = ("dagnn.Loss", "dagnn.SoftmaxCELoss", "dagnn.EuclideanLoss", "dagnn.HuberLoss", "dagnn.VerboseLoss", "dagnn.ErrorStats")
        return self.type == cls or (cls == "dagnn.Loss" and self.type in loss_family)


class _Layer:
    def __init__(self, name, block, inputs, outputs, params=()):
        self.name, self.block, self.inputs, self.outputs, self.params = name, block, list(inputs), list(outputs), list(params)


class _Var:
    def __init__(self, name):
        self.name, self.value = name, None


class Model:
    """The dagnn.DagNN object `emoVoxZoo` / `ferPlusZoo` hand to the reference scripts, reduced to what those scripts
    touch: `layers(i).{name,block}`, `removeLayer`, `getLayerIndex`, `getInputs`, `renameVar`,
    `mode`, `move`, `meta.normalization`, `layers(pool6).block.poolSize`, `eval({'data', x})`, `vars(end).value`,
    `params`.  `eval` runs the fused device program of the graph; there is no layer-by-layer interpreter."""

    def __init__(self, name, params, kind, width=None, loss_type=None):
        self.name, self.params, self.kind = name, params, kind
        self.mode = "test" if kind == "teacher" else "normal"
        self.device = "cpu"
        self.meta = {"normalization": {"imageSize": (224, 224, 3) if kind == "teacher" else (512, width, 1),
                                       "averageImage": AVERAGE_IMAGE if kind == "teacher" else None},
                     "classes": {"name": ["neutral", "happiness", "surprise", "sadness", "anger", "disgust", "fear", "contempt"]}}
        self.layers, self.vars = [], [_Var("data")]
        self._build_layers(loss_type)
        self._prog = None

    # ---- graph description
    def _add(self, name, block, inputs, outputs, params=()):
        self.layers.append(_Layer(name, block, inputs, outputs, params))
        for o in outputs:
            if o not in [v.name for v in self.vars]:
                self.vars.append(_Var(o))

    def _build_layers(self, loss_type):
        prev = "data"
        if self.kind == "student":
            width = self.meta["normalization"]["imageSize"][1]
            for name, fh, fw, cin, cout, stride, pad, has_bn in STUDENT_CONVS:
                out = "prediction" if name == "fc8" else "x_" + name
                self._add(name, _Block("dagnn.Conv", size=(fh, fw, cin, cout), stride=stride, pad=pad), [prev], [out], [name + "f", name + "b"])
                prev = out
                if has_bn:
                    i = name[-1]
                    self._add("bn" + i, _Block("dagnn.BatchNorm", epsilon=1e-5), [prev], ["x_bn" + i], ["bn%sm" % i, "bn%sb" % i, "bn%sx" % i])
                    self._add("relu" + i, _Block("dagnn.ReLU"), ["x_bn" + i], ["x_relu" + i])
                    prev = "x_relu" + i
                pool = {"conv1": ("pool1", "max", (3, 3), (2, 2)), "conv2": ("pool2", "max", (3, 3), (2, 2)),
                        "conv5": ("pool5", "max", (5, 3), (3, 2)), "fc6": ("pool6", "avg", pool6_window(width), (1, 1))}.get(name)
                if pool:
                    self._add(pool[0], _Block("dagnn.Pooling", method=pool[1], poolSize=tuple(pool[2]), stride=pool[3], pad=(0, 0, 0, 0)),
                              [prev], ["x_" + pool[0]])
                    prev = "x_" + pool[0]
            if loss_type is not None:   # configureForRegression, emoVoxZoo.m:137-169
                if loss_type == "hot-cross-ent":
                    self._add("loss", _Block("dagnn.SoftmaxCELoss", temperature=2, logitTargets=True), ["prediction", "logitTarget"], ["objective"])
                elif loss_type == "softmaxlog":
                    self._add("loss", _Block("dagnn.Loss", loss="softmaxlog"), ["prediction", "maxLabel"], ["objective"])
                elif loss_type == "euclidean":
                    self._add("loss", _Block("dagnn.EuclideanLoss"), ["prediction", "logitTarget", "instanceWeights"], ["objective"])
                    # "scale down a lot to prevent exploding gradients" (emoVoxZoo.m:141-144): the head filters / 10
                    self.params["fc8f"] = (self.params["fc8f"] / 10).astype(self.params["fc8f"].dtype)
                elif loss_type == "huber":
                    self._add("loss", _Block("dagnn.HuberLoss", sigma=1), ["prediction", "logitTarget", "instanceWeights"], ["objective"])
                else:
                    raise ValueError("unrecognised regression loss: %s" % loss_type)
                self._add("error", _Block("dagnn.VerboseLoss", loss="classerror"), ["prediction", "maxLabel"], ["classerror"])
                self._add("errorStats", _Block("dagnn.ErrorStats", numClasses=8), ["prediction", "maxLabel"], ["errorStats"])
        else:
            self._add("conv1", _Block("dagnn.Conv", size=(7, 7, 3, 64), stride=(2, 2), pad=(3, 3, 3, 3)), ["data"], ["conv1"], ["conv1f"])
            self._add("pool1", _Block("dagnn.Pooling", method="max", poolSize=(3, 3), stride=(2, 2), pad=(0, 1, 0, 1)), ["conv1"], ["pool1"])
            prev = "pool1"
            for si, (blocks, mid, cout, stride) in enumerate(TEACHER_STAGES):
                for bi in range(blocks):
                    pre = "s%db%d" % (si + 2, bi + 1)
                    self._add(pre, _Block("bottleneck", mid=mid, out=cout, stride=stride if bi == 0 else 1, se=self.params["arch"] == "senet50"),
                              [prev], [pre])
                    prev = pre
            self._add("pool5", _Block("dagnn.Pooling", method="avg", poolSize=(7, 7), stride=(1, 1), pad=(0, 0, 0, 0)), [prev], ["pool5"])
            self._add("classifier", _Block("dagnn.Conv", size=(1, 1, 2048, 8), stride=(1, 1), pad=(0, 0, 0, 0)), ["pool5"], ["prediction"],
                      ["classifierf", "classifierb"])

    # ---- the DagNN methods the scripts call
    def getLayerIndex(self, name):
        for i, l in enumerate(self.layers):
            if l.name == name:
                return i
        raise KeyError("no layer named %s" % name)

    def removeLayer(self, name):
        i = self.getLayerIndex(name)
        if not self.layers[i].block.isa("dagnn.Loss"):
            raise ValueError("only loss / metric layers can be removed from the fused graphs")
        del self.layers[i]
        used = {v for l in self.layers for v in l.inputs + l.outputs}
        self.vars = [v for v in self.vars if v.name in used]

    def getInputs(self):
        produced = {o for l in self.layers for o in l.outputs}
        seen, out = set(), []
        for l in self.layers:
            for v in l.inputs:
                if v not in produced and v not in seen:
                    seen.add(v)
                    out.append(v)
        return out

    def renameVar(self, old, new):
        for l in self.layers:
            l.inputs = [new if v == old else v for v in l.inputs]
            l.outputs = [new if v == old else v for v in l.outputs]
        for v in self.vars:
            if v.name == old:
                v.name = new

    @property
    def pool6(self):
        return self.layers[self.getLayerIndex("pool6")].block.poolSize if self.kind == "student" else None

    def move(self, device):
        if device not in ("gpu", "cpu"):
            raise ValueError("move: device must be 'gpu' or 'cpu'")
        if device == "cpu":
            raise RuntimeError("this back-end has no CPU path; the model stays on the GPU")
        self.device = device

    def eval(self, inputs):
        """dag.eval({'data', x}): runs the network through the graph-level C ABI (net.py), stores `prediction` (1 x 1 x K x N) in vars(end).value when it
        is the last variable (losses removed), and returns it as N x K (`gather(squeeze(dag.vars(end).value))'`)."""
        from .net import StudentNet, TeacherNet

        if isinstance(inputs, (list, tuple)):   # MATLAB-style {'data', x}
            inputs = dict(zip(inputs[0::2], inputs[1::2]))
        x = inputs[self.getInputs()[0]]
        n = x.shape[-1]
        if self.kind == "teacher":
            if self._prog is None or self._prog.N != n:
                self._prog = TeacherNet(self.params, n, input_mode="u8" if x.ndim == 3 else "hwcn224",
                                            face_size=x.shape[0] if x.ndim == 3 else 48)
            out = self._prog.forward(x)
        else:
            width = x.shape[1]
            if self.pool6 != pool6_window(width):
                raise ValueError("pool6.poolSize = %s does not match a %d-column input (expected %s): set it from the width bucket "
                                 "as compute_audio_feats.m:121-125 does" % (self.pool6, width, pool6_window(width)))
            if self._prog is None or self._prog.N != n or self._prog.W != width:
                self._prog = StudentNet(self.params, n, width)
            out = self._prog.forward(x, "test" if self.mode == "test" else "train")
        for v in self.vars:
            if v.name == "prediction":
                v.value = out.T.reshape(1, 1, out.shape[1], n)
        return out


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
    return Model(modelName, student_init(seed, numOutputs), "student", width, loss_type=lossType if scratch else None)


def ferPlusZoo(modelName):
    """ferPlusZoo (teacher/ferPlusZoo.m:95-114): the pretrained branch falls through to emoVoxZoo(modelName)."""
    if modelName not in TEACHERS:
        raise ValueError("unrecognised FER+ model: %s" % modelName)
    return emoVoxZoo(modelName)
