"""DagNN `.mat` model files: the on-disk format either side of the hot path.

The reference loads `<name>.mat` with `load` -> `dagnn.DagNN.loadobj` (emoVoxCeleb/emoVoxZoo.m:40-48,199): a struct
with `layers(i).{name,type,inputs,outputs,params,block}`, `params(i).{name,value}` and `meta`.  `load_dagnn` reads
such a file with scipy (MAT v5-v7; v7.3/HDF5 files are not readable without h5py) and maps the parameters onto the
zoo's dictionaries *structurally* -- by walking the Conv / BatchNorm layers in execution order and checking every
shape against the architecture -- because the upstream parameter names are not pinned anywhere in the reference.
`save_dagnn` writes the same structure, so that trained students can be handed back to MATLAB.

The released weights are remote downloads (emoVoxZoo.m:95-97) that are not available here: the reader is exercised on
files written by `save_dagnn` (tests/test_host.py), not on the originals."""
from __future__ import annotations

import numpy as np

from .programs import STUDENT_CONVS, TEACHER_STAGES


def _student_layers(params):
    layers = []
    prev = "data"
    for name, fh, fw, cin, cout, stride, pad, has_bn in STUDENT_CONVS:
        out = "x_" + name
        layers.append(dict(name=name, type="dagnn.Conv", inputs=[prev], outputs=[out], params=[name + "f", name + "b"],
                           block=dict(size=list(params[name + "f"].shape), stride=list(stride), pad=list(pad), hasBias=True)))
        prev = out
        if has_bn:
            bn = "bn" + name[-1]
            layers.append(dict(name=bn, type="dagnn.BatchNorm", inputs=[prev], outputs=["x_" + bn], params=[bn + "m", bn + "b", bn + "x"],
                               block=dict(numChannels=cout, epsilon=1e-5)))
            layers.append(dict(name="relu" + name[-1], type="dagnn.ReLU", inputs=["x_" + bn], outputs=["x_relu" + name[-1]], params=[],
                               block=dict(leak=0)))
            prev = "x_relu" + name[-1]
    return layers


def save_dagnn(path, params, kind="student", meta=None):
    """Write a parameter dictionary as a DagNN struct (`net = dagnn.DagNN.loadobj(load(path))` in MATLAB)."""
    from scipy.io import savemat

    keys = [k for k in params if isinstance(params[k], np.ndarray)]
    p = np.zeros(len(keys), dtype=[("name", object), ("value", object), ("learningRate", object), ("weightDecay", object)])
    for i, k in enumerate(keys):
        v = params[k]
        p[i] = (k, v.reshape(-1, 1) if v.ndim == 1 else v, 0.1 if k.endswith("x") else 1.0, 0.0 if k.endswith("x") else 1.0)
    out = {"params": p, "meta": meta or {"arch": params.get("arch", kind)}}
    if kind == "student":
        layers = _student_layers(params)
        L = np.zeros(len(layers), dtype=[("name", object), ("type", object), ("inputs", object), ("outputs", object), ("params", object),
                                         ("block", object)])
        for i, l in enumerate(layers):
            L[i] = (l["name"], l["type"], np.array(l["inputs"], dtype=object), np.array(l["outputs"], dtype=object),
                    np.array(l["params"], dtype=object), l["block"])
        out["layers"] = L
    savemat(path, out, do_compression=True)


def _params_of(mat):
    ps = np.atleast_1d(mat["params"]).reshape(-1)
    out = {}
    for p in ps:
        name = str(np.asarray(p["name"]).reshape(-1)[0]) if not isinstance(p["name"], str) else p["name"]
        out[name] = np.asarray(p["value"], dtype=np.float32)
    return out


def load_dagnn(path, kind="student", arch=None):
    """Read a DagNN struct and return a zoo parameter dictionary, validating every tensor against the architecture."""
    from scipy.io import loadmat

    mat = loadmat(path, squeeze_me=True, struct_as_record=True)
    if "net" in mat and "params" not in mat:       # some releases wrap the struct in a `net` variable
        mat = {k: mat["net"][k].item() for k in mat["net"].dtype.names}
    raw = _params_of(mat)
    convs = [(k, v) for k, v in raw.items() if v.ndim >= 2 and not _is_moments(k, v, raw)]
    if kind == "student":
        return _map_student(raw)
    return _map_teacher(raw, arch)


def _is_moments(name, v, raw):
    return v.ndim == 2 and v.shape[1] == 2


def _vec(v):
    return np.asarray(v, np.float32).reshape(-1)


def _filt(v, shape):
    v = np.asarray(v, np.float32)
    v = v.reshape(v.shape + (1,) * (4 - v.ndim))   # MATLAB drops trailing singleton dimensions
    if v.shape != tuple(shape):
        if v.size == int(np.prod(shape)):
            v = v.reshape(shape)
        else:
            raise ValueError("filter of shape %s does not fit the architecture's %s" % (v.shape, tuple(shape)))
    return v


def _map_student(raw):
    """Parameters in file order: per conv (filter, bias) followed by its BatchNorm (mult, bias, moments)."""
    vals = list(raw.values())
    out, i = {}, 0
    for name, fh, fw, cin, cout, _, _, has_bn in STUDENT_CONVS:
        f = vals[i]
        k = np.asarray(f).shape[-1] if np.asarray(f).ndim == 4 else cout
        if name == "fc8":
            cout = int(np.asarray(f).size // (fh * fw * cin))
        out[name + "f"] = _filt(f, (fh, fw, cin, cout))
        out[name + "b"] = _vec(vals[i + 1])
        if out[name + "b"].size != cout:
            raise ValueError("%s: bias has %d elements, expected %d" % (name, out[name + "b"].size, cout))
        i += 2
        if has_bn:
            bn = "bn" + name[-1]
            out[bn + "m"], out[bn + "b"] = _vec(vals[i]), _vec(vals[i + 1])
            mom = np.asarray(vals[i + 2], np.float32).reshape(cout, 2)
            out[bn + "x"] = mom
            if out[bn + "m"].size != cout:
                raise ValueError("%s: expected %d channels" % (bn, cout))
            i += 3
    if i != len(vals):
        raise ValueError("model file holds %d parameters, the VGGVox student has %d" % (len(vals), i))
    return out


def _map_teacher(raw, arch):
    """Files written by save_dagnn keep the zoo's key names; upstream imports (Caffe-derived names) are matched by
    walking filters in order and assigning each to the next architecture slot of that shape."""
    if arch is None:
        arch = "senet50" if any("se1" in k for k in raw) else "resnet50"
    if "conv1f" in raw and "classifierf" in raw:
        out = {"arch": arch}
        for k, v in raw.items():
            if k.endswith("f"):
                out[k] = np.asarray(v, np.float32).reshape(v.shape + (1,) * (4 - v.ndim)) if v.ndim < 4 else np.asarray(v, np.float32)
            elif k.endswith("x"):
                out[k] = np.asarray(v, np.float32).reshape(-1, 2)
            else:
                out[k] = _vec(v)
        _check_teacher(out)
        return out
    raise NotImplementedError("teacher files with upstream (Caffe-derived) parameter names: map them with a name table "
                              "once a released file is at hand; the reference pins neither names nor order")


def _check_teacher(p):
    cin = 64
    assert p["conv1f"].shape[:3] == (7, 7, 3)
    for si, (blocks, mid, cout, _) in enumerate(TEACHER_STAGES):
        for bi in range(blocks):
            pre = "s%db%d_" % (si + 2, bi + 1)
            for key, shape in ((pre + "c1f", (1, 1, cin, mid)), (pre + "c2f", (3, 3, mid, mid)), (pre + "c3f", (1, 1, mid, cout))):
                p[key] = _filt(p[key], shape)
            if bi == 0:
                p[pre + "projf"] = _filt(p[pre + "projf"], (1, 1, cin, cout))
            if p["arch"] == "senet50":
                p[pre + "se1f"] = _filt(p[pre + "se1f"], (1, 1, cout, cout // 16))
                p[pre + "se2f"] = _filt(p[pre + "se2f"], (1, 1, cout // 16, cout))
            cin = cout
    k = p["classifierf"].size // 2048
    p["classifierf"] = _filt(p["classifierf"], (1, 1, 2048, k))
