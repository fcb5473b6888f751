"""DagNN `.mat` model files: the on-disk format either side of the hot path.

The reference loads `<name>.mat` with `load` -> `dagnn.DagNN.loadobj` (emoVoxCeleb/emoVoxZoo.m:40-48,199): a struct
with `layers(i).{name,type,inputs,outputs,params,block}`, `params(i).{name,value}` and `meta`.  `load_dagnn` reads
such a file with scipy (MAT v5-v7; v7.3/HDF5 files are not readable without h5py) and maps the parameters onto the
zoo's dictionaries *structurally* -- by walking the Conv / BatchNorm layers in execution order and checking every
shape against the architecture -- because the upstream parameter names are not pinned anywhere in the reference.
`save_dagnn` writes the same structure, so that trained students can be handed back to MATLAB.

Teacher files (`resnet50-ferplus.mat`, `senet50-ferplus.mat`, emoVoxZoo.m:28-31) carry upstream, Caffe-derived parameter
names: `_walk_teacher` maps them by shape and position.  `save_logits` / `load_logits` are the cached-logits formats
(`wavLogits`, `faceLogits`).

The released weights are remote downloads (emoVoxZoo.m:95-97) that are not available here: the readers are exercised on
files written by this module and on synthetic files laid out the way the Caffe imports are (tests/test_host.py), not
on the originals."""
from __future__ import annotations

import numpy as np

from .arch import STUDENT_CONVS, TEACHER_STAGES


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
    if kind == "student":
        return _map_student(raw)
    if "conv1f" in raw and "classifierf" in raw:
        return _map_teacher(raw, arch)
    return _walk_teacher(_raw_params(path), arch)


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
    """Files written by save_dagnn keep the zoo's key names; anything else (the released models: MatConvNet imports of the
    Caffe ResNet-50 / SE-ResNet-50 graphs, whose parameter names the reference pins nowhere) is mapped by `_walk_teacher`."""
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
    raise ValueError("teacher file without the zoo's key names: read it with load_dagnn (which walks the raw parameter list)")


def _raw_params(path):
    """[(name, value)] in file order with MATLAB's dimensions intact (no singleton squeezing: a 1 x 1 x C x K filter stays
    four-dimensional, a K x 1 bias stays a column, BatchNorm moments stay C x 2)."""
    from scipy.io import loadmat

    mat = loadmat(path, squeeze_me=False, struct_as_record=True)
    if "net" in mat and "params" not in mat:
        net = mat["net"][0, 0]
        mat = {k: net[k] for k in net.dtype.names}
    out = []
    for p in np.asarray(mat["params"]).reshape(-1):
        name = str(np.asarray(p["name"]).reshape(-1)[0])
        out.append((name, np.asarray(p["value"], dtype=np.float32)))
    return out


def _walk_teacher(plist, arch=None):
    """Map an upstream-named ResNet-50 / SE-ResNet-50 DagNN parameter list onto the zoo's keys by walking it in file
    order: every filter is assigned to the open slot of its shape in the current bottleneck, and the vectors that follow
    it are its BatchNorm (mult, bias, moments C x 2) or -- for the SE fully-connected layers and the classifier -- its bias.
    Works for both layer orders seen in Caffe-derived graphs (projection branch first: ResNet-50; projection after the SE
    layers: SE-ResNet-50).  The only shape shared by two slots of one block, 1 x 1 x 64 x 256 (expand and projection of
    the first bottleneck), is resolved by position: before the block's reduce convolution it is the projection."""
    if arch is None:
        arch = "senet50" if any(v.ndim == 4 and v.shape[:2] == (1, 1) and v.shape[2] == 16 * v.shape[3] for _, v in plist) else "resnet50"
    se = arch == "senet50"
    blocks, cin = [], 64
    for si, (nb, mid, cout, _) in enumerate(TEACHER_STAGES):
        for bi in range(nb):
            slots = {"c1": (1, 1, cin, mid), "c2": (3, 3, mid, mid), "c3": (1, 1, mid, cout)}
            if bi == 0:
                slots["proj"] = (1, 1, cin, cout)
            if se:
                slots["se1"], slots["se2"] = (1, 1, cout, cout // 16), (1, 1, cout // 16, cout)
            blocks.append(("s%db%d_" % (si + 2, bi + 1), slots))
            cin = cout
    out = {"arch": arch}
    bn_of = {"c1": "bn1", "c2": "bn2", "c3": "bn3", "proj": "bnp"}
    bi, done, cur, nvec = -1, set(), None, 0      # current block, its assigned slots, current conv key, vectors seen after it
    for name, v in plist:
        if v.ndim >= 3 or (v.ndim == 2 and min(v.shape) > 2):
            f = v.reshape(v.shape + (1,) * (4 - v.ndim))
            shape = tuple(f.shape)
            if bi < 0:
                if shape[:3] != (7, 7, 3):
                    raise ValueError("%s: expected the 7 x 7 x 3 stem first, got %s" % (name, shape))
                out["conv1f"], cur, bi, nvec = f, ("conv1", "bn1"), 0, 0
                continue
            if shape[:3] == (1, 1, 2048) and bi == len(blocks) - 1 and set(blocks[bi][1]) <= done:
                out["classifierf"], cur, nvec = f, ("classifier", None), 0
                continue
            if set(blocks[bi][1]) <= done:
                bi, done = bi + 1, set()
                if bi >= len(blocks):
                    raise ValueError("%s: filter %s after the last bottleneck" % (name, shape))
            pre, slots = blocks[bi]
            cand = [k for k, sh in slots.items() if sh == shape and k not in done]
            if not cand:
                raise ValueError("%s: filter %s fits no open slot of block %s (%s)" % (name, shape, pre, sorted(set(slots) - done)))
            if len(cand) > 1:       # {c3, proj} of the first bottleneck
                key = "proj" if "c1" not in done else "c3"
            else:
                key = cand[0]
            done.add(key)
            out[pre + key + "f"] = f
            cur, nvec = (pre + key, pre + bn_of[key] if key in bn_of else None), 0
        elif v.ndim == 2 and v.shape[1] == 2 and v.shape[0] > 2:
            if cur is None or cur[1] is None:
                raise ValueError("%s: BatchNorm moments without a preceding normalised convolution" % name)
            out[cur[1] + "x"] = v.reshape(-1, 2)
        else:
            vec = v.reshape(-1)
            if cur is None:
                raise ValueError("%s: vector before the first filter" % name)
            if cur[1] is None:
                out[cur[0] + "b"] = vec                      # SE FC / classifier bias
            else:
                out[cur[1] + ("m" if nvec == 0 else "b")] = vec
            nvec += 1
    _check_teacher(out)
    missing = [k for k in ("classifierf", "classifierb", "bn1m", "bn1b", "bn1x") if k not in out]
    if missing:
        raise ValueError("teacher file is missing %s" % missing)
    return out


# ------------------------------------------------------------------------------------------------
# cached teacher logits (emoVoxCeleb/fetch_emovoxceleb_imdb.m:138-148: imdb.wavLogits, a 1 x numWavs cell of F_i x 8 single
# arrays saved with save(imdbPath, '-struct', 'imdb'); external/compute_visual_feats.m:105-117 and compute_audio_feats.m:145:
# faceLogits, one cell per track) -- what getBatchEmoVoxCeleb.m:13 reads as imdb.wavLogits(batch)
def save_logits(path, logits, field="wavLogits", extra=None):
    """Write a list of F_i x K arrays as a 1 x T MATLAB cell array named `field` (plus the other top-level imdb
    variables in `extra`), as save(path, '-struct', 'imdb') leaves it."""
    from scipy.io import savemat

    if field not in ("wavLogits", "faceLogits"):
        raise ValueError("field must be 'wavLogits' or 'faceLogits'")
    cell = np.empty((1, len(logits)), dtype=object)
    for i, lg in enumerate(logits):
        lg = np.asarray(lg, np.float32)
        cell[0, i] = lg.reshape(-1, lg.shape[-1]) if lg.ndim >= 2 else lg.reshape(1, -1)
    out = dict(extra or {})
    out[field] = cell
    savemat(path, out, do_compression=True)


def load_logits(path, field=None):
    """Read `wavLogits` / `faceLogits` back as a list of F_i x K float32 arrays (an empty cell gives a 0 x 0 array)."""
    from scipy.io import loadmat

    mat = loadmat(path, squeeze_me=False)
    if field is None:
        found = [f for f in ("wavLogits", "faceLogits") if f in mat]
        if len(found) != 1:
            raise ValueError("%s holds %s: name the field" % (path, found or "neither wavLogits nor faceLogits"))
        field = found[0]
    if field not in mat:
        raise KeyError("%s has no variable %s" % (path, field))
    cell = np.asarray(mat[field]).reshape(-1)
    return [np.asarray(c, np.float32) for c in cell]


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
