"""CPU-side checks of the product package: the C-ABI library loads and exports every symbol include/xemo.h
declares, the host-side layout transforms round-trip, the product zoo generates the same synthetic graphs as
the oracle, and the data-parallel plumbing sums gradients across two gloo ranks."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "xemo.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(xemo_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    from mcncrossmodalemotions_b200 import _lib

    lib = _lib.load_library()
    syms = header_symbols()
    assert len(syms) >= 58
    for s in syms:
        assert hasattr(lib, s), "libxemo.so does not export %s" % s
    assert sorted(_lib.SIGNATURES) == syms, "ctypes signatures and include/xemo.h disagree"
    assert lib.xemo_version() >= 100


def test_no_device_means_loud_failure_not_fallback():
    import torch

    from mcncrossmodalemotions_b200 import _lib

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(_lib.XemoError):
        _lib.Context(0)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mcncrossmodalemotions_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, fn)).read().replace("the oracle", ""), fn
    # the measurement / profiling helpers under tools/ are product-side too: scripts that need the checker live in tests/tools/
    for fn in os.listdir(os.path.join(ROOT, "tools")):
        if fn.endswith((".py", ".sh")):
            txt = open(os.path.join(ROOT, "tools", fn)).read()
            assert "import oracle" not in txt and "from oracle" not in txt, fn


def test_out_size_rule():
    import ctypes as C

    from mcncrossmodalemotions_b200 import _lib

    lib = _lib.load_library()
    oh, ow = C.c_int64(), C.c_int64()
    assert lib.xemo_out_size(512, 300, 7, 7, _lib.I4(1, 1, 1, 1), _lib.I2(2, 2), C.byref(oh), C.byref(ow)) == 0
    assert (oh.value, ow.value) == (254, 148)
    assert lib.xemo_out_size(4, 4, 5, 5, _lib.I4(0, 0, 0, 0), _lib.I2(1, 1), C.byref(oh), C.byref(ow)) != 0


def test_filter_layout_round_trips():
    from mcncrossmodalemotions_b200 import programs as P

    rng = np.random.default_rng(0)
    f = rng.standard_normal((7, 7, 1, 96)).astype(np.float32)
    g = P.student_conv1_to_s2d(f)
    assert g.shape == (96, 4, 1, 16) and np.array_equal(P.student_conv1_from_s2d(g), f)
    assert np.all(g[:, :, 0, 7] == 0) and np.all(g[:, :, 0, 15] == 0) and np.all(g[:, 3, 0, 8:] == 0)
    # the space-to-depth formulation computes the same convolution
    x = rng.standard_normal((20, 18)).astype(np.float32)
    oh, ow = (20 + 2 - 7) // 2 + 1, (18 + 2 - 7) // 2 + 1
    xp = np.pad(x, 1)
    direct = np.array([[(xp[2 * i:2 * i + 7, 2 * j:2 * j + 7] * f[:, :, 0, 5]).sum() for j in range(ow)] for i in range(oh)])
    s2d = np.zeros((oh + 3, ow, 16), np.float32)
    for hp in range(oh + 3):
        for dr in range(2):
            for s in range(7):
                h, w = 2 * hp + dr - 1, 2 * np.arange(ow) + s - 1
                ok = (0 <= h < 20) & (w >= 0) & (w < 18)
                s2d[hp, ok, dr * 8 + s] = x[h, w[ok]] if 0 <= h < 20 else 0
    via = np.array([[(s2d[i:i + 4, j, :] * g[5, :, 0, :]).sum() for j in range(ow)] for i in range(oh)])
    assert np.allclose(direct, via, atol=1e-4)
    k = rng.standard_normal((3, 3, 20, 24)).astype(np.float32)
    assert P.krsc(k).shape == (32, 3, 3, 32) and np.array_equal(P.unkrsc(P.krsc(k), 3, 3, 20, 24), k)
    t = rng.standard_normal((7, 7, 3, 64)).astype(np.float32)
    r = P.teacher_conv1_to_rows(t)
    assert r.shape == (64, 7, 1, 32) and r[9, 2, 0, 5 * 4 + 1] == t[2, 5, 1, 9] and np.all(r[:, :, 0, 3::4] == 0)


def test_zoo_matches_oracle_graphs_and_buckets():
    from mcncrossmodalemotions_b200 import zoo
    from oracle import nets

    a, b = zoo.student_init(), nets.student_init()
    assert a.keys() == b.keys() and all(np.array_equal(a[k], b[k]) for k in a)
    a, b = zoo.teacher_init("resnet50-ferplus"), nets.teacher_init("resnet50")
    assert all(np.array_equal(a[k], b[k]) for k in a if k != "arch")
    assert zoo.POOL6_BUCKETS == nets.POOL6_TABLE and zoo.pool6_window(400) == (1, 11)
    with pytest.raises(ValueError):
        zoo.pool6_window(350)


def test_bucket_bounds_cover_buffer():
    from mcncrossmodalemotions_b200.dist import bucket_bounds, shard_batch

    segs = [("a", 0, 64), ("b", 64, 1000), ("c", 1088, 10), ("d", 1152, 5000), ("e", 6208, 3)]
    b = bucket_bounds(segs, 1000)
    assert b[0][0] == 0 and b[-1][1] == 6211 and all(x[1] <= y[0] for x, y in zip(b, b[1:]))
    assert shard_batch(list(range(10)), 1, 4) == [1, 5, 9]


WORKER = r"""
import os, sys, numpy as np, torch, torch.distributed as dist
sys.path.insert(0, %r)
from mcncrossmodalemotions_b200.dist import GradientAllReducer, bucket_bounds, shard_batch
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
segs = [("w%%d" %% i, i * 128, 100 + i) for i in range(9)]
flat = torch.zeros(9 * 128)
rng = np.random.default_rng(7)
per_sample = torch.from_numpy(rng.standard_normal((8, 9 * 128)).astype(np.float32))   # gradient of each sample of the batch
mask = torch.zeros(9 * 128)
for _, off, n in segs:
    mask[off:off + n] = 1                                                              # alignment padding between segments stays zero
per_sample *= mask
mine = shard_batch(list(range(8)), rank, world)
flat += per_sample[mine].sum(0)
for async_op in (False, True):
    g = flat.clone()
    GradientAllReducer(bucket_bounds(segs, 300), async_op=async_op)(g)
    assert torch.allclose(g, per_sample.sum(0), atol=1e-5), "all-reduced gradient != full-batch gradient"
# the update divides by the GLOBAL batch: identical on both ranks
w = torch.ones(9 * 128) - 0.1 * (g / 8)
ref = [torch.zeros_like(w) for _ in range(world)]
dist.all_gather(ref, w)
assert all(torch.equal(r, ref[0]) for r in ref)
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_two_rank_gradient_sum_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", str(script)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


def test_mex_shims_compile_against_stub_header():
    """The MEX shims are source-only (no MATLAB in the image): keep them syntactically valid against mex_stub.h."""
    import glob
    import shutil

    if not shutil.which("gcc"):
        pytest.skip("gcc not available")
    shims = sorted(glob.glob(os.path.join(ROOT, "mex", "*_mex.c")))
    assert len(shims) >= 3
    for f in shims:
        r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Wno-unused-function", "-DXEMO_MEX_STUB", "-I" + os.path.join(ROOT, "include"),
                            "-I" + os.path.join(ROOT, "mex"), f], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_batch_assembly_matches_reference_rules():
    from mcncrossmodalemotions_b200 import batch as B
    from oracle import nets

    # time2idx: floor(max(25 t - 1, 0) / 6) + 1 (getBatchEmoVoxCeleb.m:210-214)
    for t in (0.0, 0.03, 0.04, 0.27, 0.28, 1.0, 3.99, 4.0, 19.9):
        assert B.time2idx(t) == nets.time2idx(t)
    assert abs(B.audio_crop_seconds(400) - 4.024) < 1e-12           # 0.01*W + 0.001*Tw - 0.001
    assert B.frame_window(20, 0.0, 4.024) == (0, 17)
    assert B.frame_window(10, 1.0, 4.0) == (4, 10)                  # end clamped to the frames that exist
    with pytest.raises(ValueError):
        B.frame_window(3, 2.0, 3.0)
    rng = np.random.default_rng(0)
    lg = rng.standard_normal((17, 8)).astype(np.float32)
    assert np.array_equal(B.aggregate(lg, "max"), nets.aggregate_logits(lg, "max"))
    assert np.allclose(B.aggregate(lg, "mean", 6), nets.aggregate_logits(lg, "mean", 6))
    with pytest.raises(FloatingPointError):
        B.aggregate(np.full((2, 8), np.nan))
    spec = rng.standard_normal((512, 300)) * 3 + 1
    assert np.allclose(B.normalize_rows(spec), nets.normalize_spectrogram(spec), atol=1e-6)
    inputs = B.get_batch([spec, spec[:, ::-1]], [lg, lg[::-1]], [(0.0, 3.0), (0.5, 3.5)])
    assert inputs["data"].shape == (512, 300, 1, 2) and inputs["logitTarget"].shape == (1, 1, 8, 2)
    assert inputs["maxLabel"][0, 0, 0, 0] == lg[0:13].max(axis=0).argmax() + 1
    assert set(B.get_batch([spec], [lg], [(0, 1)], loss_type="softmaxlog")) == {"data", "maxLabel"}
    with pytest.raises(ValueError):
        B.get_batch([spec], [lg], [(0, 1)], loss_type="nope")
    assert B.width_bucket(1234) == 1000 and B.width_bucket(399) == 300 and B.centre_crop(spec, 100).shape == (512, 100)
    # compute_audio_feats.m:183-186: rstart = round(d / 2) (1-based, 0 -> 1): W = 450 -> columns 25..424 (1-based), i.e.
    # 0-based 24..423; d = 0, 1, 2, 3 -> 0-based starts 0, 0, 0, 1
    cols = np.arange(450, dtype=np.float32)[None, :].repeat(2, 0)
    assert B.centre_crop(cols, 400)[0, 0] == 24 and B.centre_crop(cols, 400)[0, -1] == 423
    for d, start in ((0, 0), (1, 0), (2, 0), (3, 1), (4, 1), (5, 2)):
        assert B.centre_crop(cols[:, : 400 + d], 400)[0, 0] == start, d
    with pytest.raises(ValueError):
        B.width_bucket(50)


def test_training_driver_helpers():
    from mcncrossmodalemotions_b200 import train as T

    lr = T.learning_rate_schedule(300)
    assert len(lr) == 300 and np.isclose(lr[0], 1e-4) and np.isclose(lr[-1], 1e-5)
    assert T.exp_dir_name("senet50-ferplus", "emovoxceleb-student", "hot-cross-ent", 4, 8, "max", 2) == \
        "voxceleb-senet50-ferplus-emovoxceleb-student-hot-cross-ent-scratch-4sec-8emo-agg-max-temp2"
    st = T.extract_stats(dict(objective=20.0, classerror=5.0, correct=np.array([3, 0, 1, 0, 0, 0, 0, 0.]),
                              count=np.array([4, 2, 2, 0, 0, 0, 0, 2.])), 10)
    assert st["objective"] == 2.0 and st["classerror"] == 0.5 and st["neutral"] == 0.75 and st["neutralPop"] == 0.4
    assert np.isclose(st["meanAcc"], (0.75 + 0.5) / 8)
    with pytest.raises(ValueError):
        T.run_distillation(None, None, notAnOption=1)


def test_dagnn_mat_files_round_trip(tmp_path):
    from scipy.io import loadmat

    from mcncrossmodalemotions_b200 import matfile, zoo
    from oracle import nets

    sp = nets.student_randomize_bn(zoo.student_init())
    path = str(tmp_path / "emovoxceleb-student.mat")
    matfile.save_dagnn(path, sp, "student")
    raw = loadmat(path, squeeze_me=True)
    assert {"layers", "params", "meta"} <= set(raw) and raw["layers"]["type"][0] == "dagnn.Conv" and len(raw["layers"]) == 8 + 2 * 7
    back = matfile.load_dagnn(path, "student")
    assert back.keys() == sp.keys() and all(np.array_equal(back[k], sp[k]) for k in sp)
    tp = zoo.teacher_init("senet50-ferplus")
    path = str(tmp_path / "senet50-ferplus.mat")
    matfile.save_dagnn(path, tp, "teacher")
    back = matfile.load_dagnn(path, "teacher")
    assert back["arch"] == "senet50" and all(np.array_equal(back[k], tp[k]) for k in tp if k != "arch")
    # a file with the wrong shapes is rejected instead of being silently mis-mapped
    bad = dict(sp); bad["conv3f"] = bad["conv3f"][:, :, :, :100]
    matfile.save_dagnn(str(tmp_path / "bad.mat"), bad, "student")
    with pytest.raises(ValueError):
        matfile.load_dagnn(str(tmp_path / "bad.mat"), "student")


def _caffe_style_teacher_file(path, tp, order, wrap_in_net=False):
    """A DagNN file laid out the way the MatConvNet imports of the Caffe ResNet-50 / SE-ResNet-50 graphs are: Caffe-derived
    parameter names, BatchNorm (mult, bias, moments) after every convolution, projection branch first ('resnet') or after
    the SE layers ('senet')."""
    from scipy.io import savemat

    from mcncrossmodalemotions_b200.programs import TEACHER_STAGES

    entries = []

    def conv_bn(tag, fkey, bnkey):
        entries.append((tag + "_filter", tp[fkey + "f"]))
        entries.append((tag + "_bn_mult", tp[bnkey + "m"].reshape(-1, 1)))
        entries.append((tag + "_bn_bias", tp[bnkey + "b"].reshape(-1, 1)))
        entries.append((tag + "_bn_moments", tp[bnkey + "x"]))

    conv_bn("conv1_7x7_s2", "conv1", "bn1")
    for si, (nb, mid, cout, _) in enumerate(TEACHER_STAGES):
        for bi in range(nb):
            pre, tag = "s%db%d_" % (si + 2, bi + 1), "conv%d_%d" % (si + 2, bi + 1)
            proj = lambda: conv_bn(tag + "_1x1_proj", pre + "proj", pre + "bnp")
            if bi == 0 and order == "resnet":
                proj()
            conv_bn(tag + "_1x1_reduce", pre + "c1", pre + "bn1")
            conv_bn(tag + "_3x3", pre + "c2", pre + "bn2")
            conv_bn(tag + "_1x1_increase", pre + "c3", pre + "bn3")
            if tp["arch"] == "senet50":
                for fc, key in (("_1x1_down", "se1"), ("_1x1_up", "se2")):
                    entries.append((tag + fc + "_filter", tp[pre + key + "f"]))
                    entries.append((tag + fc + "_bias", tp[pre + key + "b"].reshape(-1, 1)))
            if bi == 0 and order == "senet":
                proj()
    entries.append(("classifier_filter", tp["classifierf"]))
    entries.append(("classifier_bias", tp["classifierb"].reshape(-1, 1)))
    p = np.zeros(len(entries), dtype=[("name", object), ("value", object)])
    for i, (k, v) in enumerate(entries):
        p[i] = (k, v)
    net = {"params": p, "meta": {"normalization": {"imageSize": np.array([224.0, 224, 3]), "averageImage": np.array([131.0912, 103.8827, 91.4953])}}}
    savemat(path, {"net": net} if wrap_in_net else net, do_compression=True)


@pytest.mark.parametrize("arch,order,wrap", [("resnet50", "resnet", False), ("senet50", "senet", True), ("senet50", "resnet", False)])
def test_teacher_mat_import_by_shape_walk(tmp_path, arch, order, wrap):
    """emoVoxZoo.m:28-31,40-48: the released teachers are DagNN files with upstream parameter names the reference pins
    nowhere; load_dagnn maps them by shape and position, for both Caffe layer orders, with or without a `net` wrapper."""
    from mcncrossmodalemotions_b200 import matfile, zoo

    tp = zoo.teacher_init(arch + "-ferplus")
    path = str(tmp_path / (arch + "-ferplus.mat"))
    _caffe_style_teacher_file(path, tp, order, wrap)
    back = matfile.load_dagnn(path, "teacher")
    assert back["arch"] == arch
    assert set(back) == set(tp), (set(back) ^ set(tp))
    for k in tp:
        if k != "arch":
            assert back[k].shape == tp[k].shape and np.array_equal(back[k], tp[k]), k


def test_teacher_mat_import_rejects_a_foreign_graph(tmp_path):
    from scipy.io import savemat

    from mcncrossmodalemotions_b200 import matfile

    p = np.zeros(2, dtype=[("name", object), ("value", object)])
    p[0] = ("conv1_filter", np.zeros((3, 3, 3, 64), np.float32))
    p[1] = ("conv1_bias", np.zeros((64, 1), np.float32))
    savemat(str(tmp_path / "vgg.mat"), {"params": p})
    with pytest.raises(ValueError):
        matfile.load_dagnn(str(tmp_path / "vgg.mat"), "teacher")


def test_cached_logits_formats_round_trip(tmp_path):
    """imdb.wavLogits (fetch_emovoxceleb_imdb.m:138-148) / faceLogits (compute_visual_feats.m:105-117): 1 x T cells of F_i x 8
    single arrays among the top-level variables that save(path, '-struct', 'imdb') writes; consumed by
    getBatchEmoVoxCeleb.m:13 and fed to the coupling operator."""
    from scipy.io import loadmat

    from mcncrossmodalemotions_b200 import batch, matfile

    rng = np.random.default_rng(3)
    logits = [rng.standard_normal((f, 8)).astype(np.float32) for f in (13, 1, 40, 17)]
    path = str(tmp_path / "imdb.mat")
    matfile.save_logits(path, logits, "wavLogits", extra={"images": {"id": np.arange(1, 5)}})
    raw = loadmat(path)
    assert raw["wavLogits"].shape == (1, 4) and raw["wavLogits"].dtype == object and raw["wavLogits"][0, 2].shape == (40, 8)
    back = matfile.load_logits(path)
    assert len(back) == 4 and all(np.array_equal(a, b) and a.dtype == np.float32 for a, b in zip(back, logits))
    a, b = batch.frame_window(len(back[2]), 0.5, 3.5)
    assert np.array_equal(batch.aggregate(back[2][a:b]), logits[2][a:b].max(axis=0))
    matfile.save_logits(str(tmp_path / "feats.mat"), logits[:2], "faceLogits")
    assert [x.shape for x in matfile.load_logits(str(tmp_path / "feats.mat"), "faceLogits")] == [(13, 8), (1, 8)]
    with pytest.raises(KeyError):
        matfile.load_logits(path, "faceLogits")
    with pytest.raises(ValueError):
        matfile.save_logits(path, logits, "logits")


def test_zoo_model_mirrors_the_dagnn_calls_of_the_reference_scripts():
    from mcncrossmodalemotions_b200 import zoo

    dag = zoo.emoVoxZoo("emovoxceleb-student", scratch=True, lossType="hot-cross-ent", numSeconds=4)
    assert dag.getInputs() == ["data", "logitTarget", "maxLabel"]          # emoVoxZoo.m:151-169 wiring
    assert dag.layers[dag.getLayerIndex("loss")].block.temperature == 2 and dag.layers[dag.getLayerIndex("loss")].block.logitTargets
    assert dag.layers[dag.getLayerIndex("pool6")].block.poolSize == (1, 11)
    # compute_audio_feats.m:101-110: strip the losses, test mode, single input, `prediction` is the last variable
    for name in [l.name for l in dag.layers if l.block.isa("dagnn.Loss")]:
        dag.removeLayer(name)
    dag.mode = "test"
    assert dag.getInputs() == ["data"] and dag.vars[-1].name == "prediction"
    with pytest.raises(ValueError):
        dag.removeLayer("conv3")
    with pytest.raises(KeyError):
        dag.getLayerIndex("pool7")
    dag.layers[dag.getLayerIndex("pool6")].block.poolSize = (1, 8)          # compute_audio_feats.m:125
    assert dag.pool6 == (1, 8)
    dag.renameVar("data", "input")
    assert dag.getInputs() == ["input"]
    teacher = zoo.ferPlusZoo("senet50-ferplus")
    assert teacher.mode == "test" and teacher.getInputs() == ["data"] and teacher.vars[-1].name == "prediction"
    assert teacher.meta["normalization"]["imageSize"] == (224, 224, 3) and len(teacher.meta["classes"]["name"]) == 8
    assert sum(l.block.type == "bottleneck" for l in teacher.layers) == 16


def test_bench_clock_sampler_windows_rows_by_time(tmp_path, monkeypatch):
    """bench.py's nvidia-smi sampler: rows are time-stamped, only those inside the marked window count, and the caller
    can keep the load running until two samples have landed (a fake nvidia-smi stands in for the driver tool)."""
    import time

    import bench

    fake = tmp_path / "nvidia-smi"
    fake.write_text("#!/bin/bash\nsleep 0.2\nwhile true; do echo '1900, 1965, 700.0, Not Active, Not Active, Not Active, Active'; sleep 0.1; done\n")
    fake.chmod(0o755)
    monkeypatch.setenv("PATH", str(tmp_path) + os.pathsep + os.environ["PATH"])
    c = bench.ClockSampler(0)
    c.start()
    time.sleep(0.05)
    t0 = c.mark()
    t1 = c.mark()
    assert c.count_between(t0, t1) == 0            # nothing can have landed in an empty window
    for _ in range(100):
        if c.count_between(t0, t1) >= 2:
            break
        time.sleep(0.05)
        t1 = c.mark()
    out = c.stop(t0, t1)
    assert out["samples"] >= 2 and out["sm_mhz"] == 1900.0 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"]


def _student_teacher_conv_shapes(n):
    """(name, N, H, W, Cin, Kout, R, S, sh, sw, pt, pb, pl, pr) of every convolution the fused programs launch."""
    shapes = [("t.conv1", n, 224, 56, 64, 128, 7, 1, 2, 1, 3, 3, 0, 0)]
    cin, hw = 64, 56
    for si, (blocks, mid, cout, stride) in enumerate([(3, 64, 256, 1), (4, 128, 512, 2), (6, 256, 1024, 2), (3, 512, 2048, 2)]):
        for bi in range(blocks):
            st = stride if bi == 0 else 1
            ohw = hw // st
            pre = "t.s%db%d." % (si + 2, bi + 1)
            shapes.append((pre + "c1", n, hw, hw, cin, mid, 1, 1, st, st, 0, 0, 0, 0))
            shapes.append((pre + "c2", n, ohw, ohw, mid, mid, 3, 3, 1, 1, 1, 1, 1, 1))
            if bi == 0:
                shapes.append((pre + "proj", n, hw, hw, cin, cout, 1, 1, st, st, 0, 0, 0, 0))
            shapes.append((pre + "c3", n, ohw, ohw, mid, cout, 1, 1, 1, 1, 0, 0, 0, 0))
            cin, hw = cout, ohw
    shapes += [("s.conv1", n, 257, 74, 32, 192, 4, 1, 1, 1, 0, 0, 0, 0), ("s.conv1u", n, 257, 148, 16, 96, 4, 1, 1, 1, 0, 0, 0, 0),
               ("s.conv2", n, 126, 73, 128, 256, 5, 5, 2, 2, 1, 1, 1, 1), ("s.conv2u", n, 126, 73, 96, 256, 5, 5, 2, 2, 1, 1, 1, 1),
               ("s.conv3", n, 30, 17, 256, 384, 3, 3, 1, 1, 1, 1, 1, 1), ("s.conv4", n, 30, 17, 384, 256, 3, 3, 1, 1, 1, 1, 1, 1),
               ("s.conv5", n, 30, 17, 256, 256, 3, 3, 1, 1, 1, 1, 1, 1), ("s.fc6", n, 9, 8, 256, 4096, 9, 1, 1, 1, 0, 0, 0, 0),
               ("s.fc7", n, 1, 1, 4096, 1024, 1, 1, 1, 1, 0, 0, 0, 0), ("s.fc8", n, 1, 1, 1024, 16, 1, 1, 1, 1, 0, 0, 0, 0)]
    return shapes


@pytest.mark.parametrize("n", [1, 8, 64, 256, 1024])
def test_conv_plans_respect_the_hardware_budgets(n):
    """Host-side planning of the tcgen05 kernels (no device needed): every layer of the three networks gets a tile
    shape that fits 227 KB of shared memory and 512 TMEM columns, at every batch size the sweeps use."""
    import ctypes as C

    from mcncrossmodalemotions_b200 import _lib

    lib = _lib.load_library()
    out = (C.c_int * 12)()
    for name, *g in _student_teacher_conv_shapes(n):
        N, H, W, Cin, Kout, R, S, sh, sw, pt, pb, pl, pr = g
        assert lib.xemo_debug_conv_plan(*g, 148, out) == 0, name
        bk, bn, m_tiles, n_tiles, stages, epib, bres, tma_store, smem, grid, cw, k_iters = list(out)
        assert bk in (16, 32, 64) and Cin % bk == 0, name
        assert 16 <= bn <= 256 and Kout % bn == 0 and n_tiles == Kout // bn, name
        assert stages >= 2 and epib in (1, 2) and smem <= 227 * 1024, (name, smem)
        assert 1 <= grid <= 148 and cw in (16, 32, 64) and bn % cw == 0, name
        assert not bres or (n_tiles == 1 and k_iters * bn * bk * 2 <= 112 * 1024), name
        assert k_iters == R * S * (Cin // bk), name
        if name.startswith("s."):   # the student's layers also run the filter-gradient kernel
            assert lib.xemo_debug_wgrad_plan(N, H, W, Cin, Kout, Kout, R, S, sh, sw, pt, pb, pl, pr, 148, out) == 0, name
            chunk_a, chunk_b, block_c, c_tiles, T, mt, pix, groups, splits, wstages, wsmem, wgrid = list(out)
            assert block_c * c_tiles == Cin and block_c % chunk_b == 0 and chunk_b in (16, 32, 64), name
            assert mt in (1, 2) and mt * T * block_c <= 512, (name, mt, T, block_c)       # TMEM columns
            assert pix in (32, 64, 128) and wstages >= 2 and wsmem <= 227 * 1024, (name, wsmem)
            assert groups * T >= R * S * c_tiles and splits >= 1 and 1 <= wgrid <= 148, name


def test_layer_tables_agree_with_the_library():
    """csrc/xemo_net.cu carries its own architecture tables (the library does not depend on the Python package): every
    parameter name / MatConvNet shape it enumerates for the three networks must be the zoo's, in graph order."""
    import ctypes as C

    from mcncrossmodalemotions_b200 import _lib, zoo

    lib = _lib.load_library()
    for kind, params, size in ((0, zoo.teacher_init("resnet50"), 224), (1, zoo.teacher_init("senet50"), 224), (2, zoo.student_init(), 300)):
        h = C.c_void_p()
        assert lib.xemo_net_create(None, kind, 4, size, 0, 8, C.byref(h)) == 0
        names = [lib.xemo_net_param_name(h, i).decode() for i in range(lib.xemo_net_num_params(h))]
        assert set(names) == {k for k in params if k != "arch"}, set(names) ^ set(params)
        for name in names:
            d = (C.c_int64 * 4)()
            assert lib.xemo_net_param_dims(h, name.encode(), d) == 0
            shape = tuple(params[name].shape) + (1,) * (4 - params[name].ndim)
            assert tuple(d) == shape, (name, tuple(d), shape)
        assert lib.xemo_net_finalize(h) != 0          # description only: no device
        lib.xemo_net_destroy(h)
    h = C.c_void_p()
    assert lib.xemo_net_create(None, 2, 4, 20, 0, 8, C.byref(h)) != 0       # too narrow for the pooling chain
    assert lib.xemo_net_create(None, 7, 4, 300, 0, 8, C.byref(h)) != 0


def test_cta_pair_rule_of_the_convolution_planner():
    """CTA pairs (tcgen05.mma.cta_group::2) are chosen for long reductions with wide N tiles and enough tiles to fill the
    machine; never with a resident filter; grids come in whole clusters."""
    import ctypes as C

    from mcncrossmodalemotions_b200 import _lib

    lib = _lib.load_library()
    out = (C.c_int * 13)()

    def plan(*g):
        assert lib.xemo_debug_conv_plan2(*g, 148, out) == 0, g
        return list(out)

    assert lib.xemo_debug_set_conv_pair_mode(1) == 0
    try:
        conv2 = plan(256, 126, 73, 128, 256, 5, 5, 2, 2, 1, 1, 1, 1)               # student conv2: 50 k-iterations, N tile 256
        assert conv2[12] == 2 and conv2[9] == 148 and conv2[6] == 0
        res4 = plan(256, 14, 14, 256, 256, 3, 3, 1, 1, 1, 1, 1, 1)                   # teacher 3x3 c256
        assert res4[12] == 2
        wide = plan(256, 7, 7, 512, 2048, 1, 1, 1, 1, 0, 0, 0, 0)                    # 8 k-iterations, output-bound: single CTAs
        assert wide[12] == 1
        stem = plan(256, 56, 56, 64, 64, 3, 3, 1, 1, 1, 1, 1, 1)                     # resident filter: single CTAs
        assert stem[6] == 1 and stem[12] == 1
        small = plan(16, 14, 14, 256, 256, 3, 3, 1, 1, 1, 1, 1, 1)                   # 25 M tiles x 4 N tiles < 148: single CTAs
        assert small[12] == 1 and small[1] == 64
        assert lib.xemo_debug_set_conv_pair_mode(2) == 0
        forced = plan(16, 14, 14, 256, 256, 3, 3, 1, 1, 1, 1, 1, 1)                  # 13 pairs of M tiles (the last one half empty) x 4
        assert forced[12] == 2 and forced[9] == 2 * 13 * 4 and forced[4] > small[4]  # (half the filter per CTA: deeper pipeline)
        assert lib.xemo_debug_set_conv_pair_mode(0) == 0
        assert plan(256, 126, 73, 128, 256, 5, 5, 2, 2, 1, 1, 1, 1)[12] == 1
    finally:
        lib.xemo_debug_set_conv_pair_mode(-1)


def test_loss_types_of_the_zoo():
    """emoVoxZoo.m:137-157: four loss types; 'euclidean' scales the head filters by 1/10 (:141-144); anything else is
    rejected before the device is touched."""
    from mcncrossmodalemotions_b200 import zoo
    from mcncrossmodalemotions_b200.programs import StudentProgram

    with pytest.raises(ValueError):
        StudentProgram({}, 4, 100, loss_type="nonsense")
    with pytest.raises(ValueError):
        zoo.emoVoxZoo("emovoxceleb-student", scratch=True, lossType="nonsense")
    base = zoo.emoVoxZoo("emovoxceleb-student", scratch=True, lossType="hot-cross-ent")
    for lt, block, inputs in (("euclidean", "dagnn.EuclideanLoss", ["prediction", "logitTarget", "instanceWeights"]),
                              ("huber", "dagnn.HuberLoss", ["prediction", "logitTarget", "instanceWeights"]),
                              ("softmaxlog", "dagnn.Loss", ["prediction", "maxLabel"])):
        dag = zoo.emoVoxZoo("emovoxceleb-student", scratch=True, lossType=lt)
        layer = dag.layers[dag.getLayerIndex("loss")]
        assert layer.block.type == block and layer.inputs == inputs and layer.outputs == ["objective"]
        scale = 0.1 if lt == "euclidean" else 1.0
        np.testing.assert_allclose(dag.params["fc8f"], base.params["fc8f"] * scale, rtol=1e-6)
        np.testing.assert_array_equal(dag.params["fc7f"], base.params["fc7f"])


def test_se_gate_launch_plan_follows_the_batch_but_not_the_summation_order():
    """csrc/se_gate.cuh: the cluster size K follows the batch (1 at 256 faces, up to 4 at small batches, never more CTAs
    than SMs, whole divisors of C and Cr); the ranges of hidden units that define the last phase's summation order depend
    on (C, Cr) only; shared memory stays under the 48 KB default for SENet50's four stages in both forms."""
    import ctypes as C

    from mcncrossmodalemotions_b200 import _lib

    lib = _lib.load_library()
    out = (C.c_int * 5)()
    stages = [(256, 64), (512, 128), (1024, 256), (2048, 512)]
    for c, cm in stages:
        cr = c // 16
        ranges = set()
        for n, want_k in [(256, 1), (128, 2), (100, 2), (64, 4), (32, 4), (3, 4), (1, 4)]:
            for lin in (0, 1):
                assert lib.xemo_debug_se_gate_plan(n, c, cm, cr, lin, 148, out) == 0
                k, pc, tg, smem, grid = list(out)
                assert k == want_k, (c, n, k)
                assert grid == (n + 1) // 2 * k and grid <= 148
                assert c % k == 0 and cr % k == 0 and c // k >= 32 and cr // k >= 2
                assert smem <= 48 * 1024
                assert pc >= 1 and cr % pc == 0 and cr // pc >= 8 and tg <= pc and (tg == 1 or tg * (c // k) <= 1024)
                ranges.add(pc)
        assert len(ranges) == 1, (c, ranges)            # the summation order does not know the batch
    # 148 SMs: 75 groups (150 faces) no longer fit twice
    assert lib.xemo_debug_se_gate_plan(150, 1024, 256, 64, 0, 148, out) == 0 and out[0] == 1
    assert lib.xemo_debug_se_gate_plan(148, 1024, 256, 64, 0, 148, out) == 0 and out[0] == 2


def test_fixed_channel_grids_are_whole_channel_multiples_and_never_exceed_the_resident_blocks():
    """A grid-stride kernel whose threads keep their channel group needs grid * threads to be a multiple of C / 8; when the
    grid is capped at the resident blocks it rounds DOWN (592 -> 591 on the 96-channel stem, not 594: two blocks in a
    second wave cost 20 % of the kernel), small problems round up."""
    import math

    from mcncrossmodalemotions_b200 import _lib

    lib = _lib.load_library()
    g = lib.xemo_debug_fixed_channel_grid
    assert g(10**9, 12, 256, 148, 4) == 591              # stem: 96 channels, 4 resident blocks per SM
    assert g(10**9, 12, 256, 148, 3) == 444
    assert g(10**9, 32, 256, 148, 3) == 444              # 256 channels: any grid works
    assert g(10**9, 12, 256, 148, 48) == 7104
    assert g(1000, 12, 256, 148, 8) == 6                 # 4 blocks of work, rounded up to a multiple of 3
    assert g(0, 12, 256, 148, 8) == 3
    for c8 in (1, 3, 8, 12, 48, 96, 512):
        for per_sm in (1, 2, 3, 5, 8):
            n = g(10**10, c8, 256, 148, per_sm)
            assert (n * 256) % c8 == 0
            assert n <= 148 * per_sm or n == c8 // math.gcd(c8, 256)       # (never below one channel multiple)


def test_dynamic_loss_scale_policy():
    """train.LossScaler / apply_loss_scale: halve on a non-finite gradient, grow back after a clean run, stay within
    [1, initial]; the network is only touched when the scale changes."""
    from mcncrossmodalemotions_b200.train import LossScaler, apply_loss_scale

    class FakeNet:
        def __init__(self):
            self.grad_scale, self.calls = 1024.0, []

        def set_grad_scale(self, s):
            self.grad_scale = s
            self.calls.append(s)

    net, sc, lines = FakeNet(), LossScaler(1024.0, growth_interval=3), []
    for _ in range(5):
        apply_loss_scale(net, sc, False, lines.append)
    assert net.calls == [] and sc.scale == 1024.0          # never above the initial scale
    apply_loss_scale(net, sc, True, lines.append)
    apply_loss_scale(net, sc, True, lines.append)
    assert net.calls == [512.0, 256.0] and sc.overflows == 2
    apply_loss_scale(net, sc, False, lines.append)
    apply_loss_scale(net, sc, False, lines.append)
    assert net.grad_scale == 256.0
    apply_loss_scale(net, sc, False, lines.append)          # third clean step in a row
    assert net.grad_scale == 512.0
    apply_loss_scale(net, sc, True, lines.append)           # an overflow restarts the clean run
    for _ in range(2):
        apply_loss_scale(net, sc, False, lines.append)
    assert net.grad_scale == 256.0
    for _ in range(20):
        apply_loss_scale(net, sc, True, lines.append)
    assert net.grad_scale == 1.0 and sc.scale == 1.0         # floor
    assert len(lines) == len(net.calls) and "non-finite" in lines[0]
