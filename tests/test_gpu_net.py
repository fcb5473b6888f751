"""The graph-level C ABI (include/xemo.h section C, csrc/xemo_net.cu) driven through ctypes only (net.py): whole networks
behind xemo_teacher_forward / xemo_student_forward / xemo_student_train_step / xemo_sgd_step / xemo_distill_step, against
the CPU oracle and against the Python-assembled programs that sequence the same kernels."""
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-3


@pytest.fixture(scope="module")
def nets():
    from oracle import nets

    return nets


def _f64(p):
    return {k: (v.astype(np.float64) if isinstance(v, np.ndarray) else v) for k, v in p.items()}


def test_net_module_does_not_need_the_python_program_assembly():
    code = ("import sys; sys.path.insert(0, %r); import numpy as np\n"
            "from mcncrossmodalemotions_b200 import net, zoo\n"
            "t = net.TeacherNet(zoo.teacher_init('resnet50'), 2)\n"
            "out = t.forward(np.zeros((224, 224, 3, 2), np.float32))\n"
            "assert out.shape == (2, 8) and np.isfinite(out).all()\n"
            "assert not [m for m in sys.modules if m == 'torch' or m.endswith('.programs') or m.endswith('.distill')], 'python-side assembly imported'\n"
            "print('ok')\n" % ROOT)
    # the forward above ran with neither torch nor the Python program classes in the process: ctypes + libxemo.so only
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("arch,mode", [("resnet50", "hwcn224"), ("senet50", "hwcn224"), ("senet50", "u8")])
def test_teacher_forward_through_the_c_abi(nets, arch, mode):
    from mcncrossmodalemotions_b200.net import TeacherNet
    from mcncrossmodalemotions_b200.programs import TeacherProgram

    n = 6
    p = nets.teacher_init(arch)
    if mode == "u8":
        faces = nets.synth_faces48(n)
        x = nets.faces48_to_input(faces)
    else:
        faces = x = nets.synth_faces(n)
    ref = nets.teacher_forward(_f64(p), x.astype(np.float64), nets.TorchOps).reshape(8, n).T
    net = TeacherNet(p, n, input_mode=mode)
    got = net.forward(faces)
    assert rel_err(got, ref) < TOL
    assert np.array_equal(got, net.forward(faces))                               # graph replay
    assert np.array_equal(got, TeacherProgram(p, n, input_mode=mode).forward(faces))   # the same kernels, sequenced in Python


@pytest.mark.parametrize("width,n", [(300, 8), (100, 5)])
def test_student_test_mode_forward_through_the_c_abi(nets, width, n):
    from mcncrossmodalemotions_b200.net import StudentNet
    from mcncrossmodalemotions_b200.programs import StudentProgram

    p = nets.student_randomize_bn(nets.student_init())
    spec = nets.synth_spectrograms(n, width)
    ref, _ = nets.student_forward(_f64(p), spec.astype(np.float64), "test", nets.TorchOps)
    net = StudentNet(p, n, width)
    got = net.forward(spec, "test")
    assert rel_err(got, ref.reshape(8, n).T) < TOL
    assert np.array_equal(got, net.forward(spec, "test"))
    assert np.array_equal(got, StudentProgram(p, n, width).forward(spec, "test"))


@pytest.mark.parametrize("loss_type", ["hot-cross-ent", "euclidean"])
def test_student_training_step_through_the_c_abi(nets, loss_type):
    """xemo_student_train_step + xemo_sgd_step == the Python-assembled program (same kernels; filter-gradient atomics make
    two runs differ in the last bits) and the oracle's continuous quantities."""
    from mcncrossmodalemotions_b200.net import StudentNet
    from mcncrossmodalemotions_b200.programs import StudentProgram

    n, width, lr = 8, 100, 1e-3
    p = nets.student_randomize_bn(nets.student_init())
    spec, tgt = nets.synth_spectrograms(n, width), nets.synth_teacher_logits(n)
    w = np.random.default_rng(3).uniform(0.5, 2, n).astype(np.float32) if loss_type == "euclidean" else None
    net = StudentNet(p, n, width, loss_type=loss_type)
    net.set_hyper(lr=lr, batch_size=n)
    net.reset_metrics()
    prog = StudentProgram(p, n, width, loss_type=loss_type)
    prog.set_hyper(lr=lr, batch_size=n)
    prog.reset_metrics()
    noise = lambda k: k.endswith("b") and not k.startswith("bn") and k != "fc8b"   # conv biases ahead of train-mode BN: zero gradient
    net.train_step(spec, tgt, weights=w)
    prog.train_step(spec, tgt, weights=w)
    m, mp = net.metrics(), prog.metrics()
    assert abs(m["objective"] - mp["objective"]) <= 1e-5 * abs(mp["objective"]) and m["classerror"] == mp["classerror"]
    assert np.array_equal(m["count"], mp["count"]) and np.array_equal(m["correct"], mp["correct"]) and m["count"].sum() == n
    g, gp = net.export_grads(), prog.export_grads()
    q, qp = net.export_params(), prog.export_params()
    assert set(g) == set(gp) == set(p)
    for k in g:
        assert g[k].shape == gp[k].shape == p[k].shape, k
        if noise(k):     # both runs hold cancellation noise there: negligible against the layer's filter gradient / filters
            assert np.abs(g[k] - gp[k]).max() <= 1e-3 * np.abs(gp[k[:-1] + "f"]).max(), k
            assert np.abs(q[k] - qp[k]).max() <= 1e-6 * np.abs(qp[k[:-1] + "f"]).max(), k
        else:
            assert rel_err(g[k], gp[k]) < 1e-4, (k, rel_err(g[k], gp[k]))
            assert rel_err(q[k], qp[k]) < 1e-5, k
    net.train_step(spec, tgt, weights=w)      # the second step replays the captured graphs on the updated weights
    m2 = net.metrics()
    assert m2["count"].sum() == 2 * n and np.isfinite(m2["objective"]) and m2["objective"] != m["objective"] and m2["skipped_steps"] == 0
    exact = nets.distillation_student_step(_f64(p), {}, spec.astype(np.float64), tgt.astype(np.float64), lr=lr, ops=nets.TorchOps,
                                           loss_type=loss_type, instance_weights=w, update=False)
    first = StudentNet(p, n, width, loss_type=loss_type)
    first.set_input(spec); first.set_target(tgt, w); first.grad_step()
    assert abs(first.metrics()["objective"] - exact["objective"]) <= TOL * abs(exact["objective"])
    for bn, tol in (("bn1x", TOL), ("bn4x", TOL), ("bn7x", 5e-3)):      # (bn7 normalises over 8 samples only: ill-conditioned)
        assert rel_err(first.export_grads()[bn], exact["grads"][bn]) < tol, bn
    # checkpoint round trip of the optimiser state
    mom = net.export_momentum()
    net.load_momentum(mom)
    back = net.export_momentum()
    assert all(np.array_equal(mom[k], back[k]) for k in mom)


def test_full_distillation_step_through_the_c_abi(nets):
    from mcncrossmodalemotions_b200.distill import DistillationStep
    from mcncrossmodalemotions_b200.net import DistillStep

    n, F, width = 4, 3, 100
    tp, sp = nets.teacher_init("senet50"), nets.student_init()
    faces, spec = nets.synth_faces48(n * F), nets.synth_spectrograms(n, width)
    step = DistillStep(tp, sp, n, width, frames_per_clip=F)
    step.student.set_hyper(lr=1e-4, batch_size=n)
    step.teacher.set_input(faces)
    step.student.set_input(spec)
    step.student.reset_metrics()
    step.step()
    m = step.student.metrics()
    ref = DistillationStep(tp, sp, n, width, frames_per_clip=F)
    ref.teacher.set_input(faces)
    ref.student.set_input(spec)
    ref.student.set_hyper(lr=1e-4, batch_size=n)
    ref.student.reset_metrics()
    ref.step_resident()
    mr = ref.student.metrics()
    assert abs(m["objective"] - mr["objective"]) <= 1e-5 * abs(mr["objective"]) and m["classerror"] == mr["classerror"]
    assert np.array_equal(m["count"], mr["count"]) and m["count"].sum() == n
    q, qr = step.student.export_params(), ref.student.export_params()
    for k in q:
        if k.endswith("b") and not k.startswith("bn") and k != "fc8b":      # zero-gradient biases: cancellation noise
            assert np.abs(q[k] - qr[k]).max() <= 1e-6 * np.abs(qr[k[:-1] + "f"]).max(), k
        else:
            assert rel_err(q[k], qr[k]) < 1e-5, k
    logits = nets.teacher_forward(tp, nets.faces48_to_input(faces), nets.TorchOps).reshape(8, n * F).T
    target = np.stack([nets.aggregate_logits(logits[i * F:(i + 1) * F]) for i in range(n)])
    out = nets.distillation_student_step(nets.student_init(), {}, spec, target.T.reshape(1, 1, 8, n).astype(np.float32), ops=nets.TorchOps)
    assert abs(m["objective"] - out["objective"]) <= TOL * abs(out["objective"])
    step.step()      # replay
    assert step.student.metrics()["count"].sum() == 2 * n and step.num_kernels() > 150


def test_c_abi_argument_errors(nets):
    from mcncrossmodalemotions_b200 import _lib
    from mcncrossmodalemotions_b200.net import StudentNet, TeacherNet

    p = nets.student_init()
    with pytest.raises(_lib.XemoError):
        StudentNet(p, 2, 20)                        # too narrow for the pooling chain
    bad = dict(p); bad["conv3f"] = bad["conv3f"][:, :, :, :100]
    with pytest.raises(ValueError):
        StudentNet(bad, 2, 100)
    with pytest.raises(KeyError):
        TeacherNet({"arch": "resnet50", "classifierf": np.zeros((1, 1, 2048, 8), np.float32)}, 2)
    net = StudentNet(p, 2, 100)
    with pytest.raises(_lib.XemoError):
        net._check(net.lib.xemo_net_set_input(net.handle, net.buffer("spec"), 12))     # wrong byte count


def test_data_parallel_step_on_two_gpus():
    """xemo_comm_* + the exchange inside the captured step: sum of local gradients, bit-identical parameters on all ranks
    (tests/tools/dp_check.py under torch.distributed.run; needs two visible GPUs)."""
    import os

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", os.path.join(ROOT, "tests", "tools", "dp_check.py")], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "dp_check ok" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]


def test_deterministic_option_gives_bit_identical_training_steps(nets):
    """xemo_set_deterministic: two independent runs of two training steps end with bit-identical gradients and parameters
    (the default mode's filter gradients differ in their last bits: several split-K contributors per element)."""
    from mcncrossmodalemotions_b200 import _lib
    from mcncrossmodalemotions_b200.net import StudentNet

    n, width = 8, 100
    p = nets.student_randomize_bn(nets.student_init())
    spec, tgt = nets.synth_spectrograms(n, width), nets.synth_teacher_logits(n)
    runs = []
    for _ in range(2):
        ctx = _lib.Context(0)
        ctx.set_deterministic(1)
        assert ctx.lib.xemo_get_deterministic(ctx.handle) == 1
        net = StudentNet(p, n, width, ctx=ctx)
        net.set_hyper(lr=1e-3, batch_size=n)
        for _ in range(2):
            net.train_step(spec, tgt)
        runs.append((net.export_grads(), net.export_params(), net.metrics()["objective"]))
        net.close()
    (g0, q0, o0), (g1, q1, o1) = runs
    assert o0 == o1
    for k in g0:
        assert np.array_equal(g0[k], g1[k]), k
        assert np.array_equal(q0[k], q1[k]), k
