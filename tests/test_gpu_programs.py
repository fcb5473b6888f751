"""Parity of the fused device programs (teacher forward, student forward / training step) against the
CPU oracle, through libxemo.so.

Tolerances.  north_star: logits within 1e-3 relative (max|a-b| <= 1e-3 * max|ref|) of the fp32/fp64
CPU path; pooling indices / class-error counts exact.  Convolution operands are fp16 (fp32
accumulation), so a single contraction is good to ~3e-4 and a 50-layer teacher to ~7e-4 (measured;
tests/tools/precision_emulation.py reproduces it on the CPU).  Train-mode quantities that are *discontinuous*
in the activations (ReLU masks under batch-statistics BN) cannot be held to 1e-3 by any 16-bit-operand
pipeline: they are checked (a) tightly against the oracle's fp16 number-format model
(oracle.nets.Fp16ModelOps) and (b) against the exact oracle through continuous quantities (objective,
batch moments, updated parameters) at 1e-3 and through the gradient direction (cosine)."""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-3


@pytest.fixture(scope="module")
def nets():
    from oracle import nets

    return nets


def _f64(p):
    return {k: (v.astype(np.float64) if isinstance(v, np.ndarray) else v) for k, v in p.items()}


@pytest.mark.parametrize("arch", ["resnet50", "senet50"])
def test_teacher_logits(nets, arch):
    from mcncrossmodalemotions_b200.programs import TeacherProgram

    n = 8
    p = nets.teacher_init(arch)
    x = nets.synth_faces(n)
    ref = nets.teacher_forward(_f64(p), x.astype(np.float64), nets.TorchOps).reshape(8, n).T
    eager = TeacherProgram(p, n, use_graph=False).forward(x)
    prog = TeacherProgram(p, n, use_graph=True)
    g1 = prog.forward(x)
    g2 = prog.forward(x)  # CUDA-graph replay
    assert rel_err(eager, ref) < TOL
    assert np.array_equal(eager, g1) and np.array_equal(g1, g2), "graph replay must be bit-identical to eager launches"
    # a different batch through the same captured graph
    x2 = nets.synth_faces(n, seed=10)
    ref2 = nets.teacher_forward(_f64(p), x2.astype(np.float64), nets.TorchOps).reshape(8, n).T
    assert rel_err(prog.forward(x2), ref2) < TOL


def test_teacher_48x48_plumbing_config(nets):
    """BASELINE config 1: 48 x 48 grey faces -> 224 x 224 x 3 (a2) -> SENet50 logits."""
    from mcncrossmodalemotions_b200.programs import TeacherProgram

    n = 4
    p = nets.teacher_init("senet50")
    x = nets.faces48_to_input(nets.synth_faces48(n))
    ref = nets.teacher_forward(_f64(p), x.astype(np.float64), nets.TorchOps).reshape(8, n).T
    assert rel_err(TeacherProgram(p, n).forward(x), ref) < TOL


@pytest.mark.parametrize("n", [8, 5])
def test_se_blocks_by_linearity_on_every_stage_selection(nets, n, monkeypatch):
    """SE blocks by linearity: squeeze(t2) -> gate (W3 folded in) -> expand convolution with the excite in its epilogue
    (conv_fprop_kernel<64, true>).  Default: the 56 x 56 and 28 x 28 stages; also none and all of them (n = 5 makes the
    7 x 7 stage's 128-row tiles straddle up to four images).  Every selection holds 1e-3 against the oracle."""
    from mcncrossmodalemotions_b200.programs import TeacherProgram

    p = nets.teacher_init("senet50")
    x = nets.synth_faces(n)
    ref = nets.teacher_forward(_f64(p), x.astype(np.float64), nets.TorchOps).reshape(8, n).T
    out = {}
    for min_hw in ("0", "28", "7"):
        monkeypatch.setenv("XEMO_SE_LIN_MIN_HW", min_hw)
        prog = TeacherProgram(p, n, use_graph=False)
        assert prog.se_lin(56) == (min_hw != "0") and prog.se_lin(7) == (min_hw == "7")
        out[min_hw] = prog.forward(x)
        assert rel_err(out[min_hw], ref) < TOL, (min_hw, rel_err(out[min_hw], ref))
    assert not np.array_equal(out["0"], out["28"]) and not np.array_equal(out["28"], out["7"])     # the paths do differ


@pytest.mark.parametrize("width,n", [(300, 8), (100, 5), (400, 3)])
def test_student_test_mode_forward(nets, width, n):
    from mcncrossmodalemotions_b200.programs import StudentProgram

    p = nets.student_randomize_bn(nets.student_init())
    spec = nets.synth_spectrograms(n, width)
    ref, _ = nets.student_forward(_f64(p), spec.astype(np.float64), "test", nets.TorchOps)
    prog = StudentProgram(p, n, width)
    got = prog.forward(spec, "test")
    assert rel_err(got, ref.reshape(8, n).T) < TOL
    assert np.array_equal(got, prog.forward(spec, "test"))  # graph replay


def _cos(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b)))


def _l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64).reshape(np.shape(a))
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.mark.parametrize("use_graph,stem", [(False, True), (True, True), (True, False)])
def test_student_training_step(nets, use_graph, stem):
    """stem = True: conv1 / bn1 / pool1 through the linearity path of csrc/stem_kernels.cuh and the pixel-pair conv1
    (the defaults); False: the generic per-layer kernels."""
    from mcncrossmodalemotions_b200.programs import StudentProgram

    n, width, lr = 16, 300, 1e-4
    p = nets.student_randomize_bn(nets.student_init())
    spec = nets.synth_spectrograms(n, width)
    tgt = nets.synth_teacher_logits(n)
    exact_p, model_p = _f64(p), _f64(p)
    exact = nets.distillation_student_step(exact_p, {}, spec.astype(np.float64), tgt.astype(np.float64), lr=lr, ops=nets.TorchOps)
    model = nets.distillation_student_step(model_p, {}, spec.astype(np.float64), tgt.astype(np.float64), lr=lr, ops=nets.Fp16ModelOps)
    prog = StudentProgram(p, n, width, use_graph=use_graph, stem_algebra=stem, stem_pairs=stem)
    assert prog.stem_algebra == stem and prog.stem_pairs == stem
    prog.set_hyper(lr=lr)
    prog.reset_metrics()
    prog.train_step(spec, tgt)
    m = prog.metrics()
    grads, params = prog.export_grads(), prog.export_params()
    # continuous quantities against the exact oracle at the north_star tolerance
    assert abs(m["objective"] - exact["objective"]) <= TOL * abs(exact["objective"])
    assert m["classerror"] == exact["classerror"]
    correct, count = nets.M.error_stats(exact["prediction"], exact["max_label"], 8)
    assert np.array_equal(m["count"], count) and np.array_equal(m["correct"], correct)
    for k in params:
        if k.endswith("x"):
            assert rel_err(grads[k], exact["grads"][k]) < TOL, k      # batch moments [mu sigma]
            assert rel_err(params[k], exact_p[k]) < TOL, k            # ... and their moving average
    # (the updated weights are NOT compared with the oracle's: one step at lr = 1e-4 moves a weight by ~1e-4 of its value,
    # so that check could not see a wrong gradient.  The update is checked below as the step (w' - w) / lr given the
    # device's own gradient, and the gradient itself -- at 1e-3 -- in the fp32-equivalent mode, tests/test_gpu_parity.py.)
    # the gradient against the fp16 number-format model (same rounding points, fp64 accumulation) and its
    # direction against the exact oracle
    for k in grads:
        if k.endswith("x") or (k.endswith("b") and not k.startswith("bn") and k != "fc8b"):
            continue  # moments checked above; conv biases ahead of train-mode BN have an exactly-zero gradient
        # ReLU-mask flips make the gradient chaotic at the ~1e-1 level under ANY perturbation of size 2^-11
        # (the fp16 model itself sits 0.04-0.13 from the exact oracle, see DESIGN.md "Precision"): the
        # kernels must be no further from the model than the model is from the truth, and point the same way
        assert _l2(grads[k], model["grads"][k]) < 0.15, (k, _l2(grads[k], model["grads"][k]))
        assert _cos(grads[k], exact["grads"][k]) > 0.97, (k, _cos(grads[k], exact["grads"][k]))
    # SGD-momentum bookkeeping given the gradient, as the step cnn_train_dag takes: (w' - w) / lr = -(wd w + g / B)
    for k in params:
        if k.endswith("x"):
            continue
        w0 = p[k].astype(np.float64).reshape(params[k].shape)
        step = (params[k].astype(np.float64) - w0) / lr
        expect = -(5e-4 * w0 + grads[k].reshape(params[k].shape).astype(np.float64) / n)
        if np.abs(expect).max() == 0:
            assert np.abs(step).max() == 0, k
            continue
        noise = 6e-8 * np.abs(w0).max() / lr / np.abs(expect).max()     # w' is rounded to fp32
        assert rel_err(step, expect) < 1e-5 + noise, (k, rel_err(step, expect), noise)
    assert m["skipped_steps"] == 0 and not m["nonfinite_grad"]


def test_nonfinite_gradient_skips_the_update(nets):
    """The fp16 gradient chain runs under a fixed loss scale: a non-finite flat gradient must not reach the master weights,
    the momentum, the fp16 mirror or the BN moments; the step is counted and training continues."""
    import torch
    from mcncrossmodalemotions_b200.programs import StudentProgram

    n, width = 4, 100
    p = nets.student_randomize_bn(nets.student_init())
    prog = StudentProgram(p, n, width)
    prog.set_hyper(lr=1e-2, batch_size=n)
    prog.set_input(nets.synth_spectrograms(n, width), nets.synth_teacher_logits(n))
    prog.grad_step()
    prog.sync()
    before = {k: v.clone() for k, v in (("w", prog.master), ("m", prog.momentum), ("h", prog.w16), ("x", prog.moments["bn3"]))}
    with torch.cuda.stream(prog.stream):
        prog.grad[12345] = float("inf")
    prog.update()
    m = prog.metrics()
    assert m["nonfinite_grad"] and m["skipped_steps"] == 1
    assert all(torch.equal(before[k], v) for k, v in (("w", prog.master), ("m", prog.momentum), ("h", prog.w16), ("x", prog.moments["bn3"])))
    prog.grad_step()      # a clean gradient again: the update goes through, the counter stays
    prog.update()
    m = prog.metrics()
    assert not m["nonfinite_grad"] and m["skipped_steps"] == 1
    assert not torch.equal(before["w"], prog.master) and not torch.equal(before["x"], prog.moments["bn3"])
    assert torch.isfinite(prog.master).all()


def test_stem_linearity_path_agrees_with_generic_path(nets):
    """Same step through both formulations of the first layer.  The batch statistics agree to fp16 storage rounding of
    the conv1 activation; the gradients differ by what ANY two fp16 pipelines differ by (ReLU-mask flips, DESIGN.md
    section 5: measured 0.08-0.09 relative L2 between the two paths, each 0.12-0.14 from the exact oracle and 0.08
    from the fp16 model -- tests/tools/stem_diag.py), so they are held to the same 0.15 as the oracle comparison."""
    from mcncrossmodalemotions_b200.programs import StudentProgram

    n, width = 8, 100
    p = nets.student_randomize_bn(nets.student_init())
    spec, tgt = nets.synth_spectrograms(n, width), nets.synth_teacher_logits(n)
    out = []
    for stem in (False, True):
        prog = StudentProgram(p, n, width, use_graph=False, stem_algebra=stem, stem_pairs=stem)
        prog.reset_metrics()
        prog.set_input(spec, tgt)
        prog.grad_step()
        out.append((prog.export_grads(), prog.metrics()))
    (g0, m0), (g1, m1) = out
    assert abs(m0["objective"] - m1["objective"]) <= 1e-4 * abs(m0["objective"])
    assert rel_err(g1["bn1x"], g0["bn1x"]) < 1e-4
    for k in ("conv1f", "bn1m", "bn1b", "conv2f", "fc8f"):
        assert _l2(g1[k], g0[k]) < 0.15, (k, _l2(g1[k], g0[k]))
        assert _cos(g1[k], g0[k]) > 0.98, (k, _cos(g1[k], g0[k]))
    assert np.abs(g1["conv1b"]).max() == 0


def test_student_softmaxlog_loss_type(nets):
    """lossType 'softmaxlog' (emoVoxZoo.m:147-149: dagnn.Loss('softmaxlog') on {prediction, maxLabel}) through the fused
    loss kernel's one-hot / T = 1 mode: objective, class error and the fc8 bias gradient against the oracle's vl_nnloss
    applied to the step's own predictions."""
    import torch

    from mcncrossmodalemotions_b200.programs import StudentProgram
    from oracle import mcn_ops as M

    n, width = 8, 100
    labels = np.array([1, 1, 1, 2, 2, 3, 5, 8]).reshape(1, 1, 1, n)
    prog = StudentProgram(nets.student_init(), n, width, use_graph=False, loss_type="softmaxlog")
    prog.reset_metrics()
    prog.set_input(nets.synth_spectrograms(n, width), labels)
    prog.grad_step()
    m, g = prog.metrics(), prog.export_grads()
    with torch.cuda.stream(prog.stream):
        pred = prog.a["pred32"][:, :8].cpu().numpy()
    prog.sync()
    x4 = pred.T.reshape(1, 1, 8, n).astype(np.float64)
    obj = M.vl_nnloss(x4, labels, loss="softmaxlog")
    assert abs(m["objective"] - obj) <= 1e-3 * abs(obj)            # pred32 is fp32, the loss kernel reads the fp16 copy
    assert abs(m["classerror"] - M.vl_nnloss(x4, labels, loss="classerror")) <= 1
    assert np.array_equal(m["count"], np.bincount(labels.ravel() - 1, minlength=8))
    dx = M.vl_nnloss(x4, labels, 1.0, loss="softmaxlog")             # softmax(x) - onehot
    assert rel_err(g["fc8b"], dx.sum(axis=(0, 1, 3))) < 5e-3


def test_student_bias_before_train_bn_has_zero_gradient(nets):
    from mcncrossmodalemotions_b200.programs import StudentProgram

    n = 8
    p = nets.student_init()
    prog = StudentProgram(p, n, 100, use_graph=False)
    prog.reset_metrics()
    prog.set_input(nets.synth_spectrograms(n, 100), nets.synth_teacher_logits(n))
    prog.grad_step()
    g = prog.export_grads()
    for i in range(1, 8):
        name = ("conv%d" % i) if i < 6 else ("fc%d" % i)
        assert np.abs(g[name + "b"]).max() <= 2e-2 * np.abs(g["bn%db" % i]).max(), name


def test_launch_counter_counts_graph_nodes(nets):
    from mcncrossmodalemotions_b200.programs import TeacherProgram

    n = 2
    prog = TeacherProgram(nets.teacher_init("resnet50"), n)
    prog.forward(nets.synth_faces(n))
    c0 = prog.ctx.launch_count()
    prog.run()
    prog.sync()
    assert prog.ctx.launch_count() - c0 == prog.graph.num_kernels > 50


def test_split_backward_with_overlapped_allreduce_equals_single_pass(nets):
    """The data-parallel step cuts the backward pass after fc6 so that the tail of the gradient buffer can be
    all-reduced while conv5..conv1 are differentiated: with an identity 'all-reduce' it must reproduce the single-pass
    step (up to the summation order of the split-K atomics)."""
    from mcncrossmodalemotions_b200 import zoo
    from mcncrossmodalemotions_b200.distill import DistillationStep

    n = 8
    faces, spec = nets.synth_faces48(n), nets.synth_spectrograms(n, 100)
    calls, out = [], []
    for allreduce in (None, lambda t: calls.append(t.numel())):
        step = DistillationStep(zoo.teacher_init("resnet50"), zoo.student_init(), n, 100, comm_overlap=True)
        step.student.set_hyper(lr=1e-3, batch_size=n)
        step.teacher.set_input(faces)
        step.student.set_input(spec)
        for _ in range(2):   # second call replays the captured graphs
            step.step_resident(allreduce)
        step.sync()
        out.append((step.student.export_grads(), step.student.export_params(), step.student.metrics()))
    assert len(calls) == 4 and calls[0] + calls[1] == step.student.nparam and calls[0] > calls[1]   # tail (fc6..fc8) first, then head
    (g0, p0, m0), (g1, p1, m1) = out
    assert abs(m0["objective"] - m1["objective"]) <= 1e-5 * abs(m0["objective"])
    for k in g0:
        if k.endswith("f") or k.endswith("x") or k.endswith("m"):
            assert rel_err(g1[k], g0[k]) < 1e-3, k
            assert rel_err(p1[k], p0[k]) < 1e-5, k
