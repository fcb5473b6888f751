"""The fp32-equivalent (split-operand) mode against the fp64 CPU oracle at the north_star tolerance, 1e-3 relative --
for everything a training step produces: train-mode logits, objective, batch moments, every parameter gradient (under
equal discrete decisions, see _check_step) and the update itself (reference: cnn_train_dag in `single`,
emoVoxCeleb/run_distillation.m:170-182; loss emoVoxZoo.m:137-157).

Two error measures are asserted / reported for every tensor:
  range-relative  max|a-b| / max|ref|          (conftest.rel_err; the north_star criterion)   <= 1e-3
  per-element     |a-b| / |ref| over the elements with |ref| >= 1e-2 * max|ref|              <= 5e-2 (an error of 2e-4 of the
                  range is 2e-2 of an element at 1 % of the range; measured <= 1.4e-2)
"""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-3


def elem_err(a, b, floor=1e-2):
    """max per-element relative error over the elements that are not negligible against the tensor's range."""
    a, b = np.asarray(a, np.float64).reshape(-1), np.asarray(b, np.float64).reshape(-1)
    big = np.abs(b) >= floor * np.abs(b).max()
    return float((np.abs(a - b)[big] / np.abs(b)[big]).max()) if big.any() else 0.0


@pytest.fixture(scope="module")
def nets():
    from oracle import nets

    return nets


def _f64(p):
    return {k: (v.astype(np.float64) if isinstance(v, np.ndarray) else v) for k, v in p.items()}


CONV_CASES = [
    # H, W, C, N, FH, FW, K, pad, stride
    (17, 13, 3, 2, 7, 7, 20, (3, 3, 3, 3), (2, 2)),      # teacher-stem-like, K not a multiple of 16
    (14, 14, 64, 3, 3, 3, 64, (1, 1, 1, 1), (1, 1)),
    (20, 11, 96, 2, 5, 5, 32, (1, 1, 1, 1), (2, 2)),     # student conv2-like
    (9, 8, 32, 2, 9, 1, 48, (0, 0, 0, 0), (1, 1)),       # fc6-like
    (64, 30, 1, 2, 7, 7, 96, (1, 1, 1, 1), (2, 2)),      # student conv1 (one input channel)
    (7, 7, 128, 4, 1, 1, 8, (0, 0, 0, 0), (2, 2)),       # strided 1x1
]


@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "x".join(str(v) for v in c[:7]))
def test_vl_nnconv_split_mode_is_fp32_equivalent(case):
    """vl_nnconv in XEMO_CONV_F32X3 against the fp64 oracle: 2e-5 (measured 2e-7 ... 8e-6: fp32 accumulation of up to 2400
    products whose magnitudes span three decades), forward and all gradients; the default fp16-operand mode on the same
    inputs is worse than 1e-5 -- i.e. the flag does switch the arithmetic."""
    from oracle import mcn_ops as M
    from mcncrossmodalemotions_b200 import vl_nn

    H, W, C, N, FH, FW, K, pad, stride = case
    rng = np.random.default_rng(7)
    # wide dynamic range across channels / filters: the per-tensor power-of-two scale must keep small entries accurate
    x = (rng.standard_normal((H, W, C, N)) * 10.0 ** rng.uniform(-2, 1, (1, 1, C, 1))).astype(np.float32)
    f = (rng.standard_normal((FH, FW, C, K)) * 10.0 ** rng.uniform(-3, 0, (1, 1, 1, K))).astype(np.float32)
    b = rng.standard_normal(K).astype(np.float32)
    y64 = M.vl_nnconv(x.astype(np.float64), f.astype(np.float64), b.astype(np.float64), pad=pad, stride=stride)
    dy = (rng.standard_normal(y64.shape) * 1e-4).astype(np.float32)     # gradients are small numbers
    dx64, df64, db64 = M.vl_nnconv(x.astype(np.float64), f.astype(np.float64), b.astype(np.float64), dy.astype(np.float64), pad=pad, stride=stride)
    ctx = vl_nn.default_context()
    y16 = vl_nn.vl_nnconv(x, f, b, pad=pad, stride=stride)
    ctx.set_conv_precision(1)
    try:
        assert ctx.lib.xemo_get_conv_precision(ctx.handle) == 1
        y = vl_nn.vl_nnconv(x, f, b, pad=pad, stride=stride)
        dx, df, db = vl_nn.vl_nnconv(x, f, b, dy, pad=pad, stride=stride)
        # gpuArray inputs take the same path
        yg = vl_nn.gather(vl_nn.vl_nnconv(vl_nn.gpuArray(x), vl_nn.gpuArray(f), vl_nn.gpuArray(b.reshape(-1, 1)), pad=pad, stride=stride))
    finally:
        ctx.set_conv_precision(0)
    assert rel_err(y, y64) < 2e-5, rel_err(y, y64)
    assert np.array_equal(yg, y)
    assert rel_err(dx, dx64) < 2e-5, rel_err(dx, dx64)
    assert rel_err(df, df64) < 2e-5, rel_err(df, df64)
    assert rel_err(db, db64) < 2e-6
    assert rel_err(y16, y64) > 3 * rel_err(y, y64) and rel_err(y16, y64) < 1e-3


def test_vl_nnconv_split_mode_degenerate_inputs():
    """all-zero operands (scale guard) and tiny / huge magnitudes"""
    from oracle import mcn_ops as M
    from mcncrossmodalemotions_b200 import vl_nn

    ctx = vl_nn.default_context()
    rng = np.random.default_rng(8)
    x = rng.standard_normal((6, 5, 16, 2)).astype(np.float32)
    f = rng.standard_normal((3, 3, 16, 16)).astype(np.float32)
    ctx.set_conv_precision(1)
    try:
        assert np.abs(vl_nn.vl_nnconv(np.zeros_like(x), f, None, pad=1)).max() == 0
        assert np.abs(vl_nn.vl_nnconv(x, np.zeros_like(f), None, pad=1)).max() == 0
        for sx, sf in ((1e-20, 1e10), (1e15, 1e-12), (3e-30, 7e-3)):
            ref = M.vl_nnconv(x.astype(np.float64) * sx, f.astype(np.float64) * sf, None, pad=1)
            got = vl_nn.vl_nnconv((x * np.float32(sx)), (f * np.float32(sf)), None, pad=1)
            assert rel_err(got, M.vl_nnconv((x * np.float32(sx)).astype(np.float64), (f * np.float32(sf)).astype(np.float64), None, pad=1)) < 2e-5
            assert rel_err(got, ref) < 2e-5
    finally:
        ctx.set_conv_precision(0)


def _loss_fns(nets, pred, tgt, max_label, loss_type, w):
    M = nets.M
    if loss_type == "hot-cross-ent":
        return lambda *dz: M.vl_nnsoftmaxceloss(pred, tgt, *dz, temperature=2.0, logitTargets=True)
    if loss_type == "softmaxlog":
        return lambda *dz: M.vl_nnloss(pred, max_label, *dz, loss="softmaxlog")
    if loss_type == "euclidean":
        return lambda *dz: M.vl_nneuclideanloss(pred, tgt, *dz, instanceWeights=w)
    return lambda *dz: M.vl_nnhuberloss(pred, tgt, *dz, sigma=1.0, instanceWeights=w)


def _check_step(nets, n, width, loss_type="hot-cross-ent"):
    """One training step of the fp32-equivalent program against the fp64 oracle.

    Forward quantities (train-mode logits, objective, batch moments, class error / ErrorStats counters) are compared
    directly.  Gradients are compared with the oracle's backward UNDER THE PROGRAM'S OWN DISCRETE DECISIONS (ReLU masks and
    pooling winners, exported by the program): the backward pass is a discontinuous function of the activations, and a
    SINGLE mask that differs moves a BatchNorm bias gradient of the late layers (N rows per channel) by ~1/N of its range.
    No single-precision implementation can hold 1e-3 against fp64 there -- the CPU oracle itself, run in fp32, sits
    1e-2 ... 1e-1 from its fp64 run at N = 16 with 9 masks and 7 pooling winners changed out of 9e7
    (profiles/r02_fp32_vs_fp64_oracle.txt, tests/tools/f32_vs_f64_oracle.py) -- so the number of differing decisions is
    asserted (<= 2e-6 of all) and the arithmetic is asserted under equal decisions; the unconditioned distance is printed."""
    from mcncrossmodalemotions_b200.parity import StudentProgramF32

    lr = 1e-4
    p = nets.student_randomize_bn(nets.student_init())
    spec, tgt = nets.synth_spectrograms(n, width), nets.synth_teacher_logits(n)
    w = np.random.default_rng(11).uniform(0.5, 2.0, n).astype(np.float32) if loss_type in ("euclidean", "huber") else None
    p64 = _f64(p)
    pred64, tape = nets.student_forward(p64, spec.astype(np.float64), "train", nets.TorchOps, keep=True)
    max_label = tgt.argmax(axis=2).reshape(1, 1, 1, n) + 1
    loss = _loss_fns(nets, pred64, tgt.astype(np.float64), max_label, loss_type, w)
    objective, dpred = float(loss()), loss(np.array(1.0))
    classerror = nets.M.vl_nnloss(pred64, max_label, loss="classerror")

    prog = StudentProgramF32(p, n, width, loss_type=loss_type)
    prog.set_hyper(lr=lr, batch_size=n)
    prog.reset_metrics()
    prog.train_step(spec, max_label if loss_type == "softmaxlog" else tgt, weights=w)
    m = prog.metrics()
    grads, params, dec = prog.export_grads(), prog.export_params(), prog.export_decisions()
    masks = {k: v for k, v in dec.items() if k.startswith("relu")}
    index = {k: v for k, v in dec.items() if k.startswith("pool")}
    free = nets.student_backward(p64, tape, dpred, nets.TorchOps)
    cond = nets.student_backward(p64, tape, dpred, nets.TorchOps, relu_masks=masks, pool_index=index)

    failures, rows = [], []
    pred = prog.prediction()
    ref_pred = pred64.reshape(8, n).T
    rows.append(("prediction (train mode)", rel_err(pred, ref_pred), elem_err(pred, ref_pred), None))
    if rel_err(pred, ref_pred) >= TOL:
        failures.append(("prediction", rel_err(pred, ref_pred)))
    assert abs(m["objective"] - objective) <= 1e-4 * abs(objective), (m["objective"], objective)
    assert m["classerror"] == classerror
    correct, count = nets.M.error_stats(pred64, max_label, 8)
    assert np.array_equal(m["count"], count) and np.array_equal(m["correct"], correct)
    # discrete decisions that differ from the oracle's own
    total = sum(v.size for v in dec.values())
    differ = sum(int((v != (tape[k + ":x"] > 0)).sum()) for k, v in masks.items()) + \
        sum(int((v != tape[k + ":argmax"]).sum()) for k, v in index.items())
    assert differ <= max(10, 2e-6 * total), (differ, total)     # measured: 9 of 7.2e6 (N = 4), 80 of 9.1e7 (N = 16)
    for k in sorted(grads):
        if k.endswith("x"):
            r = rel_err(grads[k], free[k])          # batch moments [mu sigma]: a forward quantity
            rows.append(("moments/" + k, r, elem_err(grads[k], free[k]), None))
            if r >= TOL:
                failures.append((k, r))
            continue
        ref = np.asarray(cond[k]).reshape(grads[k].shape)
        if np.abs(ref).max() < 1e-12 * max(1.0, np.abs(cond["fc8f"]).max()):
            # conv biases ahead of a train-mode BN: the exact gradient is zero (fp64: ~1e-17); ours is a sum of ~1e6 cancelling
            # fp32 terms -- negligible against the layer's filter gradient
            if np.abs(grads[k]).max() > TOL * np.abs(cond[k[:-1] + "f"]).max():
                failures.append((k, float(np.abs(grads[k]).max())))
            continue
        r, e, r_free = rel_err(grads[k], ref), elem_err(grads[k], ref), rel_err(grads[k], np.asarray(free[k]).reshape(grads[k].shape))
        rows.append(("grad/" + k, r, e, r_free))
        if r >= TOL or e >= 5e-2:
            failures.append((k, r, e))
    # the update as cnn_train_dag applies it, as the step (w' - w) / lr = -(wd w + g / B) given the program's own gradient
    # (not w': one step moves a weight by ~1e-4 of its value, which would hide a wrong gradient)
    for k in sorted(params):
        if k.endswith("x"):
            ref = 0.9 * p64[k] + 0.1 * np.asarray(free[k])
            if rel_err(params[k], ref) >= TOL:
                failures.append(("moments average " + k, rel_err(params[k], ref)))
            continue
        w0 = p64[k].reshape(params[k].shape)
        step = (params[k].astype(np.float64) - w0) / lr
        expect = -(5e-4 * w0 + grads[k].reshape(params[k].shape).astype(np.float64) / n)
        if np.abs(expect).max() == 0:
            continue
        noise = 6e-8 * np.abs(w0).max() / lr / np.abs(expect).max()      # w' is rounded to fp32
        if rel_err(step, expect) >= 1e-5 + noise:
            failures.append(("step/" + k, rel_err(step, expect), noise))
    print("\nfp32-equivalent student step (%s), N = %d, W = %d vs the fp64 oracle; %d of %d discrete decisions differ" % (
        loss_type, n, width, differ, total))
    print("  %-28s %-10s %-10s %s" % ("tensor", "range-rel", "per-elem", "range-rel without conditioning on the decisions"))
    for name, r, e, rf in rows:
        print("  %-28s %.2e   %.2e   %s" % (name, r, e, "" if rf is None else "%.2e" % rf))
    assert not failures, failures
    return rows


@pytest.mark.parametrize("n,width", [(4, 100), (16, 300)])
def test_student_training_step_f32x3_matches_the_oracle(nets, n, width):
    _check_step(nets, n, width)


def test_student_training_step_f32x3_reference_default_operating_point(nets):
    """4-second clips (512 x 400), the reference's default numSeconds (emoVoxCeleb/run_distillation.m:74), train mode."""
    _check_step(nets, 6, 400)


@pytest.mark.parametrize("loss_type", ["softmaxlog", "euclidean", "huber"])
def test_student_training_step_f32x3_other_loss_types(nets, loss_type):
    _check_step(nets, 5, 100, loss_type=loss_type)


def test_student_test_mode_forward_f32x3(nets):
    """dag.mode = 'test' (external/compute_audio_feats.m:106-126) at a batch where the fp16-operand program has no
    margin left (1.0e-3 at N = 32): the fp32-equivalent forward holds 2e-4 (measured 6.6e-5)."""
    from mcncrossmodalemotions_b200.parity import StudentProgramF32

    n, width = 32, 300
    p = nets.student_randomize_bn(nets.student_init())
    spec = nets.synth_spectrograms(n, width)
    ref, _ = nets.student_forward(_f64(p), spec.astype(np.float64), "test", nets.TorchOps)
    got = StudentProgramF32(p, n, width).forward(spec, "test")
    assert rel_err(got, ref.reshape(8, n).T) < 2e-4, rel_err(got, ref.reshape(8, n).T)
