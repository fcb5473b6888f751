"""Is the distance between the fp16-operand training step and the exact oracle "decision chaos" or arithmetic?

With batch-statistics BN + ReLU + max pooling the backward pass is a discontinuous function of the activations: an
activation within 2^-11 of zero flips its mask under fp16 operand rounding.  DESIGN.md section 5 claims this -- not the
kernels' arithmetic -- is why the fast program's parameter gradients sit 0.04 - 0.14 (relative L2) from the fp64 oracle.
This test isolates the claim: the device exports its ReLU masks and pooling winners, the oracle's fp64 backward is re-run
with THOSE decisions held fixed (oracle.nets.student_backward(relu_masks=, pool_index=)), and the remaining difference --
pure arithmetic: fp16 operands / fp16 activation + gradient storage, fp32 accumulation -- is asserted.

(The fp32-equivalent mode, tests/test_gpu_parity.py, is the configuration that meets 1e-3 without conditioning.)"""
import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _f64(p):
    return {k: (v.astype(np.float64) if isinstance(v, np.ndarray) else v) for k, v in p.items()}


def _l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64).reshape(np.shape(a))
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


@pytest.mark.parametrize("stem", [True, False])
def test_gradients_under_the_device_decisions_are_arithmetic_close(stem):
    from oracle import nets
    from mcncrossmodalemotions_b200.programs import StudentProgram

    n, width = 8, 100
    p = nets.student_randomize_bn(nets.student_init())
    spec, tgt = nets.synth_spectrograms(n, width), nets.synth_teacher_logits(n)
    prog = StudentProgram(p, n, width, use_graph=False, stem_algebra=stem, stem_pairs=stem)
    prog.reset_metrics()
    prog.set_input(spec, tgt)
    prog.grad_step()
    grads, dec = prog.export_grads(), prog.export_decisions()

    p64 = _f64(p)
    pred, tape = nets.student_forward(p64, spec.astype(np.float64), "train", nets.TorchOps, keep=True)
    dpred = nets.M.vl_nnsoftmaxceloss(pred, tgt.astype(np.float64), np.array(1.0), temperature=2.0, logitTargets=True)
    free = nets.student_backward(p64, tape, dpred, nets.TorchOps)
    masks = {k: v for k, v in dec.items() if k.startswith("relu")}
    index = {k: v for k, v in dec.items() if k.startswith("pool")}
    cond = nets.student_backward(p64, tape, dpred, nets.TorchOps, relu_masks=masks, pool_index=index)

    # how many decisions differ at all (fraction of elements): the size of the "chaos" input
    flips = {k: float((v != (tape[k + ":x"] > 0)).mean()) for k, v in masks.items()}
    rows = []
    for k in sorted(grads):
        if k.endswith("x") or (k.endswith("b") and not k.startswith("bn") and k != "fc8b"):
            continue   # batch moments are not gradients; conv biases ahead of train-mode BN have a zero gradient
        rows.append((k, rel_err(grads[k], np.asarray(free[k]).reshape(grads[k].shape)), _l2(grads[k], free[k]),
                     rel_err(grads[k], np.asarray(cond[k]).reshape(grads[k].shape)), _l2(grads[k], cond[k])))
    print("\nfp16-operand student step (stem by linearity: %s), N = %d, W = %d" % (stem, n, width))
    print("  ReLU decisions that differ from the oracle's: " + ", ".join("%s %.2e" % kv for kv in sorted(flips.items())))
    print("  %-8s %-23s %-23s" % ("tensor", "vs oracle (max / L2)", "vs oracle under the device decisions (max / L2)"))
    for k, r0, l0, r1, l1 in rows:
        print("  %-8s %.2e / %.2e   %.2e / %.2e" % (k, r0, l0, r1, l1))
    for k, r0, l0, r1, l1 in rows:
        # measured on B200 (profiles/r02_isolation.txt): conditioning on the decisions takes the relative-L2 distance of the
        # layer gradients from 0.04 ... 0.13 down to 7e-3 ... 9e-3 (fc8, which has no decision downstream of it, sits at
        # 7e-3 either way: that residue is the train-mode logit error of the fp16 forward, 5e-3, carried by dpred)
        assert l1 < 2.5e-2, (k, l1)
        assert r1 < 6e-2, (k, r1)      # (max-norm: fc7f 3.8e-2 under the stem-by-linearity path)
        assert l1 < 0.35 * l0 or l0 < 2.5e-2, (k, l0, l1)
