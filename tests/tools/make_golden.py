"""Generates tests/golden/*.npz from the CPU oracle (fp64): committed known-answer vectors that pin the oracle against
accidental change (tests/test_golden.py, CPU) and give the CUDA path fixed targets (tests/test_gpu_golden.py).

The reference repository has no golden vectors, tests or runnable arithmetic for this path (SURVEY.md section 8c:
parity unpinned), so these are *oracle-generated* fixtures, not reference outputs.  Inputs and weights are seeded
(oracle/nets.py synth_* / *_init), so only the small outputs are stored.

    python tests/tools/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import mcn_ops as M  # noqa: E402
from oracle import nets  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def f64(p):
    return {k: (v.astype(np.float64) if isinstance(v, np.ndarray) else v) for k, v in p.items()}


def ops():
    rng = np.random.default_rng(2024)
    g = {}
    x = rng.standard_normal((9, 8, 6, 2)); f = rng.standard_normal((3, 3, 6, 5)) / 7; b = rng.standard_normal(5)
    dy = rng.standard_normal((5, 4, 5, 2))
    g["conv_x"], g["conv_f"], g["conv_b"], g["conv_dy"] = x, f, b, dy
    g["conv_y"] = M.vl_nnconv(x, f, b, pad=(1, 1, 0, 1), stride=(2, 2))
    g["conv_dx"], g["conv_df"], g["conv_db"] = M.vl_nnconv(x, f, b, dy, pad=(1, 1, 0, 1), stride=(2, 2))
    xp = np.maximum(np.round(rng.standard_normal((9, 7, 8, 2)) * 2) / 2, 0)
    g["pool_x"] = xp
    g["pool_y"], g["pool_idx"] = M.vl_nnpool(xp, (3, 3), pad=(0, 1, 0, 1), stride=2, method="max", return_index=True)
    g["pool_dy"] = rng.standard_normal(g["pool_y"].shape)
    g["pool_dx"] = M.vl_nnpool(xp, (3, 3), g["pool_dy"], pad=(0, 1, 0, 1), stride=2, method="max")
    g["avg_y"] = M.vl_nnpool(xp, (2, 3), pad=(1, 0, 1, 1), stride=(2, 1), method="avg")
    xb = rng.standard_normal((4, 5, 8, 3)) * 2 + 1
    g["bn_x"], g["bn_g"], g["bn_b"], g["bn_dy"] = xb, rng.uniform(0.5, 1.5, 8), rng.standard_normal(8), rng.standard_normal(xb.shape)
    g["bn_y"], g["bn_mom"] = M.vl_nnbnorm(xb, g["bn_g"], g["bn_b"], epsilon=1e-5)
    g["bn_dx"], g["bn_dg"], g["bn_db"], _ = M.vl_nnbnorm(xb, g["bn_g"], g["bn_b"], g["bn_dy"], epsilon=1e-5)
    xl, tl = 3 * rng.standard_normal((1, 1, 8, 7)), 3 * rng.standard_normal((1, 1, 8, 7))
    g["loss_x"], g["loss_t"] = xl, tl
    g["loss_y"] = np.array(M.vl_nnsoftmaxceloss(xl, tl, temperature=2.0, logitTargets=True))
    g["loss_dx"] = M.vl_nnsoftmaxceloss(xl, tl, 1.0, temperature=2.0, logitTargets=True)
    np.savez_compressed(os.path.join(OUT, "ops.npz"), **g)


def networks():
    g = {}
    for arch in ("resnet50", "senet50"):
        p = nets.teacher_init(arch)
        g["teacher_%s_logits" % arch] = nets.teacher_forward(f64(p), nets.synth_faces(2).astype(np.float64), nets.TorchOps).reshape(8, 2).T
    g["teacher_senet50_logits_48"] = nets.teacher_forward(f64(nets.teacher_init("senet50")),
                                                          nets.faces48_to_input(nets.synth_faces48(2)).astype(np.float64), nets.TorchOps).reshape(8, 2).T
    p = nets.student_randomize_bn(nets.student_init())
    spec = nets.synth_spectrograms(2, 100).astype(np.float64)
    g["student_test_pred_w100"] = nets.student_forward(f64(p), spec, "test", nets.TorchOps)[0].reshape(8, 2).T
    spec4, tgt4 = nets.synth_spectrograms(4, 100).astype(np.float64), nets.synth_teacher_logits(4).astype(np.float64)
    r = nets.distillation_student_step(f64(nets.student_init()), {}, spec4, tgt4, lr=1e-4, ops=nets.TorchOps, update=False)
    g["student_step_objective"] = np.array(r["objective"])
    g["student_step_classerror"] = np.array(r["classerror"])
    g["student_step_prediction"] = r["prediction"].reshape(8, 4).T
    for k in ("bn1x", "bn4x", "bn7x"):
        g["student_step_" + k] = r["grads"][k]
    for k in ("fc8f", "fc6f", "conv3f", "conv1f"):
        g["student_step_gradnorm_" + k] = np.array(np.linalg.norm(r["grads"][k]))
    wav = 0.1 * np.random.default_rng(77).standard_normal(16384)
    g["runspec_wav"] = wav
    g["runspec_sample"] = nets.run_spec(wav)[::37, ::9]
    np.savez_compressed(os.path.join(OUT, "nets.npz"), **g)


if __name__ == "__main__":
    ops()
    networks()
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")
