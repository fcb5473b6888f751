"""GPU-box parity report: fused device programs (through the C ABI) vs the CPU oracle, per tensor.

    python tests/tools/parity_report.py [--teacher-batch 4] [--student-batch 4] > gpurun_out/parity.txt

The oracle is the checker here (test infrastructure); nothing in the product imports it."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

from oracle import nets  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    d = np.abs(b).max()
    return float(np.abs(a - b).max() / (d if d > 0 else 1.0))


def rel2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    d = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / (d if d > 0 else 1.0))


def teacher_report(arch, n, out):
    from mcncrossmodalemotions_b200.programs import TeacherProgram

    p = nets.teacher_init(arch)
    x = nets.synth_faces(n)
    t0 = time.time()
    ref = nets.teacher_forward({k: (v.astype(np.float64) if isinstance(v, np.ndarray) else v) for k, v in p.items()},
                               x.astype(np.float64), nets.TorchOps).reshape(8, n).T
    t_cpu = time.time() - t0
    for use_graph in (False, True):
        prog = TeacherProgram(p, n, use_graph=use_graph)
        got = prog.forward(x)
        if use_graph:
            got = prog.forward(x)  # replay
        out["teacher/%s/n%d/%s" % (arch, n, "graph" if use_graph else "eager")] = dict(
            rel_max=rel(got, ref), rel_l2=rel2(got, ref), max_ref=float(np.abs(ref).max()), cpu_s=t_cpu)


def student_report(n, width, out):
    from mcncrossmodalemotions_b200.programs import StudentProgram

    p = nets.student_randomize_bn(nets.student_init())
    spec = nets.synth_spectrograms(n, width)
    tgt = nets.synth_teacher_logits(n)
    p64 = {k: v.astype(np.float64) for k, v in p.items()}
    # test-mode forward
    ref, _ = nets.student_forward(p64, spec.astype(np.float64), "test", nets.TorchOps)
    prog = StudentProgram(p, n, width)
    got = prog.forward(spec, "test")
    out["student/test_fwd/n%d/w%d" % (n, width)] = dict(rel_max=rel(got, ref.reshape(8, n).T), rel_l2=rel2(got, ref.reshape(8, n).T))
    # one training step
    state = {}
    r = nets.distillation_student_step(p64, state, spec.astype(np.float64), tgt.astype(np.float64), lr=1e-2, ops=nets.TorchOps)
    prog = StudentProgram(p, n, width)
    prog.set_hyper(lr=1e-2)
    prog.reset_metrics()
    prog.train_step(spec, tgt)
    m = prog.metrics()
    grads = prog.export_grads()
    params = prog.export_params()
    key = "student/train/n%d/w%d" % (n, width)
    out[key + "/objective"] = dict(got=m["objective"], ref=r["objective"], rel=abs(m["objective"] - r["objective"]) / abs(r["objective"]))
    out[key + "/classerror"] = dict(got=m["classerror"], ref=r["classerror"])
    with __import__("torch").cuda.stream(prog.stream):
        pred = prog.a["pred32"][:, :8].cpu().numpy()
    out[key + "/prediction"] = dict(rel_max=rel(pred, r["prediction"].reshape(8, n).T))
    for k in sorted(grads):
        ref_g = r["grads"][k].reshape(grads[k].shape)
        if k.endswith("b") and not k.startswith("bn") and k != "fc8b":
            # bias of a conv followed by train-mode BN: the true gradient is exactly zero (BN removes the
            # bias); compare on the scale of the BN shift gradient of the same layer
            scale = np.abs(r["grads"]["bn" + k[-2] + "b"]).max()
            out[key + "/grad/" + k] = dict(abs_over_bn_bias_grad=float(np.abs(grads[k] - ref_g).max() / scale))
            continue
        out[key + "/grad/" + k] = dict(rel_max=rel(grads[k], ref_g), rel_l2=rel2(grads[k], ref_g))
    for k in sorted(params):
        out[key + "/param/" + k] = dict(rel_max=rel(params[k], p64[k].reshape(params[k].shape)))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--teacher-batch", type=int, default=4)
    ap.add_argument("--student-batch", type=int, default=4)
    ap.add_argument("--width", type=int, default=300)
    ap.add_argument("--skip-teacher", action="store_true")
    args = ap.parse_args()
    out = {}
    student_report(args.student_batch, args.width, out)
    if not args.skip_teacher:
        for arch in ("resnet50", "senet50"):
            teacher_report(arch, args.teacher_batch, out)
    for k, v in out.items():
        print(k, json.dumps(v))
