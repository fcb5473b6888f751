"""Distance of the first-layer gradients to the oracle for both formulations of the student stem
(generic per-layer kernels vs. csrc/stem_kernels.cuh).   python tests/tools/stem_diag.py [n] [width]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import nets  # noqa: E402  (diagnostic tool: the oracle is the checker)
from mcncrossmodalemotions_b200.programs import StudentProgram  # noqa: E402


def l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64).reshape(np.shape(a))
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
width = int(sys.argv[2]) if len(sys.argv) > 2 else 300
p = nets.student_randomize_bn(nets.student_init())
spec, tgt = nets.synth_spectrograms(n, width), nets.synth_teacher_logits(n)
f64 = lambda d: {k: (v.astype(np.float64) if isinstance(v, np.ndarray) else v) for k, v in d.items()}
exact = nets.distillation_student_step(f64(p), {}, spec.astype(np.float64), tgt.astype(np.float64), ops=nets.TorchOps, update=False)
model = nets.distillation_student_step(f64(p), {}, spec.astype(np.float64), tgt.astype(np.float64), ops=nets.Fp16ModelOps, update=False)
res = {}
for stem in (False, True):
    prog = StudentProgram(p, n, width, use_graph=False, stem_algebra=stem, stem_pairs=stem)
    prog.reset_metrics()
    prog.set_input(spec, tgt)
    prog.grad_step()
    res[stem] = prog.export_grads()
print("n=%d width=%d   (relative L2 distances)" % (n, width))
print("%-8s %12s %12s %12s %12s %12s" % ("param", "gen-exact", "stem-exact", "gen-model", "stem-model", "stem-gen"))
for k in ("conv1f", "bn1m", "bn1b", "bn1x", "conv2f", "bn2m"):
    print("%-8s %12.4g %12.4g %12.4g %12.4g %12.4g" % (k, l2(res[False][k], exact["grads"][k]), l2(res[True][k], exact["grads"][k]),
                                                  l2(res[False][k], model["grads"][k]), l2(res[True][k], model["grads"][k]),
                                                  l2(res[True][k], res[False][k])))
print("model-exact conv1f %.4g" % l2(model["grads"]["conv1f"], exact["grads"]["conv1f"]))
