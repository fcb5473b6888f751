"""Full-size fixtures for BASELINE.json's configurations, generated OFFLINE by the CPU oracle (fp32 torch CPU kernels behind
the MatConvNet operator semantics; ~15 minutes on 8 cores) and committed as tests/golden/fullsize.npz:
    C2  ResNet50-ferplus forward, all 256 faces (224 x 224 x 3)                      -> logits 256 x 8
    C3  VGGVox student step, batch 128 @512x300                                      -> train-mode logits, objective, class error,
                                                                                        batch moments, per-tensor gradient norms
    C4  full distillation step, batch 256: SENet50 on 256 48x48 faces -> max-aggregation (F = 1) -> student step
                                                                                     -> teacher logits, objective, train-mode logits, moments
    C5  embedding extraction at batch 64: SENet50 logits of 64 faces, test-mode VGGVox logits of 64 clips
Inputs and weights are seeded (oracle.nets.synth_* / *_init), so only the outputs are stored; tests/test_gpu_fullsize.py
regenerates the inputs on the GPU box and asserts the CUDA path against these vectors.  Oracle-generated, not reference
outputs (the reference has none: SURVEY.md section 8c).
    python tests/tools/make_fullsize_golden.py"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import nets  # noqa: E402

torch.set_num_threads(os.cpu_count() or 1)
PATH = os.path.join(ROOT, "tests", "golden", "fullsize.npz")
ONLY = sys.argv[1] if len(sys.argv) > 1 else None          # e.g. `c5`: add / refresh one configuration, keep the rest
g = dict(np.load(PATH)) if (ONLY and os.path.exists(PATH)) else {}
t0 = time.time()


def chunks(fn, x, size):
    return np.concatenate([fn(x[..., i:i + size]) for i in range(0, x.shape[-1], size)], axis=-1)


# C2
if ONLY in (None, "c2"):
    p = nets.teacher_init("resnet50")
    g["c2_logits"] = chunks(lambda x: nets.teacher_forward(p, x, nets.TorchOps), nets.synth_faces(256), 32).reshape(8, 256).T
    print("C2 done %.0f s" % (time.time() - t0), flush=True)


def student(tag, n, target):
    sp = nets.student_init()
    out = nets.distillation_student_step(sp, {}, nets.synth_spectrograms(n, 300), target, lr=1e-4, ops=nets.TorchOps, update=False)
    g[tag + "_prediction"] = out["prediction"].reshape(8, n).T
    g[tag + "_objective"] = np.array(out["objective"])
    g[tag + "_classerror"] = np.array(out["classerror"])
    for k, v in out["grads"].items():
        if k.endswith("x"):
            g[tag + "_moments_" + k] = np.asarray(v)
        else:
            g[tag + "_gradnorm_" + k] = np.array(np.linalg.norm(np.asarray(v, np.float64)))


# C3
if ONLY in (None, "c3"):
    student("c3", 128, nets.synth_teacher_logits(128))
    print("C3 done %.0f s" % (time.time() - t0), flush=True)
# C4
if ONLY in (None, "c4"):
    tp = nets.teacher_init("senet50")
    logits = chunks(lambda x: nets.teacher_forward(tp, x, nets.TorchOps), nets.faces48_to_input(nets.synth_faces48(256)), 32)
    g["c4_teacher_logits"] = logits.reshape(8, 256).T
    student("c4", 256, logits.reshape(1, 1, 8, 256).astype(np.float32))
    print("C4 done %.0f s" % (time.time() - t0), flush=True)
# C5: embedding extraction, batch 64: SENet50 logits of 64 faces, test-mode student logits of 64 clips (non-trivial BN state)
if ONLY in (None, "c5"):
    tp = nets.teacher_init("senet50")
    g["c5_teacher_logits"] = chunks(lambda x: nets.teacher_forward(tp, x, nets.TorchOps), nets.synth_faces(64, seed=31), 32).reshape(8, 64).T
    sp = nets.student_randomize_bn(nets.student_init())
    pred, _ = nets.student_forward(sp, nets.synth_spectrograms(64, 300, seed=32), "test", nets.TorchOps)
    g["c5_student_logits"] = pred.reshape(8, 64).T
    print("C5 done %.0f s" % (time.time() - t0), flush=True)
np.savez_compressed(PATH, **{k: np.asarray(v, np.float32) for k, v in g.items()})
print("wrote tests/golden/fullsize.npz (%d arrays)" % len(g))
