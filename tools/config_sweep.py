"""Throughput of the other BASELINE.json configs on one B200 (CUDA events, graphs, inputs resident in HBM):
  C2  ResNet50-ferplus teacher forward, batch 256 x 224 x 224 x 3
  C3  VGGVox student forward + backward (+ loss + SGD) on 512 x 300 spectrograms, batch 128
  C5  embedding-extraction sweeps, batch 64 ... 1024 (SENet50 teacher on 224x224x3; student test-mode forward @512x300)
    python tools/config_sweep.py > gpurun_out/config_sweep.json"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcncrossmodalemotions_b200 import zoo  # noqa: E402
from mcncrossmodalemotions_b200.programs import StudentProgram, TeacherProgram  # noqa: E402

GF = {"resnet50": 7.711883264, "senet50": 7.716913152, "student_fwd": 5.662228992, "student_step": 16.633}


def timed(fn, stream, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(iters):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


out = {}
tp = {a: zoo.teacher_init(a) for a in ("resnet50", "senet50")}
sp = zoo.student_init()
# C2
prog = TeacherProgram(tp["resnet50"], 256)
prog.run(); prog.sync()
ms = timed(prog.run, prog.stream)
out["C2 resnet50 fwd n256"] = dict(ms=ms, faces_per_s=256 / ms * 1e3, tflops=256 * GF["resnet50"] / ms)
del prog
# C3
prog = StudentProgram(sp, 128, 300)
prog.set_hyper(lr=1e-4, batch_size=128)
rng = np.random.default_rng(0)
prog.set_input(rng.standard_normal((512, 300, 1, 128)).astype(np.float32), rng.standard_normal((1, 1, 8, 128)).astype(np.float32))
step = lambda: (prog.grad_step(), prog.update())
ms = timed(step, prog.stream)
out["C3 student fwd+bwd n128"] = dict(ms=ms, clips_per_s=128 / ms * 1e3, tflops=128 * GF["student_step"] / ms)
del prog
# C5
for n in (64, 128, 256, 512, 1024):
    prog = TeacherProgram(tp["senet50"], n)
    prog.run(); prog.sync()
    ms = timed(prog.run, prog.stream, iters=5)
    out["C5 senet50 fwd n%d" % n] = dict(ms=ms, faces_per_s=n / ms * 1e3, tflops=n * GF["senet50"] / ms)
    del prog
    torch.cuda.empty_cache()
    prog = StudentProgram(sp, n, 300)
    f = lambda: prog._run("fwd_test", lambda: prog._record_forward(False))
    f(); prog.sync()
    ms = timed(f, prog.stream, iters=5)
    out["C5 student test fwd n%d" % n] = dict(ms=ms, clips_per_s=n / ms * 1e3, tflops=n * GF["student_fwd"] / ms)
    del prog
    torch.cuda.empty_cache()
print(json.dumps(out, indent=1))
