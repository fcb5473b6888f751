"""A/B timings of option switches on one B200 (CUDA events, graphs, inputs resident in HBM):
    python tools/ab_options.py [batch] > gpurun_out/ab_options.json
Each entry: teacher forward (SENet50) or one student step (forward + backward + update) at `batch`, per option value."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcncrossmodalemotions_b200 import zoo  # noqa: E402
from mcncrossmodalemotions_b200.programs import StudentProgram, TeacherProgram  # noqa: E402


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


def with_env(env, fn):
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        return fn()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def teacher_ms(n, arch="senet50"):
    prog = TeacherProgram(zoo.teacher_init(arch), n)
    prog.run(); prog.sync()
    ms = timed(prog.run, prog.stream)
    del prog
    torch.cuda.empty_cache()
    return ms


def student_ms(n, width=300):
    prog = StudentProgram(zoo.student_init(), n, width)
    prog.set_hyper(lr=1e-4, batch_size=n)
    rng = np.random.default_rng(0)
    prog.set_input(rng.standard_normal((512, width, 1, n)).astype(np.float32), rng.standard_normal((1, 1, 8, n)).astype(np.float32))
    ms = timed(lambda: (prog.grad_step(), prog.update()), prog.stream)
    del prog
    torch.cuda.empty_cache()
    return ms


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    which = sys.argv[2].split(",") if len(sys.argv) > 2 else ["se_lin"]
    out = {"batch": n}
    if "se_lin" in which:
        for v in ("0", "28", "14", "7"):
            out["teacher senet50 XEMO_SE_LIN_MIN_HW=" + v] = with_env({"XEMO_SE_LIN_MIN_HW": v}, lambda: teacher_ms(n))
    if "gate" in which:     # (read once per process by the library: pass XEMO_SE_GATE_CLUSTER from the outside)
        v = os.environ.get("XEMO_SE_GATE_CLUSTER", "1") + " XEMO_SE_GATE_K=" + os.environ.get("XEMO_SE_GATE_K", "0")
        out["teacher senet50 XEMO_SE_GATE_CLUSTER=" + v] = teacher_ms(n)
    if "costmodel" in which:
        v = os.environ.get("XEMO_CONV_COSTMODEL", "1")   # read once per process by the library: one value per run
        out["teacher senet50 XEMO_CONV_COSTMODEL=" + v] = teacher_ms(n)
        out["student step XEMO_CONV_COSTMODEL=" + v] = student_ms(n)
    print(json.dumps(out, indent=1))
