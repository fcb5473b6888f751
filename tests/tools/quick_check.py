"""Ten-second GPU sanity check (fits a nearly spent gpurun budget): (1) test-mode student logits of the same clips must be
bit-identical at batch 256 (conv2 planned with 256-wide N tiles) and batch 8 (128-wide) -- the tile shape must not change
any dot product; (2) lossType 'softmaxlog': objective equals -sum log softmax(pred)[label] of the step's own logits."""
import os
import sys
import time

import numpy as np

t0 = time.time()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch  # noqa: E402

from mcncrossmodalemotions_b200 import zoo  # noqa: E402
from mcncrossmodalemotions_b200.programs import StudentProgram  # noqa: E402

p = zoo.student_init()
rng = np.random.default_rng(0)
spec = rng.standard_normal((512, 300, 1, 256)).astype(np.float32)
big = StudentProgram(p, 256, 300).forward(spec, "test")
print("big done %.1fs" % (time.time() - t0), flush=True)
small = StudentProgram(p, 8, 300).forward(spec[..., :8], "test")
print("BITEXACT" if np.array_equal(big[:8], small) else "DIFF max %.3g" % np.abs(big[:8] - small).max(), flush=True)
n = 8
labels = np.array([1, 1, 1, 2, 2, 3, 5, 8]).reshape(1, 1, 1, n)
prog = StudentProgram(p, n, 100, use_graph=False, loss_type="softmaxlog")
prog.reset_metrics()
prog.set_input(spec[:, :100, :, :n].copy(), labels)
prog.grad_step()
m = prog.metrics()
with torch.cuda.stream(prog.stream):
    pred = prog.a["pred32"][:, :8].cpu().numpy().astype(np.float64)
prog.sync()
lse = np.log(np.exp(pred - pred.max(1, keepdims=True)).sum(1)) + pred.max(1)
obj = float((lse - pred[np.arange(n), labels.ravel() - 1]).sum())
print("softmaxlog objective %.6f vs %.6f  rel %.2e  count %s  %.1fs" % (m["objective"], obj, abs(m["objective"] - obj) / obj, m["count"], time.time() - t0), flush=True)
