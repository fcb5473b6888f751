"""Two output-bound convolution launches in isolation, for `ncu --set full --import-source on` (epilogue study):
the student conv1 in pixel-pair form and the teacher's 56 x 56 c64 -> k256 1x1 expand.  python tools/prof_epilogue.py"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcncrossmodalemotions_b200 import _lib  # noqa: E402

stream = torch.cuda.Stream()
ctx = _lib.Context(0, stream.cuda_stream)
p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
N = 256
with torch.cuda.stream(stream):
    x1 = torch.randn(N, 257, 74, 32, device="cuda").half()
    w1 = (torch.randn(192, 4, 1, 32, device="cuda") * 0.1).half()
    s1 = torch.randn(192, device="cuda")
    y1 = torch.zeros(N, 254, 74, 192, device="cuda", dtype=torch.float16)
    x2 = torch.randn(N, 56, 56, 64, device="cuda").half()
    w2 = (torch.randn(256, 1, 1, 64, device="cuda") * 0.1).half()
    a2, b2 = torch.rand(256, device="cuda") + 0.5, torch.randn(256, device="cuda")
    y2 = torch.zeros(N, 56, 56, 256, device="cuda", dtype=torch.float16)
    for _ in range(2):
        ctx.op_conv_fwd(p(x1), N, 257, 74, 32, p(w1), 192, 4, 1, 1, 1, 0, 0, 0, 0, None, p(s1), None, 0, p(y1), None, 0)
        ctx.op_conv_fwd(p(x2), N, 56, 56, 64, p(w2), 256, 1, 1, 1, 1, 0, 0, 0, 0, p(a2), p(b2), None, 1, p(y2), None, 0)
ctx.sync()
print("ok")
