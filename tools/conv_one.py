"""One convolution of the step in isolation (for `ncu --set full` captures and A/B timings):
    python tools/conv_one.py conv2|conv3|res4 [batch]
Runs the forward convolution three times (the third launch is the one to capture: ncu -k regex:conv_fprop --launch-skip 2 -c 1)
and prints its CUDA-event time.  XEMO_CONV_2CTA=0/1 selects single CTAs / CTA pairs."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcncrossmodalemotions_b200 import _lib  # noqa: E402

SHAPES = {   # H, W, Cin, Kout, R, S, stride, pad
    "conv2": (126, 73, 128, 256, 5, 5, 2, 1),
    "conv3": (30, 17, 256, 384, 3, 3, 1, 1),
    "res4": (14, 14, 256, 256, 3, 3, 1, 1),
    "res5": (7, 7, 512, 512, 3, 3, 1, 1),
}
name = sys.argv[1] if len(sys.argv) > 1 else "conv2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 256
h, w, cin, kout, r, s, st, pd = SHAPES[name]
oh, ow = (h + 2 * pd - r) // st + 1, (w + 2 * pd - s) // st + 1
stream = torch.cuda.Stream()
ctx = _lib.Context(0, stream.cuda_stream)
with torch.cuda.stream(stream):
    x = torch.randn(n, h, w, cin, device="cuda").half()
    wt = (torch.randn(kout, r, s, cin, device="cuda") / (r * s * cin) ** 0.5).half()
    y = torch.empty(n, oh, ow, kout, device="cuda", dtype=torch.float16)
    bias = torch.zeros(kout, device="cuda")
torch.cuda.synchronize()
vp = lambda t: C.c_void_p(t.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(3):
    if i == 2:
        e0.record(stream)
    ctx.op_conv_fwd(vp(x), n, h, w, cin, vp(wt), kout, r, s, st, st, pd, pd, pd, pd, None, vp(bias), None, 0, vp(y), None, kout)
e1.record(stream)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
fl = 2.0 * n * oh * ow * kout * r * s * cin
print("%s n=%d XEMO_CONV_2CTA=%s: %.3f ms, %.0f TFLOP/s" % (name, n, os.environ.get("XEMO_CONV_2CTA", "1"), ms, fl / ms / 1e9))
