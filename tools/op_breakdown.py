"""Per-op CUDA-event timing of one eager distillation step (every xemo_op_* call): where the step goes.
    python tools/op_breakdown.py [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mcncrossmodalemotions_b200 import zoo  # noqa: E402
from mcncrossmodalemotions_b200.distill import DistillationStep  # noqa: E402


class Prof:
    def __init__(self, stream):
        self.stream, self.rec, self.cur = stream, [], None

    def before(self, name, args):
        if not name.startswith("op_"):
            self.cur = None
            return
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        v = [x.value if hasattr(x, "value") else x for x in args]
        self.cur = (name, [a for a in v if isinstance(a, int) and a < (1 << 31)][:16], e0, e1)

    def after(self, name, args):
        if self.cur:
            self.cur[3].record(self.stream)
            self.rec.append(self.cur)
            self.cur = None


B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
step = DistillationStep(zoo.teacher_init("senet50"), zoo.student_init(), B, 300, use_graph=False, overlap=False)
step.grad_step(); step.update(); step.sync()
prof = Prof(step.stream)
step.ctx.profiler = prof
step.grad_step(); step.update()
step.ctx.profiler = None
torch.cuda.synchronize()
tot = 0.0
agg = {}
for name, ints, e0, e1 in prof.rec:
    ms = e0.elapsed_time(e1)
    tot += ms
    a = agg.setdefault(name, [0.0, 0])
    a[0] += ms; a[1] += 1
    extra = ""
    if name in ("op_conv_fwd", "op_conv_dgrad"):
        n, h, w, cin, kout, r, s, sh, sw, pt, pb, pl, pr = ints[:13]
        oh, ow = (h + pt + pb - r) // sh + 1, (w + pl + pr - s) // sw + 1
        fl = 2.0 * n * oh * ow * kout * r * s * cin
        byts = 2.0 * n * (h * w * cin + oh * ow * kout)
        extra = "%4dx%-4d c%-4d k%-4d %dx%d s%d  %7.1f TF  %6.0f GB/s(in+out)" % (h, w, cin, kout, r, s, sh, fl / ms / 1e9, byts / ms / 1e6)
    elif name == "op_conv_wgrad":
        n, h, w, cin, ldy, kout, r, s, sh, sw, pt, pb, pl, pr = ints[:14]
        oh, ow = (h + pt + pb - r) // sh + 1, (w + pl + pr - s) // sw + 1
        fl = 2.0 * n * oh * ow * kout * r * s * cin
        extra = "%4dx%-4d c%-4d k%-4d %dx%d s%d  %7.1f TF" % (h, w, cin, kout, r, s, sh, fl / ms / 1e9)
    print("%-26s %8.3f ms  %s" % (name, ms, extra))
print("---- total %.3f ms (eager, event-timed per op)" % tot)
for k, (ms, c) in sorted(agg.items(), key=lambda x: -x[1][0]):
    print("%-26s %8.3f ms %4d calls %5.1f%%" % (k, ms, c, 100 * ms / tot))
