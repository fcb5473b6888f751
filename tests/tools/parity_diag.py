"""Where does the fp32-equivalent student step (parity.StudentProgramF32) differ from the fp64 oracle?
    python tests/tools/parity_diag.py [N] [W]
Prints (a) every tape activation of the forward, (b) every gradient, (c) each backward operator run in isolation on
the ORACLE's own inputs (so that its error is not inherited from upstream)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import nets  # noqa: E402
from mcncrossmodalemotions_b200 import vl_nn  # noqa: E402
from mcncrossmodalemotions_b200.parity import StudentProgramF32  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64).reshape(np.shape(a))
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


n = int(sys.argv[1]) if len(sys.argv) > 1 else 4
w = int(sys.argv[2]) if len(sys.argv) > 2 else 100
p = nets.student_randomize_bn(nets.student_init())
spec, tgt = nets.synth_spectrograms(n, w), nets.synth_teacher_logits(n)
p64 = {k: v.astype(np.float64) for k, v in p.items()}
ex = nets.distillation_student_step(p64, {}, spec.astype(np.float64), tgt.astype(np.float64), ops=nets.TorchOps, update=False)
tape64 = ex["tape"]
prog = StudentProgramF32(p, n, w)
prog.set_input(spec, tgt)
prog._forward(True, keep=True)
names = {"conv": ":x", "bn": ":x"}
print("== forward tape (input of each block) vs oracle")
for L in prog.layers:
    nm, i = L["name"], L["name"][-1]
    pairs = [(nm + ":x", nm + ":x")]
    if L["bn"]:
        pairs += [("bn" + i + ":x", "bn" + i + ":x"), ("bn" + i + ":relu", "relu" + i + ":x")]
    if L["pool"]:
        pairs += [(nm + ":pool", "pool" + i + ":x")]
    for mine, theirs in pairs:
        a, b = vl_nn.gather(prog.tape[mine]), tape64[theirs]
        extra = ""
        if mine.endswith(":relu"):
            extra = "  mask flips %d / %d" % (int(((a > 0) != (b > 0)).sum()), a.size)
        print("  %-12s %.2e%s" % (mine, rel(a, b), extra))
print("  prediction   %.2e" % rel(vl_nn.gather(prog.pred), ex["prediction"]))
dz = prog._loss(True)
prog._backward(dz)
g = prog.export_grads()
print("== gradients vs oracle")
for k in sorted(g):
    ref = np.asarray(ex["grads"][k])
    if np.abs(ref).max() < 1e-12:
        print("  %-8s (exact zero) max|ours| %.2e" % (k, np.abs(g[k]).max()))
        continue
    print("  %-8s %.2e" % (k, rel(g[k], ref)))
print("== backward operators in isolation, on the oracle's inputs (fp32 copies)")
f32 = lambda a: np.asarray(a, np.float32)
ctx = vl_nn.default_context()
# re-run the oracle backward keeping the intermediate gradients
cur = nets.M.vl_nnsoftmaxceloss(ex["prediction"], tgt.astype(np.float64), np.array(1.0), temperature=2.0, logitTargets=True)
for name, fh, fw, cin, cout, stride, pad, has_bn in reversed(nets.STUDENT_CONVS):
    i = name[-1]
    if name in nets.STUDENT_POOLS:
        pname, method, _, pstride = nets.STUDENT_POOLS[name]
        ref = nets.TorchOps.pool(tape64[pname + ":x"], tape64[pname + ":win"], cur, pad=0, stride=pstride, method=method)
        got = vl_nn.vl_nnpool(f32(tape64[pname + ":x"]), tape64[pname + ":win"], f32(cur), stride=pstride, method=method)
        print("  %-6s bwd dx %.2e" % (pname, rel(got, ref)))
        cur = ref
    if has_bn:
        bn = "bn" + i
        ref = nets.TorchOps.relu(tape64["relu" + i + ":x"], cur)
        got = vl_nn.vl_nnrelu(f32(tape64["relu" + i + ":x"]), f32(cur))
        print("  relu%s  bwd dx %.2e" % (i, rel(got, ref)))
        cur = ref
        rdx, rdg, rdb, _ = nets.TorchOps.bnorm(tape64[bn + ":x"], p64[bn + "m"], p64[bn + "b"], cur, epsilon=nets.BN_EPS)
        gdx, gdg, gdb, _ = vl_nn.vl_nnbnorm(f32(tape64[bn + ":x"]), p[bn + "m"], p[bn + "b"], f32(cur), epsilon=nets.BN_EPS)
        print("  %-6s bwd dx %.2e dg %.2e db %.2e   (|db|max %.2e, sum|dy|max %.2e)" % (
            bn, rel(gdx, rdx), rel(gdg, rdg), rel(gdb, rdb), np.abs(rdb).max(), np.abs(cur).sum(axis=(0, 1, 3)).max()))
        cur = rdx
    rdx, rdf, rdb = nets.TorchOps.conv(tape64[name + ":x"], p64[name + "f"], p64[name + "b"], cur, pad=pad, stride=stride)
    ctx.set_conv_precision(1)
    gdx, gdf, gdb = vl_nn.vl_nnconv(f32(tape64[name + ":x"]), p[name + "f"], p[name + "b"], f32(cur), pad=pad, stride=stride)
    ctx.set_conv_precision(0)
    print("  %-6s bwd dx %.2e df %.2e" % (name, rel(gdx, rdx), rel(gdf, rdf)))
    cur = rdx
