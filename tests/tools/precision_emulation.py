"""CPU emulation of the device number format (fp16 storage, fp32 accumulate) on the teacher graphs,
against the fp64 oracle: decides whether a single fp16 pass can hold the 1e-3 logit tolerance."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from oracle import nets, mcn_ops as M

def r16(a):
    return a.astype(np.float16).astype(np.float32)

class Emu(nets.TorchOps):
    name = "emu16"
    @classmethod
    def conv(cls, x, f, b=None, dzdy=None, pad=0, stride=1):
        return nets.TorchOps.conv(r16(x), r16(f), b, None, pad, stride)
    @staticmethod
    def relu(x, dzdy=None):
        return r16(M.vl_nnrelu(x))

arch = sys.argv[1] if len(sys.argv) > 1 else "resnet50"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
p32 = nets.teacher_init(arch)
p64 = {k: (v.astype(np.float64) if isinstance(v, np.ndarray) else v) for k, v in p32.items()}
x = nets.synth_faces(n)
t0 = time.time()
taps64, taps16 = {}, {}
y64 = nets.teacher_forward(p64, x.astype(np.float64), nets.TorchOps, taps64)
y32 = nets.teacher_forward(p32, x, nets.TorchOps)
y16 = nets.teacher_forward(p32, x, Emu, taps16)
print("time", time.time() - t0)
rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
print("logits64", y64.reshape(8, n).T)
print("rel fp32 vs fp64", rel(y32, y64), " rel emu16 vs fp64", rel(y16, y64))
for k in taps64:
    print(k, "max|ref| %.3g" % np.abs(taps64[k]).max(), "rel %.3g" % rel(taps16[k], taps64[k]))
