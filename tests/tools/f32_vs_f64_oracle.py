"""How far is a single-precision run of the training step from a double-precision one?  CPU only: the oracle
(torch CPU kernels behind the MatConvNet operator semantics) in fp32 against itself in fp64, on the same inputs.
    python tests/tools/f32_vs_f64_oracle.py N W  > profiles/r02_fp32_vs_fp64_oracle.txt
Prints the range-relative distance of the train-mode logits and of every gradient, and how many ReLU masks / pooling
winners differ between the two runs.  The forward agrees to ~5e-6; the gradients differ by 1e-2 ... 1e-1 whenever a
handful of decisions flip (the backward pass is discontinuous in the activations)."""
import sys, numpy as np
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import nets
def rel_err(a,b):
    a=np.asarray(a,np.float64); b=np.asarray(b,np.float64).reshape(a.shape)
    return np.abs(a-b).max()/np.abs(b).max()
n,w=int(sys.argv[1]),int(sys.argv[2])
p=nets.student_randomize_bn(nets.student_init())
spec,tgt=nets.synth_spectrograms(n,w),nets.synth_teacher_logits(n)
p64={k:v.astype(np.float64) for k,v in p.items()}
e64=nets.distillation_student_step(p64,{},spec.astype(np.float64),tgt.astype(np.float64),ops=nets.TorchOps,update=False)
e32=nets.distillation_student_step(dict(p),{},spec,tgt,ops=nets.TorchOps,update=False)
print("pred",rel_err(e32["prediction"],e64["prediction"]))
for k in sorted(e64["grads"]):
    if np.abs(e64["grads"][k]).max()<1e-12: continue
    print(k, "%.2e"%rel_err(e32["grads"][k],e64["grads"][k]))
# decision differences
for i in "1234567":
    a=e32["tape"].get("relu"+i+":x"); b=e64["tape"].get("relu"+i+":x")
    if a is not None: print("relu"+i, int(((a>0)!=(b>0)).sum()), a.size)
for pn in ("pool1","pool2","pool5"):
    a=e32["tape"].get(pn+":argmax"); b=e64["tape"].get(pn+":argmax")
    if a is not None: print(pn, int((a!=b).sum()), a.size)
