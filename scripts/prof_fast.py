"""ncu driver: the texture-path kernels at cfg3 (bf16 maps)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multiviewhmr_b200 import synthetic as syn, aggregation as agg
dev = torch.device('cuda:0')
w = syn.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else 'cfg3']
if w.B > 8:
    w = syn.Workload(w.name, 8, w.V, w.C, w.H, w.W, w.G, w.method, w.dtype, w.joints, w.cuboid_side)
f, P, cv, c = syn.make_inputs(w)
fd, Pd, cvd = f.to(dev).bfloat16(), P.to(dev), cv.to(dev)
out = torch.empty((w.B, w.C, w.G, w.G, w.G), device=dev)
for _ in range(3):
    agg.unprojection(fd, Pd, cvd, w.method, out=out, precision='fast')
torch.cuda.synchronize()
print('done', float(out.sum()))
