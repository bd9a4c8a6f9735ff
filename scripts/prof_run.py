"""Tiny driver for ncu captures: one workload, a few launches of the hot path."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multiviewhmr_b200 import synthetic as syn, aggregation as agg

name = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
w = syn.CONFIGS[name]
if len(sys.argv) > 3:      # optional batch-size override (keeps ncu replays short)
    w = syn.Workload(w.name, int(sys.argv[3]), w.V, w.C, w.H, w.W, w.G, w.method, w.dtype, w.joints, w.cuboid_side)
dev = torch.device('cuda:0')
f, P, cv, c = syn.make_inputs(w)
fd, Pd, cvd = f.to(dev), P.to(dev), cv.to(dev)
if w.dtype == 'bf16':
    fd = fd.bfloat16()
out = torch.empty((w.B, w.C, w.G, w.G, w.G), device=dev)
for _ in range(iters):
    agg.unprojection(fd, Pd, cvd, w.method, out=out)
    if w.joints:
        agg.soft_argmax_3d(out[:, :w.joints], cvd)      # channel slice read in place
torch.cuda.synchronize()
print('done', float(out.sum()))
