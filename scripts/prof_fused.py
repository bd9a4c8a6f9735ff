"""One fused unproject+aggregate+soft-argmax call per variant at cfg3 (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multiviewhmr_b200 import synthetic as syn, aggregation as agg
dev = torch.device('cuda:0')
w = syn.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else 'cfg3']
f, P, cv, c = syn.make_inputs(w)
fd, Pd, cvd = f.to(dev), P.to(dev), cv.to(dev)
if w.dtype == 'bf16':
    fd = fd.bfloat16()
packed = agg.pack_features(fd)
out = torch.empty((w.B, w.C, w.G, w.G, w.G), device=dev)
for _ in range(2):
    v = agg.unprojection(fd, Pd, cvd, w.method, packed=packed, out=out)
    agg.soft_argmax_3d(v[:, :17], cvd)
    agg.unprojection_soft_argmax(fd, Pd, cvd, 17, w.method, packed=packed)
    agg.unprojection_soft_argmax(fd, Pd, cvd, 17, w.method, packed=packed, store_volume=False)
torch.cuda.synchronize()
