"""Driver for one ncu capture of the secondary kernels: soft-argmax (cfg3) and the softmax backward (cfg2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multiviewhmr_b200 import synthetic as syn, aggregation as agg, autograd as ag
dev = torch.device('cuda:0')
w = syn.CONFIGS['cfg3']
f, P, cv, c = syn.make_inputs(w)
fd, Pd, cvd = f.to(dev).bfloat16(), P.to(dev), cv.to(dev)
vol = agg.unprojection(fd, Pd, cvd, 'softmax')
for _ in range(2):
    agg.soft_argmax_3d(vol[:, :w.joints], cvd)
w2 = syn.CONFIGS['cfg2']
f2, P2, cv2, _ = syn.make_inputs(w2)
f2, P2, cv2 = f2.to(dev), P2.to(dev), cv2.to(dev)
g = torch.randn(w2.B, w2.C, w2.G, w2.G, w2.G, device=dev)
for _ in range(2):
    ag.unprojection_backward(g, f2, P2, cv2, 'softmax')
torch.cuda.synchronize()
print('done')
