"""Where does a cfg5 step spend its time? events around pack and the fused kernel, host time per call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multiviewhmr_b200 import synthetic as syn, aggregation as agg
dev = torch.device('cuda:0')
w = syn.CONFIGS['cfg5']
f, P, cv, c = syn.make_inputs(w)
fd, Pd, cvd = f.to(dev), P.to(dev), cv.to(dev)
outs = [torch.empty((w.B, w.C, w.G, w.G, w.G), device=dev) for _ in range(2)]
for i in range(3):
    packed = agg.pack_features(fd); agg.unprojection(fd, Pd, cvd, w.method, out=outs[i % 2], packed=packed)
torch.cuda.synchronize()
ev = []
t0 = time.perf_counter()
for i in range(10):
    a, b, c2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    h0 = time.perf_counter()
    a.record(); packed = agg.pack_features(fd); b.record()
    h1 = time.perf_counter()
    agg.unprojection(fd, Pd, cvd, w.method, out=outs[i % 2], packed=packed); c2.record()
    h2 = time.perf_counter()
    ev.append((a, b, c2, h1 - h0, h2 - h1))
torch.cuda.synchronize()
t1 = time.perf_counter()
for a, b, c2, hp, hk in ev:
    print('pack %.3f ms  kernel %.3f ms   host: pack call %.3f ms, unprojection call %.3f ms' % (a.elapsed_time(b), b.elapsed_time(c2), hp * 1e3, hk * 1e3))
print('wall per step %.3f ms' % ((t1 - t0) / 10 * 1e3))
