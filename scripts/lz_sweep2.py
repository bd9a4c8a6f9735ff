"""lz sweep on a custom shape: python scripts/lz_sweep2.py B V C H W G lz1,lz2,..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multiviewhmr_b200 import synthetic as syn, aggregation as agg
dev = torch.device('cuda:0')
B, V, C, H, W, G = (int(x) for x in sys.argv[1:7])
w = syn.Workload('t', B, V, C, H, W, G)
f, P, cv, c = syn.make_inputs(w)
fd, Pd, cvd = f.to(dev), P.to(dev), cv.to(dev)
out = torch.empty((w.B, w.C, w.G, w.G, w.G), device=dev)
packed = agg.pack_features(fd)
line = []
for lz in ['auto'] + sys.argv[7].split(','):
    if lz == 'auto': os.environ.pop('MVHMR_LZ', None)
    else: os.environ['MVHMR_LZ'] = lz
    fn = lambda: agg.unprojection(fd, Pd, cvd, w.method, out=out, packed=packed)
    for _ in range(2): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    line.append('%s=%.1f' % (lz, min(ts) * 1e3))
print(' '.join(sys.argv[1:7]), ' '.join(line), flush=True)
