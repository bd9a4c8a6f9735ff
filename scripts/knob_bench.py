"""Scheduling knobs (MVHMR_LZ, MVHMR_YCHUNK) on one config: python scripts/knob_bench.py cfg5"""
import os, sys, itertools
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multiviewhmr_b200 import synthetic as syn, aggregation as agg
dev = torch.device('cuda:0')
name = sys.argv[1] if len(sys.argv) > 1 else 'cfg5'
w = syn.CONFIGS[name]
f, P, cv, c = syn.make_inputs(w)
fd, Pd, cvd = f.to(dev), P.to(dev), cv.to(dev)
if w.dtype == 'bf16': fd = fd.bfloat16()
out = torch.empty((w.B, w.C, w.G, w.G, w.G), device=dev)
packed = agg.pack_features(fd)
for lz, yc in itertools.product(sys.argv[2].split(','), sys.argv[3].split(',')):
    os.environ.pop('MVHMR_LZ', None); os.environ.pop('MVHMR_YCHUNK', None)
    if lz != '0': os.environ['MVHMR_LZ'] = lz
    if yc != '0': os.environ['MVHMR_YCHUNK'] = yc
    fn = lambda: agg.unprojection(fd, Pd, cvd, w.method, out=out, packed=packed)
    for _ in range(2): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    print('%s lz=%s ychunk=%s min %.1f us' % (name, lz, yc, min(ts) * 1e3), flush=True)
