"""MVHMR_YCHUNK sweep (consecutive y rows a CTA takes per chunk): fused kernel over pre-packed planes, min of 6.
usage: python scripts/ychunk_sweep.py cfg5:8,cfg5:64,cfg2:8 1,2,3,4,5,6,8,10"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multiviewhmr_b200 import synthetic as syn, aggregation as agg
dev = torch.device('cuda:0')
cases = sys.argv[1].split(',')
ycs = sys.argv[2].split(',')
for case in cases:
    name, B = case.split(':')
    w = syn.CONFIGS[name]
    w = syn.Workload(w.name, int(B), w.V, w.C, w.H, w.W, w.G, w.method, w.dtype, w.joints, w.cuboid_side)
    f, P, cv, c = syn.make_inputs(w)
    fd, Pd, cvd = f.to(dev), P.to(dev), cv.to(dev)
    if w.dtype == 'bf16':
        fd = fd.bfloat16()
    out = torch.empty((w.B, w.C, w.G, w.G, w.G), device=dev)
    packed = agg.pack_features(fd)
    line = []
    for yc in ['auto'] + ycs:
        if yc == 'auto':
            os.environ.pop('MVHMR_YCHUNK', None)
        else:
            os.environ['MVHMR_YCHUNK'] = yc
        fn = lambda: agg.unprojection(fd, Pd, cvd, w.method, out=out, packed=packed)
        for _ in range(2): fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        line.append('%s=%.1f' % (yc, min(ts) * 1e3))
    print(case, ' '.join(line), flush=True)
