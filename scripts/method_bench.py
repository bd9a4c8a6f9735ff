import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multiviewhmr_b200 import synthetic as syn, aggregation as agg
dev = torch.device('cuda:0')
for name in ['cfg2', 'cfg3']:
    w = syn.CONFIGS[name]
    f, P, cv, c = syn.make_inputs(w)
    fd, Pd, cvd = f.to(dev), P.to(dev), cv.to(dev)
    if w.dtype == 'bf16': fd = fd.bfloat16()
    out = torch.empty((w.B, w.C, w.G, w.G, w.G), device=dev)
    packed = agg.pack_features(fd)
    for m in ['sum', 'mean', 'max', 'softmax']:
        fn = lambda: agg.unprojection(fd, Pd, cvd, m, out=out, packed=packed)
        for _ in range(3): fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        print(name, m, 'min %.1f us' % (min(ts) * 1e3), flush=True)
    # channels-last maps gathered in place (no pack pass) vs NCHW input (pack + kernel)
    fcl = fd.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)
    assert agg._is_channels_last(fcl)
    for nm, fn in [('nchw (pack+kernel)', lambda: agg.unprojection(fd, Pd, cvd, 'softmax', out=out)),
                   ('channels-last in place', lambda: agg.unprojection(fcl, Pd, cvd, 'softmax', out=out))]:
        for _ in range(3): fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        print(name, 'softmax', nm, 'min %.1f us' % (min(ts) * 1e3), flush=True)
