"""Fused unproject + aggregate + soft-argmax kernel against the two-kernel path (pre-packed planes, CUDA events,
L2 flushed by the 268 MB output between iterations): cfg3 (bf16 maps) and cfg2 (fp32 maps), 17 joints."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multiviewhmr_b200 import synthetic as syn, aggregation as agg
dev = torch.device('cuda:0')
J = 17
for name in sys.argv[1:] or ['cfg3', 'cfg2']:
    w = syn.CONFIGS[name]
    f, P, cv, c = syn.make_inputs(w)
    fd, Pd, cvd = f.to(dev), P.to(dev), cv.to(dev)
    if w.dtype == 'bf16':
        fd = fd.bfloat16()
    packed = agg.pack_features(fd)
    out = torch.empty((w.B, w.C, w.G, w.G, w.G), device=dev)

    def two():
        v = agg.unprojection(fd, Pd, cvd, w.method, packed=packed, out=out)
        return agg.soft_argmax_3d(v[:, :J], cvd)
    fns = {'aggregate only': lambda: agg.unprojection(fd, Pd, cvd, w.method, packed=packed, out=out),
           'two kernels (aggregate + soft-argmax)': two,
           'fused, volume stored': lambda: agg.unprojection_soft_argmax(fd, Pd, cvd, J, w.method, packed=packed)[1],
           'fused, no volume': lambda: agg.unprojection_soft_argmax(fd, Pd, cvd, J, w.method, packed=packed, store_volume=False)[1]}
    res = {}
    for label, fn in fns.items():
        for _ in range(3): r = fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(12):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r = fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        res[label] = r
        ts.sort()
        print('%s %-40s min %.1f us  median %.1f us' % (name, label, ts[0] * 1e3, ts[len(ts) // 2] * 1e3), flush=True)
    d = (res['fused, volume stored'] - res['two kernels (aggregate + soft-argmax)']).abs().max().item()
    print('   max |fused - two-kernel| joints: %.3e mm (max |coord| %.0f mm)' % (d, cv.abs().max().item()), flush=True)
