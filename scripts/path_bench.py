"""A/B of the two fused-kernel designs (MVHMR_PATH=gather|staged) on the BASELINE configs:
device time of the fused kernel alone over pre-packed planes, and bitwise equality of the results."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multiviewhmr_b200 import synthetic as syn, aggregation as agg

dev = torch.device('cuda:0')
names = sys.argv[1].split(',') if len(sys.argv) > 1 else ['cfg1', 'cfg2', 'cfg3', 'cfg5']
paths = sys.argv[2].split(',') if len(sys.argv) > 2 else ['gather', 'staged']
for name in names:
    w = syn.CONFIGS[name]
    if w.B > 8:
        w = syn.Workload(w.name, 8, w.V, w.C, w.H, w.W, w.G, w.method, w.dtype, w.joints, w.cuboid_side)
    f, P, cv, c = syn.make_inputs(w)
    fd, Pd, cvd = f.to(dev), P.to(dev), cv.to(dev)
    if w.dtype == 'bf16':
        fd = fd.bfloat16()
    res = {}
    for path in paths:
        os.environ['MVHMR_PATH'] = path
        out = torch.empty((w.B, w.C, w.G, w.G, w.G), device=dev)
        packed = agg.pack_features(fd)
        fn = lambda: agg.unprojection(fd, Pd, cvd, w.method, out=out, packed=packed)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        res[path] = out
        ts.sort()
        print('%s B%d %-7s min %.1f us  median %.1f us  (%.0f Gvcv/s)' % (name, w.B, path, ts[0] * 1e3, ts[5] * 1e3,
              w.vcv / ts[0] / 1e6), flush=True)
    if len(res) == 2:
        a, b = res[paths[0]], res[paths[1]]
        print('   bitwise equal:', bool(torch.equal(a, b)), ' max abs diff %.3g' % float((a - b).abs().max()), flush=True)
