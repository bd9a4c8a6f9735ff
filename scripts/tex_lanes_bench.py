"""Lane -> voxel mappings of the texture-path kernel (MVHMR_TEX_LANES, lane bit 0 first): device time of
layout pass + kernel, min of 10, bf16 maps."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multiviewhmr_b200 import synthetic as syn, aggregation as agg
dev = torch.device('cuda:0')
specs = sys.argv[1].split(',')
ref = {}
for name in ['cfg3', 'cfg5']:
    w = syn.CONFIGS[name]
    if w.B > 8:
        w = syn.Workload(w.name, 8, w.V, w.C, w.H, w.W, w.G, w.method, w.dtype, w.joints, w.cuboid_side)
    f, P, cv, c = syn.make_inputs(w)
    fd, Pd, cvd = f.to(dev).bfloat16(), P.to(dev), cv.to(dev)
    out = torch.empty((w.B, w.C, w.G, w.G, w.G), device=dev)
    for spec in specs:
        os.environ['MVHMR_TEX_LANES'] = spec
        fn = lambda: agg.unprojection(fd, Pd, cvd, w.method, out=out, precision='fast')
        for _ in range(3): fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        same = True
        if name in ref: same = bool(torch.equal(ref[name], out))
        else: ref[name] = out.clone()
        print('%s B%d %s: min %.1f us  (same bits as the first mapping: %s)' % (name, w.B, spec, min(ts) * 1e3, same), flush=True)
