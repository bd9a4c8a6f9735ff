"""precision='fast' (texture units, fp16 maps) against the exact path: device time incl. the layout pass, deviation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multiviewhmr_b200 import synthetic as syn, aggregation as agg
dev = torch.device('cuda:0')
for name in sys.argv[1:] or ['cfg3', 'cfg2', 'cfg5']:
    w = syn.CONFIGS[name]
    if w.B > 8:
        w = syn.Workload(w.name, 8, w.V, w.C, w.H, w.W, w.G, w.method, w.dtype, w.joints, w.cuboid_side)
    f, P, cv, c = syn.make_inputs(w)
    fd, Pd, cvd = f.to(dev).bfloat16(), P.to(dev), cv.to(dev)
    out = torch.empty((w.B, w.C, w.G, w.G, w.G), device=dev)
    res = {}
    for prec in ['exact', 'fast']:
        fn = lambda: agg.unprojection(fd, Pd, cvd, w.method, out=out, precision=prec)
        for _ in range(3): fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        res[prec] = out.clone()
        print('%s B%d bf16 maps, %-5s: min %.1f us (layout pass + fused kernel), %.0f Gvcv/s' % (name, w.B, prec, min(ts) * 1e3, w.vcv / min(ts) / 1e6), flush=True)
    d = (res['fast'].double() - res['exact'].double())
    print('   deviation fast vs exact: rel l2 %.3e, max abs %.3e' % (float(d.norm() / res['exact'].double().norm()), float(d.abs().max())), flush=True)
