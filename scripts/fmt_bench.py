"""Output formats of the fused kernel (pre-packed planes): default NCDHW, channels_last_3d, fused max_pool3d(2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multiviewhmr_b200 import synthetic as syn, aggregation as agg
dev = torch.device('cuda:0')
for name in sys.argv[1:] or ['cfg2', 'cfg4', 'cfg5']:
    w = syn.CONFIGS[name]
    if w.B > 16:
        w = syn.Workload(w.name, 16, w.V, w.C, w.H, w.W, w.G, w.method, w.dtype, w.joints, w.cuboid_side)
    f, P, cv, c = syn.make_inputs(w)
    fd, Pd, cvd = f.to(dev), P.to(dev), cv.to(dev)
    if w.dtype == 'bf16':
        fd = fd.bfloat16()
    packed = agg.pack_features(fd)
    for fmt in ['ncdhw', 'channels_last_3d', 'max_pool2']:
        out = agg.unprojection(fd, Pd, cvd, w.method, packed=packed, output=fmt)
        fn = lambda: agg.unprojection(fd, Pd, cvd, w.method, packed=packed, out=out, output=fmt)
        for _ in range(3): fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        print('%s B%d %-17s min %.1f us (%.0f Gvcv/s), output %.1f MB' % (name, w.B, fmt, min(ts) * 1e3, w.vcv / min(ts) / 1e6, out.numel() * 4 / 1e6), flush=True)
    # what the consumer pays without the fused pool: torch max_pool3d over the full volume
    full = agg.unprojection(fd, Pd, cvd, w.method, packed=packed)
    for _ in range(2): torch.nn.functional.max_pool3d(full, 2)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); torch.nn.functional.max_pool3d(full, 2); e1.record(); torch.cuda.synchronize()
    print('%s torch max_pool3d(2) of the full volume: %.1f us' % (name, e0.elapsed_time(e1) * 1e3), flush=True)
