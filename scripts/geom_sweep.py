"""Per-geometry timing: cfg5-shaped batches of 8 samples starting at sample b_offset (the synthetic cameras' yaw grows by
0.05 rad per sample), fused kernel over pre-packed planes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multiviewhmr_b200 import synthetic as syn, aggregation as agg
dev = torch.device('cuda:0')
name = sys.argv[1] if len(sys.argv) > 1 else 'cfg5'
w0 = syn.CONFIGS[name]
w = syn.Workload(w0.name, 8, w0.V, w0.C, w0.H, w0.W, w0.G, w0.method, w0.dtype, w0.joints, w0.cuboid_side)
line = []
for off in (0, 8, 16, 24, 32, 40, 48, 56):
    f, P, cv, c = syn.make_inputs(w, seed=1234 + off, b_offset=off)
    fd, Pd, cvd = f.to(dev), P.to(dev), cv.to(dev)
    out = torch.empty((w.B, w.C, w.G, w.G, w.G), device=dev)
    packed = agg.pack_features(fd)
    fn = lambda: agg.unprojection(fd, Pd, cvd, w.method, out=out, packed=packed)
    for _ in range(2): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    line.append('%d:%.0f' % (off, min(ts) * 1e3))
print(name, os.environ.get('MVHMR_SMEM_PAD', '0'), ' '.join(line), flush=True)
