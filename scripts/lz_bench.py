"""z-segment length (MVHMR_LZ) for grids whose z extent is not a multiple of 32."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multiviewhmr_b200 import synthetic as syn, aggregation as agg
dev = torch.device('cuda:0')
for G, lzs in ((40, ('0', '20', '32')), (48, ('0', '24', '32', '16')), (80, ('0', '27', '32', '20', '16')), (56, ('0', '28', '32'))):
    w = syn.Workload('t', B=4, V=4, C=32, H=96, W=96, G=G, method='softmax')
    f, P, cv, c = syn.make_inputs(w)
    fd, Pd, cvd = f.to(dev), P.to(dev), cv.to(dev)
    out = torch.empty((w.B, w.C, G, G, G), device=dev)
    packed = agg.pack_features(fd)
    for lz in lzs:
        os.environ['MVHMR_LZ'] = lz
        fn = lambda: agg.unprojection(fd, Pd, cvd, w.method, out=out, packed=packed)
        for _ in range(3): fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        print('G=%d lz=%s min %.1f us  %.0f Gvcv/s' % (G, lz, min(ts) * 1e3, w.vcv / min(ts) / 1e6), flush=True)
