"""The reference's own torch op sequence (oracle/torch_port.py, a restatement of models/aggregation.py:20-87)
run on CUDA tensors on the B200 — the baseline a user of the reference gets on the same GPU — next to the
fused path.  Measurement script only (not part of bench.py, not part of the product)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import torch_port
from multiviewhmr_b200 import synthetic as syn, aggregation as agg
dev = torch.device('cuda:0')
torch.backends.cuda.matmul.allow_tf32 = False
for name in sys.argv[1:] or ['cfg1', 'cfg2']:
    w = syn.CONFIGS[name]
    f, P, cv, c = syn.make_inputs(w)
    fd, Pd, cvd = f.to(dev), P.to(dev), cv.to(dev)
    res = {}
    for nm, fn in [('torch ops on CUDA (reference path)', lambda: torch_port.unprojection(fd, Pd, cvd, w.method)),
                   ('fused (pack + kernel)', lambda: agg.unprojection(fd, Pd, cvd, w.method))]:
        for _ in range(2): out = fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); out = fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        res[nm] = (min(ts), out)
        print('%s %-36s min %9.1f us  %8.1f Gvcv/s' % (name, nm, min(ts) * 1e3, w.vcv / min(ts) / 1e6), flush=True)
    a, b = res['torch ops on CUDA (reference path)'][1], res['fused (pack + kernel)'][1]
    print('%s rel L2 difference between the two: %.2e' % (name, float((a - b).norm() / a.norm())))
