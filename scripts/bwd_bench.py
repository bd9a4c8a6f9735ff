"""Time the backward kernels (gradient w.r.t. the feature maps): fast path vs the simple kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multiviewhmr_b200 import synthetic as syn, autograd as ag
dev = torch.device('cuda:0')
for name in sys.argv[1:] or ['cfg2']:
    w = syn.CONFIGS[name]
    f, P, cv, c = syn.make_inputs(w)
    fd, Pd, cvd = f.to(dev), P.to(dev), cv.to(dev)
    if w.dtype == 'bf16': fd = fd.bfloat16()
    g = torch.randn(w.B, w.C, w.G, w.G, w.G, device=dev)
    for m in ['sum', 'max', 'softmax']:
        for simple in (True, False):
            fn = lambda: ag.unprojection_backward(g, fd, Pd, cvd, m, simple=simple)
            for _ in range(2): r = fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); r = fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
            print(name, m, 'simple' if simple else 'fast  ', 'min %.1f us' % (min(ts) * 1e3), 'checksum %.6e' % float(r.double().abs().sum()), flush=True)
