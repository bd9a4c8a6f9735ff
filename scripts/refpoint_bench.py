"""The reference's own operating point (cfg/baseline.yaml: 256-ch 7x7 maps, G=16, V=4, B=16) — tiny maps,
tiny grid, many channels — fused path vs the reference torch op sequence on CUDA."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import torch_port
from multiviewhmr_b200 import synthetic as syn, aggregation as agg
dev = torch.device('cuda:0')
for (B, V, C, H, G) in ((16, 4, 256, 7, 16), (16, 4, 256, 14, 32), (16, 4, 32, 56, 32)):
    w = syn.Workload('ref', B=B, V=V, C=C, H=H, W=H, G=G, method='softmax')
    f, P, cv, c = syn.make_inputs(w)
    fd, Pd, cvd = f.to(dev), P.to(dev), cv.to(dev)
    out = torch.empty((B, C, G, G, G), device=dev)
    ref = torch_port.unprojection(fd, Pd, cvd, 'softmax')
    got = agg.unprojection(fd, Pd, cvd, 'softmax', out=out)
    err = float((got - ref).norm() / ref.norm())
    for nm, fn in (('torch ops on CUDA', lambda: torch_port.unprojection(fd, Pd, cvd, 'softmax')),
                   ('fused, python call', lambda: agg.unprojection(fd, Pd, cvd, 'softmax', out=out))):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        print('B%d V%d C%d %dx%d G%d  %-20s min %8.1f us  %7.1f Gvcv/s' % (B, V, C, H, H, G, nm, min(ts) * 1e3, w.vcv / min(ts) / 1e6), flush=True)
    st = torch.cuda.Stream(); gr = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        agg.unprojection(fd, Pd, cvd, 'softmax', out=out)
        with torch.cuda.graph(gr, stream=st):
            for _ in range(10): agg.unprojection(fd, Pd, cvd, 'softmax', out=out)
        gr.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st); gr.replay(); e1.record(st); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / 10
    print('B%d V%d C%d %dx%d G%d  %-20s     %8.1f us  %7.1f Gvcv/s   (rel L2 vs torch CUDA %.1e)' % (B, V, C, H, H, G, 'fused, graph replay', t * 1e3, w.vcv / t / 1e6, err), flush=True)
