"""pack_kernel alone at several batch sizes (NCHW fp32 -> pixel-major planes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multiviewhmr_b200 import aggregation as agg
dev = torch.device('cuda:0')
for (B, V, C, H, W) in [(8, 4, 32, 96, 96), (8, 8, 32, 96, 96), (16, 8, 32, 96, 96), (32, 8, 32, 96, 96), (64, 8, 32, 96, 96), (16, 8, 64, 128, 128)]:
    f = torch.randn(B, V, C, H, W, device=dev)
    for _ in range(3): p = agg.pack_features(f)
    torch.cuda.synchronize()
    ts = []
    for _ in range(8):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); p = agg.pack_features(f); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    mb = f.numel() * 4 / 1e6
    print('B%d V%d C%d %dx%d: %.1f MB in, pack min %.1f us median %.1f us -> %.2f TB/s (read+write)' % (
        B, V, C, H, W, mb, min(ts) * 1e3, sorted(ts)[4] * 1e3, (mb + p.numel() / 1e6) / min(ts) / 1e3), flush=True)
    del f, p
