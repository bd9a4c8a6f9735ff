"""Time the 3-D soft-argmax (partials + finalize) on cfg3's aggregate shape; check against float64 torch."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multiviewhmr_b200 import synthetic as syn, aggregation as agg

dev = torch.device('cuda:0')
w = syn.CONFIGS['cfg3']
N = w.G ** 3
g = torch.Generator().manual_seed(3)
vols = [(torch.randn(w.B, w.joints, w.G, w.G, w.G, generator=g) * 3).to(dev) for _ in range(3)]   # 3 x 142 MB > L2
cv = syn.make_coord_volumes(torch.randn(w.B, 3, generator=g) * 100, w.G, w.cuboid_side, 0.3).to(dev)
ref = (torch.softmax(vols[0].double().view(w.B, w.joints, N), 2) @ cv.double().view(w.B, N, 3))
got = agg.soft_argmax_3d(vols[0], cv)
print('max abs err', (got.double() - ref).abs().max().item(), 'max|coord|', cv.abs().max().item())
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
import ctypes
import numpy as np
from multiviewhmr_b200 import _lib
L = _lib.load()
gbuf = torch.cat([torch.zeros(w.B * 3), torch.eye(3).repeat(w.B, 1, 1).reshape(-1)]).to(dev)     # centres, rotations
grid = _lib.Grid(); grid.centers = gbuf.data_ptr(); grid.rot = gbuf.data_ptr() + w.B * 12
for k in range(3):
    grid.pos[k] = float(np.float32(-1250.0)); grid.step[k] = float(np.float32(2500.0 / (w.G - 1)))
sa_out = torch.empty(w.B, w.joints, 3, device=dev)
sa_ws = torch.empty(L.mvhmr_soft_argmax3d_workspace_bytes(w.B, w.joints, N), dtype=torch.uint8, device=dev)
def grid_call(v):      # raw ABI: the python wrapper stages the descriptor through pinned memory, which a graph capture cannot contain
    _lib.check(L.mvhmr_soft_argmax3d_grid(_lib.ptr(v), ctypes.byref(grid), _lib.ptr(sa_out), w.B, w.joints, w.G, w.G, w.G, w.joints * N,
                                          _lib.ptr(sa_ws), sa_ws.numel(), _lib.stream_ptr(dev)))
    return sa_out
for fn, name in [(lambda v: agg.soft_argmax_3d_records(v, cv), 'partials'), (lambda v: agg.soft_argmax_3d(v, cv), 'total'),
                 (grid_call, 'total, grid generated in-kernel')]:
  for ncall in (1, 6, 24):
    for i in range(6): fn(vols[i % 3])
    torch.cuda.synchronize()
    st = torch.cuda.Stream()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.stream(st):
        with torch.cuda.graph(gr, stream=st):
            keep = [fn(vols[i % 3]) for i in range(ncall)]   # host launch cost out of the picture
        gr.replay()
        torch.cuda.synchronize()
        ts, clk = [], []
        for i in range(20):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st); gr.replay(); e1.record(st)
            clk.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) / ncall)
    ts.sort()
    t = ts[len(ts) // 2]
    print(f'{name} x{ncall}: min {ts[0]*1e3:.1f} us med {t*1e3:.1f} us  {w.soft_argmax_bytes()/t/1e6:.0f} GB/s  roofline {w.soft_argmax_bytes()/t/1e6/6543.1*100:.1f}%  sm clk {sorted(clk)[len(clk)//2]} MHz')
