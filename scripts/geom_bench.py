"""Time the streaming geometry kernels (coord volume build, rotation, projection) at cfg2 size."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multiviewhmr_b200 import aggregation as agg, volumetric, multiview, synthetic as syn
dev = torch.device('cuda:0')
B, G = 8, 64
centers = np.zeros((B, 3), np.float32); rots = np.stack([np.eye(3, dtype=np.float32)] * B)
cv = agg.build_coord_volumes(centers, rots, G, 2500.0, dev)
P = torch.from_numpy(syn.ring_projection(0, 1, 4, 96, 96)).float().to(dev)
pts = cv.reshape(-1, 3)
def timeit(fn, name, nbytes):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    t = min(ts)
    print('%-28s min %.1f us (includes host launch path)  %.0f GB/s' % (name, t * 1e3, nbytes / t / 1e6))
n = B * G ** 3
timeit(lambda: agg.build_coord_volumes(centers, rots, G, 2500.0, dev), 'build_coord_volumes', n * 12)
timeit(lambda: volumetric.rotate_coord_volume(cv, 0.3, [0, 0, 1]), 'rotate_coord_volume', n * 24)
timeit(lambda: multiview.project_3d_points_to_image_plane_without_distortion(P, pts), 'project (euclid)', n * 20)
