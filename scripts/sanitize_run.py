"""Small end-to-end pass for compute-sanitizer: every kernel of the library once, ragged shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multiviewhmr_b200 import synthetic as syn, aggregation as agg, volumetric, multiview
dev = torch.device('cuda:0')
for (B, V, C, H, W, G, dt) in [(1, 4, 32, 24, 24, 12, 'fp32'), (2, 3, 5, 9, 13, 7, 'fp32'), (1, 8, 16, 16, 16, 10, 'bf16'), (1, 9, 8, 8, 8, 6, 'fp32')]:
    w = syn.Workload('s', B, V, C, H, W, G, dtype=dt)
    f, P, cv, c = syn.make_inputs(w, seed=1, theta=0.2, behind_views=(1,))
    fd = f.to(dev).bfloat16() if dt == 'bf16' else f.to(dev)
    for m in ('sum', 'mean', 'max', 'softmax'):
        out = agg.unprojection(fd, P.to(dev), cv.to(dev), m)
    rots = np.stack([np.eye(3, dtype=np.float32)] * B)
    agg.unprojection_grid(fd, P.to(dev), c.numpy(), rots, G, 2500.0, 'softmax')
    agg.soft_argmax_3d(out, cv.to(dev))
    fd2 = f.to(dev).requires_grad_(True)
    agg.unprojection(fd2, P.to(dev), cv.to(dev), 'softmax').sum().backward()
    volumetric.rotate_coord_volume(cv.to(dev), 0.3, [0, 0, 1])
    multiview.project_3d_points_to_image_plane_without_distortion(P[0, 0].to(dev), cv[0].reshape(-1, 3).to(dev))
torch.cuda.synchronize()
print('sanitize run done')
