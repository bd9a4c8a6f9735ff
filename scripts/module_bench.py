"""VolumeGenerator.forward at the BASELINE cfg2 operating point: wall clock per call (host prologue +
launches, GPU kept busy) next to the device time of the same call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multiviewhmr_b200 import aggregation as agg, multiview
dev = torch.device('cuda:0')
B, V, Cin, H, G = 8, 4, 256, 96, 64
rng = np.random.default_rng(0)
cams = [[multiview.Camera(np.eye(3), rng.normal(size=3) * 50 + np.array([0, 0, 4500.0]),
                          [[1100.0, 0, 192], [0, 1100.0, 192], [0, 0, 1]]) for _ in range(B)] for _ in range(V)]
batch = {'images': np.zeros((B, V, 384, 384, 3), np.float32), 'cameras': cams,
         'keypoints_3d': [rng.normal(size=(17, 3)) * 50 for _ in range(B)]}
vg = agg.VolumeGenerator(volume_size=G, input_channels=Cin, output_channels=32, device=dev).eval()
feats = torch.randn(B, V, Cin, H, H, device=dev)
proj = torch.zeros(B, V, 3, 4, device=dev)
for cl, tf in ((False, False), (True, False), (False, True), (True, True)):
    vg.channels_last = cl
    torch.backends.cuda.matmul.allow_tf32 = tf; torch.backends.cudnn.allow_tf32 = tf
    with torch.no_grad():
        for _ in range(3): vg(feats, proj, batch)
        torch.cuda.synchronize()
        n = 20
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        for _ in range(n): out = vg(feats, proj, batch)
        e1.record(); t_host = (time.perf_counter() - t0) / n
        torch.cuda.synchronize(); t_wall = (time.perf_counter() - t0) / n
    print('tf32=%s channels_last=%s: host %.0f us per call to enqueue, %.0f us wall per call, %.0f us device per call'
          % (tf, cl, t_host * 1e6, t_wall * 1e6, e0.elapsed_time(e1) * 1e3 / n), flush=True)
