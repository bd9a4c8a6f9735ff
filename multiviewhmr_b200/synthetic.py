"""Synthetic inputs for the volumetric-aggregation path (SURVEY.md §8(c,d)).

The reference ships no datasets, fixtures or tests, so every parity case and
every benchmark in this repo runs on the deterministic generator below: V ring
cameras looking at a cuboid of side `cuboid_side` mm from 4500 mm, focal length
chosen so the cuboid fills ~80 % of the feature map, N(0,1) feature maps and
N(0,100^2) mm cuboid centres.

Nothing in here touches the GPU library or the oracle; it only produces host
tensors with the layouts the reference's entry points take
(`/root/reference/models/aggregation.py:20-25`).
"""
import math
from dataclasses import dataclass

import numpy as np
import torch


@dataclass(frozen=True)
class Workload:
    """One BASELINE.json configuration."""
    name: str
    B: int
    V: int
    C: int
    H: int
    W: int
    G: int
    method: str = "softmax"
    dtype: str = "fp32"      # storage type of the feature maps
    joints: int = 0          # >0: also run the J-joint 3-D soft-argmax
    cuboid_side: float = 2500.0

    @property
    def vcv(self):
        """voxel-channel-views of one pass (BASELINE.json metric numerator)."""
        return self.B * self.G ** 3 * self.C * self.V

    def algorithmic_bytes(self, coord_in_hbm=True):
        """Compulsory HBM bytes of one unproject+aggregate pass."""
        e = 2 if self.dtype == "bf16" else 4
        n = self.G ** 3
        b = self.B * self.V * self.C * self.H * self.W * e + self.B * self.V * 48
        b += self.B * self.C * n * 4
        if coord_in_hbm:
            b += self.B * n * 12
        return b

    def soft_argmax_bytes(self):
        n = self.G ** 3
        return self.B * self.joints * n * 4 + self.B * n * 12 + self.B * self.joints * 12


# BASELINE.json `configs`, in order.
CONFIGS = {
    "cfg1": Workload("cfg1", B=1, V=4, C=32, H=64, W=64, G=32, method="sum"),
    "cfg2": Workload("cfg2", B=8, V=4, C=32, H=96, W=96, G=64),
    "cfg3": Workload("cfg3", B=8, V=4, C=32, H=96, W=96, G=64, dtype="bf16", joints=17),
    "cfg4": Workload("cfg4", B=16, V=8, C=64, H=128, W=128, G=64),
    "cfg5": Workload("cfg5", B=64, V=8, C=32, H=96, W=96, G=80),
}


def ring_projection(b, v, V, H, W, cuboid_side=2500.0, distance=4500.0, behind=False):
    """3x4 projection matrix of ring camera v for sample b, float64.

    yaw a = 2*pi*v/V + 0.3 + 0.05*b about the world y axis, t = (30,-20,distance),
    f = 0.8*W*distance/cuboid_side, principal point (W/2, H/2). `behind=True`
    pulls the camera into the cuboid so that part of the grid has depth <= 0.
    """
    a = 2.0 * math.pi * v / V + 0.3 + 0.05 * b
    R = np.array([[math.cos(a), 0.0, -math.sin(a)],
                  [0.0, 1.0, 0.0],
                  [math.sin(a), 0.0, math.cos(a)]])
    dist = 0.2 * cuboid_side if behind else distance
    t = np.array([[30.0], [-20.0], [dist]])
    f = 0.8 * W * distance / cuboid_side
    K = np.array([[f, 0.0, W / 2.0], [0.0, f, H / 2.0], [0.0, 0.0, 1.0]])
    return K @ np.hstack([R, t])


def make_projections(B, V, H, W, cuboid_side=2500.0, behind_views=(), b_offset=0):
    """(B,V,3,4) fp32 projections of samples [b_offset, b_offset + B) of the ring-camera set."""
    P = np.zeros((B, V, 3, 4), dtype=np.float64)
    for b in range(B):
        for v in range(V):
            P[b, v] = ring_projection(b + b_offset, v, V, H, W, cuboid_side, behind=(v in behind_views))
    return torch.from_numpy(P).float()


def make_coord_volumes(centers, G, cuboid_side=2500.0, theta=0.0, axis=(0, 0, 1)):
    """Host construction of the coord volume with the reference's rounding
    (`models/aggregation.py:140-187`): f32(pos) + f32(side/(G-1))*f32(idx), then
    (coord - c), optional rotation, (+ c). Used to feed `unprojection` in tests
    and benches; the product path builds the same thing on the GPU."""
    B = centers.shape[0]
    pos = np.float32(-cuboid_side / 2.0)
    step = np.float32(cuboid_side / (G - 1))
    idx = np.arange(G, dtype=np.float32)
    line = (pos + step * idx).astype(np.float32)          # separately rounded mul, add
    grid = np.stack(np.meshgrid(line, line, line, indexing="ij"), axis=-1)  # (G,G,G,3)
    out = np.empty((B, G, G, G, 3), dtype=np.float32)
    c = centers.detach().cpu().numpy().astype(np.float32)
    rot = rotation_matrix(axis, theta).astype(np.float32)
    for b in range(B):
        p = (grid - c[b]).astype(np.float32)
        if theta != 0.0:
            q = np.empty_like(p)
            for i in range(3):      # forward FMA chain == sgemm K=3
                acc = (p[..., 0].astype(np.float64) * rot[i, 0]).astype(np.float32)
                acc = (p[..., 1].astype(np.float64) * rot[i, 1] + acc).astype(np.float32)
                acc = (p[..., 2].astype(np.float64) * rot[i, 2] + acc).astype(np.float32)
                q[..., i] = acc
            p = q
        out[b] = (p + c[b]).astype(np.float32)
    return torch.from_numpy(out)


def rotation_matrix(axis, theta):
    """Euler-Rodrigues rotation matrix in float64 (same formula as
    `utils/volumetric.py:87-99`; restated, theta=0 gives the exact identity)."""
    ax = np.asarray(axis, dtype=np.float64)
    ax = ax / math.sqrt(float(np.dot(ax, ax)))
    a = math.cos(theta / 2.0)
    b, c, d = (-ax * math.sin(theta / 2.0)).tolist()
    return np.array([
        [a * a + b * b - c * c - d * d, 2 * (b * c + a * d), 2 * (b * d - a * c)],
        [2 * (b * c - a * d), a * a + c * c - b * b - d * d, 2 * (c * d + a * b)],
        [2 * (b * d + a * c), 2 * (c * d - a * b), a * a + d * d - b * b - c * c]])


def make_inputs(w, seed=1234, theta=0.0, behind_views=(), H=None, W=None, b_offset=0):
    """(features, proj, coord_volumes, centers) on the host for workload `w`.

    features (B,V,C,H,W) fp32 N(0,1) — for bf16 workloads already rounded to
    bf16 values (still returned as fp32; cast with `.bfloat16()` is lossless),
    proj (B,V,3,4) fp32, coord_volumes (B,G,G,G,3) fp32, centers (B,3) fp32.
    `b_offset` shifts the sample index the ring cameras are built from (a rank that
    owns samples [b_offset, b_offset + B) of a larger batch).
    """
    H = w.H if H is None else H
    W = w.W if W is None else W
    g = torch.Generator().manual_seed(seed)
    feats = torch.randn(w.B, w.V, w.C, H, W, generator=g)
    centers = torch.randn(w.B, 3, generator=g) * 100.0
    if w.dtype == "bf16":
        feats = feats.bfloat16().float()
    proj = make_projections(w.B, w.V, H, W, w.cuboid_side, behind_views, b_offset)
    coord = make_coord_volumes(centers, w.G, w.cuboid_side, theta)
    return feats, proj, coord, centers
