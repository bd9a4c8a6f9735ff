"""Drop-in for the reference's `utils/volumetric.py` (hot-path parts).

`Cuboid3D` (only `position` / `sides` are read on the path), the axis-angle
rotation matrix (float64 numpy on the host, exactly as the reference computes
it) and `rotate_coord_volume`, whose fp32 3x3 · 3xN product runs on the GPU
with the reference sgemm's rounding order.  The cv2 debug renderers
(`Point3D`, `Line3D`, `Cuboid3D.render`) are out of scope.
"""
import numpy as np
import torch

from . import _lib


class Cuboid3D:
    """Axis-aligned cuboid primitive — reference `utils/volumetric.py:44-47`."""

    def __init__(self, position, sides):
        self.position = position
        self.sides = sides

    def render(self, proj_matrix, canvas):
        raise NotImplementedError("debug rendering is outside the aggregation hot path")


def get_rotation_matrix(axis, theta):
    """Rotation matrix for a counter-clockwise turn of `theta` radians about
    `axis` (Euler-Rodrigues, float64) — reference `utils/volumetric.py:87-99`."""
    axis = np.asarray(axis)
    axis = axis / np.sqrt(np.dot(axis, axis))
    a = np.cos(theta / 2.0)
    b, c, d = -axis * np.sin(theta / 2.0)
    aa, bb, cc, dd = a * a, b * b, c * c, d * d
    bc, ad, ac, ab, bd, cd = b * c, a * d, a * c, a * b, b * d, c * d
    return np.array([[aa + bb - cc - dd, 2 * (bc + ad), 2 * (bd - ac)],
                     [2 * (bc - ad), aa + cc - bb - dd, 2 * (cd + ab)],
                     [2 * (bd + ac), 2 * (cd - ab), aa + dd - bb - cc]])


def rotate_coord_volume(coord_volume, theta, axis):
    """Rotate every point of a (..., 3) CUDA tensor — reference `:102-114`."""
    dev = _lib.require_cuda(coord_volume)
    if coord_volume.dtype != torch.float32:
        raise TypeError("multiviewhmr_b200: rotate_coord_volume is fp32 (got %s)" % coord_volume.dtype)
    if coord_volume.shape[-1] != 3:
        raise ValueError("expected a (..., 3) coordinate tensor, got %s" % (tuple(coord_volume.shape),))
    rot = get_rotation_matrix(axis, theta).astype(np.float32)
    pts = coord_volume.contiguous()
    out = torch.empty_like(pts)
    with torch.cuda.device(dev):
        _lib.check(_lib.load().mvhmr_rotate_points(
            _lib.ptr(out), _lib.ptr(pts), _lib.host3(rot.reshape(-1).tolist()), pts.numel() // 3,
            _lib.stream_ptr(dev)))
    return out
