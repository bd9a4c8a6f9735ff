"""Drop-in for the reference's `utils/multiview.py` (same names, argument
meaning and error behaviour).

`Camera` and the numpy branches stay host-side numpy, as they are in the
reference ("keep on host, bit-exact by construction").  The
torch branch of the 3x4 projection runs the sm_100a kernel behind
`mvhmr_project_points`; torch tensors must live on a CUDA device.
"""
import numpy as np
import torch

from . import _lib


class Camera:
    """Pinhole camera bookkeeping — reference `utils/multiview.py:5-52`."""

    def __init__(self, R, t, K, dist=None, name=""):
        self.R = np.array(R).copy()
        assert self.R.shape == (3, 3)
        self.t = np.array(t).copy()
        assert self.t.size == 3
        self.t = self.t.reshape(3, 1)
        self.K = np.array(K).copy()
        assert self.K.shape == (3, 3)
        self.dist = None if dist is None else np.array(dist).copy().flatten()
        self.name = name

    def update_after_crop(self, bbox):
        """Shift the principal point by the crop origin (`:23-31`)."""
        left, upper = bbox[0], bbox[1]
        self.K[0, 2] = self.K[0, 2] - left
        self.K[1, 2] = self.K[1, 2] - upper

    def update_after_resize(self, image_shape, new_image_shape):
        """Rescale intrinsics (`:33-44`).  NOTE the reference unpacks the new
        shape as (new_width, new_height) although callers pass (H, W); that
        argument order is part of its behaviour and is kept."""
        height, width = image_shape
        new_width, new_height = new_image_shape
        sx, sy = new_width / width, new_height / height
        fx, fy, cx, cy = self.K[0, 0], self.K[1, 1], self.K[0, 2], self.K[1, 2]
        self.K[0, 0], self.K[1, 1], self.K[0, 2], self.K[1, 2] = fx * sx, fy * sy, cx * sx, cy * sy

    @property
    def extrinsics(self):
        return np.hstack([self.R, self.t])

    @property
    def projection(self):
        return self.K.dot(self.extrinsics)


def euclidean_to_homogeneous(points):
    """(N, M) -> (N, M+1); reference `:55-69`."""
    if isinstance(points, np.ndarray):
        return np.hstack([points, np.ones((len(points), 1))])
    if torch.is_tensor(points):
        one = torch.ones((points.shape[0], 1), dtype=points.dtype, device=points.device)
        return torch.cat([points, one], dim=1)
    raise TypeError("Works only with numpy arrays and PyTorch tensors.")


def homogeneous_to_euclidean(points):
    """(N, M+1) -> (N, M); reference `:72-86`."""
    if isinstance(points, np.ndarray):
        return (points.T[:-1] / points.T[-1]).T
    if torch.is_tensor(points):
        pt = points.transpose(1, 0)
        return (pt[:-1] / pt[-1]).transpose(1, 0)
    raise TypeError("Works only with numpy arrays and PyTorch tensors.")


class _ProjectPoints(torch.autograd.Function):
    """Kernel forward (`mvhmr_project_points`), analytic backward.  The reference's torch branch
    is differentiable (`utils/loss.py:377-378` projects predictions that require grad), so the
    drop-in has to be as well."""

    @staticmethod
    def forward(ctx, proj_matrix, points_3d, euclid):
        dev = points_3d.device
        P, pts = proj_matrix.detach().contiguous(), points_3d.detach().contiguous()
        n = pts.shape[0]
        out = torch.empty((n, 2 if euclid else 3), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().mvhmr_project_points(
                _lib.ptr(out), _lib.ptr(P), _lib.ptr(pts), n, int(bool(euclid)), _lib.stream_ptr(dev)))
        ctx.save_for_backward(P, pts, out)
        ctx.euclid = bool(euclid)
        return out

    @staticmethod
    def backward(ctx, grad):
        P, pts, out = ctx.saved_tensors
        grad = grad.contiguous()
        if ctx.euclid:
            # out = h[:, :2] / w with h = [p 1] P^T, w = h[:, 2]
            w = (pts @ P[2, :3] + P[2, 3]).unsqueeze(1)
            g_xy = grad / w
            g_w = -(g_xy * out).sum(dim=1, keepdim=True)
            g_h = torch.cat([g_xy, g_w], dim=1)
        else:
            g_h = grad
        g_pts = g_h @ P[:, :3] if ctx.needs_input_grad[1] else None
        g_P = None
        if ctx.needs_input_grad[0]:
            ones = torch.ones((pts.shape[0], 1), dtype=pts.dtype, device=pts.device)
            g_P = g_h.t() @ torch.cat([pts, ones], dim=1)
        return g_P, g_pts, None


def project_3d_points_to_image_plane_without_distortion(proj_matrix, points_3d, convert_back_to_euclidean=True):
    """Project (N,3) points with a 3x4 matrix; reference `:89-110`.

    numpy in -> numpy out (host, float64 as given).  torch in -> CUDA kernel
    with the reference's fp32 rounding (K=4 FMA chain), (N,2) or (N,3) out;
    differentiable w.r.t. both arguments like the reference's torch ops."""
    if isinstance(proj_matrix, np.ndarray) and isinstance(points_3d, np.ndarray):
        result = euclidean_to_homogeneous(points_3d) @ proj_matrix.T
        return homogeneous_to_euclidean(result) if convert_back_to_euclidean else result
    if torch.is_tensor(proj_matrix) and torch.is_tensor(points_3d):
        _lib.require_cuda(points_3d, proj_matrix)
        if points_3d.dtype != torch.float32 or proj_matrix.dtype != torch.float32:
            raise TypeError("multiviewhmr_b200: projection kernel is fp32 (got %s, %s)"
                            % (proj_matrix.dtype, points_3d.dtype))
        if proj_matrix.shape != (3, 4) or points_3d.dim() != 2 or points_3d.shape[1] != 3:
            raise ValueError("expected proj_matrix (3,4) and points_3d (N,3), got %s and %s"
                             % (tuple(proj_matrix.shape), tuple(points_3d.shape)))
        return _ProjectPoints.apply(proj_matrix, points_3d, bool(convert_back_to_euclidean))
    raise TypeError("Works only with numpy arrays and PyTorch tensors.")


def triangulate_point_from_multiple_views_linear(proj_matricies, points):
    """DLT triangulation of one point, numpy; reference `:113-138`."""
    assert len(proj_matricies) == len(points)
    n_views = len(proj_matricies)
    A = np.zeros((2 * n_views, 4))
    for j in range(n_views):
        A[2 * j] = points[j][0] * proj_matricies[j][2, :] - proj_matricies[j][0, :]
        A[2 * j + 1] = points[j][1] * proj_matricies[j][2, :] - proj_matricies[j][1, :]
    _, _, vh = np.linalg.svd(A, full_matrices=False)
    return homogeneous_to_euclidean(vh[3, :])


def triangulate_point_from_multiple_views_linear_torch(proj_matricies, points, confidences=None):
    """DLT triangulation, torch; reference `:141-168`.  A 2V x 4 SVD: kept as a
    torch op on whatever device the inputs are on."""
    assert len(proj_matricies) == len(points)
    n_views = len(proj_matricies)
    if confidences is None:
        confidences = torch.ones(n_views, dtype=torch.float32, device=points.device)
    A = proj_matricies[:, 2:3].expand(n_views, 2, 4) * points.view(n_views, 2, 1)
    A = (A - proj_matricies[:, :2]) * confidences.view(-1, 1, 1)
    _, _, vh = torch.svd(A.reshape(-1, 4))
    point_3d_homo = -vh[:, 3]
    return homogeneous_to_euclidean(point_3d_homo.unsqueeze(0))[0]
