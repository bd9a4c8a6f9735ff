"""The reference's torch call sequence for the aggregation path, restated.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  This is what
`bench.py --impl reference` and the `cpu_baseline` leg time on the host cores:
`/root/reference` does not exist on the GPU box, so the stock path cannot be
imported there.  The port issues the same ATen operators in the same order as
`/root/reference/models/aggregation.py:20-87` (one `@`, one `F.grid_sample`,
one masked fill per view; one softmax per sample), so its CPU cost and its
rounding are the reference's.  `tests/test_oracle.py` asserts `torch.equal`
between this port and the imported reference whenever `/root/reference` is
present, and against the golden fixtures everywhere else.
"""
import torch
import torch.nn.functional as F


def _homog_project(P, pts):
    """utils/multiview.py:55-69,89-110 with convert_back_to_euclidean=False."""
    ones = torch.ones((pts.shape[0], 1), dtype=pts.dtype, device=pts.device)
    return torch.cat([pts, ones], dim=1) @ P.t()


def _sample_view(fmap, P, pts, vol_shape):
    """One (b, v) of models/aggregation.py:34-65 -> (C, *vol_shape)."""
    C, H, W = fmap.shape
    ph = _homog_project(P, pts)
    behind = ph[:, 2] <= 0.0
    ph[ph[:, 2] == 0.0, 2] = 1.0
    xy = (ph.transpose(1, 0)[:-1] / ph.transpose(1, 0)[-1]).transpose(1, 0)
    g = torch.zeros_like(xy)
    g[:, 0] = 2 * (xy[:, 0] / H - 0.5)     # x by feature_shape[0] — reference behaviour
    g[:, 1] = 2 * (xy[:, 1] / W - 0.5)
    s = F.grid_sample(fmap.unsqueeze(0), g.unsqueeze(1).unsqueeze(0), align_corners=True)
    s = s.view(C, -1)
    s[:, behind] = 0.0
    return s.view(C, *vol_shape)


def fuse_views(stack, method):
    """models/aggregation.py:71-85 on a (V, C, *vol) stack."""
    if method == "sum":
        return stack.sum(0)
    if method == "mean":
        return stack.mean(0)
    if method == "max":
        return stack.max(0)[0]
    if method == "softmax":
        V = stack.shape[0]
        p = F.softmax(stack.clone().view(V, -1), dim=0).view_as(stack)
        return (stack * p).sum(0)
    raise ValueError("Unknown aggregation_method: {}".format(method))


def unprojection(features, proj_matricies, coord_volumes, aggregation_method="softmax"):
    B, V, C = features.shape[:3]
    vol_shape = coord_volumes.shape[1:4]
    dev = features.device
    out = torch.zeros(B, C, *vol_shape, device=dev)
    for b in range(B):
        pts = coord_volumes[b].reshape((-1, 3))
        stack = torch.zeros(V, C, *vol_shape, device=dev)
        for v in range(V):
            stack[v] = _sample_view(features[b, v], proj_matricies[b, v], pts, vol_shape)
        out[b] = fuse_views(stack, aggregation_method)
    return out


def soft_argmax_3d(volumes, coord_volumes):
    """fp32 torch statement of the (unpinned) 3-D soft-argmax: (B,J,3).
    Upstream formula: softmax over all voxels, expectation of the coordinates."""
    B, J = volumes.shape[:2]
    p = F.softmax(volumes.reshape(B, J, -1), dim=2)
    return torch.einsum("bjn,bnc->bjc", p, coord_volumes.reshape(B, -1, 3))
