"""Autograd for the fused unprojection.

Forward = the fused kernel; backward = `mvhmr_unproject_aggregate_backward`, the gradient
w.r.t. the feature maps that torch autograd derives for the reference's
`models/aggregation.py:20-87` (the only tensor there that carries gradient: projections and
coordinates are built from constants).  Training through `VolumeGenerator` (`train.py:110`)
therefore works with the drop-in.
"""
import torch

from . import _lib


class _Unprojection(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, proj_matricies, coord_volumes, aggregation_method, window, packed):
        from .aggregation import unprojection
        with torch.no_grad():
            if window is None:
                out = unprojection(features.detach(), proj_matricies, coord_volumes, aggregation_method, packed=packed)
            else:
                # shard window: voxels outside it are zeros here (and receive no gradient)
                B, _, C = features.shape[:3]
                out = torch.zeros((B, C) + tuple(coord_volumes.shape[1:4]), dtype=torch.float32, device=features.device)
                unprojection(features.detach(), proj_matricies, coord_volumes, aggregation_method,
                             window=window, out=out, packed=packed)
        ctx.save_for_backward(features, proj_matricies, coord_volumes)
        ctx.method = aggregation_method
        ctx.window = window
        return out

    @staticmethod
    def backward(ctx, grad_out):
        features, proj_matricies, coord_volumes = ctx.saved_tensors
        g = grad_out
        if ctx.window is not None:
            b0, b1, n0, n1 = ctx.window
            B, C = g.shape[:2]
            masked = torch.zeros_like(g, memory_format=torch.contiguous_format)
            masked.view(B, C, -1)[b0:b1, :, n0:n1] = g.reshape(B, C, -1)[b0:b1, :, n0:n1]
            g = masked
        gf = unprojection_backward(g, features, proj_matricies, coord_volumes, ctx.method)
        return gf.to(features.dtype), None, None, None, None, None


def unprojection_backward(grad_out, features, proj_matricies, coord_volumes, aggregation_method, simple=False):
    """Gradient of `unprojection` w.r.t. `features`: (B,C,...) fp32 x (B,V,C,H,W) -> (B,V,C,H,W) fp32.
    `simple=True` runs the first, thread-per-voxel kernel (kept as an independent implementation
    for the tests; V <= 64)."""
    from .aggregation import _feat_dtype
    dev = _lib.require_cuda(grad_out, features, proj_matricies, coord_volumes)
    B, V, C, H, W = features.shape
    N = coord_volumes.numel() // (3 * B) if B else 0
    L = _lib.load()
    dt, m = _feat_dtype(features), _lib.METHODS[aggregation_method]
    g = grad_out.detach().float().contiguous()
    feats = features.detach().contiguous()
    proj = proj_matricies.detach().float().contiguous()
    coord = coord_volumes.detach().float().contiguous()
    with torch.cuda.device(dev):
        if simple:
            gf = torch.zeros((B, V, C, H, W), dtype=torch.float32, device=dev)
            _lib.check(L.mvhmr_unproject_aggregate_backward(
                _lib.ptr(g), _lib.ptr(feats), dt, _lib.ptr(proj), _lib.ptr(coord), _lib.ptr(gf),
                B, V, C, H, W, N, m, _lib.stream_ptr(dev)))
            return gf
        gf = torch.empty((B, V, C, H, W), dtype=torch.float32, device=dev)
        ws_bytes = L.mvhmr_unproject_backward_workspace_bytes(dt, B, V, C, H, W, m)
        ws = torch.empty(max(ws_bytes, 16), dtype=torch.uint8, device=dev)
        _lib.check(L.mvhmr_unproject_aggregate_backward_ws(
            _lib.ptr(g), _lib.ptr(feats), dt, _lib.ptr(proj), _lib.ptr(coord), _lib.ptr(gf),
            B, V, C, H, W, N, m, _lib.ptr(ws), ws_bytes, _lib.stream_ptr(dev)))
    return gf


def unprojection_with_grad(features, proj_matricies, coord_volumes, aggregation_method, window=None, packed=None):
    """Differentiable `unprojection`.  `window=(b0,b1,n0,n1)` restricts both passes to a shard
    (zeros outside); the result is always a fresh tensor (`out=` cannot be combined with
    autograd — `aggregation.unprojection` rejects it)."""
    if window is not None:
        window = tuple(int(v) for v in window)
    return _Unprojection.apply(features, proj_matricies, coord_volumes, aggregation_method, window, packed)
