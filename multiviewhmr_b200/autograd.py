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
    def forward(ctx, features, proj_matricies, coord_volumes, aggregation_method):
        from .aggregation import unprojection
        with torch.no_grad():
            out = unprojection(features.detach(), proj_matricies, coord_volumes, aggregation_method)
        ctx.save_for_backward(features, proj_matricies, coord_volumes)
        ctx.method = aggregation_method
        return out

    @staticmethod
    def backward(ctx, grad_out):
        features, proj_matricies, coord_volumes = ctx.saved_tensors
        from .aggregation import _feat_dtype
        dev = features.device
        B, V, C, H, W = features.shape
        N = coord_volumes.shape[1] * coord_volumes.shape[2] * coord_volumes.shape[3]
        g = grad_out.detach().float().contiguous()
        feats = features.detach().contiguous()
        proj = proj_matricies.detach().float().contiguous()
        coord = coord_volumes.detach().float().contiguous()
        gf = torch.zeros((B, V, C, H, W), dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.load().mvhmr_unproject_aggregate_backward(
                _lib.ptr(g), _lib.ptr(feats), _feat_dtype(features), _lib.ptr(proj), _lib.ptr(coord), _lib.ptr(gf),
                B, V, C, H, W, N, _lib.METHODS[ctx.method], _lib.stream_ptr(dev)))
        return gf.to(features.dtype), None, None, None


def unprojection_with_grad(features, proj_matricies, coord_volumes, aggregation_method):
    return _Unprojection.apply(features, proj_matricies, coord_volumes, aggregation_method)
