"""Multi-GPU sharding of the aggregation path.

The path is embarrassingly parallel: an output voxel depends only on its own
sample's feature maps and cameras (`models/aggregation.py:28-83`).  One process
per GPU; the work — the B*Gx x-planes of the batch, in (b, x) order — is cut
into `world_size` contiguous, equally sized pieces.  When B is a multiple of
world_size this is plain batch sharding (the reference's only parallelism, DDP
in `train.py:166-168`); otherwise samples are split into x-slabs, which are
contiguous runs of the flattened voxel index.  There is NO collective in the
data path.  The only exchanges offered here are optional:

  * `all_gather_volume`     — assemble the full (B,C,G,G,G) output on every
                              rank (API completeness; costs ~40x the compute
                              at cfg #5, keep outputs sharded when possible);
  * `all_gather_records`    — 5 floats per (b, joint, rank) to finish a
                              soft-argmax over a slab-sharded volume.
"""
from dataclasses import dataclass
from typing import List

import torch


@dataclass(frozen=True)
class Window:
    """Samples [b0,b1) x x-planes [x0,x1) of a (B,Gx,Gy,Gz) problem."""
    b0: int
    b1: int
    x0: int
    x1: int

    def voxels(self, gy, gz):
        return self.x0 * gy * gz, self.x1 * gy * gz

    def units(self):
        return (self.b1 - self.b0) * (self.x1 - self.x0)


def shard_windows(B: int, gx: int, rank: int, world_size: int) -> List[Window]:
    """Windows owned by `rank`: whole samples where possible, x-slabs otherwise.
    Over all ranks the windows are disjoint and cover every (b, x) exactly once."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d outside world of %d" % (rank, world_size))
    total = B * gx
    lo = (total * rank) // world_size
    hi = (total * (rank + 1)) // world_size
    out: List[Window] = []
    u = lo
    while u < hi:
        b, x = divmod(u, gx)
        if x == 0 and hi - u >= gx:                     # run of whole samples
            nb = (hi - u) // gx
            out.append(Window(b, b + nb, 0, gx))
            u += nb * gx
        else:                                           # partial sample: one slab
            x1 = min(gx, x + (hi - u))
            out.append(Window(b, b + 1, x, x1))
            u += x1 - x
    return out


def unprojection_sharded(features, proj_matricies, coord_volumes, aggregation_method,
                         rank, world_size, out=None):
    """Compute this rank's share of `unprojection` over replicated inputs.

    Returns (out, windows): `out` is a full-size (B,C,Gx,Gy,Gz) buffer in which
    only the rank's windows are written (bit-identical to the unsharded call
    there).  When the feature maps require grad the windows go through autograd
    (`out=` is then not allowed): the returned tensor is zero outside the rank's
    windows and its backward only sees the gradient inside them."""
    from .aggregation import unprojection
    B = features.shape[0]
    gx, gy, gz = (int(v) for v in coord_volumes.shape[1:4])
    wins = shard_windows(B, gx, rank, world_size)
    if torch.is_grad_enabled() and features.requires_grad:
        if out is not None:
            raise ValueError("unprojection_sharded: `out=` cannot be combined with autograd")
        total = None
        for w in wins:
            n0, n1 = w.voxels(gy, gz)
            part = unprojection(features, proj_matricies, coord_volumes, aggregation_method,
                                window=(w.b0, w.b1, n0, n1))
            total = part if total is None else total + part
        if total is None:
            total = torch.zeros((B, features.shape[2], gx, gy, gz), dtype=torch.float32, device=features.device)
        return total, wins
    if out is None:
        out = torch.zeros((B, features.shape[2], gx, gy, gz), dtype=torch.float32, device=features.device)
    for w in wins:
        n0, n1 = w.voxels(gy, gz)
        unprojection(features, proj_matricies, coord_volumes, aggregation_method,
                     window=(w.b0, w.b1, n0, n1), out=out)
    return out, wins


def all_gather_volume(out, B, gx, world_size, group=None):
    """Optional: every rank ends with the full volume.  `out` (B,C,Gx,Gy,Gz)
    holds this rank's windows; the other ranks' windows arrive with ONE
    collective.  When every rank owns the same number of whole samples (B a
    multiple of world_size — the case the scaling sweep runs) the windows are
    contiguous slices of `out` and `all_gather_into_tensor` writes them in place,
    with no staging copy; otherwise each rank packs its x-slabs into one flat
    buffer and a single `all_gather` of equally sized buffers follows.  NCCL over
    NVLink on GPUs, gloo on CPU tensors (tests).  Costs ~40x the compute at
    cfg #5: keep the outputs sharded whenever the consumer is data-parallel."""
    import torch.distributed as dist
    rank = dist.get_rank(group)
    if B % world_size == 0 and out.is_contiguous():
        per = B // world_size
        mine = out[rank * per:(rank + 1) * per]
        try:
            dist.all_gather_into_tensor(out.view(-1), mine.reshape(-1), group=group)
            return out
        except (RuntimeError, NotImplementedError):              # backend without the fused form (older gloo)
            pass
    wins = [shard_windows(B, gx, r, world_size) for r in range(world_size)]
    C, gy, gz = out.shape[1], out.shape[3], out.shape[4]
    sizes = [sum(w.units() for w in ws) * C * gy * gz for ws in wins]
    cap = max(sizes)
    flat = out.new_zeros(cap)
    pos = 0
    for w in wins[rank]:
        piece = out[w.b0:w.b1, :, w.x0:w.x1].reshape(-1)
        flat[pos:pos + piece.numel()] = piece
        pos += piece.numel()
    parts = [torch.empty_like(flat) for _ in range(world_size)]
    dist.all_gather(parts, flat, group=group)
    for r in range(world_size):
        if r == rank:
            continue
        pos = 0
        for w in wins[r]:
            n = w.units() * C * gy * gz
            out[w.b0:w.b1, :, w.x0:w.x1] = parts[r][pos:pos + n].view(w.b1 - w.b0, C, w.x1 - w.x0, gy, gz)
            pos += n
    return out


def merge_softargmax_records(records):
    """Merge online-softmax records along dim -2.
    records (..., S, 5) = (max, sum e, sum e*x, sum e*y, sum e*z) -> (..., 3).
    Same algebra as the finalize kernel; runs on whatever device the tensor is
    on (the cross-rank merge handles B*J*world records — a few hundred floats)."""
    m = records[..., 0]
    M = m.max(dim=-1, keepdim=True).values
    scale = torch.where(m == M, torch.ones_like(m), torch.exp(m - M))
    scale = torch.where(torch.isinf(m) & (m < 0), torch.zeros_like(m), scale)
    sums = (records[..., 1:] * scale.unsqueeze(-1)).sum(dim=-2)
    return sums[..., 1:] / sums[..., :1]


def all_gather_records(records, world_size, group=None):
    """records (B,J,S,5) of this rank -> (B,J,S*world,5) on every rank."""
    import torch.distributed as dist
    parts = [torch.empty_like(records) for _ in range(world_size)]
    dist.all_gather(parts, records.contiguous(), group=group)
    return torch.cat(parts, dim=2)
