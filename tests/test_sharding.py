"""Host-side multi-GPU logic on CPU: shard planning, and the N>1 exchange
helpers over a world_size-2 gloo group (the compute inside each rank is the
oracle here; on GPUs it is the CUDA path — see test_gpu_parity.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle
from multiviewhmr_b200 import sharding, synthetic as syn


@pytest.mark.parametrize("B,gx,world", [(8, 64, 1), (8, 64, 2), (8, 64, 8), (1, 64, 8), (3, 10, 4),
                                         (64, 80, 8), (2, 7, 5), (5, 3, 8), (1, 2, 4)])
def test_windows_partition_the_problem(B, gx, world):
    seen = np.zeros((B, gx), dtype=np.int32)
    sizes = []
    for r in range(world):
        wins = sharding.shard_windows(B, gx, r, world)
        sizes.append(sum(w.units() for w in wins))
        for w in wins:
            assert 0 <= w.b0 < w.b1 <= B and 0 <= w.x0 < w.x1 <= gx
            if w.b1 - w.b0 > 1:
                assert (w.x0, w.x1) == (0, gx)       # multi-sample windows are whole samples
            seen[w.b0:w.b1, w.x0:w.x1] += 1
    assert (seen == 1).all()
    assert max(sizes) - min(sizes) <= 1              # balanced to one x-plane


def test_batch_multiple_of_world_is_pure_batch_sharding():
    for r in range(8):
        (w,) = sharding.shard_windows(64, 80, r, 8)
        assert (w.b0, w.b1, w.x0, w.x1) == (8 * r, 8 * r + 8, 0, 80)
    with pytest.raises(ValueError):
        sharding.shard_windows(4, 4, 4, 4)


def test_record_merge_equals_full_soft_argmax():
    g = torch.Generator().manual_seed(1)
    vol = torch.randn(2, 3, 8, 6, 4, generator=g) * 4
    cv = torch.randn(2, 8, 6, 4, 3, generator=g) * 500
    truth = oracle.soft_argmax_3d(vol, cv)
    recs = []
    for x0, x1 in [(0, 3), (3, 4), (4, 8)]:           # ragged slabs
        v = vol[:, :, x0:x1].reshape(2, 3, -1).double()
        c = cv[:, x0:x1].reshape(2, -1, 3).double()
        m = v.max(dim=2, keepdim=True).values
        e = torch.exp(v - m)
        rec = torch.cat([m, e.sum(2, keepdim=True), torch.einsum("bjn,bnc->bjc", e, c)], dim=2)
        recs.append(rec.unsqueeze(2))
    empty = torch.tensor([-float("inf"), 0, 0, 0, 0], dtype=torch.float64).expand(2, 3, 1, 5)
    merged = sharding.merge_softargmax_records(torch.cat(recs + [empty], dim=2))
    assert np.abs(merged.numpy() - truth).max() < 1e-9 * 500


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, G, method, q):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        w = syn.Workload("t", B=B, V=3, C=4, H=16, W=16, G=G, method=method)
        f, P, cv, _ = syn.make_inputs(w, seed=9)
        full = oracle.unprojection(f, P, cv, method)
        out = torch.zeros(B, w.C, G, G, G)
        for win in sharding.shard_windows(B, G, rank, world):
            piece = oracle.unprojection(f[win.b0:win.b1], P[win.b0:win.b1],
                                        cv[win.b0:win.b1, win.x0:win.x1], method)
            out[win.b0:win.b1, :, win.x0:win.x1] = torch.from_numpy(piece)
        mine = out.clone()
        sharding.all_gather_volume(out, B, G, world)
        ok_gather = np.array_equal(out.numpy(), full)
        # slab-sharded soft-argmax: one record per (b, j) per rank, 5 floats each
        vol = torch.from_numpy(full)
        rec = torch.zeros(B, w.C, 1, 5, dtype=torch.float64)
        rec[..., 0] = -float("inf")
        for win in sharding.shard_windows(B, G, rank, world):
            v = vol[win.b0:win.b1, :, win.x0:win.x1].reshape(win.b1 - win.b0, w.C, -1).double()
            c = cv[win.b0:win.b1, win.x0:win.x1].reshape(win.b1 - win.b0, -1, 3).double()
            m = v.max(dim=2, keepdim=True).values
            e = torch.exp(v - m)
            new = torch.cat([m, e.sum(2, keepdim=True), torch.einsum("bjn,bnc->bjc", e, c)], dim=2).unsqueeze(2)
            both = torch.cat([rec[win.b0:win.b1], new], dim=2)
            M = both[..., 0].max(dim=2, keepdim=True).values
            sc = torch.where(torch.isinf(both[..., 0]), torch.zeros_like(M), torch.exp(both[..., 0] - M))
            rec[win.b0:win.b1, :, 0, 0] = M[..., 0]
            rec[win.b0:win.b1, :, 0, 1:] = (both[..., 1:] * sc.unsqueeze(-1)).sum(2)
        allrec = sharding.all_gather_records(rec, world)
        sa = sharding.merge_softargmax_records(allrec)
        ok_sa = float(np.abs(sa.numpy() - oracle.soft_argmax_3d(vol, cv)).max()) < 1e-8 * 1250
        written = int((mine != 0).sum())
        q.put((rank, ok_gather, ok_sa, written))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B,G,method", [(2, 6, "softmax"), (1, 6, "sum"), (3, 5, "max")])
def test_two_rank_gloo_shards_reassemble_bitwise(B, G, method):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, B, G, method, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in results), results
    assert all(ok for _, _, ok, _ in results), results
    assert all(w > 0 for *_, w in results)
