"""Round-2 GPU parity cases: BASELINE configs #4 and #5 at full grid size, the
shard windows of the cfg5 scaling sweep, both fused-kernel designs against each
other, and the drop-in behaviours the advisor flagged (autograd through the
projection helper, sharded training, default-flag GEMM route, shape checks)."""
import os

import numpy as np
import pytest
import torch

import oracle
from multiviewhmr_b200 import _lib, aggregation as agg, multiview, sharding, synthetic as syn
from conftest import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
OUR_TOL_SOFTMAX = 1e-6
SPEC_TOL_FP32 = 1e-5


def cuda(*arrays):
    return [torch.as_tensor(a).to(DEV) for a in arrays]


def sub_workload(w, B):
    return syn.Workload(w.name, B, w.V, w.C, w.H, w.W, w.G, w.method, w.dtype, w.joints, w.cuboid_side)


# ---------------------------------------------------------------- cfg4 / cfg5 at full grid size
@pytest.mark.parametrize("method", ["softmax", "max"])
def test_cfg4_full_grid_against_oracle(method):
    """BASELINE config #4 (V8 C64 128x128 -> 64^3): two samples of the full grid against the C
    restatement (softmax within ex2.approx noise, max bit-exact), then the whole B=16 batch
    through size-independent properties (sample b of the batch call == the same sample alone)."""
    w = syn.CONFIGS["cfg4"]
    f, P, cv, _ = syn.make_inputs(sub_workload(w, 2))
    got = agg.unprojection(*cuda(f, P, cv), method).cpu().numpy()
    ref = oracle.unprojection(f, P, cv, method)
    if method == "softmax":
        assert rel_l2(got, ref) < OUR_TOL_SOFTMAX
    else:
        assert np.array_equal(got, ref)


def test_cfg4_full_batch_is_sample_wise_consistent():
    w = syn.CONFIGS["cfg4"]
    f, P, cv, _ = syn.make_inputs(w)
    fd, Pd, cvd = cuda(f, P, cv)
    full = agg.unprojection(fd, Pd, cvd, "softmax")
    assert full.shape == (16, 64, 64, 64, 64)
    for b in (0, 7, 15):
        one = agg.unprojection(fd[b:b + 1].contiguous(), Pd[b:b + 1].contiguous(), cvd[b:b + 1].contiguous(), "softmax")
        assert torch.equal(one[0], full[b])
    ref = oracle.unprojection(f[15:16], P[15:16], cv[15:16], "softmax")
    assert rel_l2(full[15:16].cpu().numpy(), ref) < OUR_TOL_SOFTMAX
    del full


@pytest.mark.parametrize("method", ["softmax", "sum"])
def test_cfg5_full_grid_against_oracle(method):
    """BASELINE config #5 (V8 C32 96x96 -> 80^3, gz = 80 is not a multiple of 32): two samples of
    the full grid against the C restatement."""
    w = syn.CONFIGS["cfg5"]
    f, P, cv, _ = syn.make_inputs(sub_workload(w, 2))
    got = agg.unprojection(*cuda(f, P, cv), method).cpu().numpy()
    ref = oracle.unprojection(f, P, cv, method)
    if method == "softmax":
        assert rel_l2(got, ref) < OUR_TOL_SOFTMAX
    else:
        assert np.array_equal(got, ref)


def test_cfg5_shard_windows_are_bitwise_identical_to_the_unsharded_call():
    """All shard_windows of the B=64 cfg5 problem at 2/4/8 ranks (batch split) and at a world size
    that forces x-slabs, computed on one GPU, compared bit for bit with the unsharded call."""
    w = syn.CONFIGS["cfg5"]
    B = 8                                      # the windows are per (sample, x-plane): 8 samples exercise them all
    f, P, cv, _ = syn.make_inputs(sub_workload(w, B))
    fd, Pd, cvd = cuda(f, P, cv)
    full = agg.unprojection(fd, Pd, cvd, "softmax")
    for world in (2, 4, 8, 3, 16):
        out = torch.full_like(full, float("nan"))
        seen = 0
        for r in range(world):
            _, wins = sharding.unprojection_sharded(fd, Pd, cvd, "softmax", r, world, out=out)
            seen += sum(x.units() for x in wins)
        assert seen == B * w.G
        assert torch.equal(out, full), world
    # the windows the B=64 sweep really uses
    for world in (1, 2, 4, 8):
        wins = [sharding.shard_windows(64, w.G, r, world) for r in range(world)]
        assert all(len(x) == 1 and x[0].x0 == 0 and x[0].x1 == w.G and x[0].b1 - x[0].b0 == 64 // world for x in wins)
    del full, out


# ---------------------------------------------------------------- both kernel designs agree
@pytest.mark.parametrize("case", [("cfg1", "sum"), ("cfg1", "softmax"), ("cfg2s", "softmax"), ("cfg5s", "softmax"),
                                  ("cfg4s", "mean"), ("cfg3s", "softmax"), ("rot", "max")])
def test_staged_and_gather_kernels_agree_bitwise(case, monkeypatch):
    """The shared-memory-staged kernel and the L1-gather kernel do the same IEEE operations in the
    same order: their results must be equal bit for bit, in every fusion mode."""
    name, method = case
    table = {"cfg1": syn.CONFIGS["cfg1"], "cfg2s": sub_workload(syn.CONFIGS["cfg2"], 1),
             "cfg3s": sub_workload(syn.CONFIGS["cfg3"], 1), "cfg4s": sub_workload(syn.CONFIGS["cfg4"], 1),
             "cfg5s": sub_workload(syn.CONFIGS["cfg5"], 1),
             "rot": syn.Workload("rot", 2, 4, 32, 48, 40, 24)}
    w = table[name]
    f, P, cv, _ = syn.make_inputs(w, theta=0.7 if name == "rot" else 0.0, behind_views=(1,) if name == "rot" else ())
    fd, Pd, cvd = cuda(f, P, cv)
    if w.dtype == "bf16":
        fd = fd.bfloat16()
    monkeypatch.setenv("MVHMR_PATH", "gather")
    a = agg.unprojection(fd, Pd, cvd, method)
    monkeypatch.setenv("MVHMR_PATH", "staged")
    b = agg.unprojection(fd, Pd, cvd, method)
    assert torch.equal(a, b)
    pixel = w.C * fd.element_size()
    if pixel >= 16 and pixel & (pixel - 1) == 0:
        fcl = fd.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)
        assert torch.equal(agg.unprojection(fcl, Pd, cvd, method), a)


# ---------------------------------------------------------------- advisor findings
def test_projection_helper_is_differentiable_like_the_reference():
    """`utils/loss.py:377-378` projects predicted 3-D keypoints (requires_grad) through this helper; the
    2-D reprojection loss must reach them."""
    g = torch.Generator().manual_seed(0)
    P = torch.tensor(syn.ring_projection(0, 1, 4, 96, 96), dtype=torch.float32)
    pts = (torch.randn(17, 3, generator=g) * 300.0)
    target = torch.rand(17, 2, generator=g) * 96.0
    for euclid in (True, False):
        # the reference's torch ops (utils/multiview.py:89-110) on the CPU
        pr, Pr = pts.clone().requires_grad_(True), P.clone().requires_grad_(True)
        h = torch.cat([pr, torch.ones(17, 1)], dim=1) @ Pr.t()
        ref = (h.t()[:-1] / h.t()[-1]).t() if euclid else h
        tgt = target if euclid else torch.cat([target, torch.ones(17, 1)], dim=1)
        ((ref - tgt) ** 2).mean().backward()
        pd, Pd = pts.to(DEV).requires_grad_(True), P.to(DEV).requires_grad_(True)
        out = multiview.project_3d_points_to_image_plane_without_distortion(Pd, pd, euclid)
        assert out.requires_grad and out.grad_fn is not None
        ((out - tgt.to(DEV)) ** 2).mean().backward()
        assert torch.allclose(out.detach().cpu(), ref.detach(), rtol=1e-5, atol=1e-3)
        assert rel_l2(pd.grad.cpu().numpy(), pr.grad.numpy()) < 1e-4
        assert rel_l2(Pd.grad.cpu().numpy(), Pr.grad.numpy()) < 1e-4
    # without grad the helper stays a plain kernel call
    out = multiview.project_3d_points_to_image_plane_without_distortion(P.to(DEV), pts.to(DEV))
    assert not out.requires_grad


def test_sharded_unprojection_with_autograd():
    """Training over shard windows: forward equals the unsharded forward inside the windows (zeros
    outside), `out=` is rejected instead of silently ignored, and the sum of the ranks' feature
    gradients is the unsharded gradient."""
    w = syn.Workload("t", B=3, V=4, C=8, H=24, W=24, G=10)
    f, P, cv, _ = syn.make_inputs(w, seed=21)
    fd, Pd, cvd = cuda(f, P, cv)
    gout = torch.randn(3, 8, 10, 10, 10, device=DEV)
    full_in = fd.clone().requires_grad_(True)
    full = agg.unprojection(full_in, Pd, cvd, "softmax")
    full.backward(gout)
    world = 2                                   # 3 samples over 2 ranks: one rank gets an x-slab
    grads, covered = [], torch.zeros_like(full, dtype=torch.bool)
    for r in range(world):
        fi = fd.clone().requires_grad_(True)
        out, wins = sharding.unprojection_sharded(fi, Pd, cvd, "softmax", r, world)
        assert out.requires_grad
        mask = torch.zeros_like(covered)
        for x in wins:
            mask[x.b0:x.b1, :, x.x0:x.x1] = True
        assert torch.equal(out.detach()[mask], full.detach()[mask])
        assert float(out.detach()[~mask].abs().max()) == 0.0
        covered |= mask
        out.backward(gout)
        grads.append(fi.grad)
        with pytest.raises(ValueError):
            sharding.unprojection_sharded(fi, Pd, cvd, "softmax", r, world, out=torch.zeros_like(full))
    assert bool(covered.all())
    assert rel_l2((grads[0] + grads[1]).cpu().numpy(), full_in.grad.cpu().numpy()) < 1e-5
    with pytest.raises(ValueError):
        agg.unprojection(full_in, Pd, cvd, "softmax", out=torch.zeros_like(full))
    with pytest.raises(ValueError):
        agg.unprojection(full_in, Pd, cvd, "softmax", window=(0, 4, 0, 10))


def test_channels_last_gemm_route_runs_under_default_tf32_flags():
    B, V, Cin, G = 2, 3, 16, 6
    rng = np.random.default_rng(0)
    cams = [[multiview.Camera(np.eye(3), [0.0, 0.0, 4000.0 + 100 * v], [[300.0, 0, 32], [0, 300.0, 32], [0, 0, 1]])
             for _ in range(B)] for v in range(V)]
    batch = {"images": np.zeros((B, V, 64, 64, 3), np.float32), "cameras": cams,
             "keypoints_3d": [rng.normal(size=(17, 3)) * 50 for _ in range(B)]}
    vg = agg.VolumeGenerator(volume_size=G, input_channels=Cin, output_channels=8, cuboid_side=2000.0, device=DEV)
    vg.eval()
    feats = torch.randn(B, V, Cin, 16, 16, device=DEV)
    saved = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    try:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = False, True     # PyTorch's defaults
        with torch.no_grad():
            a = vg(feats, torch.zeros(B, V, 3, 4, device=DEV), batch)
        assert vg.last_route == "gemm_channels_last"
        assert torch.backends.cuda.matmul.allow_tf32 is False                                   # restored
        vg.channels_last = False
        with torch.no_grad():
            b = vg(feats, torch.zeros(B, V, 3, 4, device=DEV), batch)
        assert vg.last_route == "conv"
        assert rel_l2(a.cpu().numpy(), b.cpu().numpy()) < 5e-3                                  # both TF32 contractions
        torch.backends.cudnn.allow_tf32 = False
        vg.channels_last = True
        with torch.no_grad():
            c = vg(feats, torch.zeros(B, V, 3, 4, device=DEV), batch)
        vg.channels_last = False
        with torch.no_grad():
            d = vg(feats, torch.zeros(B, V, 3, 4, device=DEV), batch)
        assert rel_l2(c.cpu().numpy(), d.cpu().numpy()) < SPEC_TOL_FP32                         # both fp32
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = saved


def test_grid_descriptor_arrays_are_shape_checked():
    f, P = cuda(torch.zeros(2, 2, 4, 8, 8), torch.zeros(2, 2, 3, 4))
    good_c, good_r = np.zeros((2, 3), np.float32), np.stack([np.eye(3, dtype=np.float32)] * 2)
    agg.unprojection_grid(f, P, good_c, good_r, 4, 2500.0, "sum")
    with pytest.raises(ValueError):
        agg.unprojection_grid(f, P, good_c[:1], good_r, 4, 2500.0, "sum")
    with pytest.raises(ValueError):
        agg.unprojection_grid(f, P, good_c, good_r[:1], 4, 2500.0, "sum")
    with pytest.raises(ValueError):
        agg.build_coord_volumes(good_c, good_r.reshape(2, 9)[:, :6], 4, 2500.0, DEV)
    vol = torch.zeros(2, 3, 4, 4, 4, device=DEV)
    with pytest.raises(ValueError):
        agg.soft_argmax_3d_grid(vol, good_c[:1], good_r, 2500.0)


# ---------------------------------------------------------------- consumer-side output formats
FMT_SHAPES = [  # B, V, C, H, W, G, bf16
    (1, 4, 32, 24, 24, (8, 8, 32), False), (2, 8, 32, 20, 28, (6, 4, 80), False), (1, 3, 8, 12, 12, (4, 6, 10), False),
    (1, 4, 64, 16, 16, (4, 4, 12), True), (1, 8, 64, 16, 16, (2, 6, 34), False), (2, 2, 20, 9, 11, (2, 2, 6), False),
    (1, 4, 32, 16, 16, (6, 6, 6), True), (1, 11, 16, 10, 10, (4, 2, 8), False)]


@pytest.mark.parametrize("shape", FMT_SHAPES)
@pytest.mark.parametrize("method", ["sum", "mean", "max", "softmax"])
def test_channels_last_3d_output_has_the_same_values(shape, method):
    B, V, C, H, W, G, bf = shape
    g = torch.Generator().manual_seed(V * 100 + C)
    f = torch.randn(B, V, C, H, W, generator=g)
    P = syn.make_projections(B, V, H, W, behind_views=(V - 1,) if V > 2 else ())
    cv = (torch.rand(B, *G, 3, generator=g) - 0.5) * 2600.0
    fd, Pd, cvd = cuda(f, P, cv)
    if bf:
        fd = fd.bfloat16()
    ref = agg.unprojection(fd, Pd, cvd, method)
    got = agg.unprojection(fd, Pd, cvd, method, output="channels_last_3d")
    assert got.shape == ref.shape and got.is_contiguous(memory_format=torch.channels_last_3d)
    assert torch.equal(got, ref)
    # a window into a NaN-poisoned channels-last buffer
    N = G[0] * G[1] * G[2]
    out = torch.full((B, *G, C), float("nan"), device=DEV).permute(0, 4, 1, 2, 3)
    n0, n1 = N // 3, N - 1
    agg.unprojection(fd, Pd, cvd, method, window=(0, B, n0, n1), out=out, output="channels_last_3d")
    flat, rflat = out.reshape(B, C, N), ref.reshape(B, C, N)
    assert torch.equal(flat[:, :, n0:n1], rflat[:, :, n0:n1])
    assert bool(torch.isnan(flat[:, :, :n0]).all()) and bool(torch.isnan(flat[:, :, n1:]).all())


@pytest.mark.parametrize("shape", FMT_SHAPES)
@pytest.mark.parametrize("method", ["sum", "mean", "max", "softmax"])
def test_fused_max_pool_equals_pooling_the_aggregate(shape, method):
    """output='max_pool2' == F.max_pool3d(unprojection(...), 2), bit for bit (models/regressor.py:70-75
    pools the encoder's first block; heads that pool the aggregate itself never need it at full size)."""
    B, V, C, H, W, G, bf = shape
    g = torch.Generator().manual_seed(V * 100 + C + 1)
    f = torch.randn(B, V, C, H, W, generator=g)
    P = syn.make_projections(B, V, H, W, behind_views=(0,) if V > 2 else ())
    cv = (torch.rand(B, *G, 3, generator=g) - 0.5) * 2600.0
    fd, Pd, cvd = cuda(f, P, cv)
    if bf:
        fd = fd.bfloat16()
    full = agg.unprojection(fd, Pd, cvd, method)
    ref = torch.nn.functional.max_pool3d(full, 2)
    got = agg.unprojection(fd, Pd, cvd, method, output="max_pool2")
    assert got.shape == (B, C, G[0] // 2, G[1] // 2, G[2] // 2) and got.is_contiguous()
    assert torch.equal(got, ref)
    pixel = C * fd.element_size()
    if pixel >= 16 and pixel & (pixel - 1) == 0:                     # channels-last maps gathered in place
        fcl = fd.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)
        assert torch.equal(agg.unprojection(fcl, Pd, cvd, method, output="max_pool2"), ref)
    if G[0] >= 4:                                                    # a window of whole x-plane pairs
        yz = G[1] * G[2]
        out = torch.full_like(ref, float("nan"))
        agg.unprojection(fd, Pd, cvd, method, window=(0, B, 2 * yz, G[0] * yz), out=out, output="max_pool2")
        assert torch.equal(out[:, :, 1:], ref[:, :, 1:]) and bool(torch.isnan(out[:, :, 0]).all())


def test_fused_max_pool_full_size_and_errors():
    w = syn.CONFIGS["cfg2"]
    f, P, cv, centers = syn.make_inputs(sub_workload(w, 2))
    fd, Pd, cvd = cuda(f, P, cv)
    full = agg.unprojection(fd, Pd, cvd, "softmax")
    ref = torch.nn.functional.max_pool3d(full, 2)
    assert torch.equal(agg.unprojection(fd, Pd, cvd, "softmax", output="max_pool2"), ref)
    rots = np.stack([np.eye(3, dtype=np.float32)] * 2)
    got = agg.unprojection_grid(fd, Pd, centers.numpy(), rots, w.G, w.cuboid_side, "softmax", output="max_pool2")
    assert torch.equal(got, ref)
    assert torch.equal(agg.unprojection_grid(fd, Pd, centers.numpy(), rots, w.G, w.cuboid_side, "softmax",
                                             output="channels_last_3d"), full)
    # NaN propagates like torch's pooling
    fn = fd.clone()
    fn[0, 1, 3, 40:44, 40:44] = float("nan")
    full = agg.unprojection(fn, Pd, cvd, "sum")
    assert bool(torch.isnan(full).any())
    got = agg.unprojection(fn, Pd, cvd, "sum", output="max_pool2")
    assert torch.equal(torch.isnan(got), torch.isnan(torch.nn.functional.max_pool3d(full, 2)))
    with pytest.raises(ValueError):
        agg.unprojection(fd, Pd, cvd[:, :, :, :63], "sum", output="max_pool2")          # odd shape
    with pytest.raises(ValueError):
        agg.unprojection(fd, Pd, cvd, "sum", output="ndhwc")                            # unknown name
    with pytest.raises(ValueError):
        agg.unprojection(fd, Pd, cvd, "sum", window=(0, 2, 5, 100), out=torch.empty_like(ref), output="max_pool2")
    fr = fd.clone().requires_grad_(True)
    with pytest.raises(ValueError):
        agg.unprojection(fr, Pd, cvd, "sum", output="max_pool2")


# ---------------------------------------------------------------- reduced-precision texture path (bf16 tolerance)
SPEC_TOL_BF16 = 1e-2      # north_star: aggregated volumes within 1e-2 relative with bf16 features


def test_fast_path_cfg3_is_inside_the_bf16_tolerance(capsys):
    """BASELINE config #3 (bf16 features, 17-joint soft-argmax) through precision='fast': the texture
    units sample fp16 copies of the maps with 1.8 fixed-point weights.  The deviation from the
    reference (C restatement on the bf16-rounded maps) must stay inside the stated 1e-2; what is
    measured is printed (about 4e-3 for softmax fusion)."""
    w = syn.CONFIGS["cfg3"]
    f, P, cv, _ = syn.make_inputs(sub_workload(w, 2))
    fd, Pd, cvd = cuda(f, P, cv)
    fb = fd.bfloat16()
    errs = {}
    for method in ("sum", "mean", "max", "softmax"):
        ref = oracle.unprojection(f, P, cv, method)
        got = agg.unprojection(fb, Pd, cvd, method, precision="fast")
        assert got.shape == ref.shape and got.dtype == torch.float32
        errs[method] = rel_l2(got.cpu().numpy(), ref)
        assert errs[method] < SPEC_TOL_BF16, (method, errs[method])
        exact = agg.unprojection(fb, Pd, cvd, method)
        assert rel_l2(exact.cpu().numpy(), ref) < OUR_TOL_SOFTMAX          # the default path stays exact
    with capsys.disabled():
        print("\n[fast path, cfg3 shape] rel. l2 deviation from the reference: " +
              ", ".join("%s %.2e" % kv for kv in errs.items()))
    # soft-argmax over the fast volume: joints move by well under a voxel (39.7 mm)
    vol = agg.unprojection(fb, Pd, cvd, "softmax", precision="fast")
    ex = agg.unprojection(fb, Pd, cvd, "softmax")
    ja = agg.soft_argmax_3d(vol[:, :w.joints], cvd)
    jb = agg.soft_argmax_3d(ex[:, :w.joints], cvd)
    assert float((ja - jb).abs().max()) < 2.0


@pytest.mark.parametrize("shape", [(1, 4, 32, 24, 24, (8, 8, 32), True), (2, 8, 32, 20, 28, (6, 4, 80), True),
                                   (1, 3, 5, 12, 14, (4, 6, 10), False), (2, 2, 20, 9, 11, (2, 2, 6), True),
                                   (1, 7, 64, 16, 16, (3, 5, 34), False)])
@pytest.mark.parametrize("method", ["sum", "max", "softmax"])
def test_fast_path_shapes_windows_and_invalid_views(shape, method):
    B, V, C, H, W, G, bf = shape
    g = torch.Generator().manual_seed(V * 10 + C)
    f = torch.randn(B, V, C, H, W, generator=g)
    if bf:
        f = f.bfloat16().float()
    P = syn.make_projections(B, V, H, W, behind_views=(V - 1,) if V > 2 else ())
    cv = (torch.rand(B, *G, 3, generator=g) - 0.5) * 2600.0
    ref = oracle.unprojection(f, P, cv, method)
    fd, Pd, cvd = cuda(f, P, cv)
    if bf:
        fd = fd.bfloat16()
    got = agg.unprojection(fd, Pd, cvd, method, precision="fast")
    # max over views of noise amplifies the sampling error a little: 2e-2 on these tiny, noisy maps
    assert rel_l2(got.cpu().numpy(), ref) < (2e-2 if method == "max" else SPEC_TOL_BF16)
    assert np.array_equal(np.isnan(got.cpu().numpy()), np.isnan(ref))
    N = G[0] * G[1] * G[2]
    out = torch.full_like(got, float("nan"))
    n0, n1 = N // 4, N - 3
    agg.unprojection(fd, Pd, cvd, method, window=(0, B, n0, n1), out=out, precision="fast")
    flat = out.reshape(B, C, N)
    assert torch.equal(flat[:, :, n0:n1], got.reshape(B, C, N)[:, :, n0:n1])
    assert bool(torch.isnan(flat[:, :, :n0]).all()) and bool(torch.isnan(flat[:, :, n1:]).all())


def test_fast_path_rejects_what_it_cannot_do():
    f, P, cv = cuda(torch.zeros(1, 9, 4, 8, 8), torch.zeros(1, 9, 3, 4), torch.zeros(1, 2, 2, 2, 3))
    with pytest.raises(ValueError):
        agg.unprojection(f, P, cv, "sum", precision="fast")                 # more than 8 views
    with pytest.raises(ValueError):
        agg.unprojection(f[:, :4], P[:, :4], cv, "sum", precision="fast", output="max_pool2")
    with pytest.raises(ValueError):
        agg.unprojection(f[:, :4], P[:, :4], cv, "sum", precision="quick")
    fr = f[:, :4].clone().requires_grad_(True)
    with pytest.raises(ValueError):
        agg.unprojection(fr, P[:, :4], cv, "sum", precision="fast")


@pytest.mark.parametrize("V", [4, 8])
def test_task_chunking_does_not_change_results(monkeypatch, V):
    """The persistent CTAs take their tasks in chunks of `ychunk` consecutive y rows (chosen per launch from the
    tail it leaves; MVHMR_YCHUNK overrides): any chunking, also one that does not divide the row count, gives
    the same bits."""
    w = syn.Workload("t", B=3, V=V, C=32, H=40, W=40, G=36)
    f, P, cv, _ = syn.make_inputs(w)
    fd, Pd, cvd = cuda(f, P, cv)
    monkeypatch.delenv("MVHMR_YCHUNK", raising=False)
    ref = agg.unprojection(fd, Pd, cvd, "softmax")
    assert rel_l2(ref.cpu().numpy(), oracle.unprojection(f, P, cv, "softmax")) < OUR_TOL_SOFTMAX
    for yc in ("1", "3", "5", "7", "16", "1000"):
        monkeypatch.setenv("MVHMR_YCHUNK", yc)
        assert torch.equal(agg.unprojection(fd, Pd, cvd, "softmax"), ref), yc


def test_eight_view_launches_on_concurrent_streams_and_threads():
    """The eight-view kernels take their work from a per-launch counter inside the library (a ring of slots):
    launches in flight at the same time — several streams, several host threads — must not share one."""
    import threading
    w = syn.Workload("t", B=2, V=8, C=32, H=48, W=48, G=40)
    f, P, cv, _ = syn.make_inputs(w)
    fd, Pd, cvd = cuda(f, P, cv)
    ref = agg.unprojection(fd, Pd, cvd, "softmax")
    packed = agg.pack_features(fd)
    torch.cuda.synchronize()
    results, errors = {}, []

    def worker(k):
        try:
            st = torch.cuda.Stream()
            outs = []
            with torch.cuda.stream(st):
                for _ in range(6):
                    outs.append(agg.unprojection(fd, Pd, cvd, "softmax", packed=packed))
            st.synchronize()
            results[k] = outs
        except Exception as exc:          # noqa: BLE001
            errors.append(exc)

    threads = [threading.Thread(target=worker, args=(k,)) for k in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for outs in results.values():
        for o in outs:
            assert torch.equal(o, ref)


def test_eight_view_launch_replayed_from_a_cuda_graph():
    """The work counter of an eight-view launch is zeroed by a memset node captured in front of the kernel: a graph
    replays to the same bits every time."""
    w = syn.Workload("t", B=2, V=8, C=32, H=48, W=48, G=40)
    f, P, cv, _ = syn.make_inputs(w)
    fd, Pd, cvd = cuda(f, P, cv)
    ref = agg.unprojection(fd, Pd, cvd, "softmax")
    packed = agg.pack_features(fd)
    out = torch.empty_like(ref)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        agg.unprojection(fd, Pd, cvd, "softmax", packed=packed, out=out)
    for _ in range(3):
        out.fill_(float("nan"))
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(out, ref)
