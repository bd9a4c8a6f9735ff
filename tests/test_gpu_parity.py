"""Parity of the CUDA path (through the Python boundary -> ctypes -> C ABI)
with the oracle and the golden vectors of the real reference.

Tolerances (BASELINE.json north_star): grids / indices bit-exact; volumes
||out-ref||/||ref|| <= 1e-5 in fp32, 1e-2 with bf16 features; soft-argmax
|d| <= 1e-5 * max|coord|.  What the kernels actually deliver is tighter and is
asserted as such: sum / mean / max and every grid are BIT-IDENTICAL to the
reference's CPU torch path; softmax differs only through ex2.approx (<1e-6).
"""
import ctypes
import os

import numpy as np
import pytest
import torch

import oracle
from oracle import torch_port
from multiviewhmr_b200 import _lib, aggregation as agg, multiview, sharding, synthetic as syn, volumetric
from conftest import rel_l2

pytestmark = pytest.mark.gpu
METHODS = ("sum", "mean", "max", "softmax")
SPEC_TOL_FP32 = 1e-5      # north_star
OUR_TOL_SOFTMAX = 1e-6    # what ex2.approx leaves us
DEV = "cuda:0"


def cuda(*arrays):
    return [torch.as_tensor(a).to(DEV) for a in arrays]


def check_volume(got, ref, method):
    got = got.cpu().numpy() if torch.is_tensor(got) else got
    if method == "softmax":
        assert rel_l2(got, ref) < OUR_TOL_SOFTMAX < SPEC_TOL_FP32
        assert np.array_equal(np.isnan(got), np.isnan(ref))
    else:
        assert np.array_equal(got, ref, equal_nan=True)


@pytest.mark.parametrize("case", ["unproj_ragged", "unproj_edge", "unproj_bf16"])
@pytest.mark.parametrize("method", METHODS)
def test_golden_vectors(golden, case, method):
    z = golden(case)
    if "out_" + method not in z.files:
        pytest.skip("mode not stored for this case")
    f, P, cv = cuda(z["features"], z["proj"], z["coord_volumes"])
    check_volume(agg.unprojection(f, P, cv, method), z["out_" + method], method)
    if case == "unproj_bf16":     # bf16 storage: same values, half the bytes
        got = agg.unprojection(f.bfloat16(), P, cv, method)
        assert got.dtype == torch.float32
        check_volume(got, z["out_" + method], method)
        assert rel_l2(got.cpu().numpy(), z["out_" + method]) < 1e-2


def channels_last(f):
    """Same (B,V,C,H,W) values, physically (B,V,H,W,C): what a channels_last 1x1 conv emits."""
    return f.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)


@pytest.mark.parametrize("method", METHODS)
def test_channels_last_maps_are_gathered_in_place(golden, method):
    """MVHMR_LAYOUT_NHWC: no pack pass, no zero border — corners outside the map
    are realised by moving the cell inside and zeroing weights.  Must equal the NCHW path bit
    for bit, including cells that straddle or leave the map (unproj_edge)."""
    z = golden("unproj_edge")
    f, P, cv = cuda(z["features"], z["proj"], z["coord_volumes"])
    B, V, C, H, W = f.shape
    Cp = 4
    while Cp < C:
        Cp *= 2
    fpad = torch.zeros(B, V, Cp, H, W, device=DEV)
    fpad[:, :, :C] = f
    fcl = channels_last(fpad)
    assert agg._is_channels_last(fcl) and not agg._is_channels_last(fpad)
    got = agg.unprojection(fcl, P, cv, method)
    assert torch.equal(got, agg.unprojection(fpad, P, cv, method))
    if "out_" + method in z.files:
        check_volume(got[:, :C], z["out_" + method], method)
    # bf16 storage and the fused-grid entry point go through the same cell logic
    w = syn.CONFIGS["cfg1"]
    f1, P1, cv1, c1 = syn.make_inputs(w, theta=0.4)
    P1 = P1.clone()
    P1[:, :, :2] *= 1.7                                    # pixel coordinates x1.7: many voxels leave the maps
    f1, P1, cv1 = cuda(f1, P1, cv1)
    for ft in (f1, f1.bfloat16()):
        fcl = channels_last(ft)
        assert agg._is_channels_last(fcl)
        assert torch.equal(agg.unprojection(fcl, P1, cv1, method), agg.unprojection(ft, P1, cv1, method))
    # tiny maps: every cell touches an edge
    g = torch.Generator().manual_seed(5)
    ft = torch.randn(2, 3, 8, 2, 3, generator=g).to(DEV)
    Pt = syn.make_projections(2, 3, 2, 3).to(DEV)
    cvt = cv1[:1, ::4, ::4, ::4].expand(2, -1, -1, -1, -1).contiguous()
    assert torch.equal(agg.unprojection(channels_last(ft), Pt, cvt, method), agg.unprojection(ft, Pt, cvt, method))


@pytest.mark.parametrize("method", METHODS)
def test_cfg1_against_oracle_and_golden_slice(golden, method):
    w = syn.CONFIGS["cfg1"]
    f, P, cv, _ = syn.make_inputs(w)
    got = agg.unprojection(*cuda(f, P, cv), method)
    assert got.shape == (1, 32, 32, 32, 32) and got.is_contiguous()
    check_volume(got, oracle.unprojection(f, P, cv, method), method)
    check_volume(got[:, ::4, ::3, ::3, ::3], golden("unproj_cfg1")["slice_" + method], method)


def test_cfg2_full_size_against_oracle_and_fp64_truth():
    w = syn.CONFIGS["cfg2"]
    f, P, cv, _ = syn.make_inputs(w)
    got = agg.unprojection(*cuda(f, P, cv), "softmax").cpu().numpy()
    ref = oracle.unprojection(f, P, cv, "softmax")
    assert rel_l2(got, ref) < OUR_TOL_SOFTMAX
    truth = oracle.unprojection(f[:2], P[:2], cv[:2], "softmax", truth=True)
    ours, theirs = rel_l2(got[:2], truth), rel_l2(ref[:2], truth)
    assert ours < SPEC_TOL_FP32 and abs(ours - theirs) < 1e-6     # same noise floor as the reference


def test_cfg3_bf16_and_soft_argmax():
    w = syn.CONFIGS["cfg3"]
    f, P, cv, _ = syn.make_inputs(w)
    fd, Pd, cvd = cuda(f, P, cv)
    vol = agg.unprojection(fd.bfloat16(), Pd, cvd, "softmax")
    ref = oracle.unprojection(f, P, cv, "softmax")
    assert rel_l2(vol.cpu().numpy(), ref) < OUR_TOL_SOFTMAX
    heat = vol[:, :w.joints].contiguous()
    got = agg.soft_argmax_3d(heat, cvd).cpu().numpy()
    truth = oracle.soft_argmax_3d(heat.cpu(), cv)
    assert np.abs(got - truth).max() <= 1e-5 * float(cv.abs().max())
    port = torch_port.soft_argmax_3d(heat.cpu(), cv).numpy()
    assert np.abs(got - truth).max() <= 2 * np.abs(port - truth).max() + 1e-6 * float(cv.abs().max())


@pytest.mark.parametrize("B,V,C,H,W,G", [
    (1, 1, 1, 5, 7, (3, 4, 5)), (2, 2, 3, 9, 6, (4, 4, 4)), (1, 3, 5, 16, 12, (2, 9, 33)),
    (1, 4, 13, 8, 8, (7, 3, 16)), (2, 5, 8, 12, 20, (5, 5, 40)), (1, 8, 64, 10, 10, (3, 3, 32)),
    (1, 9, 4, 10, 10, (2, 2, 16)), (1, 17, 12, 10, 10, (2, 3, 8)), (3, 4, 32, 24, 24, (16, 16, 16)),
    # the reference's real operating point (cfg/baseline.yaml): 256 channels, 7x7 maps, 16^3 grid
    (2, 4, 256, 7, 7, (16, 16, 16)), (1, 3, 200, 6, 5, (4, 4, 20)), (1, 2, 32, 16, 16, (3, 2, 80)),
    # many views (the voxel records shrink the z segment down to one voxel), many channels (five passes), wide maps
    (1, 64, 4, 8, 8, (3, 3, 9)), (1, 300, 4, 6, 6, (2, 2, 5)), (1, 2, 520, 4, 4, (2, 2, 6)), (1, 1, 4, 3, 700, (2, 3, 40))])
@pytest.mark.parametrize("method", METHODS)
def test_ragged_shapes(B, V, C, H, W, G, method):
    g = torch.Generator().manual_seed(B * 1000 + V * 100 + C)
    f = torch.randn(B, V, C, H, W, generator=g)
    P = syn.make_projections(B, V, H, W, behind_views=(V - 1,) if V > 2 else ())
    cv = (torch.rand(B, *G, 3, generator=g) - 0.5) * 2600.0
    got = agg.unprojection(*cuda(f, P, cv), method)
    assert got.shape == (B, C) + G
    check_volume(got, oracle.unprojection(f, P, cv, method), method)


@pytest.mark.parametrize("seed", range(12))
def test_fuzzed_shapes_layouts_and_windows(seed):
    """Seeded random problems (non-cubic grids, odd channel counts, behind-camera views, far-away
    voxels, bf16 storage, channels-last maps, shard windows) against the C restatement."""
    rng = np.random.RandomState(1000 + seed)
    B, V = int(rng.randint(1, 4)), int(rng.choice([1, 2, 3, 4, 5, 8, 11]))
    C = int(rng.choice([1, 3, 4, 8, 17, 32, 40]))
    H, W = int(rng.randint(2, 40)), int(rng.randint(2, 40))
    G = tuple(int(x) for x in rng.randint(1, 37, size=3))
    method = METHODS[seed % 4]
    bf16 = bool(rng.randint(2))
    g = torch.Generator().manual_seed(seed)
    f = torch.randn(B, V, C, H, W, generator=g)
    if bf16:
        f = f.bfloat16().float()
    P = syn.make_projections(B, V, H, W, behind_views=(0,) if rng.randint(2) else ())
    cv = (torch.rand(B, *G, 3, generator=g) - 0.5) * float(rng.choice([800.0, 2600.0, 9000.0]))
    ref = oracle.unprojection(f, P, cv, method)
    fd, Pd, cvd = cuda(f, P, cv)
    if bf16:
        fd = fd.bfloat16()
    check_volume(agg.unprojection(fd, Pd, cvd, method), ref, method)
    pixel = C * fd.element_size()
    if pixel >= 16 and pixel & (pixel - 1) == 0:                     # eligible for the in-place layout
        fcl = channels_last(fd)
        assert agg._is_channels_last(fcl)
        check_volume(agg.unprojection(fcl, Pd, cvd, method), ref, method)
    # a random shard window into a NaN-poisoned buffer: inside == reference, outside untouched
    N = G[0] * G[1] * G[2]
    b0 = int(rng.randint(0, B)); b1 = int(rng.randint(b0 + 1, B + 1))
    n0 = int(rng.randint(0, N)); n1 = int(rng.randint(n0 + 1, N + 1))
    out = torch.full((B, C) + G, float("nan"), device=DEV)
    src = channels_last(fd) if (pixel >= 16 and pixel & (pixel - 1) == 0 and seed % 2) else fd
    agg.unprojection(src, Pd, cvd, method, window=(b0, b1, n0, n1), out=out)
    got = out.cpu().numpy().reshape(B, C, N)
    check_volume(got[b0:b1, :, n0:n1], ref.reshape(B, C, N)[b0:b1, :, n0:n1], method)
    mask = np.ones((B, 1, N), bool)
    mask[b0:b1, :, n0:n1] = False
    assert np.isnan(got[np.broadcast_to(mask, got.shape)]).all()


@pytest.mark.parametrize("lz", ["1", "5", "8", "24", "32"])
def test_z_segment_length_does_not_change_results(monkeypatch, lz):
    w = syn.Workload("t", B=2, V=4, C=8, H=32, W=32, G=24)
    f, P, cv, _ = syn.make_inputs(w, seed=3)
    base = agg.unprojection(*cuda(f, P, cv), "softmax")
    monkeypatch.setenv("MVHMR_LZ", lz)
    assert torch.equal(agg.unprojection(*cuda(f, P, cv), "softmax"), base)


def test_shards_are_bitwise_identical_to_the_full_call():
    w = syn.Workload("t", B=3, V=4, C=16, H=32, W=32, G=20)
    f, P, cv, _ = syn.make_inputs(w, seed=4)
    fd, Pd, cvd = cuda(f, P, cv)
    full = agg.unprojection(fd, Pd, cvd, "softmax")
    for world in (2, 4, 8):
        out = torch.full_like(full, float("nan"))
        for r in range(world):
            sharding.unprojection_sharded(fd, Pd, cvd, "softmax", r, world, out=out)
        assert torch.equal(out, full)
    # compact slab buffers straight through the C ABI (n_origin / n_extent)
    G, N = w.G, w.G ** 3
    x0, x1 = 7, 13
    n0, n1 = x0 * G * G, x1 * G * G
    L = _lib.load()
    coord_slab = cvd[:, x0:x1].contiguous()
    out_slab = torch.empty(3, 16, x1 - x0, G, G, device=DEV)
    ws = torch.empty(L.mvhmr_unproject_workspace_bytes(_lib.F32, _lib.LAYOUT_NCHW, 3, 4, 16, 32, 32),
                     dtype=torch.uint8, device=DEV)
    _lib.check(L.mvhmr_unproject_aggregate(
        _lib.ptr(fd), _lib.F32, _lib.LAYOUT_NCHW, _lib.ptr(Pd), _lib.ptr(coord_slab), _lib.ptr(out_slab),
        3, 4, 16, 32, 32, G, G, G, _lib.SOFTMAX, 0, 3, n0, n1, n0, n1 - n0, 0,
        _lib.ptr(ws), ws.numel(), _lib.stream_ptr(torch.device(DEV))))
    assert torch.equal(out_slab, full[:, :, x0:x1])


def test_prepacked_features_and_out_buffer():
    w = syn.Workload("t", B=2, V=4, C=32, H=24, W=24, G=16)
    f, P, cv, _ = syn.make_inputs(w, seed=6)
    fd, Pd, cvd = cuda(f, P, cv)
    base = agg.unprojection(fd, Pd, cvd, "softmax")
    packed = agg.pack_features(fd)
    out = torch.empty_like(base)
    assert agg.unprojection(fd, Pd, cvd, "softmax", packed=packed, out=out) is out
    assert torch.equal(out, base)
    with pytest.raises(ValueError):
        agg.unprojection(fd, Pd, cvd, "softmax", out=torch.empty(1, device=DEV))


def test_properties_at_full_size():
    """cfg2 size (268 M voxel-channel-views): identities that need no oracle."""
    w = syn.CONFIGS["cfg2"]
    f, P, cv, _ = syn.make_inputs(w)
    fd, Pd, cvd = cuda(f, P, cv)
    s = agg.unprojection(fd, Pd, cvd, "sum")
    assert torch.equal(agg.unprojection(fd * 4.0, Pd, cvd, "sum"), s * 4.0)        # exact linearity (power of two)
    assert torch.equal(agg.unprojection(fd, Pd, cvd, "mean"), s / 4.0)             # mean == sum / V
    mx = agg.unprojection(fd, Pd, cvd, "max")
    sm = agg.unprojection(fd, Pd, cvd, "softmax")
    mean = s / 4.0
    assert bool((sm <= mx + 1e-5).all()) and bool((sm >= mean - 1e-5).all())       # max >= softmax-fused >= mean
    perm = [2, 0, 3, 1]
    assert torch.equal(agg.unprojection(fd[:, perm], Pd[:, perm], cvd, "max"), mx)  # view order is irrelevant to max
    assert torch.equal(agg.unprojection(fd, Pd, cvd, "softmax"), sm)                # deterministic
    del s, mx, sm, mean


def test_non_finite_positions_follow_the_reference():
    g = torch.Generator().manual_seed(2)
    f = torch.randn(1, 2, 4, 8, 8, generator=g)
    P = torch.tensor([[[1.0, 0, 0, 0], [0, 1.0, 0, 0], [0, 0, 1.0, 0]]]).repeat(1, 2, 1, 1)
    pts = torch.tensor([[3.0, 4.0, 1.0], [float("inf"), 2.0, 1.0], [2.0, float("nan"), 1.0],
                        [1e38, 1e38, 1e-38], [3.0, 3.0, 1e-45], [2.5, 2.5, 1.0], [1.0, 1.0, -1.0], [5.0, 5.0, 1.0]])
    cv = pts.view(1, 2, 2, 2, 3)
    for m in METHODS:
        ref = torch_port.unprojection(f, P, cv, m).numpy()
        got = agg.unprojection(*cuda(f, P, cv), m).cpu().numpy()
        assert np.array_equal(np.isnan(got), np.isnan(ref)), m
        assert np.allclose(got, ref, rtol=1e-6, atol=1e-6, equal_nan=True), m


def test_errors_on_cuda_inputs():
    f, P, cv = cuda(torch.zeros(1, 2, 4, 8, 8), torch.zeros(1, 2, 3, 4), torch.zeros(1, 2, 2, 2, 3))
    with pytest.raises(ValueError, match="Unknown aggregation_method: median"):
        agg.unprojection(f, P, cv, "median")
    with pytest.raises(ValueError):
        agg.unprojection(f, P[:, :1], cv, "sum")
    with pytest.raises(TypeError):
        agg.unprojection(f.half(), P, cv, "sum")
    with pytest.raises(ValueError):
        agg.unprojection(f, P, cv, "sum", window=(0, 2, 0, 8))
    assert agg.unprojection(f[:0], P[:0], cv[:0], "sum").shape == (0, 4, 2, 2, 2)


@pytest.mark.parametrize("case", ["vg_eval_mpii", "vg_train_coco", "vg_train_mpii_rect", "vg_eval_dlt"])
def test_volume_generator_module(golden, case):
    z = golden(case)
    B, V = z["features"].shape[:2]
    cams = [[multiview.Camera(z["cam_R"][v, b], z["cam_t"][v, b], z["cam_K"][v, b]) for b in range(B)]
            for v in range(V)]
    Hi, Wi = (int(x) for x in z["image_hw"])
    batch = {"images": np.zeros((B, V, Hi, Wi, 3), np.float32), "cameras": cams,
             "keypoints_3d": [k for k in z["keypoints_3d"]]}
    vg = agg.VolumeGenerator(volume_size=z["used_coord_volumes"].shape[1], input_channels=z["features"].shape[2],
                             output_channels=z["volumes"].shape[1], cuboid_side=2500.0,
                             use_triangulation=bool(z["use_triangulation"]), kind=str(z["kind"]), device=DEV)
    vg.load_state_dict({"process_feature.0.weight": torch.from_numpy(z["conv_weight"]),
                        "process_feature.0.bias": torch.from_numpy(z["conv_bias"])})
    vg.train(bool(z["training"]))
    vg.fuse_grid = False                             # two-step path first: exposes the coord volume
    torch.backends.cudnn.allow_tf32 = False          # the 1x1 conv is the reference's own torch op: keep it fp32
    torch.backends.cuda.matmul.allow_tf32 = False
    seen = {}
    real = agg.unprojection

    def spy(features, proj, coord, aggregation_method="softmax", **kw):
        seen.update(proj=proj.clone(), coord=coord.clone(), feats=features.clone(), method=aggregation_method)
        return real(features, proj, coord, aggregation_method=aggregation_method, **kw)
    agg.unprojection = spy
    try:
        np.random.seed(int(z["np_seed"]))
        K_before = cams[0][0].K.copy()
        with torch.no_grad():
            vol = vg(*cuda(z["features"], z["proj_in"]), batch)
        assert np.array_equal(cams[0][0].K, K_before)             # caller's cameras untouched
    finally:
        agg.unprojection = real
    assert seen["method"] == str(z["used_method"]) == "softmax"
    assert np.array_equal(seen["proj"].cpu().numpy(), z["used_proj"])                    # bit-exact
    if bool(z["use_triangulation"]):
        assert np.allclose(seen["coord"].cpu().numpy(), z["used_coord_volumes"], rtol=0, atol=1e-3)
    else:
        assert np.array_equal(seen["coord"].cpu().numpy(), z["used_coord_volumes"])      # bit-exact grid
    # the 1x1 conv is cuDNN (fp32, but not bit-equal to MKL): compare the fused op on the
    # reference's own squeezed features, then the module end to end
    fused = agg.unprojection(*cuda(z["used_features"], z["used_proj"], z["used_coord_volumes"]), "softmax")
    assert rel_l2(fused.cpu().numpy(), z["volumes"]) < OUR_TOL_SOFTMAX
    assert rel_l2(vol.cpu().numpy(), z["volumes"]) < SPEC_TOL_FP32
    assert vol.shape == z["volumes"].shape and vol.dtype == torch.float32
    # default path: the grid is generated inside the fused kernel — same bits
    vg.fuse_grid = True
    np.random.seed(int(z["np_seed"]))
    with torch.no_grad():
        vol_fused = vg(*cuda(z["features"], z["proj_in"]), batch)
    assert torch.equal(vol_fused, vol)
    # the squeeze normally runs as a batched GEMM that emits channels-last maps (gathered in place);
    # the cuDNN conv + pack route must agree to fp32 rounding of the 1x1 contraction
    assert seen["feats"].shape == z["used_features"].shape
    channels = seen["feats"].shape[2] * 4
    assert agg._is_channels_last(seen["feats"]) == (channels >= 16 and channels & (channels - 1) == 0)
    vg.channels_last = False
    np.random.seed(int(z["np_seed"]))
    with torch.no_grad():
        vol_conv = vg(*cuda(z["features"], z["proj_in"]), batch)
    assert rel_l2(vol_conv.cpu().numpy(), z["volumes"]) < SPEC_TOL_FP32
    assert rel_l2(vol_conv.cpu().numpy(), vol.cpu().numpy()) < SPEC_TOL_FP32


def test_geometry_kernels_are_bit_exact(golden):
    z = golden("geometry")
    (vol,) = cuda(z["vol"])
    for i in range(len(z["thetas"])):
        got = volumetric.rotate_coord_volume(vol, float(z["thetas"][i]), list(z["axes"][i]))
        assert got.shape == vol.shape and np.array_equal(got.cpu().numpy(), z["rotated"][i])
    P, pts = cuda(z["P"], z["pts"])
    for v in range(4):
        h = multiview.project_3d_points_to_image_plane_without_distortion(P[v], pts, convert_back_to_euclidean=False)
        e = multiview.project_3d_points_to_image_plane_without_distortion(P[v], pts)
        assert np.array_equal(h.cpu().numpy(), z["homog"][v]) and np.array_equal(e.cpu().numpy(), z["eucl"][v])
    tri = multiview.triangulate_point_from_multiple_views_linear_torch(P, cuda(z["tri_uv"])[0])
    assert np.allclose(tri.cpu().numpy(), z["tri_torch"], rtol=0, atol=1e-2)


def test_coord_volume_kernel_full_size_bit_exact():
    G, B = 64, 4
    g = np.random.default_rng(0)
    centers = (g.normal(size=(B, 3)) * 100).astype(np.float32)
    rots = np.stack([syn.rotation_matrix([0, 0, 1], t) for t in (0.0, 0.3, 2.0, 5.5)]).astype(np.float32)
    got = agg.build_coord_volumes(centers, rots, G, 2500.0, torch.device(DEV)).cpu().numpy()
    ref = oracle.build_coord_volumes(centers, rots, np.float32(-1250.0), np.float32(2500.0 / 63), G)
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("n", [1, 3, 4, 5, 1023, 4099])
def test_geometry_kernels_ragged_sizes_and_raw_abi(n):
    """The streaming kernels take 4 points per thread with 16-byte accesses; point counts that are
    not multiples of 4, grids with N % 4 != 0 and misaligned buffers take the scalar paths."""
    L = _lib.load()
    dev = torch.device(DEV)
    g = np.random.default_rng(n)
    pts = (g.normal(size=(n + 1, 3)) * 700).astype(np.float32)
    rot = syn.rotation_matrix([0.3, -0.5, 0.8], 1.1).astype(np.float32)
    P = syn.ring_projection(0, 1, 4, 96, 96, 2500.0).astype(np.float32)
    dpts = torch.from_numpy(pts).to(dev)
    for skip in (0, 1):                              # skip=1: a view that starts 12 bytes into the buffer
        src = dpts.reshape(-1)[3 * skip:3 * skip + 3 * n]
        ref_src = pts[skip:skip + n]
        out = torch.empty(n + 1, 3, device=dev).reshape(-1)[3 * skip:3 * skip + 3 * n]
        _lib.check(L.mvhmr_rotate_points(_lib.ptr(out), _lib.ptr(src), _lib.host3(rot.reshape(-1).tolist()), n,
                                         _lib.stream_ptr(dev)))
        assert np.array_equal(out.cpu().numpy().reshape(n, 3), oracle.rotate_points(ref_src, rot))
        Pd = torch.from_numpy(P).to(dev)
        for euclid in (0, 1):
            o = torch.empty(n * (2 if euclid else 3) + 4, device=dev)[(4 * skip):]
            _lib.check(L.mvhmr_project_points(_lib.ptr(o), _lib.ptr(Pd), _lib.ptr(src), n, euclid, _lib.stream_ptr(dev)))
            got = o[: n * (2 if euclid else 3)].cpu().numpy().reshape(n, -1)
            assert np.array_equal(got, oracle.project_points(P, ref_src, euclid=bool(euclid)), equal_nan=True)
    # coord volumes on a grid whose voxel count is not a multiple of 4 (and one that is)
    for G in ((3, 5, 7), (2, 6, 9), (4, 3, 8)):
        B = 3
        centers = (g.normal(size=(B, 3)) * 100).astype(np.float32)
        rots = np.stack([syn.rotation_matrix([0, 1, 0], t) for t in (0.0, 0.7, 4.0)]).astype(np.float32)
        out = torch.empty(B, *G, 3, device=dev)
        cen, rt = torch.from_numpy(centers).to(dev), torch.from_numpy(rots).to(dev)
        pos, step = [-1250.0, -1000.0, -700.0], [33.0, 41.5, 57.25]
        _lib.check(L.mvhmr_build_coord_volumes(_lib.ptr(out), _lib.ptr(cen), _lib.ptr(rt), _lib.host3(pos), _lib.host3(step),
                                               B, G[0], G[1], G[2], _lib.stream_ptr(dev)))
        ref = oracle.build_coord_volumes(centers, rots, np.float32(pos), np.float32(step), G)
        assert np.array_equal(out.cpu().numpy(), ref)


@pytest.mark.parametrize("B,J,G", [(2, 17, (16, 16, 16)), (1, 3, (5, 7, 9)), (1, 70, (4, 8, 8)), (2, 1, (3, 3, 3)),
                                   (1, 2, (40, 40, 40))])
def test_soft_argmax(B, J, G):
    g = torch.Generator().manual_seed(J)
    vol = torch.randn(B, J, *G, generator=g) * 5.0
    cv = (torch.rand(B, *G, 3, generator=g) - 0.5) * 2500.0 + 300.0
    got = agg.soft_argmax_3d(*cuda(vol, cv)).cpu().numpy()
    truth = oracle.soft_argmax_3d(vol, cv)
    assert got.shape == (B, J, 3)
    assert np.abs(got - truth).max() <= 1e-5 * float(cv.abs().max())
    # slab records concatenate: three ragged x-slabs == whole volume
    cuts = [0, max(1, G[0] // 3), max(2, G[0] // 2), G[0]] if G[0] >= 3 else [0, G[0]]
    recs = [agg.soft_argmax_3d_records(*cuda(vol[:, :, a:b].contiguous(), cv[:, a:b].contiguous()))
            for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
    merged = agg.soft_argmax_3d_from_records(torch.cat(recs, dim=2)).cpu().numpy()
    assert np.abs(merged - truth).max() <= 1e-5 * float(cv.abs().max())
    host = sharding.merge_softargmax_records(torch.cat(recs, dim=2).cpu().double()).numpy()
    assert np.abs(host - truth).max() <= 1e-5 * float(cv.abs().max())


def test_soft_argmax_reads_a_channel_slice_in_place():
    """`vol[:, :J]` of a wider aggregate (cfg3: 17 joints out of 32 channels) is not copied: the ABI takes
    the sample stride."""
    g = torch.Generator().manual_seed(2)
    vol = (torch.randn(3, 8, 6, 5, 4, generator=g) * 4).to(DEV)
    cv = ((torch.rand(3, 6, 5, 4, 3, generator=g) - 0.5) * 2000).to(DEV)
    for J in (1, 3, 8):
        sl = vol[:, :J]
        assert sl.is_contiguous() == (J == 8)
        assert torch.equal(agg.soft_argmax_3d(sl, cv), agg.soft_argmax_3d(sl.contiguous(), cv))
    assert torch.equal(agg.soft_argmax_3d(vol[:1, 2:5], cv[:1]), agg.soft_argmax_3d(vol[:1, 2:5].contiguous(), cv[:1]))
    # a slice that is not a leading-channel window is still correct (copied)
    assert torch.equal(agg.soft_argmax_3d(vol[:, ::2], cv), agg.soft_argmax_3d(vol[:, ::2].contiguous(), cv))
    L = _lib.load()
    one = torch.zeros(4, device=DEV)
    rc = L.mvhmr_soft_argmax3d_strided(_lib.ptr(vol), _lib.ptr(cv), _lib.ptr(one), 3, 8, 120, 100, _lib.ptr(vol), 1 << 20, None)
    assert rc == _lib.ERR_INVALID_ARGUMENT and b"sample stride" in L.mvhmr_last_error()


def test_soft_argmax_over_the_generated_grid_has_the_bits_of_the_coord_volume_path():
    g = np.random.default_rng(3)
    for B, J, G in ((3, 5, 12), (2, 17, 33), (1, 2, 7)):
        centers = (g.normal(size=(B, 3)) * 120).astype(np.float32)
        rots = np.stack([syn.rotation_matrix([0, 0, 1], t) for t in g.uniform(0, 6.28, size=B)]).astype(np.float32)
        vol = torch.from_numpy(g.normal(size=(B, J + 3, G, G, G)).astype(np.float32) * 4).to(DEV)
        cv = agg.build_coord_volumes(centers, rots, G, 2500.0, torch.device(DEV))
        ref = agg.soft_argmax_3d(vol[:, :J], cv)
        got = agg.soft_argmax_3d_grid(vol[:, :J], centers, rots, 2500.0)
        assert torch.equal(got, ref)


def test_soft_argmax_extreme_logits():
    vol = torch.full((1, 2, 4, 4, 4), -1e4)
    vol[0, 0, 1, 2, 3] = 80.0            # one-hot after softmax
    vol[0, 1] = 1e4                      # uniform, large
    cv = syn.make_coord_volumes(torch.tensor([[10.0, -20.0, 30.0]]), 4)
    got = agg.soft_argmax_3d(*cuda(vol, cv)).cpu().numpy()
    assert np.allclose(got[0, 0], cv[0, 1, 2, 3].numpy(), atol=1e-3)
    assert np.allclose(got[0, 1], cv[0].reshape(-1, 3).mean(0).numpy(), atol=1e-2)


def test_cuda_graph_capture_and_side_stream():
    w = syn.Workload("t", B=2, V=4, C=16, H=24, W=24, G=16)
    f, P, cv, _ = syn.make_inputs(w, seed=8)
    fd, Pd, cvd = cuda(f, P, cv)
    base = agg.unprojection(fd, Pd, cvd, "softmax")
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        packed = agg.pack_features(fd)
        out = torch.empty_like(base)
        agg.unprojection(fd, Pd, cvd, "softmax", packed=packed, out=out)      # warm-up on the side stream
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            agg.unprojection(fd, Pd, cvd, "softmax", packed=packed, out=out)
    torch.cuda.current_stream().wait_stream(side)
    out.zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, base)


def test_deviation_from_the_torch_cuda_path_is_reported():
    """The reference run on a GPU goes through cuBLAS / ATen-CUDA, which round
    differently from its CPU path (tensor / python-int becomes a multiply by the
    reciprocal there).  Our contract is the CPU path; this documents the gap."""
    w = syn.Workload("t", B=1, V=4, C=16, H=64, W=64, G=24)
    f, P, cv, _ = syn.make_inputs(w, seed=10)
    fd, Pd, cvd = cuda(f, P, cv)
    theirs = torch_port.unprojection(fd, Pd, cvd, "softmax").cpu().numpy()
    ours = agg.unprojection(fd, Pd, cvd, "softmax").cpu().numpy()
    truth = oracle.unprojection(f, P, cv, "softmax", truth=True)
    gap = rel_l2(ours, theirs)
    print("torch-CUDA path vs ours %.3g | torch-CUDA vs fp64 %.3g | ours vs fp64 %.3g"
          % (gap, rel_l2(theirs, truth), rel_l2(ours, truth)))
    assert gap < 5e-5


@pytest.mark.parametrize("method", METHODS)
def test_backward_matches_torch_autograd(method):
    """Gradient w.r.t. the features vs autograd through the reference's torch ops (CPU)."""
    w = syn.Workload("t", B=2, V=3, C=5, H=14, W=18, G=9)
    f, P, cv, _ = syn.make_inputs(w, seed=12, theta=0.4, behind_views=(2,))
    g = torch.Generator().manual_seed(3)
    gout = torch.randn(2, 5, 9, 9, 9, generator=g)
    fr = f.clone().requires_grad_(True)
    torch_port.unprojection(fr, P, cv, method).backward(gout)
    fd, Pd, cvd, gd = cuda(f, P, cv, gout)
    fd.requires_grad_(True)
    out = agg.unprojection(fd, Pd, cvd, method)
    assert out.requires_grad
    out.backward(gd)
    assert fd.grad.shape == f.shape and fd.grad.dtype == torch.float32
    assert rel_l2(fd.grad.cpu().numpy(), fr.grad.numpy()) < 1e-5
    if method in ("sum", "mean"):      # views behind the camera receive no gradient
        assert float(fd.grad[:, 2].abs().max()) < float(fd.grad[:, 0].abs().max())


@pytest.mark.parametrize("shape", [(2, 4, 32, 24, 24, 12, False), (1, 3, 5, 9, 13, 7, False), (1, 11, 8, 16, 16, 6, False),
                                   (2, 4, 16, 16, 20, 9, True), (1, 6, 24, 12, 12, 5, True), (1, 2, 136, 8, 8, 5, False)])
@pytest.mark.parametrize("method", METHODS)
def test_backward_fast_path_against_simple_kernel_and_autograd(shape, method):
    """The packed / vector-red backward against (a) the thread-per-voxel kernel, (b) torch autograd
    through the reference's op sequence; V > 8 exercises the re-sampling pass, C = 136 two channel
    passes, odd shapes the ragged tiles."""
    from multiviewhmr_b200 import autograd as ag
    B, V, C, H, W, G, bf16 = shape
    g = torch.Generator().manual_seed(B * 100 + V * 10 + C)
    f = torch.randn(B, V, C, H, W, generator=g)
    if bf16:
        f = f.bfloat16().float()
    P = syn.make_projections(B, V, H, W, behind_views=(V - 1,) if V > 2 else ())
    cv = (torch.rand(B, G, G + 1, G + 2, 3, generator=g) - 0.5) * 2400.0
    gout = torch.randn(B, C, G, G + 1, G + 2, generator=g)
    fr = f.clone().requires_grad_(True)
    torch_port.unprojection(fr, P, cv, method).backward(gout)
    fd, Pd, cvd, gd = cuda(f, P, cv, gout)
    if bf16:
        fd = fd.bfloat16()
    fast = ag.unprojection_backward(gd, fd, Pd, cvd, method)
    simple = ag.unprojection_backward(gd, fd, Pd, cvd, method, simple=True)
    assert fast.shape == f.shape and fast.dtype == torch.float32
    assert rel_l2(fast.cpu().numpy(), fr.grad.numpy()) < 1e-5
    assert rel_l2(fast.cpu().numpy(), simple.cpu().numpy()) < 1e-5
    if V > 2 and method in ("sum", "mean"):      # views behind the camera receive (almost) no gradient
        assert float(fast[:, V - 1].abs().max()) < float(fast[:, 0].abs().max())


def test_backward_bf16_features_and_module_training_step():
    w = syn.Workload("t", B=1, V=4, C=8, H=16, W=16, G=8)
    f, P, cv, _ = syn.make_inputs(w, seed=13)
    fb = f.bfloat16()
    fr = fb.float().clone().requires_grad_(True)
    torch_port.unprojection(fr, P, cv, "softmax").sum().backward()
    fd, Pd, cvd = cuda(fb, P, cv)
    fd.requires_grad_(True)
    agg.unprojection(fd, Pd, cvd, "softmax").sum().backward()
    assert fd.grad.dtype == torch.bfloat16
    assert rel_l2(fd.grad.float().cpu().numpy(), fr.grad.numpy()) < 1e-2
    # a training step through the module: gradients reach the 1x1 conv
    B, V, Cin, G = 2, 3, 6, 6
    rng = np.random.default_rng(0)
    cams = [[multiview.Camera(np.eye(3), [0.0, 0.0, 4000.0 + 100 * v], [[300.0, 0, 32], [0, 300.0, 32], [0, 0, 1]])
             for _ in range(B)] for v in range(V)]
    batch = {"images": np.zeros((B, V, 64, 64, 3), np.float32), "cameras": cams,
             "keypoints_3d": [rng.normal(size=(17, 3)) * 50 for _ in range(B)]}
    vg = agg.VolumeGenerator(volume_size=G, input_channels=Cin, output_channels=4, cuboid_side=2000.0, device=DEV)
    vg.train()
    np.random.seed(5)
    feats = torch.randn(B, V, Cin, 16, 16, device=DEV)
    vol = vg(feats, torch.zeros(B, V, 3, 4, device=DEV), batch)
    vol.square().mean().backward()
    gw = vg.process_feature[0].weight.grad
    assert gw is not None and bool(torch.isfinite(gw).all()) and float(gw.abs().sum()) > 0


@pytest.mark.parametrize("d", [7.0, 10.0, 12.0, 14.0, 16.0, 24.0, 56.0, 64.0, 96.0, 100.0, 128.0, 224.0, 3.0, 4.0, 8.0, 4503.7])
def test_exact_division_shortcuts_exhaustively(d):
    """The kernel divides by H / W / V with reciprocal + FMA correction and shares one reciprocal
    between x/w and y/w.  Both must equal IEEE division bit for bit: checked for all 2^32 numerators."""
    bad = torch.zeros(1, dtype=torch.int64, device=DEV)
    _lib.check(_lib.load().mvhmr_selftest_division(d, _lib.ptr(bad), _lib.stream_ptr(torch.device(DEV))))
    assert int(bad.item()) == 0


def _guarded(nbytes, dev, fill=0x5A, guard=4096):
    """A uint8 buffer of nbytes with guard bands on both sides (16-byte aligned payload)."""
    total = torch.full((guard + nbytes + guard + 16,), fill, dtype=torch.uint8, device=dev)
    return total, total[guard:guard + nbytes]


def _guards_intact(total, nbytes, fill=0x5A, guard=4096):
    return bool((total[:guard] == fill).all()) and bool((total[guard + nbytes:] == fill).all())


@pytest.mark.parametrize("shape", [(2, 3, 5, 9, 13, (3, 5, 7), False), (1, 4, 32, 16, 16, (8, 8, 33), False),
                                   (2, 8, 12, 10, 6, (5, 4, 9), True), (1, 2, 136, 4, 4, (2, 3, 5), False)])
def test_no_write_outside_the_callers_buffers(shape):
    """compute-sanitizer is not available on this pool: every output / workspace buffer of the raw C ABI
    is placed between guard bands instead, which must come back untouched (forward with NCHW / packed /
    channels-last input, soft-argmax, coord volume, both backward kernels)."""
    B, V, C, H, W, G, bf16 = shape
    L = _lib.load()
    dev = torch.device(DEV)
    st = _lib.stream_ptr(dev)
    g = torch.Generator().manual_seed(C)
    f = torch.randn(B, V, C, H, W, generator=g)
    P = syn.make_projections(B, V, H, W)
    cv = (torch.rand(B, *G, 3, generator=g) - 0.5) * 2600.0
    fd, Pd, cvd = cuda(f, P, cv)
    if bf16:
        fd = fd.bfloat16()
    dt = _lib.BF16 if bf16 else _lib.F32
    N = G[0] * G[1] * G[2]
    ref = agg.unprojection(fd, Pd, cvd, "softmax")
    # forward, NCHW input: output + workspace guarded
    ws_bytes = L.mvhmr_unproject_workspace_bytes(dt, _lib.LAYOUT_NCHW, B, V, C, H, W)
    out_t, out_b = _guarded(B * C * N * 4, dev)
    ws_t, ws_b = _guarded(ws_bytes, dev)
    tail = (B, V, C, H, W, G[0], G[1], G[2], _lib.SOFTMAX, 0, B, 0, N, 0, N, 0)
    _lib.check(L.mvhmr_unproject_aggregate(_lib.ptr(fd), dt, _lib.LAYOUT_NCHW, _lib.ptr(Pd), _lib.ptr(cvd), _lib.ptr(out_b),
                                           *tail, _lib.ptr(ws_b), ws_bytes, st))
    torch.cuda.synchronize()
    assert _guards_intact(out_t, B * C * N * 4) and _guards_intact(ws_t, ws_bytes)
    assert torch.equal(out_b.view(torch.float32).view(B, C, *G), ref)
    # forward from the packed planes (no workspace) and, where eligible, channels-last in place
    out_t.fill_(0x5A)
    _lib.check(L.mvhmr_unproject_aggregate(_lib.ptr(ws_b), dt, _lib.LAYOUT_PACKED, _lib.ptr(Pd), _lib.ptr(cvd), _lib.ptr(out_b),
                                           *tail, None, 0, st))
    torch.cuda.synchronize()
    assert _guards_intact(out_t, B * C * N * 4) and torch.equal(out_b.view(torch.float32).view(B, C, *G), ref)
    pixel = C * fd.element_size()
    if pixel >= 16 and pixel & (pixel - 1) == 0:
        fcl = fd.permute(0, 1, 3, 4, 2).contiguous()
        out_t.fill_(0x5A)
        _lib.check(L.mvhmr_unproject_aggregate(_lib.ptr(fcl), dt, _lib.LAYOUT_NHWC, _lib.ptr(Pd), _lib.ptr(cvd), _lib.ptr(out_b),
                                               *tail, None, 0, st))
        torch.cuda.synchronize()
        assert _guards_intact(out_t, B * C * N * 4) and torch.equal(out_b.view(torch.float32).view(B, C, *G), ref)
    # soft-argmax: records workspace and the (B,C,3) result
    sa_ws = L.mvhmr_soft_argmax3d_workspace_bytes(B, C, N)
    saw_t, saw_b = _guarded(sa_ws, dev)
    sao_t, sao_b = _guarded(B * C * 3 * 4, dev)
    _lib.check(L.mvhmr_soft_argmax3d(_lib.ptr(ref), _lib.ptr(cvd), _lib.ptr(sao_b), B, C, N, _lib.ptr(saw_b), sa_ws, st))
    torch.cuda.synchronize()
    assert _guards_intact(saw_t, sa_ws) and _guards_intact(sao_t, B * C * 3 * 4)
    # coord volume
    cvo_t, cvo_b = _guarded(B * N * 3 * 4, dev)
    cen = torch.zeros(B, 3, device=dev)
    rot = torch.eye(3, device=dev).repeat(B, 1, 1).contiguous()
    _lib.check(L.mvhmr_build_coord_volumes(_lib.ptr(cvo_b), _lib.ptr(cen), _lib.ptr(rot), _lib.host3([-1250.0] * 3), _lib.host3([40.0] * 3),
                                           B, G[0], G[1], G[2], st))
    torch.cuda.synchronize()
    assert _guards_intact(cvo_t, B * N * 3 * 4)
    # backward: fast path (gradient + workspace) and the simple kernel
    gout = torch.randn(B, C, N, device=dev)
    bws = L.mvhmr_unproject_backward_workspace_bytes(dt, B, V, C, H, W, _lib.SOFTMAX)
    gw_t, gw_b = _guarded(bws, dev)
    gf_t, gf_b = _guarded(B * V * C * H * W * 4, dev)
    _lib.check(L.mvhmr_unproject_aggregate_backward_ws(_lib.ptr(gout), _lib.ptr(fd), dt, _lib.ptr(Pd), _lib.ptr(cvd), _lib.ptr(gf_b),
                                                       B, V, C, H, W, N, _lib.SOFTMAX, _lib.ptr(gw_b), bws, st))
    torch.cuda.synchronize()
    assert _guards_intact(gw_t, bws) and _guards_intact(gf_t, B * V * C * H * W * 4)
    fast = gf_b.view(torch.float32).clone()
    gf_b.zero_()
    _lib.check(L.mvhmr_unproject_aggregate_backward(_lib.ptr(gout), _lib.ptr(fd), dt, _lib.ptr(Pd), _lib.ptr(cvd), _lib.ptr(gf_b),
                                                    B, V, C, H, W, N, _lib.SOFTMAX, st))
    torch.cuda.synchronize()
    assert _guards_intact(gf_t, B * V * C * H * W * 4)
    assert rel_l2(fast.cpu().numpy(), gf_b.view(torch.float32).cpu().numpy()) < 1e-5


def test_voxel_indices_beyond_int32_through_a_compact_slab():
    """Maximum sizes: a 2000 x 2000 x 1000 grid has 4e9 voxels (> 2^32 / 2).  The ABI lets a shard pass
    only its slab of the coordinate / output buffers (n_origin, n_extent), so the 64-bit index path can be
    exercised without 50 GB of memory: a ragged window near the end of the grid must give the bits of a
    small call over the same coordinates."""
    L = _lib.load()
    dev = torch.device(DEV)
    gx, gy, gz = 2000, 2000, 1000
    n0 = (1999 * gy + 1000) * gz + 123
    ext = 70001
    assert n0 > 2 ** 31 and n0 + ext <= gx * gy * gz
    B, V, C, H, W = 1, 2, 4, 8, 8
    g = torch.Generator().manual_seed(11)
    f = torch.randn(B, V, C, H, W, generator=g).to(dev)
    P = syn.make_projections(B, V, H, W).to(dev)
    cv = ((torch.rand(B, ext, 3, generator=g) - 0.5) * 2400.0).to(dev)
    ref = agg.unprojection(f, P, cv.view(B, 1, 1, ext, 3), "softmax").view(B, C, ext)
    out = torch.full((B, C, ext), float("nan"), device=dev)
    ws_bytes = L.mvhmr_unproject_workspace_bytes(_lib.F32, _lib.LAYOUT_NCHW, B, V, C, H, W)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    _lib.check(L.mvhmr_unproject_aggregate(
        _lib.ptr(f), _lib.F32, _lib.LAYOUT_NCHW, _lib.ptr(P), _lib.ptr(cv), _lib.ptr(out),
        B, V, C, H, W, gx, gy, gz, _lib.SOFTMAX, 0, B, n0, n0 + ext, n0, ext, 0,
        _lib.ptr(ws), ws_bytes, _lib.stream_ptr(dev)))
    assert torch.equal(out, ref)
