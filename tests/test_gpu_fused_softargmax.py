"""The fused unproject + aggregate + 3-D soft-argmax kernel (`mvhmr_unproject_aggregate_softargmax`,
BASELINE.json's target path): the stored volume has the bits of `unprojection`, the joints agree with the
two-kernel path and with the float64 truth of the oracle, with and without the volume store."""
import ctypes

import numpy as np
import pytest
import torch

import oracle
from multiviewhmr_b200 import _lib, aggregation as agg, synthetic as syn

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
SPEC_TOL = 1e-5          # soft-argmax coordinates: 1e-5 * max|coord| (BASELINE.json, fp32)


def cuda(*arrays):
    return [torch.as_tensor(a).to(DEV) for a in arrays]


def workload(B, V, C, H, W, G, dtype="fp32"):
    return syn.Workload("t", B, V, C, H, W, G, "softmax", dtype, 0, 2500.0)


def check_joints(joints, vol, cv, J, scale=SPEC_TOL):
    """against the float64 truth over the SAME stored volume"""
    truth = oracle.soft_argmax_3d(vol[:, :J].cpu(), cv)
    bound = scale * float(cv.abs().max())
    assert np.abs(joints.cpu().numpy() - truth).max() <= bound, (np.abs(joints.cpu().numpy() - truth).max(), bound)


@pytest.mark.parametrize("method", ["sum", "mean", "max", "softmax"])
def test_fused_path_cfg1_all_methods(method):
    w = syn.CONFIGS["cfg1"]
    f, P, cv, _ = syn.make_inputs(w)
    fd, Pd, cvd = cuda(f, P, cv)
    ref_vol = agg.unprojection(fd, Pd, cvd, method)
    vol, joints = agg.unprojection_soft_argmax(fd, Pd, cvd, 17, method)
    assert torch.equal(vol, ref_vol)
    two = agg.soft_argmax_3d(ref_vol[:, :17].contiguous(), cvd)
    bound = SPEC_TOL * float(cv.abs().max())
    assert (joints - two).abs().max().item() <= bound
    check_joints(joints, ref_vol, cv, 17)
    none, joints2 = agg.unprojection_soft_argmax(fd, Pd, cvd, 17, method, store_volume=False)
    assert none is None
    assert torch.equal(joints2, joints)          # the same arithmetic, only the stores are skipped


def test_fused_path_cfg3_full_size_bf16():
    """BASELINE config #3: B8 V4 C32 96x96 -> 64^3, bf16 maps, 17 joints."""
    w = syn.CONFIGS["cfg3"]
    f, P, cv, _ = syn.make_inputs(w)
    fd, Pd, cvd = cuda(f.bfloat16(), P, cv)
    ref_vol = agg.unprojection(fd, Pd, cvd, "softmax")
    vol, joints = agg.unprojection_soft_argmax(fd, Pd, cvd, 17, "softmax")
    assert torch.equal(vol, ref_vol)
    two = agg.soft_argmax_3d(ref_vol[:, :17], cvd)
    bound = SPEC_TOL * float(cv.abs().max())
    print("\n[fused soft-argmax, cfg3] max |fused - two-kernel| = %.3e mm (bound %.3e)" % ((joints - two).abs().max().item(), bound))
    assert (joints - two).abs().max().item() <= bound
    check_joints(joints[:2], ref_vol[:2], cv[:2], 17)
    _, j2 = agg.unprojection_soft_argmax(fd, Pd, cvd, 17, "softmax", store_volume=False)
    assert torch.equal(j2, joints)


@pytest.mark.parametrize("B,V,C,H,W,G,J,dtype", [
    (2, 3, 17, 40, 56, 20, 17, "fp32"),      # ragged channels, generic V, gz < 32
    (1, 8, 32, 48, 48, 40, 32, "fp32"),      # uncached V = 8 path, J = C = 32, two z segments
    (3, 4, 64, 32, 32, 33, 5, "fp32"),       # 256-byte pixels: 2 voxels per warp step
    (2, 2, 8, 24, 24, 16, 8, "bf16"),        # one 16-byte vector per pixel
    (1, 5, 40, 30, 34, 24, 21, "bf16"),
    (1, 4, 160, 16, 16, 12, 9, "fp32"),      # two channel passes (C > 128): only the first carries joints
])
def test_fused_path_ragged_shapes(B, V, C, H, W, G, J, dtype):
    w = workload(B, V, C, H, W, G, dtype)
    f, P, cv, _ = syn.make_inputs(w)
    if dtype == "bf16":
        f = f.bfloat16()
    fd, Pd, cvd = cuda(f, P, cv)
    ref_vol = agg.unprojection(fd, Pd, cvd, "softmax")
    vol, joints = agg.unprojection_soft_argmax(fd, Pd, cvd, J, "softmax")
    assert torch.equal(vol, ref_vol)
    check_joints(joints, ref_vol, cv, J)
    _, j2 = agg.unprojection_soft_argmax(fd, Pd, cvd, J, "softmax", store_volume=False)
    assert torch.equal(j2, joints)


def test_fused_path_peaky_heat_maps():
    """Feature maps scaled so that the softmax over the voxels is dominated by a handful of them:
    the running-max rescale is exercised with values far apart."""
    w = workload(2, 4, 32, 64, 64, 32)
    f, P, cv, _ = syn.make_inputs(w)
    fd, Pd, cvd = cuda(f * 60.0, P, cv)
    ref_vol = agg.unprojection(fd, Pd, cvd, "sum")
    vol, joints = agg.unprojection_soft_argmax(fd, Pd, cvd, 17, "sum")
    assert torch.equal(vol, ref_vol)
    assert torch.isfinite(joints).all()
    check_joints(joints, ref_vol, cv, 17)


def test_fused_path_over_the_generated_grid_and_packed_maps():
    w = syn.CONFIGS["cfg1"]
    f, P, cv, centers = syn.make_inputs(w)
    rots = np.stack([np.eye(3, dtype=np.float32)] * w.B)
    fd, Pd, cvd = cuda(f, P, cv)
    ref_vol = agg.unprojection(fd, Pd, cvd, "softmax")
    vol, joints = agg.unprojection_soft_argmax(fd, Pd, None, 17, "softmax",
                                               grid=(centers.numpy(), rots, w.G, w.cuboid_side))
    assert torch.equal(vol, ref_vol)
    _, j_coord = agg.unprojection_soft_argmax(fd, Pd, cvd, 17, "softmax")
    assert torch.equal(joints, j_coord)          # generated coordinates have the bits of the built volume
    _, j_packed = agg.unprojection_soft_argmax(fd, Pd, cvd, 17, "softmax", packed=agg.pack_features(fd))
    assert torch.equal(j_packed, joints)


def test_fused_path_argument_errors():
    w = syn.CONFIGS["cfg1"]
    f, P, cv, _ = syn.make_inputs(w)
    fd, Pd, cvd = cuda(f, P, cv)
    for bad in (0, 33, 40):
        with pytest.raises(ValueError):
            agg.unprojection_soft_argmax(fd, Pd, cvd, bad)
    with pytest.raises(ValueError):
        agg.unprojection_soft_argmax(fd, Pd, None, 4)
    # raw ABI: a record workspace that is too small is refused, nothing is launched
    L = _lib.load()
    B, V, C, H, Wd = fd.shape
    ws = torch.empty(L.mvhmr_unproject_workspace_bytes(_lib.F32, _lib.LAYOUT_NCHW, B, V, C, H, Wd), dtype=torch.uint8, device=DEV)
    out = torch.empty((B, 4, 3), device=DEV)
    rc = L.mvhmr_unproject_aggregate_softargmax(
        _lib.ptr(fd), _lib.F32, _lib.LAYOUT_NCHW, _lib.ptr(Pd), _lib.ptr(cvd), None, None, _lib.ptr(out), 4,
        B, V, C, H, Wd, w.G, w.G, w.G, _lib.SOFTMAX, 0, _lib.ptr(ws), ws.numel(), _lib.ptr(ws), 16,
        ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == _lib.ERR_WORKSPACE


def test_pack_features_into_a_caller_buffer():
    w = syn.CONFIGS["cfg1"]
    f, P, cv, _ = syn.make_inputs(w)
    fd, Pd, cvd = cuda(f, P, cv)
    fresh = agg.pack_features(fd)
    buf = torch.full_like(fresh, 0xAB)
    assert agg.pack_features(fd, out=buf) is buf
    assert torch.equal(buf, fresh)
    assert torch.equal(agg.unprojection(fd, Pd, cvd, "softmax", packed=buf), agg.unprojection(fd, Pd, cvd, "softmax"))
    with pytest.raises(ValueError):
        agg.pack_features(fd, out=buf[:-16])
    with pytest.raises(ValueError):
        agg.pack_features(fd, out=buf.float())


def test_fused_path_channels_last_maps_and_many_views():
    """Channels-last maps gathered in place (no pack pass) and V = 12 (view blocks with carried fusion state)."""
    w = workload(2, 4, 32, 40, 48, 24)
    f, P, cv, _ = syn.make_inputs(w)
    fd, Pd, cvd = cuda(f, P, cv)
    fcl = fd.permute(0, 1, 3, 4, 2).contiguous().permute(0, 1, 4, 2, 3)
    assert agg._is_channels_last(fcl)
    ref_vol, ref_j = agg.unprojection_soft_argmax(fd, Pd, cvd, 17, "softmax")
    vol, joints = agg.unprojection_soft_argmax(fcl, Pd, cvd, 17, "softmax")
    assert torch.equal(vol, ref_vol) and torch.equal(joints, ref_j)
    w = workload(1, 12, 16, 32, 32, 20)
    f, P, cv, _ = syn.make_inputs(w)
    fd, Pd, cvd = cuda(f, P, cv)
    ref_vol = agg.unprojection(fd, Pd, cvd, "softmax")
    vol, joints = agg.unprojection_soft_argmax(fd, Pd, cvd, 16, "softmax")
    assert torch.equal(vol, ref_vol)
    check_joints(joints, ref_vol, cv, 16)
