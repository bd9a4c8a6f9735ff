"""The oracle is only trustworthy once pinned: C restatement and torch port
against the golden vectors produced by the real reference, and — where
/root/reference is mounted — against the live reference."""
import numpy as np
import pytest
import torch

import oracle
from oracle import torch_port
from multiviewhmr_b200 import synthetic as syn
from conftest import rel_l2

METHODS = ("sum", "mean", "max", "softmax")
# softmax: the C oracle uses libm expf, ATen uses Sleef -> last-bit differences
SOFTMAX_TOL = 3e-7


@pytest.mark.parametrize("case", ["unproj_ragged", "unproj_edge", "unproj_bf16"])
@pytest.mark.parametrize("method", METHODS)
def test_c_oracle_matches_golden(golden, case, method):
    z = golden(case)
    if "out_" + method not in z.files:
        pytest.skip("mode not stored for this case")
    got = oracle.unprojection(z["features"], z["proj"], z["coord_volumes"], method)
    ref = z["out_" + method]
    if method == "softmax":
        assert rel_l2(got, ref) < SOFTMAX_TOL
    else:
        assert np.array_equal(got, ref, equal_nan=True)      # bit-exact


@pytest.mark.parametrize("method", METHODS)
def test_c_oracle_cfg1_slice_and_checksums(golden, method):
    z = golden("unproj_cfg1")
    f, P, cv, c = syn.make_inputs(syn.CONFIGS["cfg1"], seed=1234)
    assert np.array_equal(c.numpy(), z["centers"]) and np.array_equal(P.numpy(), z["proj"])
    out = oracle.unprojection(f, P, cv, method)
    sl = out[:, ::4, ::3, ::3, ::3]
    if method == "softmax":
        assert rel_l2(sl, z["slice_" + method]) < SOFTMAX_TOL
    else:
        assert np.array_equal(sl, z["slice_" + method])
    assert abs(out.astype(np.float64).sum() - float(z["sum_" + method])) <= 1e-6 * abs(float(z["l2_" + method]))


def test_sampling_positions_are_bit_exact(golden):
    """Homogeneous projection of the reference (utils/multiview.py) vs the FMA chain."""
    z = golden("unproj_ragged")
    pts = z["coord_volumes"][0].reshape(-1, 3)
    got = oracle.project_points(z["proj"][0, 1], pts, euclid=False)
    assert np.array_equal(got, z["homog_b0_v1"])


@pytest.mark.parametrize("case", ["unproj_ragged", "unproj_edge", "unproj_bf16"])
@pytest.mark.parametrize("method", METHODS)
def test_torch_port_matches_golden_bitwise(golden, case, method):
    z = golden(case)
    if "out_" + method not in z.files:
        pytest.skip("mode not stored for this case")
    got = torch_port.unprojection(torch.from_numpy(z["features"]), torch.from_numpy(z["proj"]),
                                  torch.from_numpy(z["coord_volumes"]), method).numpy()
    assert np.array_equal(got, z["out_" + method], equal_nan=True)


def test_torch_port_equals_live_reference(reference):
    w = syn.Workload("t", B=2, V=3, C=8, H=20, W=28, G=12)
    f, P, cv, _ = syn.make_inputs(w, seed=5, theta=0.3, behind_views=(2,))
    for m in METHODS:
        a = torch_port.unprojection(f, P, cv, m)
        b = reference["aggregation"].unprojection(f, P, cv, m)
        assert torch.equal(a, b)
    with pytest.raises(ValueError, match="Unknown aggregation_method"):
        torch_port.unprojection(f, P, cv, "median")


def test_c_oracle_equals_live_reference_cfg1(reference):
    f, P, cv, _ = syn.make_inputs(syn.CONFIGS["cfg1"])
    for m in ("sum", "max"):
        a = oracle.unprojection(f, P, cv, m)
        b = reference["aggregation"].unprojection(f, P, cv, m).numpy()
        assert np.array_equal(a, b)
    a = oracle.unprojection(f, P, cv, "softmax")
    b = reference["aggregation"].unprojection(f, P, cv, "softmax").numpy()
    assert rel_l2(a, b) < SOFTMAX_TOL


def test_fp32_noise_floor_is_reported():
    """Reference-order fp32 vs float64 truth: the 1e-5 tolerance is norm-wise
    because the reference itself sits at ~5e-6 from the truth."""
    f, P, cv, _ = syn.make_inputs(syn.CONFIGS["cfg1"])
    o32 = oracle.unprojection(f, P, cv, "softmax")
    o64 = oracle.unprojection(f, P, cv, "softmax", truth=True)
    floor = rel_l2(o32, o64)
    assert 1e-7 < floor < 1e-5


@pytest.mark.parametrize("case", ["vg_eval_mpii", "vg_train_coco", "vg_train_mpii_rect", "vg_eval_dlt"])
def test_coord_volume_oracle_matches_volume_generator(golden, case):
    z = golden(case)
    G = z["used_coord_volumes"].shape[1]
    B = z["used_coord_volumes"].shape[0]
    training, kind = bool(z["training"]), str(z["kind"])
    np.random.seed(int(z["np_seed"]))
    axis = [0, 1, 0] if kind == "coco" else [0, 0, 1]
    rots = np.stack([syn.rotation_matrix(axis, np.random.uniform(0.0, 2 * np.pi) if training else 0.0)
                     for _ in range(B)]).astype(np.float32)
    if bool(z["use_triangulation"]):
        pytest.skip("centre comes from the DLT; covered by the GPU module test")
    centers = z["keypoints_3d"][:, 6, :3].astype(np.float32)
    got = oracle.build_coord_volumes(centers, rots, np.float32(-1250.0), np.float32(2500.0 / (G - 1)), G)
    assert np.array_equal(got, z["used_coord_volumes"])


def test_geometry_oracles(golden):
    z = golden("geometry")
    for i in range(len(z["thetas"])):
        rot = syn.rotation_matrix(z["axes"][i], float(z["thetas"][i]))
        assert np.array_equal(rot, z["rots"][i])
        got = oracle.rotate_points(z["vol"].reshape(-1, 3), rot.astype(np.float32)).reshape(z["vol"].shape)
        assert np.array_equal(got, z["rotated"][i])
    for v in range(4):
        assert np.array_equal(oracle.project_points(z["P"][v], z["pts"], euclid=False), z["homog"][v])
        assert np.array_equal(oracle.project_points(z["P"][v], z["pts"], euclid=True), z["eucl"][v])


def test_soft_argmax_truth_vs_torch_restatement():
    g = torch.Generator().manual_seed(0)
    vol = torch.randn(2, 5, 6, 7, 8, generator=g) * 3
    cv = syn.make_coord_volumes(torch.zeros(2, 3), 8)[:, :6, :7, :8].contiguous()
    a = oracle.soft_argmax_3d(vol, cv)
    b = torch_port.soft_argmax_3d(vol.double(), cv.double()).numpy()
    assert np.abs(a - b).max() < 1e-9 * 1250


def test_unknown_method_raises():
    f, P, cv, _ = syn.make_inputs(syn.Workload("t", 1, 2, 4, 8, 8, 4))
    with pytest.raises(ValueError, match="Unknown aggregation_method"):
        oracle.unprojection(f, P, cv, "median")


@pytest.mark.parametrize("method", ["sum", "mean", "max", "softmax"])
def test_c_oracle_propagates_nan_like_torch(method):
    """torch.max / sum / softmax over the view axis propagate NaN (models/aggregation.py:71-83);
    the C restatement has to as well, whichever view carries it."""
    w = syn.Workload("t", B=1, V=3, C=2, H=8, W=8, G=4)
    f, P, cv, _ = syn.make_inputs(w, seed=7)
    for v in range(3):
        fn = f.clone()
        fn[0, v, 1, 2:6, 2:6] = float("nan")
        ref = torch_port.unprojection(fn, P, cv, method).numpy()
        got = oracle.unprojection(fn, P, cv, method)
        assert np.isnan(ref).any()
        assert np.array_equal(np.isnan(got), np.isnan(ref)), (method, v)
        assert np.allclose(got, ref, rtol=1e-6, atol=1e-6, equal_nan=True)
