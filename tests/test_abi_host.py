"""CPU-side checks: the C-ABI library loads and exports everything the header
declares, argument validation works without a GPU, and the host-side mirror of
the reference interface (cameras, numpy geometry, cfg plumbing) matches the
golden vectors."""
import ctypes
import os
import re
import sys
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from multiviewhmr_b200 import _lib, aggregation, dropin, multiview, volumetric

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "mvhmr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mvhmr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built_lib):
    names = header_functions()
    assert len(names) >= 14
    L = ctypes.CDLL(built_lib)
    for n in names:
        assert hasattr(L, n), "missing export " + n
    assert sorted(_lib.SIGNATURES) == names, "python binding and header disagree"


def test_abi_version_and_sizes(built_lib):
    L = _lib.load()
    assert L.mvhmr_abi_version() == 1
    # (B*V, H+4, W+4, vectors) with 16-byte channel vectors, count padded to a power of two
    assert L.mvhmr_packed_bytes(_lib.F32, 32, 32, 96, 96) == 32 * 8 * 100 * 100 * 16
    assert L.mvhmr_packed_bytes(_lib.BF16, 32, 32, 96, 96) == 32 * 4 * 100 * 100 * 16
    assert L.mvhmr_packed_bytes(_lib.F32, 2, 5, 7, 9) == 2 * 2 * 11 * 13 * 16
    assert L.mvhmr_packed_bytes(_lib.F32, 1, 20, 7, 9) == 8 * 11 * 13 * 16
    assert L.mvhmr_packed_bytes(7, 2, 5, 7, 9) == 0
    assert L.mvhmr_unproject_workspace_bytes(_lib.F32, _lib.LAYOUT_PACKED, 8, 4, 32, 96, 96) == 0
    assert L.mvhmr_soft_argmax3d_num_slices(64 ** 3) == 512
    assert L.mvhmr_soft_argmax3d_workspace_bytes(8, 17, 64 ** 3) == 8 * 17 * 512 * 5 * 4


def test_argument_validation_needs_no_gpu(built_lib):
    L = _lib.load()
    one = ctypes.c_void_p(16)          # never dereferenced: validation fails first
    args = [one, _lib.F32, _lib.LAYOUT_NCHW, one, one, one, 1, 4, 32, 64, 64, 32, 32, 32]
    rc = L.mvhmr_unproject_aggregate(*args, 9, 0, 1, 0, 32 ** 3, 0, 32 ** 3, 0, one, 1 << 30, None)
    assert rc == _lib.ERR_INVALID_ARGUMENT
    assert b"Unknown aggregation_method" in L.mvhmr_last_error()
    with pytest.raises(ValueError, match="Unknown aggregation_method"):
        _lib.check(rc)
    rc = L.mvhmr_unproject_aggregate(*args, _lib.SUM, 0, 2, 0, 32 ** 3, 0, 32 ** 3, 0, one, 1 << 30, None)
    assert rc == _lib.ERR_INVALID_ARGUMENT and b"shard window" in L.mvhmr_last_error()
    rc = L.mvhmr_unproject_aggregate(*args, _lib.SUM, 0, 1, 0, 32 ** 3, 0, 32 ** 3, 0, None, 0, None)
    assert rc == _lib.ERR_WORKSPACE
    with pytest.raises(RuntimeError, match="workspace"):
        _lib.check(rc)
    rc = L.mvhmr_unproject_aggregate(*args, _lib.SUM, 0, 1, 0, 32 ** 3, 0, 32 ** 3, 65,
                                     one, 1 << 30, None)
    assert rc == _lib.ERR_INVALID_ARGUMENT and b"tile_hint" in L.mvhmr_last_error()
    # more views than the per-warp voxel records can hold in shared memory: refused, not mis-launched
    many = [one, _lib.F32, _lib.LAYOUT_NCHW, one, one, one, 1, 1024, 4, 8, 8, 4, 4, 4]
    rc = L.mvhmr_unproject_aggregate(*many, _lib.SUM, 0, 1, 0, 64, 0, 64, 0, one, 1 << 30, None)
    assert rc == _lib.ERR_INVALID_ARGUMENT and b"shared memory" in L.mvhmr_last_error()
    # channels-last input: gathered in place, so the pixel must be 16 * 2^k bytes and the map >= 2x2
    assert L.mvhmr_unproject_workspace_bytes(_lib.F32, _lib.LAYOUT_NHWC, 8, 4, 32, 96, 96) == 0
    nhwc = [one, _lib.F32, _lib.LAYOUT_NHWC, one, one, one, 1, 4, 17, 64, 64, 32, 32, 32]
    rc = L.mvhmr_unproject_aggregate(*nhwc, _lib.SUM, 0, 1, 0, 32 ** 3, 0, 32 ** 3, 0, None, 0, None)
    assert rc == _lib.ERR_INVALID_ARGUMENT and b"channels-last" in L.mvhmr_last_error()
    nhwc[8:11] = [32, 1, 64]
    rc = L.mvhmr_unproject_aggregate(*nhwc, _lib.SUM, 0, 1, 0, 32 ** 3, 0, 32 ** 3, 0, None, 0, None)
    assert rc == _lib.ERR_INVALID_ARGUMENT and b"H, W >= 2" in L.mvhmr_last_error()
    assert L.mvhmr_soft_argmax3d(one, one, one, 1, 1, 64, None, 0, None) == _lib.ERR_WORKSPACE
    assert L.mvhmr_build_coord_volumes(one, one, one, None, None, 1, 4, 4, 4, None) == _lib.ERR_INVALID_ARGUMENT
    # empty problems succeed without touching the device
    assert L.mvhmr_rotate_points(None, None, None, 0, None) == _lib.OK
    assert L.mvhmr_unproject_aggregate(*args, _lib.SUM, 0, 0, 0, 32 ** 3, 0, 32 ** 3, 0, None, 0, None) == _lib.OK


def test_cpu_tensors_fail_loudly(built_lib):
    f = torch.zeros(1, 2, 4, 8, 8)
    P = torch.zeros(1, 2, 3, 4)
    cv = torch.zeros(1, 2, 2, 2, 3)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        aggregation.unprojection(f, P, cv, "sum")
    with pytest.raises(ValueError, match="Unknown aggregation_method: median"):
        aggregation.unprojection(f, P, cv, "median")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        volumetric.rotate_coord_volume(cv, 0.3, [0, 0, 1])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        multiview.project_3d_points_to_image_plane_without_distortion(P[0, 0], torch.zeros(5, 3))
    with pytest.raises(TypeError, match="Works only with numpy arrays and PyTorch tensors"):
        multiview.project_3d_points_to_image_plane_without_distortion([[1.0]], [[1.0]])
    with pytest.raises(TypeError):
        multiview.euclidean_to_homogeneous([1, 2, 3])


def test_missing_library_is_an_error(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "multiviewhmr_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "mvhmr_oracle" not in text, f


def test_camera_and_numpy_geometry_match_golden(golden):
    z = golden("geometry")
    cam = multiview.Camera(np.eye(3), [1.0, 2.0, 3.0], [[1100.0, 0, 512.0], [0, 1090.0, 384.0], [0, 0, 1]])
    cam.update_after_crop((100, 50, 900, 700))
    cam.update_after_resize((650, 800), (96, 128))
    assert np.array_equal(cam.K, z["cam_K"])
    assert np.array_equal(cam.projection, z["cam_proj"]) and np.array_equal(cam.extrinsics, z["cam_extr"])
    for i in range(len(z["thetas"])):
        assert np.array_equal(volumetric.get_rotation_matrix(z["axes"][i], float(z["thetas"][i])), z["rots"][i])
    assert np.array_equal(volumetric.get_rotation_matrix([0, 0, 1], 0.0), np.eye(3))
    got = multiview.project_3d_points_to_image_plane_without_distortion(
        z["P"][0].astype(np.float64), z["pts"].astype(np.float64))
    assert np.array_equal(got, z["eucl_np"])
    tri = multiview.triangulate_point_from_multiple_views_linear(z["P"].astype(np.float64), z["tri_uv"].astype(np.float64))
    assert np.allclose(tri, z["tri_numpy"], rtol=0, atol=1e-9)
    tri_t = multiview.triangulate_point_from_multiple_views_linear_torch(
        torch.from_numpy(z["P"]), torch.from_numpy(z["tri_uv"]))
    assert np.array_equal(tri_t.numpy(), z["tri_torch"])
    pts = torch.from_numpy(z["pts"])
    h = multiview.euclidean_to_homogeneous(pts)
    assert h.shape == (257, 4) and torch.equal(multiview.homogeneous_to_euclidean(h), pts)


def _cfg(method="mean", deconv_layers=3):
    return SimpleNamespace(
        MODEL=SimpleNamespace(
            BACKBONE=SimpleNamespace(DECONV_FILTERS=[256, 256, 64], DECONV_LAYERS=deconv_layers),
            AGGREGATION=SimpleNamespace(VOLUME_SIZE=16, OUTPUT_CHANNELS=8, CUBOID_SIDE=2000.0,
                                        USE_TRIANGULATION=False, METHOD=method)),
        DATASET=SimpleNamespace(KIND="coco", TYPE="human36m"))


def test_build_volume_generator_keeps_reference_quirk():
    """cfg METHOD is swallowed by **kwargs in the reference: always 'softmax'."""
    torch.manual_seed(0)
    vg = aggregation.VolumeGenerator(volume_size=4, input_channels=6, output_channels=2, device="cpu")
    assert sorted(vg.state_dict()) == ["process_feature.0.bias", "process_feature.0.weight"]
    real_init = aggregation.VolumeGenerator.__init__

    def cpu_init(self, *a, **k):
        k["device"] = "cpu"
        real_init(self, *a, **k)
    aggregation.VolumeGenerator.__init__ = cpu_init
    try:
        vg = aggregation.build_volume_generator(_cfg("mean"))
        assert vg.aggregation_method == "softmax"
        assert vg.volume_size == 16 and vg.cuboid_side == 2000.0 and vg.kind == "coco"
        assert vg.process_feature[0].in_channels == 64 and vg.process_feature[0].out_channels == 8
        assert aggregation.build_volume_generator(_cfg("sum", deconv_layers=0)).process_feature[0].in_channels == 2048
    finally:
        aggregation.VolumeGenerator.__init__ = real_init


def test_build_volume_generator_matches_reference_ctor(reference):
    real_init = reference["aggregation"].VolumeGenerator.__init__

    def cpu_init(self, *a, **k):
        k["device"] = "cpu"
        real_init(self, *a, **k)
    reference["aggregation"].VolumeGenerator.__init__ = cpu_init
    try:
        ref_vg = reference["aggregation"].build_volume_generator(_cfg("mean"))
    finally:
        reference["aggregation"].VolumeGenerator.__init__ = real_init
    assert ref_vg.aggregation_method == "softmax"
    ours = aggregation.VolumeGenerator(volume_size=16, input_channels=64, output_channels=8, device="cpu")
    assert sorted(ours.state_dict()) == sorted(ref_vg.state_dict())
    ours.load_state_dict(ref_vg.state_dict())


def test_dropin_registers_reference_module_names():
    saved = {k: v for k, v in sys.modules.items() if k in ("models", "utils") or k.startswith(("models.", "utils."))}
    try:
        names = dropin.install()
        assert names == ["models.aggregation", "utils.multiview", "utils.volumetric"]
        from models.aggregation import build_volume_generator, unprojection   # noqa: F401
        from utils import multiview as mv2, volumetric as vol2
        assert unprojection is aggregation.unprojection
        assert mv2 is multiview and vol2 is volumetric
    finally:
        dropin.uninstall()
        for k in list(sys.modules):
            if k in ("models", "utils") or k.startswith(("models.", "utils.")):
                sys.modules.pop(k)
        sys.modules.update(saved)


def test_batched_projections_equal_the_camera_loop():
    """VolumeGenerator._projections: the batched float64 path (element-wise K scaling + per-camera
    dgemm through np.matmul) must give the bits of the reference's per-camera loop."""
    from multiviewhmr_b200 import aggregation as agg, multiview
    rng = np.random.default_rng(7)
    B, V = 5, 4
    cams = [[multiview.Camera(np.linalg.qr(rng.normal(size=(3, 3)))[0], rng.normal(size=3) * 900 + [0, 0, 4000.0],
                              [[1145.04 + rng.normal(), 0.0, 512.54 + rng.normal()], [0.0, 1143.78, 515.45 + rng.normal()], [0, 0, 1.0]])
             for _ in range(B)] for _ in range(V)]
    batch = {"cameras": cams}
    vg = agg.VolumeGenerator.__new__(agg.VolumeGenerator)
    for img, feat in (((384, 384), (96, 96)), ((1000, 1002), (7, 9)), ((224, 256), (56, 64))):
        fast = agg.VolumeGenerator._projections(vg, batch, img, feat, V, B)
        slow = agg.VolumeGenerator._projections(vg, batch, img, feat, V, B, batched=False)
        assert fast.dtype == np.float32 and fast.shape == (B, V, 3, 4)
        assert np.array_equal(fast, slow)
    K0 = cams[0][0].K.copy()
    agg.VolumeGenerator._projections(vg, batch, (384, 384), (96, 96), V, B)
    assert np.array_equal(cams[0][0].K, K0)                       # the caller's cameras are not modified
    # cameras that are not plain float64 take the loop
    cams[1][2].K = cams[1][2].K.astype(np.float32)
    assert np.array_equal(agg.VolumeGenerator._projections(vg, batch, (384, 384), (96, 96), V, B),
                          agg.VolumeGenerator._projections(vg, batch, (384, 384), (96, 96), V, B, batched=False))


def test_header_is_c99_and_library_links_from_plain_c(built_lib, tmp_path):
    """What a cgo / JNI / FFI binding sees: include/mvhmr_b200.h compiled as strict C99 by gcc and the
    shared library linked into a C program (argument validation only, no device work)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.dirname(_lib.LIB_PATH)
    exe = str(tmp_path / "c_abi_smoke")
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(root, "include"),
                    os.path.join(root, "tests", "c_abi", "c_abi_smoke.c"), "-o", exe,
                    "-L", libdir, "-lmvhmr_b200", "-Wl,-rpath," + libdir], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "c abi ok" in out.stdout
