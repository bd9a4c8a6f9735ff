"""Generate tests/golden/*.npz by running the REAL reference.

Run once, in the build container (where /root/reference is mounted):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

The reference (`models.aggregation`, `utils.volumetric`, `utils.multiview`) is
imported unchanged from /root/reference; nothing from it is copied.  Outputs are
small arrays (inputs + the reference's outputs) that travel to the GPU box,
where /root/reference does not exist.  torch 2.11.0+cu128 CPU path, MKL.
"""
import math
import os
import sys
import warnings

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

import numpy as np
import torch

from models import aggregation as ref_agg          # noqa: E402  (the reference)
from utils import multiview as ref_mv              # noqa: E402
from utils import volumetric as ref_vol            # noqa: E402

from multiviewhmr_b200 import synthetic as syn     # noqa: E402

METHODS = ("sum", "mean", "max", "softmax")


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in arrays.items()})
    print("%-28s %7.1f KB" % (name, os.path.getsize(path) / 1024))


def t2n(t):
    return t.detach().cpu().numpy()


def ragged_case():
    """Non-square maps, non-cubic rotated grid, one camera inside the cuboid
    (depth <= 0 for part of the grid)."""
    g = torch.Generator().manual_seed(7)
    B, V, C, H, W = 2, 3, 8, 12, 16
    feats = torch.randn(B, V, C, H, W, generator=g)
    centers = torch.randn(B, 3, generator=g) * 100.0
    proj = syn.make_projections(B, V, H, W, behind_views=(1,))
    Gx, Gy, Gz = 6, 5, 8
    coords = torch.zeros(B, Gx, Gy, Gz, 3)
    for b in range(B):
        xs = torch.linspace(-1250.0, 1250.0, Gx)
        ys = torch.linspace(-1250.0, 1250.0, Gy)
        zs = torch.linspace(-1250.0, 1250.0, Gz)
        grid = torch.stack(torch.meshgrid(xs, ys, zs, indexing="ij"), dim=-1)
        grid = ref_vol.rotate_coord_volume(grid - centers[b], 0.4 + b, [0, 0, 1]) + centers[b]
        coords[b] = grid
    out = {m: t2n(ref_agg.unprojection(feats, proj, coords, m)) for m in METHODS}
    ph = ref_mv.project_3d_points_to_image_plane_without_distortion(
        proj[0, 1], coords[0].reshape(-1, 3), convert_back_to_euclidean=False)
    save("unproj_ragged", features=t2n(feats), proj=t2n(proj), coord_volumes=t2n(coords),
         homog_b0_v1=t2n(ph), **{"out_" + m: out[m] for m in METHODS})


def edge_case():
    """Hand-placed points: depth exactly 0, negative depth, exactly on the map
    border, half a pixel outside, far outside, huge finite coordinates."""
    g = torch.Generator().manual_seed(11)
    B, V, C, H, W = 1, 2, 4, 8, 8
    feats = torch.randn(B, V, C, H, W, generator=g)
    # view 0: pixel = (X, Y) / Z ; view 1: shifted pinhole
    P0 = torch.tensor([[1.0, 0, 0, 0], [0, 1.0, 0, 0], [0, 0, 1.0, 0]])
    P1 = torch.tensor([[2.0, 0, 1.0, 3.0], [0, 2.0, 0.5, -1.0], [0, 0, 1.0, 0.5]])
    proj = torch.stack([P0, P1]).unsqueeze(0)
    pts = torch.tensor([
        [3.0, 4.0, 0.0],        # depth == 0 in view 0 -> w:=1, invalid
        [3.0, 4.0, -1.0],       # negative depth
        [0.0, 0.0, 1.0],        # top-left texel centre
        [8.0, 8.0, 1.0],        # x/H = 1 -> ix = W-1 exactly
        [8.0000005, 4.0, 1.0],  # one ulp outside
        [-0.5, 3.3, 1.0],       # partial corners on the left border
        [4.25, -0.75, 1.0],     # partial corners on the top border
        [9.5, 9.5, 1.0],        # all corners outside
        [1e30, 2.0, 1.0],       # huge finite
        [2.0, -1e30, 1.0],
        [3.7, 2.2, 2.0],
        [5.5, 6.5, 1.0],
        [7.999, 7.999, 1.0],
        [1e-3, 1e-3, 1e-2],
        [-3.0, -4.0, -0.5],     # valid depth in view 1 only if 0.5-0.5 > 0: it is 0 -> invalid
        [100.0, 100.0, 50.0],
    ])
    coords = pts.view(1, 2, 2, 4, 3).contiguous()
    out = {m: t2n(ref_agg.unprojection(feats, proj, coords, m)) for m in METHODS}
    save("unproj_edge", features=t2n(feats), proj=t2n(proj), coord_volumes=t2n(coords),
         **{"out_" + m: out[m] for m in METHODS})


def cfg1_case():
    """BASELINE config #1 (B1 V4 C32 64x64 -> 32^3): strided slice of each mode
    plus float64 checksums.  Inputs are regenerated from the seed in the tests."""
    w = syn.CONFIGS["cfg1"]
    feats, proj, coords, centers = syn.make_inputs(w, seed=1234)
    arrays = {"centers": t2n(centers), "proj": t2n(proj)}
    for m in METHODS:
        o = ref_agg.unprojection(feats, proj, coords, m)
        arrays["slice_" + m] = t2n(o[:, ::4, ::3, ::3, ::3])
        arrays["sum_" + m] = np.float64(o.double().sum().item())
        arrays["l2_" + m] = np.float64(o.double().norm().item())
    save("unproj_cfg1", **arrays)


def bf16_case():
    """cfg #3 semantics: reference fp32 path on bf16-rounded maps."""
    g = torch.Generator().manual_seed(3)
    B, V, C, H, W, G = 1, 4, 16, 24, 24, 10
    feats = torch.randn(B, V, C, H, W, generator=g).bfloat16().float()
    centers = torch.randn(B, 3, generator=g) * 100.0
    proj = syn.make_projections(B, V, H, W)
    coords = syn.make_coord_volumes(centers, G)
    save("unproj_bf16", features=t2n(feats), proj=t2n(proj), coord_volumes=t2n(coords),
         out_softmax=t2n(ref_agg.unprojection(feats, proj, coords, "softmax")),
         out_sum=t2n(ref_agg.unprojection(feats, proj, coords, "sum")))


def make_cameras(B, V, Hi, Wi, rng):
    cams = []
    for v in range(V):
        row = []
        for b in range(B):
            a = 2.0 * math.pi * v / V + 0.3 + 0.05 * b
            R = np.array([[math.cos(a), 0.0, -math.sin(a)], [0.0, 1.0, 0.0],
                          [math.sin(a), 0.0, math.cos(a)]])
            t = np.array([30.0, -20.0, 4500.0]) + rng.normal(size=3) * 10.0
            f = 0.8 * Wi * 4500.0 / 2500.0
            K = np.array([[f, 0.0, Wi / 2.0 + 1.5], [0.0, f * 1.01, Hi / 2.0 - 2.0], [0.0, 0.0, 1.0]])
            row.append((R, t, K))
        cams.append(row)
    return cams


def volume_generator_case(tag, Hf, Wf, kind, training, use_triangulation, seed):
    """VolumeGenerator.forward end to end; the arguments it hands to
    `unprojection` (projections, coord volumes, squeezed features) are captured
    by temporarily wrapping the module-level function."""
    rng = np.random.default_rng(seed)
    torch.manual_seed(seed)
    B, V, Cin, C, G, Hi, Wi = 2, 3, 12, 8, 8, 4 * Hf, 4 * Wf
    vg = ref_agg.VolumeGenerator(volume_size=G, input_channels=Cin, output_channels=C,
                                 cuboid_side=2500.0, use_triangulation=use_triangulation,
                                 kind=kind, device="cpu")
    vg.train(training)
    raw = make_cameras(B, V, Hi, Wi, rng)
    cameras = [[ref_mv.Camera(R, t, K) for (R, t, K) in row] for row in raw]
    kp3d = [rng.normal(size=(17, 4)) * 150.0 for _ in range(B)]
    batch = {"images": np.zeros((B, V, Hi, Wi, 3), np.float32), "cameras": cameras, "keypoints_3d": kp3d}
    feats = torch.randn(B, V, Cin, Hf, Wf)
    proj_in = torch.stack([torch.stack([torch.from_numpy(c.projection).float() for c in row])
                           for row in cameras]).transpose(1, 0).contiguous()
    captured = {}
    real = ref_agg.unprojection

    def spy(features, proj_matricies, coord_volumes, aggregation_method="softmax"):
        captured.update(features=features.clone(), proj=proj_matricies.clone(),
                        coord_volumes=coord_volumes.clone(), method=aggregation_method)
        return real(features, proj_matricies, coord_volumes, aggregation_method=aggregation_method)

    ref_agg.unprojection = spy
    try:
        np.random.seed(seed)         # training-mode theta comes from the global numpy RNG
        with torch.no_grad():
            vol = vg(feats, proj_in, batch)
    finally:
        ref_agg.unprojection = real
    save("vg_" + tag,
         features=t2n(feats), proj_in=t2n(proj_in),
         cam_R=np.array([[r[0] for r in row] for row in raw]),
         cam_t=np.array([[r[1] for r in row] for row in raw]),
         cam_K=np.array([[r[2] for r in row] for row in raw]),
         keypoints_3d=np.array(kp3d), image_hw=np.array([Hi, Wi]),
         conv_weight=t2n(vg.process_feature[0].weight), conv_bias=t2n(vg.process_feature[0].bias),
         used_proj=t2n(captured["proj"]), used_coord_volumes=t2n(captured["coord_volumes"]),
         used_features=t2n(captured["features"]), used_method=np.array(captured["method"]),
         volumes=t2n(vol), np_seed=np.array(seed), training=np.array(training),
         kind=np.array(kind), use_triangulation=np.array(use_triangulation))


def geometry_case():
    """utils/volumetric.py + utils/multiview.py helpers."""
    g = torch.Generator().manual_seed(5)
    axes = np.array([[0, 0, 1], [0, 1, 0], [1, 2, 3], [0, 0, 1]], dtype=np.float64)
    thetas = np.array([0.0, 0.7, 2.5, 6.0])
    rots = np.stack([ref_vol.get_rotation_matrix(a, t) for a, t in zip(axes, thetas)])
    vol = torch.randn(4, 5, 6, 3, generator=g) * 800.0
    rotated = np.stack([t2n(ref_vol.rotate_coord_volume(vol, float(t), list(a)))
                        for a, t in zip(axes, thetas)])
    P = syn.make_projections(1, 4, 96, 96)[0]
    pts = torch.randn(257, 3, generator=g) * 700.0
    homog = np.stack([t2n(ref_mv.project_3d_points_to_image_plane_without_distortion(P[v], pts, False))
                      for v in range(4)])
    eucl = np.stack([t2n(ref_mv.project_3d_points_to_image_plane_without_distortion(P[v], pts, True))
                     for v in range(4)])
    eucl_np = ref_mv.project_3d_points_to_image_plane_without_distortion(
        P[0].double().numpy(), pts.double().numpy())
    # DLT: project one 3-D point with the four cameras, triangulate it back
    X = torch.tensor([[120.0, -340.0, 510.0]])
    uv = torch.stack([ref_mv.project_3d_points_to_image_plane_without_distortion(P[v], X)[0]
                      for v in range(4)])
    tri_t = t2n(ref_mv.triangulate_point_from_multiple_views_linear_torch(P, uv))
    tri_n = ref_mv.triangulate_point_from_multiple_views_linear(P.double().numpy(), uv.double().numpy())
    # Camera bookkeeping
    cam = ref_mv.Camera(np.eye(3), [1.0, 2.0, 3.0], [[1100.0, 0, 512.0], [0, 1090.0, 384.0], [0, 0, 1]])
    cam.update_after_crop((100, 50, 900, 700))
    cam.update_after_resize((650, 800), (96, 128))
    save("geometry", axes=axes, thetas=thetas, rots=rots, vol=t2n(vol), rotated=rotated,
         P=t2n(P), pts=t2n(pts), homog=homog, eucl=eucl, eucl_np=eucl_np,
         tri_uv=t2n(uv), tri_torch=tri_t, tri_numpy=tri_n,
         cam_K=cam.K, cam_proj=cam.projection, cam_extr=cam.extrinsics)


def main():
    torch.set_num_threads(8)
    ragged_case()
    edge_case()
    cfg1_case()
    bf16_case()
    volume_generator_case("eval_mpii", 12, 12, "mpii", False, False, 21)
    volume_generator_case("train_coco", 12, 12, "coco", True, False, 22)
    volume_generator_case("train_mpii_rect", 10, 14, "mpii", True, False, 23)
    volume_generator_case("eval_dlt", 12, 12, "mpii", False, True, 24)
    geometry_case()


if __name__ == "__main__":
    main()
