"""CPU oracle for the volumetric-aggregation path — TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl
reference` legs of `bench.py` may import this package.  The product
(`multiviewhmr_b200/`) never does: it fails loudly when its CUDA library is
missing instead of falling back to anything in here.

Two restatements live here:
  * `mvhmr_oracle.c`  — plain C, op-for-op in the reference's fp32 rounding
    order plus a float64 "truth" variant (wrapped by the functions below);
  * `torch_port.py`   — the reference's own sequence of torch calls
    (per-view `F.grid_sample` loop), used as the timed CPU baseline.
Parity status: pinned against `tests/golden/*.npz` (generated from the real
reference by `tests/golden/make_golden.py`) for everything except the 3-D
soft-argmax, which the reference does not contain (parity unpinned).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "mvhmr_oracle.c")
_SO = os.path.join(_HERE, "_build", "libmvhmr_oracle.so")

METHODS = {"sum": 0, "mean": 1, "max": 2, "softmax": 3}


def build(force=False):
    """gcc the C restatement into oracle/_build/ (git-ignored, travels with gpurun)."""
    if (not force and os.path.exists(_SO)
            and os.path.getmtime(_SO) >= os.path.getmtime(_SRC)):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    cmd = ["gcc", "-std=c11", "-O2", "-fPIC", "-shared", "-fopenmp", "-ffp-contract=off",
           "-fno-fast-math", "-o", _SO, _SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build())
        fp = ctypes.POINTER(ctypes.c_float)
        dp = ctypes.POINTER(ctypes.c_double)
        i, sz = ctypes.c_int, ctypes.c_size_t
        L.orc_num_threads.restype = i
        L.orc_set_num_threads.argtypes = [i]
        L.orc_build_coord_volumes.argtypes = [fp, fp, fp, fp, fp, i, i, i, i]
        L.orc_rotate_points.argtypes = [fp, fp, fp, sz]
        L.orc_project_points.argtypes = [fp, fp, fp, sz, i]
        L.orc_sample_positions.argtypes = [fp, fp, ctypes.POINTER(ctypes.c_uint8), fp, fp, sz, i, i]
        L.orc_unproject_aggregate_f32.argtypes = [fp, fp, fp, fp, i, i, i, i, i, sz, i]
        L.orc_unproject_aggregate_f32.restype = i
        L.orc_unproject_aggregate_f64.argtypes = [fp, fp, fp, dp, i, i, i, i, i, sz, i]
        L.orc_unproject_aggregate_f64.restype = i
        L.orc_soft_argmax3d_f64.argtypes = [fp, fp, dp, i, i, sz]
        _lib = L
    return _lib


def _f32(a):
    if hasattr(a, "detach"):
        a = a.detach().cpu().float().numpy()
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t=ctypes.c_float):
    return a.ctypes.data_as(ctypes.POINTER(t))


def num_threads():
    return lib().orc_num_threads()


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


def build_coord_volumes(centers, rot, pos, step, G):
    """(B,Gx,Gy,Gz,3) fp32.  centers (B,3), rot (B,3,3) fp32, pos/step 3-vectors
    already rounded to fp32.  Follows models/aggregation.py:140-187."""
    centers, rot = _f32(centers), _f32(rot)
    pos, step = _f32(np.broadcast_to(pos, (3,))), _f32(np.broadcast_to(step, (3,)))
    Gx, Gy, Gz = (G, G, G) if np.isscalar(G) else G
    B = centers.shape[0]
    out = np.empty((B, Gx, Gy, Gz, 3), np.float32)
    lib().orc_build_coord_volumes(_p(out), _p(centers), _p(rot), _p(pos), _p(step), B, Gx, Gy, Gz)
    return out


def rotate_points(pts, rot):
    pts, rot = _f32(pts), _f32(rot)
    out = np.empty_like(pts)
    lib().orc_rotate_points(_p(out), _p(pts), _p(rot), pts.size // 3)
    return out


def project_points(P, pts, euclid=True):
    P, pts = _f32(P), _f32(pts)
    n = pts.shape[0]
    out = np.empty((n, 2 if euclid else 3), np.float32)
    lib().orc_project_points(_p(out), _p(P), _p(pts), n, int(euclid))
    return out


def sample_positions(P, coord, H, W):
    """(ix, iy, invalid) of every point of `coord` (N,3) in one view."""
    P, coord = _f32(P), _f32(coord).reshape(-1, 3)
    n = coord.shape[0]
    ix, iy = np.empty(n, np.float32), np.empty(n, np.float32)
    inv = np.empty(n, np.uint8)
    lib().orc_sample_positions(_p(ix), _p(iy), _p(inv, ctypes.c_uint8), _p(P), _p(coord), n, H, W)
    return ix, iy, inv.astype(bool)


def unprojection(features, proj, coord_volumes, aggregation_method="softmax", truth=False):
    """CPU statement of models/aggregation.py:20-87.  Returns numpy (B,C,*vol)
    fp32 (reference rounding order) or, with truth=True, float64."""
    if aggregation_method not in METHODS:
        raise ValueError("Unknown aggregation_method: {}".format(aggregation_method))
    f, P, cv = _f32(features), _f32(proj), _f32(coord_volumes)
    B, V, C, H, W = f.shape
    vol = cv.shape[1:-1]
    N = int(np.prod(vol))
    out = np.empty((B, C) + tuple(vol), np.float64 if truth else np.float32)
    if truth:
        rc = lib().orc_unproject_aggregate_f64(_p(f), _p(P), _p(cv), _p(out, ctypes.c_double),
                                               B, V, C, H, W, N, METHODS[aggregation_method])
    else:
        rc = lib().orc_unproject_aggregate_f32(_p(f), _p(P), _p(cv), _p(out),
                                               B, V, C, H, W, N, METHODS[aggregation_method])
    if rc != 0:
        raise ValueError("oracle rejected arguments (V=%d)" % V)
    return out


def soft_argmax_3d(volumes, coord_volumes):
    """float64 truth of the (unpinned) 3-D soft-argmax: (B,J,3)."""
    vol, cv = _f32(volumes), _f32(coord_volumes)
    B, J = vol.shape[:2]
    N = int(np.prod(vol.shape[2:]))
    out = np.empty((B, J, 3), np.float64)
    lib().orc_soft_argmax3d_f64(_p(vol), _p(cv), _p(out, ctypes.c_double), B, J, N)
    return out
