"""ctypes binding of the C ABI declared in include/mvhmr_b200.h.

There is no CPU or eager-PyTorch fallback anywhere in this package: if the
CUDA library is missing or a tensor is not on a CUDA device, the call raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MVHMR_LIB") or os.path.join(_HERE, "lib", "libmvhmr_b200.so")   # MVHMR_LIB: tuning builds

OK, ERR_INVALID_ARGUMENT, ERR_WORKSPACE, ERR_CUDA = 0, -1, -2, -3
SUM, MEAN, MAX, SOFTMAX = 0, 1, 2, 3
F32, BF16 = 0, 1
LAYOUT_NCHW, LAYOUT_PACKED, LAYOUT_NHWC = 0, 1, 2
OUT_NDHWC, OUT_POOL2 = 1, 2
METHODS = {"sum": SUM, "mean": MEAN, "max": MAX, "softmax": SOFTMAX}

_vp, _i, _ll, _sz, _u = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_size_t, ctypes.c_uint
_fp = ctypes.POINTER(ctypes.c_float)

# name -> (restype, argtypes); mirrors include/mvhmr_b200.h one to one
SIGNATURES = {
    "mvhmr_abi_version": (_i, []),
    "mvhmr_last_error": (ctypes.c_char_p, []),
    "mvhmr_build_coord_volumes": (_i, [_vp, _vp, _vp, _fp, _fp, _i, _i, _i, _i, _vp]),
    "mvhmr_rotate_points": (_i, [_vp, _vp, _fp, _sz, _vp]),
    "mvhmr_project_points": (_i, [_vp, _vp, _vp, _sz, _i, _vp]),
    "mvhmr_packed_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "mvhmr_pack_features": (_i, [_vp, _i, _vp, _i, _i, _i, _i, _vp]),
    "mvhmr_unproject_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i, _i]),
    "mvhmr_unproject_aggregate": (_i, [_vp, _i, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i,
                                       _i, _i, _ll, _ll, _ll, _ll, _u, _vp, _sz, _vp]),
    "mvhmr_unproject_aggregate_grid": (_i, [_vp, _i, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i,
                                            _i, _i, _ll, _ll, _ll, _ll, _u, _vp, _sz, _vp]),
    "mvhmr_unproject_aggregate_fmt": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _u, _i, _i, _i, _i, _i, _i, _i, _i, _i,
                                           _i, _i, _ll, _ll, _ll, _ll, _u, _vp, _sz, _vp]),
    "mvhmr_unproject_tex_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "mvhmr_unproject_aggregate_tex": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i,
                                           _i, _i, _ll, _ll, _ll, _ll, _vp, _sz, _vp]),
    "mvhmr_unproject_aggregate_backward": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _ll, _i, _vp]),
    "mvhmr_unproject_backward_workspace_bytes": (_sz, [_i, _i, _i, _i, _i, _i, _i]),
    "mvhmr_unproject_aggregate_backward_ws": (_i, [_vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _ll, _i, _vp, _sz, _vp]),
    "mvhmr_selftest_division": (_i, [ctypes.c_float, _vp, _vp]),
    "mvhmr_soft_argmax3d_num_slices": (_i, [_ll]),
    "mvhmr_soft_argmax3d_workspace_bytes": (_sz, [_i, _i, _ll]),
    "mvhmr_soft_argmax3d": (_i, [_vp, _vp, _vp, _i, _i, _ll, _vp, _sz, _vp]),
    "mvhmr_soft_argmax3d_strided": (_i, [_vp, _vp, _vp, _i, _i, _ll, _ll, _vp, _sz, _vp]),
    "mvhmr_soft_argmax3d_grid": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _ll, _vp, _sz, _vp]),
    "mvhmr_soft_argmax3d_partials": (_i, [_vp, _vp, _vp, _i, _i, _ll, _ll, _ll, _vp]),
    "mvhmr_soft_argmax3d_finalize": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "mvhmr_unproject_softargmax_workspace_bytes": (_sz, [_i, _i]),
    "mvhmr_unproject_aggregate_softargmax": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i,
                                                  _i, _u, _vp, _sz, _vp, _sz, _vp]),
}



class Grid(ctypes.Structure):
    """mvhmr_grid_t"""
    _fields_ = [("centers", ctypes.c_void_p), ("rot", ctypes.c_void_p),
                ("pos", ctypes.c_float * 3), ("step", ctypes.c_float * 3)]


_lib = None


def load():
    """dlopen the library (no CUDA call is made); raises if it was never built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "multiviewhmr_b200: CUDA library %s is missing — run "
                "`python -m multiviewhmr_b200.build` (there is no CPU fallback)" % LIB_PATH)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        if L.mvhmr_abi_version() != 1:
            raise RuntimeError("multiviewhmr_b200: ABI version mismatch in %s" % LIB_PATH)
        _lib = L
    return _lib


def check(rc):
    if rc == OK:
        return
    msg = load().mvhmr_last_error().decode("utf-8", "replace")
    if rc == ERR_INVALID_ARGUMENT:
        raise ValueError(msg)
    raise RuntimeError("multiviewhmr_b200 (code %d): %s" % (rc, msg))


def require_cuda(*tensors):
    for t in tensors:
        if not torch.is_tensor(t):
            raise TypeError("Works only with numpy arrays and PyTorch tensors.")
        if not t.is_cuda:
            raise RuntimeError(
                "multiviewhmr_b200 runs on CUDA tensors only (got a %s tensor); "
                "there is no CPU fallback" % t.device)
    dev = tensors[0].device
    for t in tensors[1:]:
        if t.device != dev:
            raise RuntimeError("multiviewhmr_b200: tensors on different devices (%s vs %s)" % (dev, t.device))
    return dev


def stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def host3(values):
    return (ctypes.c_float * len(values))(*[float(v) for v in values])
