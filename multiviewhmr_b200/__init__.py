"""multiviewhmr_b200 — B200-native (sm_100a) volumetric aggregation for
multi-view human mesh recovery.

Host-side mirror of the three reference modules on the aggregation hot path:

    multiviewhmr_b200.aggregation  <->  models/aggregation.py
    multiviewhmr_b200.volumetric   <->  utils/volumetric.py
    multiviewhmr_b200.multiview    <->  utils/multiview.py

backed by hand-written CUDA kernels behind the C ABI in include/mvhmr_b200.h.
`multiviewhmr_b200.dropin.install()` registers them under the reference's
module names.
"""
from . import aggregation, multiview, volumetric  # noqa: F401
from .aggregation import (VolumeGenerator, build_coord_volumes, build_volume_generator, pack_features,  # noqa: F401
                          soft_argmax_3d, soft_argmax_3d_grid, unprojection, unprojection_grid,
                          unprojection_soft_argmax)

__all__ = ["aggregation", "multiview", "volumetric", "unprojection", "unprojection_grid", "VolumeGenerator",
           "build_volume_generator", "build_coord_volumes", "soft_argmax_3d", "soft_argmax_3d_grid", "pack_features",
           "unprojection_soft_argmax"]
