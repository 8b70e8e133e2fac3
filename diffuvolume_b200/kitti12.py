"""Drop-in mirror of the volume ops PCWNet uses — KITTI12/models/submodule.py (and the identical orphan
SceneFlow/submodule.py).

    build_gwc_volume          KITTI12/models/submodule.py:109-119
    groupwise_correlation     KITTI12/models/submodule.py:100-106
    build_concat_volume       KITTI12/models/submodule.py:86-97     (variant T: BOTH halves zero for x < d)
    build_corrleation_volume  KITTI12/models/submodule.py:121-135   (sic; two-sided, with the negative-shift quirk)
    disparity_regression      KITTI12/models/submodule.py:33-37
    warp                      KITTI12/models/submodule.py:137-176   (grid_sample + validity mask, align_corners quirk)
"""
from .functional import build_concat_volume_t as build_concat_volume
from .functional import build_corrleation_volume, build_gwc_volume, groupwise_correlation, warp


def disparity_regression(x, maxdisp):
    from .functional import disparity_regression as _dr
    return _dr(x, maxdisp, keepdim=False)


__all__ = ["build_gwc_volume", "groupwise_correlation", "build_concat_volume", "build_corrleation_volume",
           "disparity_regression", "warp"]
