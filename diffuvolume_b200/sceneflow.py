"""Drop-in mirror of the volume ops ACVNet uses — same names, argument meaning and error behaviour as
SceneFlow/models/submodule.py (the module `from models.submodule import *` binds in acv.py / acv_ddim.py).

    build_gwc_volume       SceneFlow/models/submodule.py:228-238
    groupwise_correlation  SceneFlow/models/submodule.py:209-215
    build_concat_volume    SceneFlow/models/submodule.py:180-191   (variant M: left half NOT zero-masked)
    disparity_regression   SceneFlow/models/submodule.py:173-177

The orphan top-level SceneFlow/submodule.py is byte-identical to KITTI12/models/submodule.py for these
functions: use diffuvolume_b200.kitti12 for it.
"""
from .functional import build_concat_volume_m as build_concat_volume
from .functional import build_gwc_volume, groupwise_correlation


def disparity_regression(x, maxdisp):
    from .functional import disparity_regression as _dr
    return _dr(x, maxdisp, keepdim=False)


__all__ = ["build_gwc_volume", "groupwise_correlation", "build_concat_volume", "disparity_regression"]
