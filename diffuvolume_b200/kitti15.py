"""Drop-in mirror of the IGEV-Stereo hot-path ops — KITTI15/core/submodule.py, core/geometry.py,
core/geometry_ddim.py and core/utils/utils.py:bilinear_sampler.

    build_gwc_volume / groupwise_correlation   KITTI15/core/submodule.py:151-169
    build_concat_volume                        KITTI15/core/submodule.py:206-217   (variant M)
    disparity_regression                       KITTI15/core/submodule.py:219-223   (keepdim=True)
    Combined_Geo_Encoding_Volume               KITTI15/core/geometry.py:6-68  (`__call__(disp, coords)`)
    Combined_Geo_Encoding_Volume_DDIM          KITTI15/core/geometry_ddim.py:6-80 (`__call__(disp, coords, noisy)`)
"""
from __future__ import annotations

import torch

from . import ops
from .functional import build_concat_volume_m as build_concat_volume
from .functional import build_gwc_volume, groupwise_correlation


def disparity_regression(x, maxdisp):
    from .functional import disparity_regression as _dr
    return _dr(x, maxdisp, keepdim=True)


class Combined_Geo_Encoding_Volume:
    """Same constructor, attributes (`geo_volume_pyramid`, `init_corr_pyramid`, `num_levels`, `radius`,
    `channel`) and call convention as the reference class; `noisy` is optional so that one class serves
    both geometry.py (2 arguments) and geometry_ddim.py (3 arguments)."""

    def __init__(self, init_fmap1, init_fmap2, geo_volume, num_levels=2, radius=4):
        self.num_levels = num_levels
        self.radius = radius
        self.geo_volume_pyramid = []
        self.init_corr_pyramid = []
        init_corr = Combined_Geo_Encoding_Volume.corr(init_fmap1, init_fmap2)
        b, h, w, _, w2 = init_corr.shape
        b, c, d, h, w = geo_volume.shape
        self.channel = c
        geo = ops.geo_permute(geo_volume.float())                 # [b*h*w, c, 1, d]
        init_corr = init_corr.reshape(b * h * w, 1, 1, w2)
        self.geo_volume_pyramid.append(geo)
        self.init_corr_pyramid.append(init_corr)
        for _ in range(self.num_levels - 1):
            geo = ops.avgpool_w2(geo)
            self.geo_volume_pyramid.append(geo)
        for _ in range(self.num_levels - 1):
            init_corr = ops.avgpool_w2(init_corr)
            self.init_corr_pyramid.append(init_corr)

    def __call__(self, disp, coords, noisy=None):
        b, _, h, w = disp.shape
        if noisy is not None:
            # the reference reshapes the [B,D,h,w] buffer to [b*h*w, 1, 1, D] WITHOUT a permute
            # (geometry_ddim.py:37); the kernel reads the same raw layout
            noisy = noisy.float().contiguous()
        return ops.geo_lookup(self.geo_volume_pyramid, self.init_corr_pyramid, disp.float(), coords.float(), noisy,
                              self.radius)

    @staticmethod
    def corr(fmap1, fmap2):
        B, D, H, W1 = fmap1.shape
        _, _, _, W2 = fmap2.shape
        return ops.corr1d_allpairs(fmap1.float(), fmap2.float())


Combined_Geo_Encoding_Volume_DDIM = Combined_Geo_Encoding_Volume

__all__ = ["build_gwc_volume", "groupwise_correlation", "build_concat_volume", "disparity_regression",
           "Combined_Geo_Encoding_Volume", "Combined_Geo_Encoding_Volume_DDIM"]
