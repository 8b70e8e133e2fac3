"""Reference-named, autograd-aware entry points shared by the per-sub-project mirrors
(diffuvolume_b200/{sceneflow,kitti12,kitti15}.py).

Forward passes are the sm_100a kernels (diffuvolume_b200.ops).  The reference's training scripts
back-propagate through these functions (SceneFlow/main.py:154), so each one is a
torch.autograd.Function; the backward formulas are composed from PyTorch ops for now (SURVEY.md §8f
row f1 — dedicated backward kernels are the next row, inference never takes this path).

dtype handling follows the reference's `new_zeros`: the result has the input's dtype.  Inputs that
are not float32 (IGEV runs these under autocast, KITTI15/core/igev_stereo_ddim.py:366) are computed in
float32 and cast back.
"""
from __future__ import annotations

import torch

from . import ops


def _as_f32(t: torch.Tensor) -> torch.Tensor:
    return t if t.dtype == torch.float32 else t.float()


class _GwcVolume(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ref, tgt, maxdisp, num_groups):
        ctx.save_for_backward(ref, tgt)
        ctx.cfg = (maxdisp, num_groups)
        return ops.gwc_volume(_as_f32(ref), _as_f32(tgt), maxdisp, num_groups).to(ref.dtype)

    @staticmethod
    def backward(ctx, g):
        ref, tgt = ctx.saved_tensors
        D, G = ctx.cfg
        B, C, H, W = ref.shape
        cpg = C // G
        g = g.float()
        r32, t32 = ref.float(), tgt.float()
        gref = torch.zeros_like(r32) if ctx.needs_input_grad[0] else None
        gtgt = torch.zeros_like(t32) if ctx.needs_input_grad[1] else None
        for d in range(min(D, W)):
            gd = g[:, :, d, :, d:].repeat_interleave(cpg, dim=1) / cpg      # [B,C,H,W-d]
            if gref is not None:
                gref[..., d:] += gd * t32[..., : W - d]
            if gtgt is not None:
                gtgt[..., : W - d] += gd * r32[..., d:]
        return (None if gref is None else gref.to(ref.dtype), None if gtgt is None else gtgt.to(tgt.dtype), None, None)


class _ConcatVolume(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ref, tgt, maxdisp, mask_left):
        ctx.cfg = (maxdisp, mask_left, ref.dtype, tgt.dtype)
        return ops.concat_volume(_as_f32(ref), _as_f32(tgt), maxdisp, mask_left=mask_left).to(ref.dtype)

    @staticmethod
    def backward(ctx, g):
        D, mask_left, dt_r, dt_t = ctx.cfg
        C = g.shape[1] // 2
        W = g.shape[-1]
        g = g.float()
        gl, gr = g[:, :C], g[:, C:]
        gref = gtgt = None
        if ctx.needs_input_grad[0]:
            if mask_left:
                gref = torch.zeros_like(gl[:, :, 0])
                for d in range(min(D, W)):
                    gref[..., d:] += gl[:, :, d, :, d:]
            else:
                gref = gl.sum(dim=2)
            gref = gref.to(dt_r)
        if ctx.needs_input_grad[1]:
            gtgt = torch.zeros_like(gr[:, :, 0])
            for d in range(min(D, W)):
                gtgt[..., : W - d] += gr[:, :, d, :, d:]
            gtgt = gtgt.to(dt_t)
        return gref, gtgt, None, None


class _CorrVolume2Sided(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ref, tgt, maxdisp, num_groups):
        ctx.save_for_backward(ref, tgt)
        ctx.cfg = (maxdisp, num_groups)
        return ops.corr_volume_2sided(_as_f32(ref), _as_f32(tgt), maxdisp, num_groups).to(ref.dtype)

    @staticmethod
    def backward(ctx, g):
        ref, tgt = ctx.saved_tensors
        m, G = ctx.cfg
        B, C, H, W = ref.shape
        cpg = C // G
        g = g.float()
        r32, t32 = ref.float(), tgt.float()
        gref, gtgt = torch.zeros_like(r32), torch.zeros_like(t32)
        for i in range(-m, m + 1):
            gi = g[:, :, i + m].repeat_interleave(cpg, dim=1) / cpg
            if i >= 0:
                if i < W:
                    gref[..., i:] += gi[..., i:] * t32[..., : W - i]
                    gtgt[..., : W - i] += gi[..., i:] * r32[..., i:]
            else:
                k = min(-i, W)
                gref[..., :k] += gi[..., :k] * t32[..., W - k:]
                gtgt[..., W - k:] += gi[..., :k] * r32[..., :k]
        return gref.to(ref.dtype), gtgt.to(tgt.dtype), None, None


class _GroupwiseCorrelation(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f1, f2, num_groups):
        ctx.save_for_backward(f1, f2)
        ctx.G = num_groups
        return ops.groupwise_correlation(_as_f32(f1), _as_f32(f2), num_groups).to(f1.dtype)

    @staticmethod
    def backward(ctx, g):
        f1, f2 = ctx.saved_tensors
        cpg = f1.shape[1] // ctx.G
        ge = g.float().repeat_interleave(cpg, dim=1) / cpg
        return (ge * f2.float()).to(f1.dtype), (ge * f1.float()).to(f2.dtype), None


class _DisparityRegression(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, maxdisp, keepdim):
        ctx.cfg = (maxdisp, keepdim, x.dtype)
        return ops.disparity_regression(_as_f32(x), maxdisp, keepdim=keepdim).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        maxdisp, keepdim, dt = ctx.cfg
        if not keepdim:
            g = g.unsqueeze(1)
        dv = torch.arange(0, maxdisp, dtype=torch.float32, device=g.device).view(1, maxdisp, 1, 1)
        return (g.float() * dv).to(dt), None, None


def groupwise_correlation(fea1, fea2, num_groups):
    B, C, H, W = fea1.shape
    assert C % num_groups == 0
    return _GroupwiseCorrelation.apply(fea1, fea2, num_groups)


def build_gwc_volume(refimg_fea, targetimg_fea, maxdisp, num_groups):
    B, C, H, W = refimg_fea.shape
    assert C % num_groups == 0
    return _GwcVolume.apply(refimg_fea, targetimg_fea, maxdisp, num_groups)


def build_concat_volume_m(refimg_fea, targetimg_fea, maxdisp):
    return _ConcatVolume.apply(refimg_fea, targetimg_fea, maxdisp, False)


def build_concat_volume_t(refimg_fea, targetimg_fea, maxdisp):
    return _ConcatVolume.apply(refimg_fea, targetimg_fea, maxdisp, True)


def build_corrleation_volume(refimg_fea, targetimg_fea, maxdisp, num_groups):
    B, C, H, W = refimg_fea.shape
    assert C % num_groups == 0
    return _CorrVolume2Sided.apply(refimg_fea, targetimg_fea, maxdisp, num_groups)


def disparity_regression(x, maxdisp, keepdim=False):
    assert len(x.shape) == 4
    return _DisparityRegression.apply(x, maxdisp, keepdim)
