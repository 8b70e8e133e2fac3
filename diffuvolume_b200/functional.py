"""Reference-named, autograd-aware entry points shared by the per-sub-project mirrors
(diffuvolume_b200/{sceneflow,kitti12,kitti15}.py).

Forward passes are the sm_100a kernels (diffuvolume_b200.ops).  The reference's training scripts
back-propagate through these functions (SceneFlow/main.py:154), so each one is a
torch.autograd.Function whose backward is a dedicated kernel too (SURVEY.md §8f row f1,
csrc/volume_backward.cu): one pass over the gradient volume instead of autograd's D x {slice, mul, sum, index_put}.

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
        _, G = ctx.cfg
        gref, gtgt = ops.gwc_volume_bwd(g.float().contiguous(), ref.float(), tgt.float(), G,
                                        need_ref=ctx.needs_input_grad[0], need_tgt=ctx.needs_input_grad[1])
        return (None if gref is None else gref.to(ref.dtype), None if gtgt is None else gtgt.to(tgt.dtype), None, None)


class _ConcatVolume(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ref, tgt, maxdisp, mask_left):
        ctx.cfg = (maxdisp, mask_left, ref.dtype, tgt.dtype)
        return ops.concat_volume(_as_f32(ref), _as_f32(tgt), maxdisp, mask_left=mask_left).to(ref.dtype)

    @staticmethod
    def backward(ctx, g):
        _, mask_left, dt_r, dt_t = ctx.cfg
        gref, gtgt = ops.concat_volume_bwd(g.float().contiguous(), mask_left, need_ref=ctx.needs_input_grad[0],
                                           need_tgt=ctx.needs_input_grad[1])
        return (None if gref is None else gref.to(dt_r), None if gtgt is None else gtgt.to(dt_t), None, None)


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
        gref, gtgt = ops.gwc_volume_bwd(g.float().contiguous(), ref.float(), tgt.float(), G, two_sided_maxdisp=m,
                                        need_ref=ctx.needs_input_grad[0], need_tgt=ctx.needs_input_grad[1])
        return (None if gref is None else gref.to(ref.dtype), None if gtgt is None else gtgt.to(tgt.dtype), None, None)


class _GroupwiseCorrelation(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f1, f2, num_groups):
        ctx.save_for_backward(f1, f2)
        ctx.G = num_groups
        return ops.groupwise_correlation(_as_f32(f1), _as_f32(f2), num_groups).to(f1.dtype)

    @staticmethod
    def backward(ctx, g):
        f1, f2 = ctx.saved_tensors
        g1, g2 = ops.groupwise_correlation_bwd(g.float().contiguous(), f1.float(), f2.float(), ctx.G)
        return g1.to(f1.dtype), g2.to(f2.dtype), None


class _DisparityRegression(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, maxdisp, keepdim):
        ctx.cfg = (maxdisp, keepdim, x.dtype)
        return ops.disparity_regression(_as_f32(x), maxdisp, keepdim=keepdim).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        maxdisp, keepdim, dt = ctx.cfg
        return ops.disparity_regression_bwd(g.float().contiguous(), maxdisp).to(dt), None, None


class _Warp(torch.autograd.Function):
    """warp(x, disp) with kernel-backed gradients to both inputs (PCWNet's training differentiates through grid_sample
    into the right features and the disparity, KITTI12/models/submodule.py:169-176)."""

    @staticmethod
    def forward(ctx, x, disp):
        ctx.save_for_backward(x, disp)
        return ops.warp(_as_f32(x), _as_f32(disp)).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        x, disp = ctx.saved_tensors
        gx, gd = ops.warp_bwd(g.float().contiguous(), _as_f32(x), _as_f32(disp), ctx.needs_input_grad[0],
                              ctx.needs_input_grad[1])
        return (None if gx is None else gx.to(x.dtype)), (None if gd is None else gd.reshape(disp.shape).to(disp.dtype))


def warp(x, disp):
    """warp(x, disp) — KITTI12/models/submodule.py:137-176."""
    return _Warp.apply(x, disp)


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


class _ContextUpsample(torch.autograd.Function):
    @staticmethod
    def forward(ctx, disp_low, up_weights):
        ctx.save_for_backward(disp_low, up_weights)
        return ops.context_upsample(_as_f32(disp_low), _as_f32(up_weights))

    @staticmethod
    def backward(ctx, g):
        disp_low, up_weights = ctx.saved_tensors
        glow, gw = ops.context_upsample_bwd(g.float().contiguous(), disp_low.float(), up_weights.float(),
                                            need_low=ctx.needs_input_grad[0], need_weights=ctx.needs_input_grad[1])
        return (None if glow is None else glow.to(disp_low.dtype), None if gw is None else gw.to(up_weights.dtype))


def context_upsample(disp_low, up_weights):
    """KITTI15/core/submodule.py:241-253 — result is fp32 when either input is (the reference multiplies the fp32
    unfolded disparity by the weights), else the inputs' common dtype."""
    out = _ContextUpsample.apply(disp_low, up_weights)
    rt = torch.result_type(disp_low, up_weights)
    return out if rt == torch.float32 else out.to(rt)


# ------------------------------------------------------------------------------------------------------------------
# Fused ops with fused backward passes — used by the forward-level (tier-3) drop-ins (diffuvolume_b200/sampler.py)
# in place of the unnamed ops the reference runs BETWEEN its named functions.
# ------------------------------------------------------------------------------------------------------------------
class _AcvVolume(torch.autograd.Function):
    """(concat(cl, cr) * softmax(att_logits, dim=2)) * n — SceneFlow/models/acv_ddim.py:388-390 and, with n, the training
    branch's filter multiply :446-451 in ONE producer pass; backward: csrc/fused_backward.cu."""

    @staticmethod
    def forward(ctx, cl, cr, att_logits, maxdisp, n):
        att_w = ops.att_softmax(_as_f32(att_logits))
        ctx.save_for_backward(cl, cr, att_w, n)
        ctx.att_shape = att_logits.shape
        out = ops.concat_volume_weighted(_as_f32(cl), _as_f32(cr), maxdisp, mask_left=False, att_weights=att_w, n=n)
        return out.to(cl.dtype)

    @staticmethod
    def backward(ctx, g):
        cl, cr, att_w, n = ctx.saved_tensors
        gcl, gcr, gatt = ops.acv_volume_bwd(g.float().contiguous(), cl.float(), cr.float(), att_w, n,
                                            need_cl=ctx.needs_input_grad[0], need_cr=ctx.needs_input_grad[1],
                                            need_att=ctx.needs_input_grad[2])
        return (None if gcl is None else gcl.to(cl.dtype), None if gcr is None else gcr.to(cr.dtype),
                None if gatt is None else gatt.reshape(ctx.att_shape), None, None)


def acv_attention_volume(concat_left, concat_right, att_logits, maxdisp, n=None):
    """`F.softmax(att_weights, dim=2) * build_concat_volume(cl, cr, maxdisp)` (acv_ddim.py:388-390, acv.py:201-203),
    optionally times the DDIM filter factor n [B,D,h,w] fp32 (acv_ddim.py:446-451) — differentiable w.r.t. the concat
    features and the attention logits.  H*W % 4 != 0 (no reference shape) takes the unfused kernels."""
    B, C, H, W = concat_left.shape
    if (H * W) % 4 != 0 or W < 4:
        vol = build_concat_volume_m(concat_left, concat_right, maxdisp)
        vol = torch.softmax(att_logits, dim=2) * vol
        return vol if n is None else vol * n.unsqueeze(1)
    if n is not None:
        n = n.detach().float().contiguous()
    return _AcvVolume.apply(concat_left, concat_right, att_logits, maxdisp, n)


class _SoftmaxRegress(torch.autograd.Function):
    """disparity_regression(F.softmax(cost, 1), maxdisp) without materialising the probability volume, forward
    (csrc/softmax_regress.cu) and backward (csrc/fused_backward.cu)."""

    @staticmethod
    def forward(ctx, cost, keepdim):
        ctx.save_for_backward(cost)
        ctx.keepdim = keepdim
        disp = ops.softmax_regress(_as_f32(cost))["disp"]
        return (disp.unsqueeze(1) if keepdim else disp).to(cost.dtype)

    @staticmethod
    def backward(ctx, g):
        (cost,) = ctx.saved_tensors
        return ops.softmax_regress_bwd(cost.float(), g.float().contiguous()).to(cost.dtype), None


def softmax_disparity_regression(cost, maxdisp, keepdim=False):
    """`disparity_regression(F.softmax(cost, dim=1), maxdisp)` (acv_ddim.py:460-480, acv.py:213-236) fused."""
    assert len(cost.shape) == 4
    if cost.shape[1] != maxdisp:
        raise RuntimeError(
            f"The size of tensor a ({cost.shape[1]}) must match the size of tensor b ({maxdisp}) at non-singleton dimension 1")
    return _SoftmaxRegress.apply(cost, keepdim)


class _VolumeFilter(torch.autograd.Function):
    """volume * ((clamp(noise + shift, -s, s)/s + 1)/2).unsqueeze(1).float() (acv_ddim.py:254-260, pwcnet_ddim.py:466-472);
    gradient w.r.t. the volume only (the factor is detached in the reference's training branches)."""

    @staticmethod
    def forward(ctx, volume, xt, shift, scale):
        ctx.save_for_backward(xt, shift)
        ctx.scale = scale
        return ops.volume_filter(_as_f32(volume), xt, shift, scale).to(volume.dtype)

    @staticmethod
    def backward(ctx, g):
        xt, shift = ctx.saved_tensors
        return ops.volume_filter(g.float().contiguous(), xt, shift, ctx.scale).to(g.dtype), None, None, None


def volume_filter(volume, xt, shift=None, scale=1.0):
    return _VolumeFilter.apply(volume, xt.detach(), None if shift is None else shift.detach(), scale)
