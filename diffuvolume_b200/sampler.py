"""Tier-2 and tier-3 drop-ins: the DiffuVolume sampler methods AND the `forward` methods of the reference's model
classes, re-expressed on the fused sm_100a kernels.  They are bound onto the reference's OWN nn.Modules by
diffuvolume_b200.install (same signatures, same return values, same RNG draws in the same order and
dtypes), so the conv stacks, parameters and checkpoints stay the reference's.

ACVNet_DDIM   (SceneFlow/models/acv_ddim.py):  q_sample :241-246, predict_noise_from_start :248-252,
                                               model_predictions :254-296, ddim_sample :298-370
IGEVStereo_ddim (KITTI15/core/igev_stereo_ddim.py): q_sample :213-218, predict_noise_from_start :220-224,
                                               model_predictions :226-292, ddim_sample :294-359 (the GRU iteration
                                               loop keeps the reference's update block / upsampler modules)
PWCNet_ddim   (KITTI12/models/pwcnet_ddim.py): q_sample :453-458, predict_noise_from_start :460-464,
                                               model_predictions :466-528, ddim_sample :530-602

Tier 3 (forward level): ACVNet_DDIM.forward (acv_ddim.py:372-482, eval and training branch) and ACVNet.forward
(acv.py:167-247).  Only at this level can the ops the reference leaves unnamed inside `forward` be fused: the attention
volume `F.softmax(att, 2) * concat` (:390), the initial x_start scatter (:403-419 / :425-440), the training branch's
filter multiply (:446-451) and the `F.softmax` + `disparity_regression` heads (:460-480) — and only here do the
quarter-res concat features reach the sampler, so that every DDIM step re-produces the filtered volume from 21 MB of
features and factor maps instead of re-reading the 398 MB ac_volume (`regenerate` mode, what bench.py times).

What stays PyTorch inside these methods is exactly what is out of scope: the 3-D conv
stacks (dres*/classif*), the DynamicHead MLP (a [B,48] GEMV), F.upsample, and torch's RNG.
"""
from __future__ import annotations

from typing import Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import ops


# ------------------------------------------------------------------------------------------------
# shared helpers
# ------------------------------------------------------------------------------------------------
def _host_buffers(self):
    """The module's float64 schedule buffers as host arrays (cached; they never change)."""
    cache = getattr(self, "_dv_sched_cache", None)
    if cache is None:
        ac = self.alphas_cumprod.detach().double().cpu().numpy()
        cache = {
            "alphas_cumprod": ac,
            "sqrt_ac": self.sqrt_alphas_cumprod.detach().double().cpu().numpy(),
            "sqrt_1m_ac": self.sqrt_one_minus_alphas_cumprod.detach().double().cpu().numpy(),
            "sqrt_recip": self.sqrt_recip_alphas_cumprod.detach().double().cpu().numpy(),
            "sqrt_recipm1": self.sqrt_recipm1_alphas_cumprod.detach().double().cpu().numpy(),
        }
        object.__setattr__(self, "_dv_sched_cache", cache)
    return cache


def _time_index(t: torch.Tensor) -> int:
    """All call sites pass one timestep for the whole batch (acv_ddim.py:316, :358, :441)."""
    tt = t.reshape(-1)
    return int(tt[0].item())


def _time_shift(self, t: torch.Tensor, B: int, D: int, device) -> torch.Tensor:
    """DynamicHead adds a per-(b,d) shift to its input (head.py:74-77; IGEV interpolates it to D first,
    KITTI15/core/head.py:76-83).  Feeding zeros returns the shift itself; the MLP stays PyTorch."""
    z = torch.zeros((B, D, 1, 1), dtype=torch.float32, device=device)
    return self.time_embedding(z, t).reshape(B, D).float().contiguous()


def _update_coefficients(self, time: int, time_next: int):
    ac = _host_buffers(self)["alphas_cumprod"]
    a, an = ac[time], ac[time_next]
    sigma = self.ddim_sampling_eta * np.sqrt((1 - a / an) * (1 - an) / (1 - a))
    c = np.sqrt(1 - an - sigma ** 2)
    return float(np.sqrt(an)), float(c), float(sigma)


def _time_pairs(self):
    times = torch.linspace(-1, self.num_timesteps - 1, steps=self.sampling_timesteps + 1)
    times = list(reversed(times.int().tolist()))
    return list(zip(times[:-1], times[1:]))


# ------------------------------------------------------------------------------------------------
# methods common to the three model classes
# ------------------------------------------------------------------------------------------------
def q_sample(self, x_start, t, noise=None):
    """acv_ddim.py:241-246 (= pwcnet_ddim.py:453-458, igev_stereo_ddim.py:213-218)."""
    if noise is None:
        noise = torch.randn_like(x_start)
    hb = _host_buffers(self)
    ti = _time_index(t)
    return ops.q_sample(x_start, noise, hb["sqrt_ac"][ti], hb["sqrt_1m_ac"][ti])


def predict_noise_from_start(self, x_t, t, x0):
    """acv_ddim.py:248-252."""
    hb = _host_buffers(self)
    ti = _time_index(t)
    return ops.predict_noise_from_start(x_t, x0, hb["sqrt_recip"][ti], hb["sqrt_recipm1"][ti])


# ------------------------------------------------------------------------------------------------
# ACVNet_DDIM
# ------------------------------------------------------------------------------------------------
def _acv_convs(self, volume_f):
    """The reference's 3-D conv stack (acv_ddim.py:261-266) — unchanged PyTorch; returns classif2's [B,1,D,h,w]."""
    cost0 = self.dres0(volume_f)
    cost0 = self.dres1(cost0) + cost0
    out1 = self.dres2(cost0)
    out2 = self.dres3(out1)
    return self.classif2(out2)


def _acv_aggregate(self, volume_f, h, w):
    """Conv stack + trilinear upsample to the full-resolution logits (acv_ddim.py:261-268)."""
    cost2 = F.interpolate(_acv_convs(self, volume_f), [self.maxdisp, h * 4, w * 4], mode="trilinear")
    return torch.squeeze(cost2, 1)


def acv_model_predictions(self, volume, noise, t):
    """ACVNet_DDIM.model_predictions (acv_ddim.py:254-296): returns (pred_noise, x_start, pred, pred_volume2)."""
    b, c, d, h, w = volume.shape
    ti = _time_index(t)
    hb = _host_buffers(self)
    shift = _time_shift(self, t, b, d, volume.device)
    vol_f, n = ops.volume_filter(volume, noise, shift, self.scale, return_n=True)
    cost2 = _acv_aggregate(self, vol_f, h, w)
    r = ops.softmax_regress(cost2, return_prob=True)
    pred = r["disp"]
    H, W = pred.shape[-2:]
    disp_q = ops.downsample_bilinear(pred, (H // 4, W // 4), clamp=(0, self.maxdisp - 1), post_scale=0.25)
    x_start = ops.xstart_from_disp(disp_q, d, self.scale)
    pred_noise = ops.predict_noise_from_start(n, x_start, hb["sqrt_recip"][ti], hb["sqrt_recipm1"][ti])
    return pred_noise, x_start, pred, r["prob"]


def _acv_ddim_sample_impl(self, volume, used, asd, regen=None):
    """ddim_sample of ACVNet_DDIM.  `regen = (concat_left, concat_right, att_softmax)` switches the per-step filter from
    `volume * n` (reads the 398 MB volume) to the streaming producer that re-creates (concat * att) * n from the
    quarter-res features; both give bit-identical filtered volumes.  With `regen`, `volume` is only consulted for its shape."""
    batch, channel, depth, h, w = volume.shape
    dev = asd.device
    hb = _host_buffers(self)
    shape = (batch, 48, h, w)
    torch.randn(shape, device=dev)                         # acv_ddim.py:310 — drawn and never used
    img = asd
    used = used.float().contiguous()
    H, W = used.shape[-2:]
    cof = [0.5, 0.0, 0.0, 0.0, 0.2, 0.3]
    pairs = _time_pairs(self)
    disps = [used.reshape(batch, H, W)]
    ens = ops.ensemble([disps[0]], [cof[0]]) if self.use_ensemble and len(pairs) + 1 == len(cof) else None
    mask = torch.zeros((batch, h, w), dtype=torch.float32, device=dev)
    disp = None
    shifts = [_time_shift(self, torch.full((batch,), time, device=dev, dtype=torch.long), batch, depth, dev)
              for time, _ in pairs]
    n = ops.filter_factor(img, shifts[0], self.scale) if regen is not None else None
    vol_f = None
    for i, (time, time_next) in enumerate(pairs):
        shift = shifts[i]
        if regen is not None:
            cl, cr, att_w = regen
            vol_f = ops.concat_volume_weighted(cl, cr, depth, mask_left=False, att_weights=att_w, n=n, out=vol_f)
        else:
            vol_f = ops.volume_filter(volume, img, shift, self.scale, out=vol_f)
        cost_v = _acv_convs(self, vol_f)
        # F.upsample(trilinear) + softmax + regression + uncertainty + vote in one kernel: the [B,192,H,W] logits and
        # the probability volume of acv_ddim.py:267-270, :324-329 are never materialised (ddim_sample does not return them)
        r = ops.upsample_softmax_regress(cost_v, (self.maxdisp, h * 4, w * 4), used=used if self.renewal else None,
                                         vote_thresholds=(1.0, 3.0) if self.renewal else None,
                                         ens_acc=ens, ens_coef=cof[i + 1] if ens is not None else 0.0)
        disp = r["disp"]
        disps.append(disp)
        last = time_next < 0
        kw = {}
        if not last:
            san, c, sigma = _update_coefficients(self, time, time_next)
            noise = torch.randn_like(img)
            torch.randint(time, time + 1, (1,), device=dev)
            torch.randn_like(asd)                          # q_sample's draw (acv_ddim.py:243); its result only
            no = torch.rand_like(asd, dtype=torch.float64)  # ... types rand_like: float64 uniform (acv_ddim.py:360)
            kw = dict(sqrt_alpha_next=san, c=c, sigma=sigma, step_noise=noise, renoise=no)
            if regen is not None:
                kw.update(shift_next=shifts[i + 1], want_n_next=True)
        st = ops.ddim_step(disp=disp, xt=img, shift=shift, scale=self.scale, sqrt_recip=hb["sqrt_recip"][time],
                           sqrt_recipm1=hb["sqrt_recipm1"][time], last_step=last,
                           disp_clamp_hi=float(self.maxdisp - 1), vote=r.get("vote"), mask=mask, **kw)
        img = st["x_next"]
        n = st["n_next"]
    if self.use_ensemble:
        final = torch.stack(disps, dim=0)
        if ens is None:
            ens = ops.ensemble(disps, cof[: len(disps)])
        return ens, final
    return disp


@torch.no_grad()
def acv_ddim_sample(self, volume, used, asd):
    """ACVNet_DDIM.ddim_sample (acv_ddim.py:298-370): returns (final_prediction, final) when
    use_ensemble else the last disparity.  RNG draws (order, shapes, dtypes) follow the reference:
    randn(shape) once (unused there too), then per non-final step randn_like(img), randint,
    randn_like(asd) [inside q_sample, result unused], rand_like(float64)."""
    return _acv_ddim_sample_impl(self, volume, used, asd)


# ------------------------------------------------------------------------------------------------
# ACVNet_DDIM.forward / ACVNet.forward (tier 3)
# ------------------------------------------------------------------------------------------------
# In eval mode the reference pushes ac_volume through dres0..dres3 / classif2 / upsample / softmax / regression
# (acv_ddim.py:392-401) and never uses the result (`pred2` is dead: the branch returns the sampler's prediction).  The
# drop-in does not run that pass (no module there has eval-mode side effects).  Set to False to run it anyway, e.g. when
# forward hooks on those modules must fire once more per call.
SKIP_UNUSED_EVAL_PASS = True


def _is_ours(self, name, fn):
    return getattr(type(self), name, None) is fn


def _acv_attention_logits(self, features_left, features_right):
    """gwc volume -> patch convs -> attention hourglass (acv_ddim.py:375-386, acv.py:170-197): the volume op and, when no
    gradient is needed, the depth-wise patch chain are the fused kernels; the 3-D hourglass is the reference's."""
    from . import functional as Fn
    gwc_volume = Fn.build_gwc_volume(features_left["gwc_feature"], features_right["gwc_feature"], self.maxdisp // 4,
                                     self.num_groups)
    convs = (self.patch, self.patch_l1, self.patch_l2, self.patch_l3)
    needs_grad = torch.is_grad_enabled() and (gwc_volume.requires_grad or any(m.weight.requires_grad for m in convs))
    plain = all(m.bias is None and type(m).forward is torch.nn.Conv3d.forward for m in convs)
    if not needs_grad and plain and gwc_volume.is_cuda and gwc_volume.dtype == torch.float32:
        patch_volume = ops.acv_patch_volume(gwc_volume, self.patch.weight.detach(), self.patch_l1.weight.detach(),
                                            self.patch_l2.weight.detach(), self.patch_l3.weight.detach())
    else:
        gwc_volume = self.patch(gwc_volume)
        patch_volume = torch.cat((self.patch_l1(gwc_volume[:, :8]), self.patch_l2(gwc_volume[:, 8:24]),
                                  self.patch_l3(gwc_volume[:, 24:40])), dim=1)
    cost_attention = self.dres1_att_(patch_volume)
    cost_attention = self.dres2_att_(cost_attention)
    return self.classif_att_(cost_attention)


def _initial_xstart(self, disp, mask_gt):
    """The 2-tap x_start volume of the quarter-res initial disparity, `(disp_volume_final * 2 - 1) * self.scale`
    (acv_ddim.py:403-419 eval, :425-440 training — the two differ only in how mask_gt is broadcast)."""
    b, c, h, w = disp.shape
    asd = ops.xstart_from_disp(disp.float().reshape(b, h, w).contiguous(), 48, self.scale)
    if mask_gt is not None:
        allone = ((torch.ones((), dtype=torch.float32, device=asd.device) / 48) * 2 - 1) * self.scale
        m = mask_gt if mask_gt.dim() == 4 or self.training is False else mask_gt.unsqueeze(1)
        asd = torch.where(m == 0, allone, asd)
    return asd


def _regress_head(cost_q, size, maxdisp):
    """F.upsample(cost, size, 'trilinear') -> squeeze -> F.softmax(dim=1) -> disparity_regression (acv_ddim.py:396-401,
    :460-480).  Without autograd everything is one kernel (the 398 MB logits never exist); with autograd the trilinear
    upsample stays ATen (out of scope) and softmax + regression run fused, forward and backward."""
    from . import functional as Fn
    if torch.is_grad_enabled() and cost_q.requires_grad:
        cost = torch.squeeze(F.interpolate(cost_q, list(size), mode="trilinear"), 1)
        return Fn.softmax_disparity_regression(cost, maxdisp)
    return ops.upsample_softmax_regress(cost_q.float(), size)["disp"]


def acv_ddim_forward(self, left, right, used, disp, mask_gt=None):
    """ACVNet_DDIM.forward (acv_ddim.py:372-482).  Same arguments, same return values ([pred] in eval mode,
    [pred_attention, pred0, pred1, pred2] in training mode), same RNG draws in the same order."""
    from . import functional as Fn
    features_left = self.feature_extraction(left)
    features_right = self.feature_extraction(right)
    att_weights = _acv_attention_logits(self, features_left, features_right)
    concat_feature_left = self.concatconv(features_left["gwc_feature"])
    concat_feature_right = self.concatconv(features_right["gwc_feature"])
    D = self.maxdisp // 4
    size = (self.maxdisp, left.size()[2], left.size()[3])
    if not self.training:
        fused = (not torch.is_grad_enabled() or not (concat_feature_left.requires_grad or att_weights.requires_grad)) \
            and _is_ours(self, "ddim_sample", acv_ddim_sample)
        asd = _initial_xstart(self, disp, mask_gt)
        if fused:
            cl, cr = concat_feature_left.detach().float().contiguous(), concat_feature_right.detach().float().contiguous()
            att_w = ops.att_softmax(att_weights.detach().float())
            if not SKIP_UNUSED_EVAL_PASS:
                ac_volume = ops.concat_volume_weighted(cl, cr, D, mask_left=False, att_weights=att_w)
                _regress_head(_acv_convs(self, ac_volume), size, self.maxdisp)
                del ac_volume
            b, c, h, w = cl.shape
            shape_only = torch.empty((), dtype=torch.float32, device=cl.device).expand(b, 2 * c, D, h, w)
            with torch.no_grad():
                pred, pred_all = _acv_ddim_sample_impl(self, shape_only, used, asd, regen=(cl, cr, att_w))
        else:
            ac_volume = Fn.acv_attention_volume(concat_feature_left, concat_feature_right, att_weights, D)
            if not SKIP_UNUSED_EVAL_PASS:
                _regress_head(_acv_convs(self, ac_volume), size, self.maxdisp)
            pred, pred_all = self.ddim_sample(ac_volume, used, asd)
        return [pred]

    asd = _initial_xstart(self, disp, mask_gt)
    t = torch.randint(0, self.num_timesteps, (1,), device=asd.device).long()
    noisy = self.q_sample(asd, t)                               # draws randn_like(asd); float64
    with torch.no_grad():                                       # the reference re-wraps the factor with torch.tensor()
        shift = _time_shift(self, t, asd.shape[0], asd.shape[1], asd.device)
        n32 = ops.filter_factor(noisy, shift, self.scale)
    ac_volume = Fn.acv_attention_volume(concat_feature_left, concat_feature_right, att_weights, D, n=n32)
    cost0 = self.dres0(ac_volume)
    cost0 = self.dres1(cost0) + cost0
    out1 = self.dres2(cost0)
    out2 = self.dres3(out1)
    pred_attention = _regress_head(att_weights, size, self.maxdisp)
    pred0 = _regress_head(self.classif0(cost0), size, self.maxdisp)
    pred1 = _regress_head(self.classif1(out1), size, self.maxdisp)
    pred2 = _regress_head(self.classif2(out2), size, self.maxdisp)
    return [pred_attention, pred0, pred1, pred2]


def acvnet_forward(self, left, right):
    """ACVNet.forward (acv.py:167-247): freeze_attn_weights / attn_weights_only switches and the training / eval return
    lists exactly as the reference."""
    from . import functional as Fn
    if self.freeze_attn_weights:
        with torch.no_grad():
            features_left = self.feature_extraction(left)
            features_right = self.feature_extraction(right)
            att_weights = _acv_attention_logits(self, features_left, features_right)
    else:
        features_left = self.feature_extraction(left)
        features_right = self.feature_extraction(right)
        att_weights = _acv_attention_logits(self, features_left, features_right)
    size = (self.maxdisp, left.size()[2], left.size()[3])
    if not self.attn_weights_only:
        concat_feature_left = self.concatconv(features_left["gwc_feature"])
        concat_feature_right = self.concatconv(features_right["gwc_feature"])
        ac_volume = Fn.acv_attention_volume(concat_feature_left, concat_feature_right, att_weights, self.maxdisp // 4)
        cost0 = self.dres0(ac_volume)
        cost0 = self.dres1(cost0) + cost0
        out1 = self.dres2(cost0)
        out2 = self.dres3(out1)
    if self.training:
        preds = []
        if not self.freeze_attn_weights:
            preds.append(_regress_head(att_weights, size, self.maxdisp))
        if not self.attn_weights_only:
            preds += [_regress_head(self.classif0(cost0), size, self.maxdisp),
                      _regress_head(self.classif1(out1), size, self.maxdisp),
                      _regress_head(self.classif2(out2), size, self.maxdisp)]
        return preds
    if self.attn_weights_only:
        return [_regress_head(att_weights, size, self.maxdisp)]
    return [_regress_head(self.classif2(out2), size, self.maxdisp)]


# ------------------------------------------------------------------------------------------------
# PWCNet_ddim
# ------------------------------------------------------------------------------------------------
def pcw_model_predictions(self, volume, noise, t, features_left, features_right):
    """PWCNet_ddim.model_predictions (pwcnet_ddim.py:466-528), the reference's signature and return values."""
    return _pcw_model_predictions_impl(self, volume, noise, t, features_left, features_right, want_prob=True)[:4]


def _pcw_model_predictions_impl(self, volume, noise, t, features_left, features_right, want_prob):
    """Returns (pred_noise, x_start, disp_finetune, pred3_volume or None, cost3 logits [B,maxdisp,H,W]).
    want_prob=False is what pcw_ddim_sample uses on its non-final steps: there the probability volume is only ever reduced to
    the uncertainty of the refined disparity (pwcnet_ddim.py:553-558), which `ops.softmax_uncertainty_vote` takes from the
    logits — the 398 MB/pair write of pred3_volume and its read back never happen.
    Filter multiply :468-472, softmax + regression :483-484 (the probability volume IS returned here, so it is written
    once by the same kernel), warp + the +-24 correlation volume :493-494, x_start scatter :504-524 and pred_noise :526
    are the CUDA ops; the 3-D hourglasses, `dispupsample`, `refinenet3` and the two F.upsample calls are the reference's
    own modules / ATen (out of scope)."""
    b, c, d, h, w = volume.shape
    ti = _time_index(t)
    hb = _host_buffers(self)
    shift = _time_shift(self, t, b, d, volume.device)
    volume_noise, n = ops.volume_filter(volume, noise, shift, self.scale, return_n=True)
    out1 = self.dres2(volume_noise)
    out2 = self.dres3(out1)
    out3 = self.dres4(out2)
    cost3 = self.classif3(out3)
    cost3 = F.interpolate(cost3, [self.maxdisp, h * 4, w * 4], mode="trilinear", align_corners=True)
    cost3 = torch.squeeze(cost3, 1)
    r = ops.softmax_regress(cost3, return_prob=want_prob)
    pred3_volume = r["prob"] if want_prob else None
    pred3 = torch.unsqueeze(r["disp"], 1)
    refinenet_feature_left = F.interpolate(features_left["finetune_feature"], [h * 4, w * 4], mode="bilinear",
                                           align_corners=True)
    refinenet_feature_right = F.interpolate(features_right["finetune_feature"], [h * 4, w * 4], mode="bilinear",
                                            align_corners=True)
    # warp + (left - warped) + copy of left + the +-24 volume written straight into the refinement input: the reference's
    # torch.cat((left - right_warp, left, dispupsample(pred3), pred3, cost), 1) (:497-499) without the subtraction and
    # concatenation passes over 146 full-resolution channels
    pred3feature = self.dispupsample(pred3)
    B_, Cf, Hf, Wf = refinenet_feature_left.shape
    Cd, S = pred3feature.shape[1], 2 * 24 + 1
    refinenet_combine = torch.empty((B_, 2 * Cf + Cd + 1 + S, Hf, Wf), dtype=torch.float32, device=volume.device)
    ops.refine_input_assemble(refinenet_feature_left.float(), refinenet_feature_right.float(), pred3, 24, 1,
                              diff_out=refinenet_combine[:, :Cf], copy_out=refinenet_combine[:, Cf:2 * Cf],
                              corr_out=refinenet_combine[:, 2 * Cf + Cd + 1:])
    refinenet_combine[:, 2 * Cf:2 * Cf + Cd] = pred3feature
    refinenet_combine[:, 2 * Cf + Cd:2 * Cf + Cd + 1] = pred3
    disp_finetune = self.refinenet3(refinenet_combine, pred3)
    disp_finetune = torch.squeeze(disp_finetune, 1)
    H, W = disp_finetune.shape[-2:]
    disp_q = ops.downsample_bilinear(disp_finetune.float().contiguous(), (H // 4, W // 4), clamp=(0, self.maxdisp - 1),
                                     post_scale=0.25)
    x_start = ops.xstart_from_disp(disp_q, d, self.scale)
    pred_noise = ops.predict_noise_from_start(n, x_start, hb["sqrt_recip"][ti], hb["sqrt_recipm1"][ti])
    return pred_noise, x_start, disp_finetune, pred3_volume, cost3


@torch.no_grad()
def pcw_ddim_sample(self, volume, used, asd, features_left, features_right):
    """PWCNet_ddim.ddim_sample (pwcnet_ddim.py:530-602): T=3, thresholds 1 / (last step) 2, cumulative
    `asd = q_sample(asd, t)`, ensemble [0.9, 0, 0, 0.1]; returns (final, pred3_volume)."""
    batch, channel, depth, h, w = volume.shape
    dev = volume.device
    hb = _host_buffers(self)
    shape = (batch, 48, h, w)
    img = torch.randn(shape, device=dev)                   # pwcnet_ddim.py:541 — the start state
    used = used.float().contiguous()
    H, W = used.shape[-2:]
    cof = [0.9, 0.0, 0.0, 0.1]
    pairs = _time_pairs(self)
    disps = [used.reshape(batch, H, W)]
    mask = torch.zeros((batch, h, w), dtype=torch.float32, device=dev)
    pred3_volume = None
    disp = None
    # with our own model_predictions bound, the probability volume is materialised on the LAST step only (the one the method
    # returns); a user-supplied model_predictions is called as is and its probability volume is used
    ours = _is_ours(self, "model_predictions", pcw_model_predictions)
    for i, (time, time_next) in enumerate(pairs):
        time_cond = torch.full((batch,), time, device=dev, dtype=torch.long)
        last = time_next < 0
        cost3 = None
        if ours:
            pred_noise, x_start, disp, pred3_volume, cost3 = _pcw_model_predictions_impl(
                self, volume, img, time_cond, features_left, features_right, want_prob=last)
        else:
            pred_noise, x_start, disp, pred3_volume = self.model_predictions(volume, img, time_cond, features_left,
                                                                             features_right)
        disp = disp.float().contiguous()
        disps.append(disp)
        if last:
            img = x_start
            continue
        vote = None
        if self.renewal:
            # uncertainty of the REFINED disparity against the pre-refinement distribution (pwcnet_ddim.py:553-558)
            if pred3_volume is None:
                vote = ops.softmax_uncertainty_vote(disp, cost3, used, 1.0, 1.0)
            else:
                vote = ops.uncertainty_vote(disp, pred3_volume, used, 1.0, 1.0)
        san, c, sigma = _update_coefficients(self, time, time_next)
        noise = torch.randn_like(img)
        torch.randint(time, time + 1, (1,), device=dev)
        qn = torch.randn_like(asd)
        shift = _time_shift(self, time_cond, batch, depth, dev)
        st = ops.ddim_step(disp=disp, xt=img, shift=shift, scale=self.scale, sqrt_recip=hb["sqrt_recip"][time],
                           sqrt_recipm1=hb["sqrt_recipm1"][time], last_step=False,
                           disp_clamp_hi=float(self.maxdisp - 1), vote=vote, mask=mask,
                           sqrt_alpha_next=san, c=c, sigma=sigma, step_noise=noise,
                           asd=asd, q_noise=qn, sqrt_ac=hb["sqrt_ac"][time], sqrt_1m_ac=hb["sqrt_1m_ac"][time],
                           want_asd_out=True)
        img = st["x_next"]
        asd = st["asd_out"]                                # cumulative re-noising (pwcnet_ddim.py:591)
    if self.use_ensemble:
        final = ops.ensemble(disps, cof[: len(disps)])
        return final, pred3_volume
    return disp, pred3_volume


# ------------------------------------------------------------------------------------------------
# IGEVStereo_ddim
# ------------------------------------------------------------------------------------------------
def _igev_autocast(self):
    """`autocast(enabled=self.args.mixed_precision)` as the reference module defines it (igev_stereo_ddim.py:13-22)."""
    import sys
    mod = sys.modules.get(type(self).__module__)
    ac = getattr(mod, "autocast", None)
    if ac is not None:
        return ac(enabled=self.args.mixed_precision)
    return torch.autocast("cuda", enabled=bool(self.args.mixed_precision))


def _igev_gru_loop(self, coords0, coords1, flow_init, iters, net_list, inp_list, corr_fn, n32, stem_2x):
    """The GRU iteration loop of IGEVStereo_ddim.model_predictions (igev_stereo_ddim.py:233-264): the update block and the
    convex upsampler are the reference's own modules; every `corr_fn` call is the fused geometry lookup kernel."""
    if flow_init is not None:
        coords1 = coords1 + flow_init
    flow_up = None
    for itr in range(iters):
        coords1 = coords1.detach()
        flow = coords1 - coords0
        corr = corr_fn(flow, coords1, n32)
        with _igev_autocast(self):
            if self.args.n_gru_layers == 3 and self.args.slow_fast_gru:
                net_list = self.update_block(net_list, inp_list, iter32=True, iter16=False, iter08=False, update=False)
            if self.args.n_gru_layers >= 2 and self.args.slow_fast_gru:
                net_list = self.update_block(net_list, inp_list, iter32=self.args.n_gru_layers == 3, iter16=True, iter08=False,
                                             update=False)
            net_list, up_mask, delta_flow = self.update_block(net_list, inp_list, corr, flow,
                                                              iter16=self.args.n_gru_layers == 3,
                                                              iter08=self.args.n_gru_layers >= 2)
        coords1 = coords1 + delta_flow
        if itr < iters - 1:
            continue
        if up_mask is None:
            import sys
            upflow8 = getattr(sys.modules.get(type(self).__module__), "upflow8")
            flow_up = upflow8(coords1 - coords0)
        else:
            flow_up = self.upsample_disp(coords1 - coords0, up_mask, stem_2x)
        flow_up = flow_up[:, :1]
    return flow_up, coords1


def igev_model_predictions(self, coords0, coords1, flow_init, iters, net_list, inp_list, corr_fn, noise, t, stem_2x):
    """IGEVStereo_ddim.model_predictions (igev_stereo_ddim.py:226-292): returns (pred_noise, x_start, pred, coords1)."""
    b, D = noise.shape[0], noise.shape[1]
    ti = _time_index(t)
    hb = _host_buffers(self)
    shift = _time_shift(self, t, b, D, noise.device)
    n, n32 = ops.filter_factor_pair(noise, shift, self.scale)
    pred, coords1 = _igev_gru_loop(self, coords0, coords1, flow_init, iters, net_list, inp_list, corr_fn, n32, stem_2x)
    H, W = pred.shape[-2:]
    disp_q = ops.downsample_bilinear(pred.float().reshape(b, H, W), (H // 4, W // 4), clamp=(0, D - 1), post_scale=0.25)
    true_coords1 = torch.clamp(coords0.reshape(b, H // 4, W // 4).float() + disp_q, 0, D - 1)
    x_start = ops.xstart_from_disp(true_coords1, D, self.scale)
    pred_noise = ops.predict_noise_from_start(n, x_start, hb["sqrt_recip"][ti], hb["sqrt_recipm1"][ti])
    return pred_noise, x_start, pred, coords1


@torch.no_grad()
def igev_ddim_sample(self, coords0, coords1, flow_init, iters, net_list, inp_list, corr_fn, used, asd, stem_2x):
    """IGEVStereo_ddim.ddim_sample (igev_stereo_ddim.py:294-359): T = 2 DDIM steps around the GRU loop; renewal vote
    |disp - used| < 5, fallback to `used` where |disp - used| >= 3, non-cumulative re-noising q_sample(asd, t), ensemble
    [0.6, 0.1, 0.3].  RNG draws in the reference's order: randn_like(asd) (start state), then per non-final step
    randn_like(img), randint, randn_like(asd) (inside q_sample).  The reference is only shape-consistent for batch 1
    (its `where(mask.unsqueeze(1) == 0, used, disp)` broadcasts across the batch); this follows it exactly at B = 1
    and treats every sample independently otherwise."""
    batch, d, h, w = asd.shape
    dev = asd.device
    hb = _host_buffers(self)
    pairs = _time_pairs(self)
    img = torch.randn_like(asd, device=dev)
    used_map = used.float().reshape(batch, used.shape[-2], used.shape[-1]).contiguous()
    H, W = used_map.shape[-2:]
    c0 = coords0.float().reshape(batch, h, w).contiguous()
    disps = [used_map]
    mask = torch.zeros((batch, h, w), dtype=torch.float32, device=dev)
    for time, time_next in pairs:
        time_cond = torch.full((batch,), time, device=dev, dtype=torch.long)
        shift = _time_shift(self, time_cond, batch, d, dev)
        n32 = ops.filter_factor(img, shift, self.scale)
        pred, coords1 = _igev_gru_loop(self, coords0, coords1, flow_init, iters, net_list, inp_list, corr_fn, n32, stem_2x)
        disp = pred.float().reshape(batch, H, W).contiguous()
        # fallback to the initial disparity where the sampled one strays (igev_stereo_ddim.py:323-325)
        disps.append(ops.select_close(disp, used_map, 3.0))
        last = time_next < 0
        kw = {}
        if not last:
            san, c, sigma = _update_coefficients(self, time, time_next)
            noise = torch.randn_like(img)
            torch.randint(time, time + 1, (1,), device=dev)
            qn = torch.randn_like(asd)
            kw = dict(sqrt_alpha_next=san, c=c, sigma=sigma, step_noise=noise, asd=asd, q_noise=qn,
                      sqrt_ac=hb["sqrt_ac"][time], sqrt_1m_ac=hb["sqrt_1m_ac"][time])
        st = ops.ddim_step(disp=disp, xt=img, shift=shift, scale=self.scale, sqrt_recip=hb["sqrt_recip"][time],
                           sqrt_recipm1=hb["sqrt_recipm1"][time], last_step=last, disp_clamp_hi=float(d - 1), coords0=c0,
                           used=used_map if self.renewal else None, vote_thr_dif=5.0, mask=mask, **kw)
        img = st["x_next"]
    if self.use_ensemble:
        cof = [0.6, 0.1, 0.3]
        return ops.ensemble(disps, cof[: len(disps)])
    return disps[-1]
