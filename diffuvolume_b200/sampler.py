"""Tier-2 drop-ins: the DiffuVolume sampler methods of the reference's model classes, re-expressed on
the fused sm_100a kernels.  They are bound onto the reference's OWN nn.Modules by
diffuvolume_b200.install (same signatures, same return values, same RNG draws in the same order and
dtypes), so the conv stacks, parameters and checkpoints stay the reference's.

ACVNet_DDIM   (SceneFlow/models/acv_ddim.py):  q_sample :241-246, predict_noise_from_start :248-252,
                                               model_predictions :254-296, ddim_sample :298-370
IGEVStereo_ddim (KITTI15/core/igev_stereo_ddim.py): q_sample :213-218, predict_noise_from_start :220-224,
                                               model_predictions :226-292, ddim_sample :294-359 (the GRU iteration
                                               loop keeps the reference's update block / upsampler modules)
PWCNet_ddim   (KITTI12/models/pwcnet_ddim.py): q_sample :453-458, predict_noise_from_start :460-464,
                                               ddim_sample :530-602 (model_predictions keeps the reference's
                                               conv / warp / refinement code, with the fused filter,
                                               softmax-regression and corr-volume ops underneath)

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


@torch.no_grad()
def acv_ddim_sample(self, volume, used, asd):
    """ACVNet_DDIM.ddim_sample (acv_ddim.py:298-370): returns (final_prediction, final) when
    use_ensemble else the last disparity.  RNG draws (order, shapes, dtypes) follow the reference:
    randn(shape) once (unused there too), then per non-final step randn_like(img), randint,
    randn_like(asd) [inside q_sample, result unused], rand_like(float64)."""
    batch, channel, depth, h, w = volume.shape
    dev = volume.device
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
    for i, (time, time_next) in enumerate(pairs):
        time_cond = torch.full((batch,), time, device=dev, dtype=torch.long)
        shift = _time_shift(self, time_cond, batch, depth, dev)
        vol_f = ops.volume_filter(volume, img, shift, self.scale)
        cost_v = _acv_convs(self, vol_f)
        del vol_f
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
        st = ops.ddim_step(disp=disp, xt=img, shift=shift, scale=self.scale, sqrt_recip=hb["sqrt_recip"][time],
                           sqrt_recipm1=hb["sqrt_recipm1"][time], last_step=last,
                           disp_clamp_hi=float(self.maxdisp - 1), vote=r.get("vote"), mask=mask, **kw)
        img = st["x_next"]
    if self.use_ensemble:
        final = torch.stack(disps, dim=0)
        if ens is None:
            ens = ops.ensemble(disps, cof[: len(disps)])
        return ens, final
    return disp


# ------------------------------------------------------------------------------------------------
# PWCNet_ddim
# ------------------------------------------------------------------------------------------------
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
    for i, (time, time_next) in enumerate(pairs):
        time_cond = torch.full((batch,), time, device=dev, dtype=torch.long)
        pred_noise, x_start, disp, pred3_volume = self.model_predictions(volume, img, time_cond, features_left,
                                                                         features_right)
        disp = disp.float().contiguous()
        disps.append(disp)
        last = time_next < 0
        if last:
            img = x_start
            continue
        vote = None
        if self.renewal:
            # uncertainty of the REFINED disparity against the pre-refinement distribution (pwcnet_ddim.py:553-558)
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
        disps.append(torch.where(torch.abs(disp - used_map) < 3, disp, used_map))
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
