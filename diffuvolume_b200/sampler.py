"""Tier-2 drop-ins: the DiffuVolume sampler methods of the reference's model classes, re-expressed on
the fused sm_100a kernels.  They are bound onto the reference's OWN nn.Modules by
diffuvolume_b200.install (same signatures, same return values, same RNG draws in the same order and
dtypes), so the conv stacks, parameters and checkpoints stay the reference's.

ACVNet_DDIM   (SceneFlow/models/acv_ddim.py):  q_sample :241-246, predict_noise_from_start :248-252,
                                               model_predictions :254-296, ddim_sample :298-370
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
