"""The ACVNet+DiffuVolume hot path as one device-resident kernel sequence.

This is the unit BASELINE.json's metric counts ("volume+DDIM-filter pairs/s"): for a batch of
stereo pairs, 1x group-wise correlation volume, 1x concat volume with the ACV attention weights
fused, then T DDIM steps of {filter multiply, softmax + disparity regression + uncertainty + renewal
vote + ensemble accumulate, fused x_start / pred_noise / DDIM update / re-noise}, i.e. everything
ACVNet_DDIM.forward (eval) + ddim_sample (SceneFlow/models/acv_ddim.py:372-422, :298-370) do
outside the 2-D/3-D convolutions.  The convolutions stay on PyTorch and are not part of this
path: their outputs (features, attention logits, per-step cost logits) are inputs here.

Every op is a hand-written sm_100a kernel reached through the C-ABI (diffuvolume_b200.ops).
"""
from __future__ import annotations

import contextlib
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops


def cosine_alphas_cumprod(timesteps: int = 1000, s: float = 0.008) -> np.ndarray:
    """Host-side float64 cosine schedule (cosine_beta_schedule, SceneFlow/models/acv_ddim.py:113-119,
    followed by the cumprod of :134-135).  Host scalars only — nothing here touches the device."""
    x = np.linspace(0, timesteps, timesteps + 1, dtype=np.float64)
    ac = np.cos(((x / timesteps) + s) / (1 + s) * math.pi * 0.5) ** 2
    ac = ac / ac[0]
    betas = np.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)
    return np.cumprod(1.0 - betas)


def ddim_time_pairs(num_timesteps: int, sampling_timesteps: int) -> List[Tuple[int, int]]:
    """acv_ddim.py:306-308 — torch.linspace(-1, T-1, S+1) (float32) truncated to int, reversed, paired."""
    times = torch.linspace(-1, num_timesteps - 1, steps=sampling_timesteps + 1)
    times = list(reversed(times.int().tolist()))
    return list(zip(times[:-1], times[1:]))


@dataclass
class DdimSchedule:
    """The float64 constants the sampler needs, as host scalars."""
    num_timesteps: int = 1000
    sampling_timesteps: int = 5
    eta: float = 1.0
    scale: float = 1.0
    alphas_cumprod: np.ndarray = field(default=None, repr=False)

    def __post_init__(self):
        if self.alphas_cumprod is None:
            self.alphas_cumprod = cosine_alphas_cumprod(self.num_timesteps)
        self.alphas_cumprod = np.asarray(self.alphas_cumprod, dtype=np.float64)

    def time_pairs(self):
        return ddim_time_pairs(self.num_timesteps, self.sampling_timesteps)

    def sqrt_recip(self, t):
        return float(np.sqrt(1.0 / self.alphas_cumprod[t]))

    def sqrt_recipm1(self, t):
        return float(np.sqrt(1.0 / self.alphas_cumprod[t] - 1))

    def sqrt_ac(self, t):
        return float(np.sqrt(self.alphas_cumprod[t]))

    def sqrt_1m_ac(self, t):
        return float(np.sqrt(1.0 - self.alphas_cumprod[t]))

    def update_coefficients(self, t, t_next):
        """acv_ddim.py:347-351 -> (sqrt(alpha_next), c, sigma)."""
        a, an = self.alphas_cumprod[t], self.alphas_cumprod[t_next]
        sigma = self.eta * np.sqrt((1 - a / an) * (1 - an) / (1 - a))
        c = np.sqrt(1 - an - sigma ** 2)
        return float(np.sqrt(an)), float(c), float(sigma)


ACV_ENSEMBLE = (0.5, 0.0, 0.0, 0.0, 0.2, 0.3)   # acv_ddim.py:367


class AcvHotPath:
    """Runs the hot path for a batch on one device.

    filter_mode:
      'regenerate' — each DDIM step re-produces the filtered volume straight from the 1/4-res concat
                     features and two precomputed fp32 factor maps (softmax of the attention logits, once per
                     pair; the filter factor n, emitted by the previous fused DDIM step): reads 21 MB, writes
                     398 MB per pair and step;
      'volume'     — each step multiplies the materialised ac_volume (reads 404 MB, writes 398 MB):
                     the reference's own op boundary (acv_ddim.py:260).
    Both produce bit-identical volumes (same roundings in the same order).

    regress_mode:
      'logits'         — costs[i] are the full-resolution logits [B,maxdisp,H,W] (the output of F.upsample in
                         acv_ddim.py:267): the op boundary BASELINE.json's metric is defined on;
      'fused_upsample' — costs[i] are the quarter-resolution outputs [B,1,D,h,w] of the last 3-D conv and the
                         trilinear x4 upsample is fused into the regression kernel.
    """

    def __init__(self, schedule: Optional[DdimSchedule] = None, num_groups: int = 40, maxdisp: int = 192,
                 filter_mode: str = "regenerate", ensemble: Sequence[float] = ACV_ENSEMBLE,
                 thr_dif: float = 1.0, thr_unc: float = 3.0, regress_mode: str = "logits"):
        assert filter_mode in ("regenerate", "volume")
        assert regress_mode in ("logits", "fused_upsample")
        self.regress_mode = regress_mode
        self.sched = schedule or DdimSchedule()
        self.G = num_groups
        self.maxdisp = maxdisp
        self.D = maxdisp // 4
        self.filter_mode = filter_mode
        self.cof = tuple(ensemble)
        self.thr = (thr_dif, thr_unc)
        self._buf: Dict[str, torch.Tensor] = {}
        assert len(self.cof) == self.sched.sampling_timesteps + 1

    def _out(self, name, shape, device):
        t = self._buf.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.device != device:
            t = torch.empty(shape, dtype=torch.float32, device=device)
            self._buf[name] = t
        return t

    def __call__(self, feat_l, feat_r, cfeat_l, cfeat_r, att_logits, costs: Sequence[torch.Tensor], used, disp_q,
                 shifts: Sequence[torch.Tensor], step_noises: Sequence[torch.Tensor],
                 renoises: Sequence[torch.Tensor], keep_volumes: bool = False, timer=None):
        """All arguments are device tensors:
        feat_* [B,C,h,w] gwc features; cfeat_* [B,Cc,h,w] concat features; att_logits [B,1,D,h,w];
        costs[i] [B,maxdisp,H,W] logits of step i (len T, or len 1 to reuse one buffer), or a callable
        i -> tensor invoked when step i needs its logits (e.g. to stage them from the host);
        used [B,H,W] initial disparity; disp_q [B,h,w] its quarter-res version (/4);
        shifts[i] [B,D] DynamicHead shift at step i; step_noises[i] / renoises[i] the injected
        randn_like / rand_like tensors of step i (i < T-1).  `timer(name)` may return a context manager
        wrapped around each kernel category (bench.py uses CUDA events).  Returns dict(pred, mask, x_last)."""
        tm = timer if timer is not None else (lambda name: contextlib.nullcontext())
        dev = feat_l.device
        B, _, h, w = feat_l.shape
        D, sched = self.D, self.sched
        Cc = cfeat_l.shape[1]
        H, W = used.shape[-2:]
        with tm("gwc_volume"):
            gwc = ops.gwc_volume(feat_l, feat_r, D, self.G, out=self._out("gwc", (B, self.G, D, h, w), dev))
        regen = self.filter_mode == "regenerate"
        with tm("concat_acv"):
            if regen:
                # softmax over D of the attention logits: once per pair, reused by all T filter passes
                att_w = ops.att_softmax(att_logits)
                ac = ops.concat_volume_weighted(cfeat_l, cfeat_r, D, mask_left=False, att_weights=att_w,
                                                out=self._out("ac", (B, 2 * Cc, D, h, w), dev))
            else:
                ac = ops.concat_volume(cfeat_l, cfeat_r, D, mask_left=False, att_logits=att_logits,
                                       out=self._out("ac", (B, 2 * Cc, D, h, w), dev))
        img = ops.xstart_from_disp(disp_q, D, sched.scale)
        ens = ops.ensemble([used], [self.cof[0]])
        mask = torch.zeros((B, h, w), dtype=torch.float32, device=dev)
        vol_f = self._out("vol_f", (B, 2 * Cc, D, h, w), dev)
        pairs = sched.time_pairs()
        n = None
        if regen:
            with tm("filter_factor"):
                n = ops.filter_factor(img, shifts[0], sched.scale)
        for i, (t, t_next) in enumerate(pairs):
            shift = shifts[i]
            with tm("filter"):
                if regen:
                    ops.concat_volume_weighted(cfeat_l, cfeat_r, D, mask_left=False, att_weights=att_w, n=n, out=vol_f)
                else:
                    ops.volume_filter(ac, img, shift, sched.scale, out=vol_f)
            # (3-D conv aggregation of vol_f happens here in the full network — out of scope)
            cost = costs(i) if callable(costs) else costs[i if len(costs) > 1 else 0]
            with tm("softmax_regress"):
                if self.regress_mode == "fused_upsample":
                    r = ops.upsample_softmax_regress(cost, (self.maxdisp, H, W), used=used, vote_thresholds=self.thr,
                                                     ens_acc=ens, ens_coef=self.cof[i + 1])
                else:
                    r = ops.softmax_regress(cost, used=used, vote_thresholds=self.thr, ens_acc=ens,
                                            ens_coef=self.cof[i + 1])
            last = t_next < 0
            kw = {}
            if not last:
                san, c, sigma = sched.update_coefficients(t, t_next)
                kw = dict(sqrt_alpha_next=san, c=c, sigma=sigma, step_noise=step_noises[i], renoise=renoises[i],
                          shift_next=shifts[i + 1], want_n_next=regen)
            with tm("ddim_step"):
                st = ops.ddim_step(disp=r["disp"], xt=img, shift=shift, scale=sched.scale,
                                   sqrt_recip=sched.sqrt_recip(t), sqrt_recipm1=sched.sqrt_recipm1(t),
                                   last_step=last, disp_clamp_hi=float(self.maxdisp - 1), vote=r["vote"], mask=mask,
                                   **kw)
            img = st["x_next"]
            n = st["n_next"]
        out = {"pred": ens, "mask": mask, "x_last": img}
        if keep_volumes:
            out.update(gwc=gwc, ac=ac, vol_f=vol_f)
        return out

    def graphed(self, **inputs):
        """Capture one call into a CUDA graph (B200 rule: launch-bound loops are replayed, not re-issued).  At batch 1
        — the reference's own evaluation batch (SceneFlow/test_sceneflow_ddim.py:48) — the 21 launches of a pair take
        longer to issue from Python than to execute.  Returns (replay, outputs): `replay()` re-runs the captured
        sequence on the current stream reading the SAME input tensors (update them in place), `outputs` are the
        tensors every replay overwrites.  The C-ABI is capture-safe: no allocation, no synchronisation, the tile
        counters of the persistent kernels re-arm themselves on the device.  (A captured launch keeps its counter
        slot, so replays must not overlap launches of the same kernels on other streams.)"""
        if callable(inputs.get("costs")):
            raise ValueError("graphed() needs device-resident cost tensors, not a staging callback")
        dev = inputs["feat_l"].device
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(2):          # warm-up outside the capture: output buffers, function attributes
                self(**inputs)
        torch.cuda.current_stream(dev).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = self(**inputs)
        return graph.replay, out

    def launches_per_call(self) -> int:
        T = self.sched.sampling_timesteps
        return (6 if self.filter_mode == "regenerate" else 4) + 3 * T


PCW_ENSEMBLE = (0.9, 0.0, 0.0, 0.1)     # pwcnet_ddim.py:599
IGEV_ENSEMBLE = (0.6, 0.1, 0.3)         # igev_stereo_ddim.py:356


class PcwHotPath:
    """BASELINE.json configs[2]: the kernel sequence PWCNet_ddim.forward (eval) + ddim_sample + model_predictions
    (KITTI12/models/pwcnet_ddim.py:604-625, :530-602, :466-528) issue outside their 2-D / 3-D convolutions, for a batch of
    pairs: 4-scale group-wise correlation volumes (D = 48/24/12/6) + concat volumes (variant T), then T = 3 x
    {filter multiply on `combine` [B,32,48,h,w], softmax + regression over [B,192,H,W] (the probability volume is written on
    the LAST step only — the one ddim_sample returns; on the other steps it is only ever reduced to an uncertainty, which
    comes from a second read of the logits: `prob="every_step"` restores the reference's own materialisation), warp of the
    full-res right refinement features by the regressed disparity, the +-24 two-sided correlation volume, x_start /
    pred_noise, the uncertainty of the refined disparity against the pre-refinement distribution + renewal vote, fused DDIM
    step with cumulative re-noising}, ensemble.  The refinement network itself is a
    convolution stack (out of scope): its output is stood in for by the regressed disparity."""

    def __init__(self, schedule: Optional[DdimSchedule] = None, num_groups: int = 40, maxdisp: int = 192,
                 ensemble: Sequence[float] = PCW_ENSEMBLE, prob: str = "last_step"):
        assert prob in ("last_step", "every_step")
        self.prob_mode = prob
        self.sched = schedule or DdimSchedule(sampling_timesteps=3)
        self.G, self.maxdisp, self.D = num_groups, maxdisp, maxdisp // 4
        self.cof = tuple(ensemble)
        self._buf: Dict[str, torch.Tensor] = {}
        assert len(self.cof) == self.sched.sampling_timesteps + 1

    def _out(self, name, shape, device):
        t = self._buf.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.device != device:
            t = torch.empty(shape, dtype=torch.float32, device=device)
            self._buf[name] = t
        return t

    def __call__(self, scales, combine, costs, used, feat_l_full, feat_r_full, start, asd, shifts, step_noises, q_noises,
                 keep: bool = False, timer=None):
        """scales: [(gw_l, gw_r, cat_l, cat_r, D_s)] x 4; combine [B,32,48,h,w]; costs[i] [B,maxdisp,H,W] logits of step i
        (len T or 1); used [B,H,W]; feat_*_full [B,32,H,W]; start [B,48,h,w] fp32 (the randn start state, pwcnet_ddim.py:541);
        asd [B,48,h,w] x_start of the initial disparity; shifts[i] [B,48]; step_noises[i] = randn_like(img) (fp32 for i = 0,
        fp64 after), q_noises[i] = randn_like(asd) fp32.  Returns dict(pred, x_last, mask, ...)."""
        tm = timer if timer is not None else (lambda name: contextlib.nullcontext())
        sched, D = self.sched, self.D
        B, _, _, h, w = combine.shape
        dev = combine.device
        vols = []
        for gl, gr, cl, cr, Ds in scales:
            with tm("gwc_volume"):
                gv = ops.gwc_volume(gl, gr, Ds, self.G)
            with tm("concat_volume"):
                cv = ops.concat_volume(cl, cr, Ds, mask_left=True)
            if keep:
                vols.append((gv, cv))
        img = start
        disps = [used]
        mask = torch.zeros((B, h, w), dtype=torch.float32, device=dev)
        pairs = sched.time_pairs()
        corr = prob = None
        for i, (t, t_next) in enumerate(pairs):
            with tm("filter"):
                vol_f, n = ops.volume_filter(combine, img, shifts[i], sched.scale, return_n=True)
            cost = costs[i if len(costs) > 1 else 0]
            last = t_next < 0
            want_prob = last or self.prob_mode == "every_step"
            with tm("softmax_regress"):
                r = ops.softmax_regress(cost, return_prob=want_prob)
            disp, prob = r["disp"], (r["prob"] if want_prob else None)
            # the refinement network's input assembled as the product does (sampler.pcw_model_predictions): warp, left - warped,
            # copy of left and the +-24 volume straight into the concat buffer — no subtraction / torch.cat passes
            Cf = feat_l_full.shape[1]
            combine_in = self._out("refine_in", (B, 2 * Cf + 32 + 1 + 49, feat_l_full.shape[2], feat_l_full.shape[3]), dev)
            with tm("refine_input"):
                _, corr = ops.refine_input_assemble(feat_l_full, feat_r_full, disp.unsqueeze(1), 24, 1,
                                                    diff_out=combine_in[:, :Cf], copy_out=combine_in[:, Cf:2 * Cf],
                                                    corr_out=combine_in[:, 2 * Cf + 33:])
            # (dispupsample / refinenet3 run here in the full network — out of scope; disp stands in for disp_finetune)
            disps.append(disp)
            last = t_next < 0
            if last:
                with tm("ddim_step"):
                    st = ops.ddim_step(disp=disp, xt=img, shift=shifts[i], scale=sched.scale, sqrt_recip=sched.sqrt_recip(t),
                                       sqrt_recipm1=sched.sqrt_recipm1(t), last_step=True,
                                       disp_clamp_hi=float(self.maxdisp - 1), mask=mask)
                img = st["x_next"]
                continue
            with tm("uncertainty_vote"):
                vote = (ops.uncertainty_vote(disp, prob, used, 1.0, 1.0) if prob is not None
                        else ops.softmax_uncertainty_vote(disp, cost, used, 1.0, 1.0))
            san, c, sigma = sched.update_coefficients(t, t_next)
            with tm("ddim_step"):
                st = ops.ddim_step(disp=disp, xt=img, shift=shifts[i], scale=sched.scale, sqrt_recip=sched.sqrt_recip(t),
                                   sqrt_recipm1=sched.sqrt_recipm1(t), last_step=False,
                                   disp_clamp_hi=float(self.maxdisp - 1), vote=vote, mask=mask, sqrt_alpha_next=san, c=c,
                                   sigma=sigma, step_noise=step_noises[i], asd=asd, q_noise=q_noises[i],
                                   sqrt_ac=sched.sqrt_ac(t), sqrt_1m_ac=sched.sqrt_1m_ac(t), want_asd_out=True)
            img, asd = st["x_next"], st["asd_out"]
        with tm("ensemble"):
            pred = ops.ensemble(disps, self.cof[: len(disps)])
        out = {"pred": pred, "x_last": img, "mask": mask, "prob": prob}
        if keep:
            out.update(volumes=vols, corr=corr, disp_last=disps[-1])
        return out


class IgevHotPath:
    """BASELINE.json configs[3]: the kernel sequence IGEVStereo_ddim.forward (eval) + ddim_sample + model_predictions
    (KITTI15/core/igev_stereo_ddim.py:361-427, :294-359, :226-292) issue outside their convolutions / GRU: gwc volume
    (C = 96, G = 8, D = 48 at 1/4 res), softmax + regression of the initial D = 48 classifier, the all-pairs correlation +
    geometry pyramid (Combined_Geo_Encoding_Volume.__init__), then T = 2 x {filter factor, geometry filter (once per
    step), `iters` x pyramid lookup, convex upsampling of the last iteration, fused DDIM step with the renewal vote},
    ensemble.  The GRU that would move the disparity between lookups is out of scope: every lookup of a step samples at
    that step's initial disparity plus a fixed per-iteration offset."""

    def __init__(self, schedule: Optional[DdimSchedule] = None, iters: int = 32, num_groups: int = 8, D: int = 48,
                 ensemble: Sequence[float] = IGEV_ENSEMBLE):
        self.sched = schedule or DdimSchedule(sampling_timesteps=2)
        self.iters, self.G, self.D = iters, num_groups, D
        self.cof = tuple(ensemble)

    def __call__(self, fmap_l, fmap_r, geo, cost48, up_weights, coords, used, start, asd, shifts, step_noises, q_noises,
                 keep: bool = False, timer=None):
        from . import kitti15
        tm = timer if timer is not None else (lambda name: contextlib.nullcontext())
        sched, D = self.sched, self.D
        B, _, h, w = fmap_l.shape
        dev = fmap_l.device
        with tm("gwc_volume"):
            gwc = ops.gwc_volume(fmap_l, fmap_r, D, self.G)
        with tm("softmax_regress"):
            disp0 = ops.softmax_regress(cost48)["disp"].unsqueeze(1)            # init_disp (igev_stereo_ddim.py:384-385)
        with tm("geo_init"):
            fn = kitti15.Combined_Geo_Encoding_Volume(fmap_l, fmap_r, geo, num_levels=2, radius=4)
        c0 = coords.reshape(B, h, w).contiguous()
        used_map = used.reshape(B, used.shape[-2], used.shape[-1])
        img = start
        disps = [used_map]
        mask = torch.zeros((B, h, w), dtype=torch.float32, device=dev)
        look = None
        for i, (t, t_next) in enumerate(sched.time_pairs()):
            with tm("filter_factor"):
                n32 = ops.filter_factor(img, shifts[i], sched.scale)
            for it in range(self.iters):
                with tm("geo_filter" if it == 0 else "geo_lookup"):              # the first lookup of a step builds the filtered pyramid
                    look = fn(disp0 + 0.125 * it, coords, n32)
            with tm("context_upsample"):
                up = ops.context_upsample(disp0 * 4.0, up_weights)
            with tm("fallback"):
                disps.append(ops.select_close(up, used_map, 3.0))                         # igev_stereo_ddim.py:323-325
            last = t_next < 0
            kw = {}
            if not last:
                san, c, sigma = sched.update_coefficients(t, t_next)
                kw = dict(sqrt_alpha_next=san, c=c, sigma=sigma, step_noise=step_noises[i], asd=asd, q_noise=q_noises[i],
                          sqrt_ac=sched.sqrt_ac(t), sqrt_1m_ac=sched.sqrt_1m_ac(t))
            with tm("ddim_step"):
                st = ops.ddim_step(disp=up, xt=img, shift=shifts[i], scale=sched.scale, sqrt_recip=sched.sqrt_recip(t),
                                   sqrt_recipm1=sched.sqrt_recipm1(t), last_step=last, disp_clamp_hi=float(D - 1), coords0=c0,
                                   used=used_map, vote_thr_dif=5.0, mask=mask, **kw)
            img = st["x_next"]
        with tm("ensemble"):
            pred = ops.ensemble(disps, self.cof[: len(disps)])
        out = {"pred": pred, "x_last": img, "mask": mask, "lookup": look}
        if keep:
            out.update(gwc=gwc, geo=fn)
        return out
