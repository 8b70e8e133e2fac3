#!/usr/bin/env python
"""Hot-path time per stereo pair for BASELINE.json's configs[2] (PCWNet+DiffuVolume, KITTI12 384x1248) and configs[3]
(IGEV+DiffuVolume, KITTI15 384x1248): the kernel sequence each model's forward + ddim_sample issues OUTSIDE its 2-D/3-D
convolutions, on synthetic inputs, B pairs per launch.  (configs[1]/[4] are bench.py's headline.)

    python scripts/bench_configs.py [--batch 8] [--steps 10] [--out gpurun_out/configs.json]

PCWNet (KITTI12/models/pwcnet_ddim.py:604-758, :530-602): 4-scale gwc (C=320, G=40, D=48/24/12/6) + concat variant T
(C=12), then T=3 x {filter on [B,32,48,96,312], softmax+regression over [B,192,384,1248] with the uncertainty/vote,
warp of the C=32 full-res right features, +-24 two-sided correlation volume, fused DDIM step}.
IGEV (KITTI15/core/igev_stereo_ddim.py:361-427, :294-359): gwc (C=96, G=8, D=48 @96x312), softmax+regression (D=48),
all-pairs correlation + packed geo pyramid, then T=2 x {filter factor, geo filter (once per step), 32 x lookup,
context_upsample, fused DDIM step}.
Timing: CUDA events around `steps` back-to-back passes after 3 warm-up passes; bytes = algorithmic bytes of the sequence.
"""
from __future__ import annotations

import argparse
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

from diffuvolume_b200 import kitti15, ops  # noqa: E402
from diffuvolume_b200.pipeline import DdimSchedule  # noqa: E402


def timed(fn, steps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--out", default="gpurun_out/configs.json")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    B = args.batch
    g = torch.Generator(device=dev)
    g.manual_seed(11)
    rn = lambda *s, dt=torch.float32: torch.randn(*s, generator=g, device=dev, dtype=dt)
    ru = lambda *s, dt=torch.float32: torch.rand(*s, generator=g, device=dev, dtype=dt)
    peak = 6650.0
    pk = ROOT / "MEASURED_PEAKS.json"
    if pk.exists():
        peak = float(json.loads(pk.read_text())["hbm_gbs"])
    res = {"batch": B, "steps": args.steps, "peak_GBs": peak, "gpu": torch.cuda.get_device_name(0)}
    F4 = 4

    # ---------------------------------------------------------------- configs[2]: PCWNet 384x1248
    H, W, D = 384, 1248, 48
    h, w = H // 4, W // 4
    sched = DdimSchedule(sampling_timesteps=3)
    feats = [(rn(B, 320, H // s, W // s), rn(B, 320, H // s, W // s), rn(B, 12, H // s, W // s), rn(B, 12, H // s, W // s), D * 4 // s)
             for s in (4, 8, 16, 32)]
    vol = rn(B, 32, D, h, w)                       # `combine` after dres0/dres1 (convs, out of scope)
    cost = rn(B, 192, H, W) * 4.0
    used = ru(B, H, W) * 191.0
    fr_full = rn(B, 32, H, W)
    fl_full = rn(B, 32, H, W)
    shifts = [rn(B, D) * 0.1 for _ in range(3)]
    noises = [rn(B, D, h, w, dt=torch.float64) for _ in range(3)]
    noises32 = [n.float() for n in noises]     # randn_like(img): img is fp32 on the first step, fp64 afterwards
    asd = ops.xstart_from_disp(ru(B, h, w) * 47.0, D, 1.0)
    mask = torch.zeros(B, h, w, device=dev)
    pairs = sched.time_pairs()

    def pcw():
        for fl, fr, cl, cr, Ds in feats:
            ops.gwc_volume(fl, fr, Ds, 40)
            ops.concat_volume(cl, cr, Ds, mask_left=True)
        img = rn(B, D, h, w)
        for i, (t, tn) in enumerate(pairs):
            ops.volume_filter(vol, img, shifts[i], 1.0)
            r = ops.softmax_regress(cost, used=used, vote_thresholds=(1.0, 1.0))
            warped = ops.warp(fr_full, r["disp"].unsqueeze(1))
            ops.corr_volume_2sided(fl_full, warped, 24, 1)
            last = tn < 0
            kw = {}
            if not last:
                san, c, sigma = sched.update_coefficients(t, tn)
                kw = dict(sqrt_alpha_next=san, c=c, sigma=sigma,
                          step_noise=noises[i] if img.dtype == torch.float64 else noises32[i], asd=asd, q_noise=noises[i],
                          sqrt_ac=sched.sqrt_ac(t), sqrt_1m_ac=sched.sqrt_1m_ac(t))
            st = ops.ddim_step(disp=r["disp"], xt=img, shift=shifts[i], scale=1.0, sqrt_recip=sched.sqrt_recip(t),
                               sqrt_recipm1=sched.sqrt_recipm1(t), last_step=last, disp_clamp_hi=191.0, vote=r["vote"],
                               mask=mask, **kw)
            img = st["x_next"]

    ms = timed(pcw, args.steps)
    vol_bytes = sum((2 * 320 + 40 * Ds + 24 + 24 * Ds) * (H // s) * (W // s) for s, Ds in ((4, 48), (8, 24), (16, 12), (32, 6))) * F4
    step_bytes = (2 * 32 * D * h * w + 193 * H * W + 2 * H * W + (2 * 32 + 1) * H * W + (2 * 32 + 49) * H * W) * F4 + 6 * D * h * w * 8
    alg = B * (vol_bytes + 3 * step_bytes)
    res["pcwnet_kitti12_384x1248"] = {"ms_per_batch": round(ms, 4), "pairs_per_s": round(B / (ms / 1e3), 1),
                                      "algorithmic_GB": round(alg / 1e9, 3), "GBs": round(alg / 1e9 / (ms / 1e3), 1),
                                      "frac_of_peak": round(alg / 1e9 / (ms / 1e3) / peak, 4),
                                      "sequence": "4-scale gwc + concat(T); T=3 x {filter, softmax-regress+vote, warp, +-24 corr, ddim_step}"}
    print(json.dumps(res["pcwnet_kitti12_384x1248"]), flush=True)
    del feats, vol, cost, fr_full, fl_full

    # ---------------------------------------------------------------- configs[3]: IGEV 384x1248 (1/4 = 96x312)
    sched2 = DdimSchedule(sampling_timesteps=2)
    f1, f2 = rn(B, 96, h, w), rn(B, 96, h, w)
    geo = rn(B, 8, D, h, w)                        # after the 3-D hourglass (conv, out of scope)
    cost48 = rn(B, D, h, w)
    upw = torch.softmax(rn(B, 9, H, W), 1)
    coords = torch.arange(w, device=dev, dtype=torch.float32).view(1, 1, 1, w).expand(B, 1, h, w).contiguous()
    used2 = ru(B, H, W) * 47.0
    shifts2 = [rn(B, D) * 0.1 for _ in range(2)]
    noise2 = rn(B, D, h, w, dt=torch.float64)
    noise2_32 = noise2.float()
    mask2 = torch.zeros(B, h, w, device=dev)
    c0 = coords.reshape(B, h, w).contiguous()
    pairs2 = sched2.time_pairs()
    iters = 32

    def igev():
        ops.gwc_volume(f1, f2, D, 8)
        r0 = ops.softmax_regress(cost48)
        fn = kitti15.Combined_Geo_Encoding_Volume(f1, f2, geo, num_levels=2, radius=4)
        disp = r0["disp"].unsqueeze(1)
        img = rn(B, D, h, w)
        for i, (t, tn) in enumerate(pairs2):
            n32 = ops.filter_factor(img, shifts2[i], 1.0)
            for _ in range(iters):
                fn(disp, coords, n32)              # (GRU update of disp happens here in the full network)
            up = ops.context_upsample(disp * 4.0, upw)
            last = tn < 0
            kw = {}
            if not last:
                san, c, sigma = sched2.update_coefficients(t, tn)
                kw = dict(sqrt_alpha_next=san, c=c, sigma=sigma, step_noise=noise2 if img.dtype == torch.float64 else noise2_32,
                          asd=asd2, q_noise=noise2,
                          sqrt_ac=sched2.sqrt_ac(t), sqrt_1m_ac=sched2.sqrt_1m_ac(t))
            st = ops.ddim_step(disp=up, xt=img, shift=shifts2[i], scale=1.0, sqrt_recip=sched2.sqrt_recip(t),
                               sqrt_recipm1=sched2.sqrt_recipm1(t), last_step=last, disp_clamp_hi=float(D - 1), coords0=c0,
                               used=used2, vote_thr_dif=5.0, mask=mask2, **kw)
            img = st["x_next"]

    asd2 = ops.xstart_from_disp(ru(B, h, w) * 47.0, D, 1.0)
    ms = timed(igev, args.steps)
    hw = h * w
    init_bytes = ((2 * 96 + 8 * D) * hw + 49 * hw + (2 * 96 * hw + 1.5 * hw * w) + 2.5 * 8 * D * hw) * F4
    step_bytes = (3 * 1.5 * 8 * D * hw + iters * ((2 * (10 * 8 + 10) + 162 + 2) * hw) + 10 * H * W) * F4 + 6 * D * hw * 8
    alg = B * (init_bytes + 2 * step_bytes)
    res["igev_kitti15_384x1248"] = {"ms_per_batch": round(ms, 4), "pairs_per_s": round(B / (ms / 1e3), 1),
                                    "algorithmic_GB": round(alg / 1e9, 3), "GBs": round(alg / 1e9 / (ms / 1e3), 1),
                                    "frac_of_peak": round(alg / 1e9 / (ms / 1e3) / peak, 4),
                                    "sequence": "gwc + regress(D=48) + all-pairs corr + geo pack; T=2 x {filter factor, geo filter, "
                                                "32 x lookup, context_upsample, ddim_step}"}
    print(json.dumps(res["igev_kitti15_384x1248"]), flush=True)
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
