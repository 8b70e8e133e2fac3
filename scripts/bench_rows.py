#!/usr/bin/env python
"""Per-row roofline table: every C-ABI kernel of SURVEY.md §8a/§8f timed alone at the size BASELINE.json's configs
give it (configs[1] ACV 540x960, configs[2] PCWNet 384x1248, configs[3] IGEV 384x1248), B pairs per launch.

    python scripts/bench_rows.py [--batch 8] [--iters 20] [--out gpurun_out/rows.json]

Timing: CUDA events on torch's current stream (the stream the C-ABI launches on), 3 warm-up launches, `iters` timed
launches; between launches an L2 flush (a 256 MB fill) unless the operands exceed the 126 MB L2 anyway.  Bytes are the
ALGORITHMIC bytes of the op (compulsory unique reads + writes at the reference's op boundary), so `frac` is comparable
with bench.py's roofline object.  Peak: MEASURED_PEAKS.json hbm_gbs, else the 6 650 GB/s fallback.
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

from diffuvolume_b200 import ops  # noqa: E402


def peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class Rows:
    def __init__(self, iters: int, dev):
        self.iters, self.dev = iters, dev
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        self.rows = []
        self.peak, self.peak_src = peak_gbs()

    def time(self, name, config, fn, nbytes, flops=0.0, note=""):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(self.iters):
            self.flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        ms = ts[len(ts) // 2]
        gbs = nbytes / 1e9 / (ms / 1e3)
        row = {"row": name, "config": config, "ms": round(ms, 4), "ms_min": round(ts[0], 4),
               "algorithmic_MB": round(nbytes / 1e6, 2), "GBs": round(gbs, 1), "frac": round(gbs / self.peak, 4)}
        if flops:
            row["TFLOPs"] = round(flops / 1e12 / (ms / 1e3), 3)
        if note:
            row["note"] = note
        self.rows.append(row)
        print(json.dumps(row), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default="gpurun_out/rows.json")
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    B = args.batch
    R = Rows(args.iters, dev)
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    rn = lambda *s, dt=torch.float32: torch.randn(*s, generator=g, device=dev, dtype=dt)
    ru = lambda *s, dt=torch.float32: torch.rand(*s, generator=g, device=dev, dtype=dt)
    want = lambda k: (not args.only) or any(s in k for s in args.only.split(","))
    F4 = 4

    # ---------------- configs[1]: ACV 540x960 -------------------------------------------------------------------
    h, w, D, C, G, Cc = 135, 240, 48, 320, 40, 32
    hw = h * w
    if want("acv"):
        fl, fr = rn(B, C, h, w), rn(B, C, h, w)
        R.time("a2 gwc_volume", f"acv B={B} C=320 G=40 D=48 135x240", lambda: ops.gwc_volume(fl, fr, D, G),
               B * (2 * C * hw + G * D * hw) * F4, flops=2.0 * B * C * D * hw)
        R.time("a2 gwc_volume -> bf16 volume", f"acv B={B} C=320 G=40 D=48 135x240",
               lambda: ops.gwc_volume(fl, fr, D, G, out_dtype=torch.bfloat16), B * (2 * C * hw * F4 + G * D * hw * 2),
               note="bytes = fp32 features in + bf16 volume out")
        R.time("a1 groupwise_correlation", f"acv B={B} C=320 G=40 135x240", lambda: ops.groupwise_correlation(fl, fr, G),
               B * (2 * C * hw + G * hw) * F4)
        gv = rn(B, G, D, h, w)
        R.time("f1 gwc_volume_bwd", f"acv B={B}", lambda: ops.gwc_volume_bwd(gv, fl, fr, G),
               B * (G * D * hw + 4 * C * hw) * F4, flops=4.0 * B * C * D * hw)
        del gv
        cl, cr = rn(B, Cc, h, w), rn(B, Cc, h, w)
        R.time("a3 concat_volume M", f"acv B={B} C=32 D=48", lambda: ops.concat_volume(cl, cr, D, mask_left=False),
               B * (2 * Cc * hw + 2 * Cc * D * hw) * F4)
        att = rn(B, 1, D, h, w)
        R.time("a3+a4 concat+ACV", f"acv B={B}", lambda: ops.concat_volume(cl, cr, D, mask_left=False, att_logits=att),
               B * (2 * Cc * hw + D * hw + 2 * Cc * D * hw) * F4)
        attw, nf = ops.att_softmax(att), ru(B, D, h, w)
        R.time("a9 filter, regenerate mode (features + 2 factor maps -> volume)", f"acv B={B}",
               lambda: ops.concat_volume_weighted(cl, cr, D, mask_left=False, att_weights=attw, n=nf),
               B * (2 * Cc * hw + 2 * D * hw + 2 * Cc * D * hw) * F4)
        R.time("a9 filter, regenerate mode -> bf16 volume", f"acv B={B}",
               lambda: ops.concat_volume_weighted(cl, cr, D, mask_left=False, att_weights=attw, n=nf, out_dtype=torch.bfloat16),
               B * ((2 * Cc * hw + 2 * D * hw) * F4 + 2 * Cc * D * hw * 2), note="bytes = fp32 inputs + bf16 volume out")
        del attw, nf
        gvol = rn(B, 2 * Cc, D, h, w)
        R.time("f1 concat_volume_bwd", f"acv B={B}", lambda: ops.concat_volume_bwd(gvol, False),
               B * (2 * Cc * D * hw + 2 * Cc * hw) * F4)
        xt = rn(B, D, h, w, dt=torch.float64)
        shift = rn(B, D) * 0.1
        R.time("a9 volume_filter", f"acv B={B} [B,64,48,135,240]", lambda: ops.volume_filter(gvol, xt, shift),
               B * (2 * 2 * Cc * D * hw * F4 + D * hw * 8))
        del gvol, xt
        cost = rn(B, 192, 540, 960) * 4.0
        used = ru(B, 540, 960) * 191.0
        R.time("a6 softmax_regress (disp only)", f"acv B={B} [B,192,540,960]", lambda: ops.softmax_regress(cost),
               B * (192 + 1) * 540 * 960 * F4)
        R.time("a6 disparity_regression", f"acv B={B} [B,192,540,960]", lambda: ops.disparity_regression(cost, 192),
               B * (192 + 1) * 540 * 960 * F4)
        gd = rn(B, 540, 960)
        R.time("f1 disparity_regression_bwd", f"acv B={B}", lambda: ops.disparity_regression_bwd(gd, 192),
               B * (192 + 1) * 540 * 960 * F4)
        del cost, used, gd
        gwcv = rn(B, G, D, h, w)
        wp, wl = rn(40, 9), rn(40, 9)
        R.time("f4 acv_patch_volume (2 chained depth-wise 3x3 + cat)", f"acv B={B} [B,40,48,135,240]",
               lambda: ops.acv_patch_volume(gwcv, wp, wl[:8], wl[8:24], wl[24:]), 2 * B * G * D * hw * F4)
        del gwcv
        # the [B,48,h,w] sampler-state ops (a7, a8, a10, a11, a12, a13) as stand-alone entry points
        x0 = ru(B, D, h, w) * 2.0 - 1.0
        nz64, xt64 = rn(B, D, h, w, dt=torch.float64), rn(B, D, h, w, dt=torch.float64)
        st_b = D * hw
        R.time("a7 q_sample (fp32 x_start, fp32 noise -> fp64)", f"acv B={B} [B,48,135,240]",
               lambda: ops.q_sample(x0, x0, 0.7, 0.7), B * st_b * (4 + 4 + 8))
        R.time("a8 predict_noise_from_start (fp64)", f"acv B={B}", lambda: ops.predict_noise_from_start(xt64, x0, 1.5, 1.1),
               B * st_b * (8 + 4 + 8))
        dq = ru(B, h, w) * 47.75
        R.time("a10 xstart_from_disp", f"acv B={B}", lambda: ops.xstart_from_disp(dq, D, 1.0), B * (hw + st_b) * F4)
        dfull = ru(B, 540, 960) * 191.0
        R.time("a10 downsample_bilinear (x1/4, clamp, /4)", f"acv B={B} 540x960", lambda: ops.downsample_bilinear(dfull, (h, w), clamp=(0, 191), post_scale=0.25),
               B * (540 * 960 // 4 + hw) * F4, note="reads the 2x2 centre of every 4x4 block")
        R.time("a9 filter_factor (fp64 x_t -> fp32 n)", f"acv B={B}", lambda: ops.filter_factor(xt64, shift, 1.0), B * st_b * (8 + 4))
        R.time("a4 att_softmax", f"acv B={B}", lambda: ops.att_softmax(att), 2 * B * st_b * F4)
        maps = [ru(B, 540, 960) for _ in range(6)]
        R.time("a13 ensemble (6 maps)", f"acv B={B} 540x960", lambda: ops.ensemble(maps, [0.5, 0.0, 0.0, 0.0, 0.2, 0.3]),
               7 * B * 540 * 960 * F4)
        vote, mask = ru(B, 540, 960).round(), torch.zeros(B, h, w, device=dev)
        R.time("a8+a10+a11+a12 ddim_step (fp64 state, re-noise)", f"acv B={B}",
               lambda: ops.ddim_step(disp=dfull, xt=xt64, shift=shift, scale=1.0, sqrt_recip=1.5, sqrt_recipm1=1.1, last_step=False,
                                     disp_clamp_hi=191.0, vote=vote, mask=mask, sqrt_alpha_next=0.8, c=0.3, sigma=0.5,
                                     step_noise=nz64, renoise=nz64, shift_next=shift, want_n_next=True),
               B * (2 * 540 * 960 * F4 // 4 * 2 + st_b * (8 + 8 + 8 + 4 + 8 + 4)))
        del maps, dfull, vote, x0, nz64, xt64
        cq = rn(B, 1, D, h, w) * 4.0
        R.time("f2 upsample_softmax_regress", f"acv B={B} [B,1,48,135,240]->540x960",
               lambda: ops.upsample_softmax_regress(cq, (192, 540, 960)), B * (D * hw + 540 * 960) * F4,
               note="MUFU-bound, bytes are tiny by construction")

    # ---------------- configs[2]: PCWNet 384x1248 ---------------------------------------------------------------
    if want("pcw"):
        for s, Ds in ((4, 48), (8, 24), (16, 12), (32, 6)):
            hs, ws = 384 // s, 1248 // s
            fl, fr = rn(B, 320, hs, ws), rn(B, 320, hs, ws)
            R.time(f"a2 gwc_volume 1/{s}", f"pcw B={B} C=320 G=40 D={Ds} {hs}x{ws}", lambda: ops.gwc_volume(fl, fr, Ds, 40),
                   B * (2 * 320 * hs * ws + 40 * Ds * hs * ws) * F4)
            cl, cr = rn(B, 12, hs, ws), rn(B, 12, hs, ws)
            R.time(f"a3 concat_volume T 1/{s}", f"pcw B={B} C=12 D={Ds} {hs}x{ws}",
                   lambda: ops.concat_volume(cl, cr, Ds, mask_left=True), B * (24 * hs * ws + 24 * Ds * hs * ws) * F4)
        vol = rn(B, 32, 48, 96, 312)
        xt = rn(B, 48, 96, 312, dt=torch.float64)
        shift = rn(B, 48) * 0.1
        R.time("a9 volume_filter", f"pcw B={B} [B,32,48,96,312]", lambda: ops.volume_filter(vol, xt, shift),
               B * (2 * 32 * 48 * 96 * 312 * F4 + 48 * 96 * 312 * 8))
        del vol, xt
        Bc = min(B, 4)
        fl, fr = rn(Bc, 32, 384, 1248), rn(Bc, 32, 384, 1248)
        dsp = ru(Bc, 1, 384, 1248) * 100.0
        R.time("f3 warp (random disparities: worst-case gather)", f"pcw B={Bc} C=32 384x1248", lambda: ops.warp(fr, dsp),
               Bc * (2 * 32 + 1) * 384 * 1248 * F4)
        dsm = (torch.linspace(2, 90, 1248, device=dev).view(1, 1, 1, -1).expand(Bc, 1, 384, 1248) + ru(Bc, 1, 384, 1248)).contiguous()
        R.time("f3 warp (smooth disparities)", f"pcw B={Bc} C=32 384x1248", lambda: ops.warp(fr, dsm),
               Bc * (2 * 32 + 1) * 384 * 1248 * F4)
        R.time("a5 corr_volume_2sided", f"pcw B={Bc} C=32 m=24 384x1248", lambda: ops.corr_volume_2sided(fl, fr, 24, 1),
               Bc * (2 * 32 + 49) * 384 * 1248 * F4, flops=2.0 * Bc * 32 * 49 * 384 * 1248)
        R.time("a2 gwc_volume C=32 G=1 D=25 (positive half of a5)", f"pcw B={Bc} 384x1248", lambda: ops.gwc_volume(fl, fr, 25, 1),
               Bc * (2 * 32 + 25) * 384 * 1248 * F4)
        R.time("f3 refine_input_assemble (warp + left-warped + copy + +-24 volume into the concat buffer)", f"pcw B={Bc} 384x1248",
               (lambda buf=torch.empty(Bc, 2 * 32 + 5 + 49, 384, 1248, device=dev): ops.refine_input_assemble(
                   fl, fr, dsm, 24, 1, corr_out=buf[:, 69:], diff_out=buf[:, :32], copy_out=buf[:, 32:64])),
               Bc * (2 * 32 + 1 + 2 * 32 + 32 + 32 + 49) * 384 * 1248 * F4,
               note="bytes: ref, src, disp read; warped written + read back by the volume; diff, copy, volume written")
        gw = rn(Bc, 32, 384, 1248)
        R.time("f1 warp_bwd (grad_x scatter-add + grad_disp, smooth disparities)", f"pcw B={Bc} C=32 384x1248",
               lambda: ops.warp_bwd(gw, fr, dsm), Bc * (3 * 32 + 2 + 32) * 384 * 1248 * F4,
               note="bytes: grad_out + features read, grad_x cleared then accumulated (2 x), disp read, grad_disp written")
        R.time("f1 warp_bwd (grad_x only)", f"pcw B={Bc} C=32 384x1248", lambda: ops.warp_bwd(gw, fr, dsm, True, False),
               Bc * (3 * 32 + 1) * 384 * 1248 * F4)
        del gw
        gv = rn(Bc, 1, 49, 384, 1248)
        R.time("f1 corr_volume_2sided_bwd", f"pcw B={Bc}", lambda: ops.gwc_volume_bwd(gv, fl, fr, 1, two_sided_maxdisp=24),
               Bc * (49 + 4 * 32) * 384 * 1248 * F4)
        del gv, fl, fr
        cost = rn(B, 192, 384, 1248) * 4.0
        R.time("a6 softmax_regress (disp only)", f"pcw B={B} [B,192,384,1248]", lambda: ops.softmax_regress(cost),
               B * 193 * 384 * 1248 * F4)
        usedp = ru(B, 384, 1248) * 191.0
        R.time("a6 softmax_regress (+ prob volume out, the tier-2 model_predictions boundary)", f"pcw B={B}",
               lambda: ops.softmax_regress(cost, return_prob=True), B * (2 * 192 + 1) * 384 * 1248 * F4)
        rr = ops.softmax_regress(cost, return_prob=True)
        R.time("a11 uncertainty_vote (refined disparity vs prob volume)", f"pcw B={B}",
               lambda: ops.uncertainty_vote(rr["disp"], rr["prob"], usedp, 1.0, 1.0), B * (192 + 3) * 384 * 1248 * F4)
        del cost, rr, usedp

    # ---------------- configs[3]: IGEV 384x1248 (1/4 = 96x312) --------------------------------------------------
    if want("igev"):
        h, w, D, Cg = 96, 312, 48, 8
        hw = h * w
        fl, fr = rn(B, 96, h, w), rn(B, 96, h, w)
        R.time("a2 gwc_volume", f"igev B={B} C=96 G=8 D=48 96x312", lambda: ops.gwc_volume(fl, fr, D, 8),
               B * (2 * 96 * hw + 8 * D * hw) * F4)
        R.time("a14 corr1d_allpairs", f"igev B={B} C=96 96x312", lambda: ops.corr1d_allpairs(fl, fr),
               B * (2 * 96 * hw + h * w * w) * F4, flops=2.0 * B * 96 * h * w * w)
        geo = rn(B, Cg, D, h, w)
        R.time("a14 geo_permute", f"igev B={B} [B,8,48,96,312]", lambda: ops.geo_permute(geo), 2 * B * Cg * D * hw * F4)
        rows = ops.geo_permute(geo)
        R.time("a14 avgpool_w2 (geo)", f"igev B={B}", lambda: ops.avgpool_w2(rows), int(1.5 * B * Cg * D * hw * F4))
        corr = ops.corr1d_allpairs(fl, fr).reshape(B * hw, 1, 1, w)
        R.time("a14 avgpool_w2 (corr)", f"igev B={B}", lambda: ops.avgpool_w2(corr), int(1.5 * B * hw * w * F4))
        gp = [rows, ops.avgpool_w2(rows)]
        cp = [corr, ops.avgpool_w2(corr)]
        disp = ru(B, 1, h, w) * 47.0
        coords = torch.arange(w, device=dev, dtype=torch.float32).view(1, 1, 1, w).expand(B, 1, h, w).contiguous()
        noisy = ru(B, D, h, w)
        # gathered reads: per pixel and level 10 hypotheses x 8 channels + 10 corr columns + noise; writes 162 channels
        lookup_bytes = B * hw * (2 * (10 * Cg + 10) + 162 + 2) * F4
        R.time("a15 geo_lookup (+noise)", f"igev B={B} r=4 L=2 -> [B,162,96,312]",
               lambda: ops.geo_lookup(gp, cp, disp, coords, noisy, 4), lookup_bytes + B * hw * 30 * F4,
               note="random disparities: worst-case gather")
        R.time("a15 geo_lookup (origin)", f"igev B={B} r=4 L=2", lambda: ops.geo_lookup(gp, cp, disp, coords, None, 4),
               lookup_bytes)
        smooth = (torch.linspace(2, 40, w, device=dev).view(1, 1, 1, w).expand(B, 1, h, w)).contiguous()
        R.time("a15 geo_lookup (+noise, smooth disp)", f"igev B={B}", lambda: ops.geo_lookup(gp, cp, smooth, coords, noisy, 4),
               lookup_bytes + B * hw * 30 * F4)
        R.time("a14 geo_pack (permute+pool fused, 2 levels)", f"igev B={B} [B,8,48,96,312]", lambda: ops.geo_pack(geo, 2),
               int(2.5 * B * Cg * D * hw * F4))
        pk = ops.geo_pack(geo, 2)
        R.time("a15 geo_lookup_packed (+noise)", f"igev B={B} r=4 L=2 -> [B,162,96,312]",
               lambda: ops.geo_lookup_packed(pk, cp, disp, coords, noisy, 4), lookup_bytes + B * hw * 30 * F4,
               note="random disparities: worst-case gather")
        R.time("a9 geo_filter_packed (once per DDIM step)", f"igev B={B}", lambda: ops.geo_filter_packed(pk, noisy),
               int(B * hw * (2 * 1.5 * Cg * D + D) * F4))
        R.time("a15 geo_lookup_packed (origin)", f"igev B={B} r=4 L=2", lambda: ops.geo_lookup_packed(pk, cp, disp, coords, None, 4),
               lookup_bytes)
        R.time("a15 geo_lookup_packed (+noise, smooth disp)", f"igev B={B}",
               lambda: ops.geo_lookup_packed(pk, cp, smooth, coords, noisy, 4), lookup_bytes + B * hw * 30 * F4)
        low, upw = ru(B, 1, h, w) * 190.0, ru(B, 9, 4 * h, 4 * w)
        R.time("f4 context_upsample", f"igev B={B} -> [B,384,1248]", lambda: ops.context_upsample(low, upw),
               B * (hw + 10 * 16 * hw) * F4)
        del upw
        cost = rn(B, 48, h, w)
        R.time("a6 softmax_regress D=48", f"igev B={B} [B,48,96,312]", lambda: ops.softmax_regress(cost),
               B * 49 * hw * F4)

    out = {"peak_GBs": R.peak, "peak_source": R.peak_src, "batch": B, "iters": args.iters,
           "gpu": torch.cuda.get_device_name(0), "rows": R.rows}
    Path(args.out).parent.mkdir(parents=True, exist_ok=True)
    Path(args.out).write_text(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
