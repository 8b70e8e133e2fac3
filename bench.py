#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on B200: "volume+DDIM-filter pairs/s @540x960 D=192".

A step = one pass of the ACVNet+DiffuVolume hot path (diffuvolume_b200.pipeline.AcvHotPath) over a
batch of B synthetic stereo pairs per GPU: 1x gwc volume (C=320, G=40, D=48 at 135x240), 1x concat
volume + ACV softmax weights, T=5 x {DDIM filter, softmax + regression + uncertainty + vote +
ensemble over [B,192,540,960], fused DDIM state update}.  The 2-D/3-D convolutions are out of scope;
their outputs are the synthetic inputs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch 8]
                    [--filter regenerate|volume]

N > 1: launched by torchrun, one rank per GPU, batch-sharded (weak scaling: B pairs per rank, no
data-path collective; one all_reduce of the timing / checksum at the end).  Rank 0 prints ONE JSON
line.  `--impl reference` times the reference's op sequence on the host CPU (oracle/torch_port.py,
all ATen threads) — the only place where the oracle is the thing measured.
"""
from __future__ import annotations

import argparse
import contextlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

METRIC = "volume+DDIM-filter pairs/s @540x960 D=192"
H, W, MAXDISP, C_GWC, G, C_CAT, T_STEPS = 540, 960, 192, 320, 40, 32, 5


# The contract is ONE JSON line on stdout.  Libraries print there too (NCCL announces its version on communicator creation),
# so fd 1 is pointed at stderr for the whole run and the result line is written to the saved original stdout.
_REAL_STDOUT = None


def protect_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(obj):
    line = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(line.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, line)


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


class KernelTimer:
    """CUDA-event pairs around each kernel category, recorded on torch's current stream (the stream the
    C-ABI launches on), resolved after the timed region."""

    def __init__(self):
        self.pairs = {}

    @contextlib.contextmanager
    def __call__(self, name):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        yield
        b.record()
        self.pairs.setdefault(name, []).append((a, b))

    def resolve(self):
        return {k: [a.elapsed_time(b) for a, b in v] for k, v in self.pairs.items()}


# ------------------------------------------------------------------------------------------------
def make_inputs(B: int, device, seed: int, regress: str = "logits"):
    """Synthetic inputs of SURVEY.md §8d, generated on the device with a seeded torch generator."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    h, w, D = H // 4, W // 4, MAXDISP // 4
    rn = lambda *s, dt=torch.float32: torch.randn(*s, generator=g, device=device, dtype=dt)
    ru = lambda *s, dt=torch.float32: torch.rand(*s, generator=g, device=device, dtype=dt)
    inp = dict(
        feat_l=rn(B, C_GWC, h, w), feat_r=rn(B, C_GWC, h, w),
        cfeat_l=rn(B, C_CAT, h, w), cfeat_r=rn(B, C_CAT, h, w),
        att_logits=rn(B, 1, D, h, w),
        # logits: one 3.2 GB [B,192,H,W] buffer, re-read by every step (>> L2); fused: the quarter-res conv output
        costs=[rn(B, MAXDISP, H, W) * 4.0] if regress == "logits" else [rn(B, 1, D, h, w) * 4.0],
        used=ru(B, H, W) * 191.0,
        disp_q=ru(B, h, w) * 47.75,
        shifts=[rn(B, D) * 0.1 for _ in range(T_STEPS)],
        step_noises=[rn(B, D, h, w, dt=torch.float32 if i == 0 else torch.float64) for i in range(T_STEPS - 1)],
        renoises=[ru(B, D, h, w, dt=torch.float64) for _ in range(T_STEPS - 1)],
    )
    return inp


def algorithmic_bytes(B: int, filter_mode: str, regress: str = "logits"):
    """Compulsory unique reads + writes per launch, fp32, for a batch of B pairs."""
    hw, HW, D = (H // 4) * (W // 4), H * W, MAXDISP // 4
    vol = 2 * C_CAT * D * hw * 4
    state64 = D * hw * 8
    fmap = D * hw * 4                       # one fp32 [D,h,w] factor map
    per_pair = {
        "gwc_volume": 2 * C_GWC * hw * 4 + G * D * hw * 4,
        "concat_acv": 2 * C_CAT * hw * 4 + D * hw * 4 + vol,            # op boundary: features + att logits in, volume out
        # regenerate: features + the two fp32 factor maps (softmax(att), n) in, volume out; volume: ac_volume + x_t in
        "filter": (2 * C_CAT * hw * 4 + 2 * fmap + vol) if filter_mode == "regenerate" else (2 * vol + state64),
        "filter_factor": 2 * fmap,                                      # first step only: x_start (fp32) -> n
        # cost read (full-res logits, or the quarter-res conv output when the upsample is fused); used read; disp, vote
        # written; ens read+write
        "softmax_regress": (MAXDISP * HW * 4 if regress == "logits" else D * hw * 4) + 5 * HW * 4,
        # disp+vote taps; xt, noise, renoise read; x0, x_next (+ next step's n in regenerate mode) written
        "ddim_step": 4 * HW * 2 + D * hw * (8 + 8 + 8 + 4 + 8) + (fmap if filter_mode == "regenerate" else 0),
    }
    return {k: v * B for k, v in per_pair.items()}


def workload_config(args, B, world):
    """The `config` object of the JSON line — the same for both arms (the reference arm times a bounded sample of it)."""
    return {"workload": f"configs[1]/[4]: gwc G=40 D=48 @{H // 4}x{W // 4} + concat/ACV + T=5 DDIM filter + "
                        + (f"softmax/regression over [B,192,{H},{W}]" if args.regress == "logits" else
                           f"trilinear x4 upsample fused into softmax/regression (input [B,1,48,{H // 4},{W // 4}])"),
            "pairs_per_gpu": B, "global_batch": B * world, "filter_mode": args.filter, "regress": args.regress,
            "l2": "inputs larger than L2 (3.2 GB logits, 3.2 GB volumes per step)",
            "parallelism": f"batch-sharded x{world}"}


def run_ours(args):
    from diffuvolume_b200 import _lib
    from diffuvolume_b200.pipeline import AcvHotPath

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("--gpus N>1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    B = args.batch
    regress = "logits" if args.regress == "logits" else "fused_upsample"
    inp = make_inputs(B, dev, seed=1234 + rank, regress=args.regress)   # rank-offset seeds: every rank owns different pairs
    path = AcvHotPath(filter_mode=args.filter, regress_mode=regress)
    timer = KernelTimer()

    def step(t=None):
        return path(**inp, timer=t)

    for _ in range(args.warmup):
        out = step()
    barrier()
    launches0 = _lib.launch_count()
    gpu_index = int(os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")[local]) if os.environ.get("CUDA_VISIBLE_DEVICES") else local
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(gpu_index) as clk:
        barrier()
        ev0.record()
        for _ in range(args.steps):
            out = step(timer)
        ev1.record()
        barrier()
    launches = _lib.launch_count() - launches0
    ms_total = ev0.elapsed_time(ev1)
    checksum = float(out["pred"].double().sum())
    if dist is not None:
        tt = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
        # the one collective of the path: the end-of-sweep metric reduction
        cs = torch.tensor([checksum, float(B)], device=dev, dtype=torch.float64)
        dist.all_reduce(cs, op=dist.ReduceOp.SUM)
        checksum = float(cs[0].item())
    ms_per_step = ms_total / args.steps
    value = (B * world) / (ms_per_step / 1e3)
    # ---- the one collective of the path: the end-of-sweep metric vector — per-batch EPE / D1 / Thres sums
    # of this rank's shard against a synthetic ground truth — reduced with ONE all_reduce(SUM) over NCCL
    from diffuvolume_b200.distributed import MetricSums
    msums = MetricSums()
    gt = inp["used"] + 0.75
    msums.update(out["pred"], gt, (gt < MAXDISP) & (gt > 0))
    metrics = msums.reduce(device=dev)
    metrics["note"] = ("EPE/D1/Thres of the ensemble prediction vs a synthetic ground truth, MetricSums.reduce(): one "
                       + ("NCCL all_reduce(SUM) of 8 float64" if world > 1 else "local sum (world size 1)"))
    # ---- sustained leg: the same step back to back for >= 2 s with its own clock record (the headline's timed region
    # is a burst of `steps` x ~7 ms)
    sustained = None
    if not args.no_sustained:
        n_sus = max(args.steps, int(args.sustained_seconds / (ms_per_step / 1e3)) + 1)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with ClockSampler(gpu_index) as clk2:
            barrier()
            s0.record()
            for _ in range(n_sus):
                step()
            s1.record()
            barrier()
        sms = s0.elapsed_time(s1)
        if dist is not None:
            tt = torch.tensor([sms], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            sms = float(tt.item())
        sustained = {"value": round(B * world * n_sus / (sms / 1e3), 2), "unit": "pairs/s", "steps": n_sus,
                     "seconds": round(sms / 1e3, 3), "ms_per_step": round(sms / n_sus, 4), "clocks": clk2.summary()}

    result = None
    if rank == 0:
        kt = timer.resolve()
        ab = algorithmic_bytes(B, args.filter, args.regress)
        peak, peak_src = measured_peak_gbs()
        kernels = {}
        for name, times in kt.items():
            avg = sum(times) / len(times)
            kernels[name] = {"launches_per_step": len(times) // args.steps, "avg_ms": round(avg, 4),
                             "ms_per_step": round(sum(times) / args.steps, 4),
                             "algorithmic_GB": round(ab[name] / 1e9, 4), "achieved_GBs": round(ab[name] / 1e9 / (avg / 1e3), 1),
                             "frac_of_peak": round(ab[name] / 1e9 / (avg / 1e3) / peak, 4)}
        dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
        step_bytes = sum(ab[k] * kernels[k]["launches_per_step"] for k in kernels)
        traffic = None
        tp = ROOT / "profiles" / "traffic.json"
        if tp.exists():
            try:
                kname = {"filter": "concat_stream_kernel", "concat_acv": "concat_stream_kernel",
                         "softmax_regress": "softmax_regress_tma_kernel", "gwc_volume": "gwc_volume_kernel",
                         "ddim_step": "ddim_step_kernel", "filter_factor": "filter_factor_kernel"}.get(dom, dom)
                traffic = json.loads(tp.read_text()).get(kname, {}).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        result = {
            "metric": METRIC, "value": round(value, 2), "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, B, world),
            "clocks": clk.summary(),
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_GBs"], "peak": peak,
                         "unit": "GB/s", "frac": kernels[dom]["frac_of_peak"], "traffic": traffic, "peak_source": peak_src,
                         "whole_step_achieved": round(step_bytes / 1e9 / (ms_per_step / 1e3), 1),
                         "whole_step_frac": round(step_bytes / 1e9 / (ms_per_step / 1e3) / peak, 4)},
            "kernels": kernels,
            "checksum": checksum,
            "metrics": metrics,
            "sustained": sustained,
        }
    # ---- the same step replayed from a CUDA graph (no per-kernel events possible inside a graph, hence a separate
    # leg): at B = 8 the step is bandwidth-bound and the two agree; at B = 1 — the reference's own evaluation batch —
    # issuing 21 launches from Python takes longer than executing them
    graph_leg = None
    if not args.no_graph:
        graph_leg = {}
        for gb in sorted({1, B}):
            ginp = inp if gb == B else make_inputs(gb, dev, seed=4321 + rank, regress=args.regress)
            gpath = path if gb == B else AcvHotPath(filter_mode=args.filter, regress_mode=regress)
            # eager first: capturing moves the allocator's cached blocks into the graph's private pool, and the eager
            # calls that follow would pay fresh cudaMallocs
            for _ in range(args.warmup):
                gpath(**ginp)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                gpath(**ginp)
            e1.record()
            barrier()
            ems = e0.elapsed_time(e1) / args.steps
            replay, _ = gpath.graphed(**ginp)
            for _ in range(args.warmup):
                replay()
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(args.steps):
                replay()
            g1.record()
            barrier()
            gms = g0.elapsed_time(g1) / args.steps
            graph_leg[f"batch_{gb}"] = {"graph_ms_per_step": round(gms, 4), "eager_ms_per_step": round(ems, 4),
                                        "graph_pairs_per_s": round(gb / (gms / 1e3), 1),
                                        "eager_pairs_per_s": round(gb / (ems / 1e3), 1)}
            del replay
        graph_leg["note"] = "rank 0's GPU, AcvHotPath.graphed(): one captured launch sequence per step"
    # ---- e2e: the same step through the public API with HOST buffers ------------------------
    e2e = None if args.no_e2e else run_e2e(args, path, inp, dev, barrier, dist, world)
    # ---- SURVEY.md §8f row f2, reported beside the headline: the same step with F.upsample(trilinear) fused into the
    # regression kernel (its input is the quarter-res conv output; the full-res logits never exist)
    fused = None
    if args.regress == "logits" and not args.no_fused:
        del inp, path, out
        torch.cuda.empty_cache()
        finp = make_inputs(B, dev, seed=1234 + rank, regress="fused")
        fpath = AcvHotPath(filter_mode=args.filter, regress_mode="fused_upsample")
        for _ in range(args.warmup):
            fpath(**finp)
        barrier()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for _ in range(args.steps):
            fpath(**finp)
        f1.record()
        barrier()
        fms = f0.elapsed_time(f1)
        if dist is not None:
            tt = torch.tensor([fms], device=dev, dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            fms = float(tt.item())
        fe2e = None if args.no_e2e else run_e2e(args, fpath, finp, dev, barrier, dist, world)
        fused = {"value": round(B * world / (fms / args.steps / 1e3), 2), "unit": "pairs/s",
                 "ms_per_step": round(fms / args.steps, 4), "e2e": fe2e,
                 "note": "same step with the trilinear x4 upsample fused into softmax/regression (SURVEY.md 8f row f2): "
                         "per-step input [B,1,48,135,240] instead of [B,192,540,960]"}
    legs = None
    if not args.no_legs:
        inp = path = out = finp = fpath = gt = None      # release the headline leg's buffers (rebinding also clears `step`'s cells)
        torch.cuda.empty_cache()
        legs = extra_legs(args, dev, rank, world, barrier, dist)
    if rank == 0:
        # primary end-to-end number: the quarter-resolution boundary (the conv stack's real output; F.upsample fused into the
        # regression kernel) — the full-resolution logits boundary the metric's `value` is defined on ships 10x the bytes
        # over PCIe for tensors that only ever exist on the device, and stays in the line as the stated worst case
        if fused is not None and fused.get("e2e") is not None:
            result["e2e"] = dict(fused["e2e"], boundary="quarter-res cost [B,1,48,h,w] per step, trilinear x4 upsample fused "
                                                        "(AcvHotPath(regress_mode='fused_upsample'))")
            if e2e is not None:
                result["e2e_logits_boundary"] = dict(e2e, boundary="full-res logits [B,192,H,W] per step (worst case)")
        else:
            result["e2e"] = e2e
        if legs:
            result.update(legs)
        if graph_leg is not None:
            result["cuda_graph"] = graph_leg
        if fused is not None:
            result["fused_upsample"] = fused
        if world == 1 and not args.no_cpu_baseline:
            cb, cpu_in, cpu_pred0, cpu_out = cpu_reference(steps=args.cpu_steps, warmup=1, regress=args.regress, keep_io=True)
            result["cpu_baseline"] = cb
            result["parity"] = parity_against_cpu_leg(cpu_in, cpu_pred0, cpu_out, args.filter, args.regress, dev)
            result["gpu_aten_baseline"] = gpu_aten_baseline(cpu_in, args.regress, dev)
        emit(result)
    if dist is not None:
        dist.destroy_process_group()



# ------------------------------------------------------------------------------------------------
# Extra legs of the default run: the ACV step at the north_star's second resolution and the hot-path kernel
# sequences of configs[2] (PCWNet) and configs[3] (IGEV), each with its own roofline, a parity gate against the
# reference's op sequence (oracle/torch_port.py) run on CUDA tensors on the same inputs at B = 1, and that run's time
# as the `gpu_aten_baseline` (what a user of the reference gets on this very GPU today).
# ------------------------------------------------------------------------------------------------
def _timed(fn, steps, warmup, barrier, dist, dev):
    for _ in range(warmup):
        fn(None)
    barrier()
    timer = KernelTimer()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn(timer)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        tt = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    return ms / steps, timer.resolve()


def _leg_result(name, B, world, ms_per_step, kt, steps, ab, extra_cfg):
    peak, peak_src = measured_peak_gbs()
    kernels = {}
    for k, times in kt.items():
        avg = sum(times) / len(times)
        per_launch = ab.get(k, 0) / max(1, len(times) // steps)
        kernels[k] = {"launches_per_step": len(times) // steps, "avg_ms": round(avg, 4),
                      "ms_per_step": round(sum(times) / steps, 4), "algorithmic_GB_per_step": round(ab.get(k, 0) / 1e9, 4),
                      "achieved_GBs": round(per_launch / 1e9 / (avg / 1e3), 1),
                      "frac_of_peak": round(per_launch / 1e9 / (avg / 1e3) / peak, 4)}
    dom = max(kernels, key=lambda k: kernels[k]["ms_per_step"])
    step_bytes = sum(ab.get(k, 0) for k in kernels)
    return {"value": round(B * world / (ms_per_step / 1e3), 2), "unit": "pairs/s", "ms_per_step": round(ms_per_step, 4),
            "config": dict(extra_cfg, pairs_per_gpu=B),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": kernels[dom]["achieved_GBs"], "peak": peak, "unit": "GB/s",
                         "frac": kernels[dom]["frac_of_peak"], "peak_source": peak_src,
                         "whole_step_achieved": round(step_bytes / 1e9 / (ms_per_step / 1e3), 1),
                         "whole_step_frac": round(step_bytes / 1e9 / (ms_per_step / 1e3) / peak, 4)},
            "kernels": kernels}


def _aten_time(fn, steps=3):
    """CUDA-event time of the reference op sequence (torch_port on CUDA tensors), B = 1."""
    out = fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps, out


def _relerr(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def pcw_inputs(B, dev, seed, Hp=384, Wp=1248):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    rn = lambda *s, dt=torch.float32: torch.randn(*s, generator=g, device=dev, dtype=dt)
    ru = lambda *s: torch.rand(*s, generator=g, device=dev)
    D, h, w = 48, Hp // 4, Wp // 4
    from diffuvolume_b200 import ops
    return dict(
        scales=[(rn(B, 320, Hp // s, Wp // s), rn(B, 320, Hp // s, Wp // s), rn(B, 12, Hp // s, Wp // s),
                 rn(B, 12, Hp // s, Wp // s), D * 4 // s) for s in (4, 8, 16, 32)],
        combine=rn(B, 32, D, h, w), costs=[rn(B, 192, Hp, Wp) * 4.0], used=ru(B, Hp, Wp) * 191.0,
        feat_l_full=rn(B, 32, Hp, Wp), feat_r_full=rn(B, 32, Hp, Wp), start=rn(B, D, h, w),
        asd=ops.xstart_from_disp(ru(B, h, w) * 47.75, D, 1.0), shifts=[rn(B, D) * 0.1 for _ in range(3)],
        step_noises=[rn(B, D, h, w, dt=torch.float32 if i == 0 else torch.float64) for i in range(2)],
        q_noises=[rn(B, D, h, w) for _ in range(2)])


def pcw_bytes(B, Hp=384, Wp=1248):
    F4, D, HW, hw = 4, 48, Hp * Wp, (Hp // 4) * (Wp // 4)
    sc = [((Hp // s) * (Wp // s), D * 4 // s) for s in (4, 8, 16, 32)]
    T = 3
    per_pair = {
        "gwc_volume": sum((2 * 320 + 40 * Ds) * p for p, Ds in sc) * F4,
        "concat_volume": sum((2 * 12 + 24 * Ds) * p for p, Ds in sc) * F4,
        "filter": T * (2 * 32 * D * hw * F4 + 2 * D * hw * 8),                       # volume in/out, x_t in, n out
        # logits in, disparity out on every step; the probability volume is written on the last step only (the one ddim_sample
        # returns) — on the other steps it is only reduced to the uncertainty, which re-reads the logits instead
        "softmax_regress": (T * (192 * HW + HW) + 192 * HW) * F4,
        # warp (+ left - warped, + copy of left) and the +-24 volume, written into the refinement network's concat buffer:
        # right, disparity, left in; warped, difference, copy out; then left + warped in, 49 planes out
        "refine_input": T * ((32 + 1 + 32 + 3 * 32) + (2 * 32 + 49)) * HW * F4,
        "uncertainty_vote": (T - 1) * (192 + 3) * HW * F4,
        "ddim_step": T * (2 * HW * F4 + D * hw * (8 + 8 + 4 + 4 + 8 + 8 + 4)),
        "ensemble": (T + 2) * HW * F4,
    }
    return {k: v * B for k, v in per_pair.items()}


def igev_inputs(B, dev, seed, Hp=384, Wp=1248):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    rn = lambda *s, dt=torch.float32: torch.randn(*s, generator=g, device=dev, dtype=dt)
    ru = lambda *s: torch.rand(*s, generator=g, device=dev)
    D, h, w = 48, Hp // 4, Wp // 4
    from diffuvolume_b200 import ops
    return dict(
        fmap_l=rn(B, 96, h, w), fmap_r=rn(B, 96, h, w), geo=rn(B, 8, D, h, w), cost48=rn(B, D, h, w) * 4.0,
        up_weights=torch.softmax(rn(B, 9, Hp, Wp), 1),
        coords=torch.arange(w, device=dev, dtype=torch.float32).view(1, 1, 1, w).expand(B, 1, h, w).contiguous(),
        used=ru(B, Hp, Wp) * 47.0, start=rn(B, D, h, w), asd=ops.xstart_from_disp(ru(B, h, w) * 47.0, D, 1.0),
        shifts=[rn(B, D) * 0.1 for _ in range(2)], step_noises=[rn(B, D, h, w)], q_noises=[rn(B, D, h, w)])


def igev_bytes(B, iters=32, Hp=384, Wp=1248):
    F4, D, HW, h, w = 4, 48, Hp * Wp, Hp // 4, Wp // 4
    hw, T = h * w, 2
    look = (2 * (10 * 8 + 10) + 162 + 2) * hw * F4          # per call: two levels x 10 hypotheses x (8 geo + 1 corr), out, disp/coords
    per_pair = {
        "gwc_volume": (2 * 96 + 8 * D) * hw * F4,
        "softmax_regress": (D + 1) * hw * F4,
        "geo_init": (2 * 96 * hw + 1.5 * hw * w + 2.5 * 8 * D * hw) * F4,           # all-pairs corr (+ pooled level), geo pack
        "filter_factor": T * D * hw * (8 + 4),
        # packed pyramid (two levels = 1.5 x [8, D] per pixel) in and out, the raw noise rows once, then the step's first lookup
        "geo_filter": T * (2 * 1.5 * 8 * D * hw * F4 + D * hw * F4 + look),
        "geo_lookup": T * (iters - 1) * look,
        "context_upsample": T * (10 * HW + hw) * F4,
        "fallback": T * 3 * HW * F4,
        # T = 2: step 0 reads the fp32 state, step noise, asd and its noise (4 x 4 B) and writes x0 (4 B), x_next and eps (fp64);
        # the last step reads the fp64 state and writes x0 / x_next only; both read the upsampled + initial disparity maps
        "ddim_step": T * 3 * HW * F4 + D * hw * ((4 * 4 + 4 + 8 + 8) + (8 + 4 + 8)),
        "ensemble": (T + 2) * HW * F4,
    }
    return {k: v * B for k, v in per_pair.items()}


def extra_legs(args, dev, rank, world, barrier, dist):
    """size_384x1248 / pcwnet / igev legs (rank 0 returns the dict)."""
    from diffuvolume_b200.pipeline import AcvHotPath, IgevHotPath, PcwHotPath
    B, steps, warmup = args.batch, max(3, args.steps // 2), args.warmup
    legs = {}
    do_parity = rank == 0 and not args.no_cpu_baseline
    if do_parity:
        from oracle import dv_oracle as O
        from oracle import torch_port as P

    # ---- the ACV step at 384x1248 (376x1248 padded): north_star's second resolution
    global H, W
    H0, W0 = H, W
    if (H0, W0) == (540, 960):
        H, W = 384, 1248
        inp = make_inputs(B, dev, seed=2234 + rank, regress="logits")
        path = AcvHotPath(filter_mode=args.filter, regress_mode="logits")
        ms, kt = _timed(lambda t: path(**inp, timer=t), steps, warmup, barrier, dist, dev)
        per_launch = algorithmic_bytes(B, args.filter, "logits")
        ab = {k: v * (T_STEPS if k in ("filter", "softmax_regress", "ddim_step") else 1) for k, v in per_launch.items()}
        legs["size_384x1248"] = _leg_result("size_384x1248", B, world, ms, kt, steps, ab,
                                            {"workload": "the headline ACV step at 384x1248 (KITTI, 376 rows padded), D=192",
                                             "filter_mode": args.filter})
        del inp, path
        torch.cuda.empty_cache()
        H, W = H0, W0

    # ---- configs[2]: PCWNet + DiffuVolume, 384x1248
    pin = pcw_inputs(B, dev, 3234 + rank)
    ppath = PcwHotPath()
    ms, kt = _timed(lambda t: ppath(**pin, timer=t), steps, warmup, barrier, dist, dev)
    leg = _leg_result("pcwnet", B, world, ms, kt, steps, pcw_bytes(B),
                      {"workload": "configs[2]: PCWNet+DiffuVolume KITTI12 384x1248: 4-scale gwc + concat(T); T=3 x {filter, "
                                   "softmax/regression (+prob on the last step), refinement input (warp, left - warped, +-24 corr volume, assembled in the concat buffer), "
                                   "uncertainty vote of the refined disparity from the logits, DDIM step}; ensemble"})
    if do_parity:
        one = pcw_inputs(1, dev, 777)
        got = ppath(**one, keep=True)
        sched = O.Schedule(sampling_timesteps=3)
        ams, (pred, (x_last, mask, prob, corr, vols)) = _aten_time(lambda: P.pcw_hot_path_pair(**one, sched=sched))
        err = (got["pred"] - pred).abs()
        leg["parity"] = {"epe_px": float(err.mean()), "max_err_px": float(err.max()),
                         # warp + the +-24 volume on the SAME disparity map (ours), against the port run on the HOST: ATen's CUDA
                         # grid_sample contracts its un-normalisation into an FMA, so its sampling positions differ from the
                         # reference's CPU arithmetic (which warp.cu reproduces) by one ulp — 3e-5 in a tap weight at W = 1248 —
                         # and a pixel at the 0.999 validity threshold can flip; that difference is reported separately
                         "corr_volume_max_rel_err": _relerr(got["corr"].cpu(), torch.squeeze(P.corr_volume_2sided(
                             one["feat_l_full"].cpu(), P.warp(one["feat_r_full"].cpu(), got["disp_last"].unsqueeze(1).cpu()), 24, 1), 1)),
                         "corr_volume_max_rel_err_vs_aten_cuda_grid_sample": _relerr(got["corr"], torch.squeeze(P.corr_volume_2sided(
                             one["feat_l_full"], P.warp(one["feat_r_full"], got["disp_last"].unsqueeze(1)), 24, 1), 1)),
                         "corr_volume_chained_frac_within_1e-4": float(((got["corr"] - corr).abs()
                                                                        <= 1e-4 * corr.abs().max()).float().mean()),
                         "gwc_volume_max_rel_err": max(_relerr(gv, v[:, :40]) for (gv, _), v in zip(got["volumes"], vols)),
                         "concat_volume_bit_exact": all(bool(torch.equal(cv, v[:, 40:])) for (_, cv), v in zip(got["volumes"], vols)),
                         "prob_max_abs_err": float((got["prob"] - prob).abs().max()),
                         "state_rel_err": _relerr(got["x_last"].double(), x_last.double()),
                         "renewal_mask_agreement": float((got["mask"] == mask).float().mean()),
                         "gates": {"epe_px": 0.01, "volume_max_rel_err": 1e-4},
                         "checker": "oracle/torch_port.py:pcw_hot_path_pair on CUDA tensors, same inputs, B=1"}
        leg["gpu_aten_baseline"] = {"value": round(1e3 / ams, 2), "unit": "pairs/s", "ms_per_pair": round(ams, 3),
                                    "what": "the reference's op sequence (ATen kernels) for this leg on this GPU, B=1"}
        del one, got, pred, x_last, mask, prob, corr, vols
    legs["pcwnet"] = leg
    del pin
    torch.cuda.empty_cache()

    # ---- configs[3]: IGEV + DiffuVolume, 384x1248
    iin = igev_inputs(B, dev, 4234 + rank)
    ipath = IgevHotPath()
    ms, kt = _timed(lambda t: ipath(**iin, timer=t), steps, warmup, barrier, dist, dev)
    leg = _leg_result("igev", B, world, ms, kt, steps, igev_bytes(B, ipath.iters),
                      {"workload": "configs[3]: IGEV+DiffuVolume KITTI15 384x1248: gwc + regression(D=48) + all-pairs corr + geo "
                                   "pyramid; T=2 x {filter factor, geo filter, 32 x pyramid lookup, context_upsample, DDIM step}; ensemble",
                       "gru_iters": ipath.iters})
    if do_parity:
        one = igev_inputs(1, dev, 778)
        got = ipath(**one, keep=True)
        sched = O.Schedule(sampling_timesteps=2)
        ams, (pred, (x_last, mask, look, gwc)) = _aten_time(lambda: P.igev_hot_path_pair(**one, sched=sched, iters=ipath.iters), steps=2)
        err = (got["pred"] - pred).abs()
        leg["parity"] = {"epe_px": float(err.mean()), "max_err_px": float(err.max()),
                         "lookup_max_rel_err": _relerr(got["lookup"], look), "gwc_volume_max_rel_err": _relerr(got["gwc"], gwc),
                         "state_rel_err": _relerr(got["x_last"].double(), x_last.double()),
                         "renewal_mask_agreement": float((got["mask"] == mask).float().mean()),
                         "gates": {"epe_px": 0.01, "volume_max_rel_err": 1e-4},
                         "checker": "oracle/torch_port.py:igev_hot_path_pair on CUDA tensors, same inputs, B=1"}
        leg["gpu_aten_baseline"] = {"value": round(1e3 / ams, 2), "unit": "pairs/s", "ms_per_pair": round(ams, 3),
                                    "what": "the reference's op sequence (ATen kernels) for this leg on this GPU, B=1"}
    legs["igev"] = leg
    del iin
    torch.cuda.empty_cache()
    return legs


def run_e2e(args, path, inp, dev, barrier, dist, world):
    """Host-resident inputs (pinned) -> H2D copies -> hot path -> D2H of the prediction, every step inside the
    timed region.  The per-step cost tensors are part of the step's inputs, so they are copied too.

    The DDIM noise is NOT an input: the reference draws it on the device inside ddim_sample (acv_ddim.py:310,354-360), so
    this leg draws it on the device too, inside the timed region, in the reference's order and dtypes (SURVEY.md §8a row a16,
    including the draws whose values the reference never uses) — `--e2e-noise host` restores the round-1 behaviour of
    shipping pre-drawn noise over PCIe.  Pinned buffers are allocated after the rank has been bound to the CPUs local to
    its GPU (diffuvolume_b200.distributed.bind_to_gpu_numa_node), so the pages sit on the GPU's NUMA node."""
    from diffuvolume_b200.distributed import bind_to_gpu_numa_node, gpu_numa_info
    B = args.batch
    # the binding only matters while the pinned buffers are allocated and first touched; the original mask is restored at the
    # end of the leg so that the CPU-baseline legs that follow still see every host core
    affinity0 = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else None
    placement = dict(gpu_numa_info(dev.index), bound_cpus=None if args.no_numa_bind else bind_to_gpu_numa_node(dev.index))
    device_noise = args.e2e_noise == "device"
    names = ["feat_l", "feat_r", "cfeat_l", "cfeat_r", "att_logits", "used", "disp_q"]
    lists = ("shifts",) if device_noise else ("shifts", "step_noises", "renoises")
    host = {k: inp[k].cpu().pin_memory() for k in names}
    host["costs"] = [inp["costs"][0].cpu().pin_memory()]
    for k in lists:
        host[k] = [t.cpu().pin_memory() for t in inp[k]]
    pred_host = torch.empty((B, H, W), dtype=torch.float32).pin_memory()
    h2d = sum(host[k].numel() * host[k].element_size() for k in names)
    h2d += T_STEPS * host["costs"][0].numel() * 4      # the cost tensor of each of the T steps is a step input
    h2d += sum(t.numel() * t.element_size() for k in lists for t in host[k])
    d2h = pred_host.numel() * 4
    # Two device-side input sets: the H2D copies of step i+1 run on copy streams while step i computes (every step's
    # inputs are still copied inside the timed region; the first copy of the region is not overlapped with anything).
    sets = []
    for _ in range(2):
        d = {k: torch.empty_like(inp[k]) for k in names}
        for k in lists:
            d[k] = [torch.empty_like(t) for t in inp[k]]
        sets.append(d)
    # the T cost tensors of a step are inputs of that step like everything else: both device-side sets hold all T of them and
    # they are prefetched on the copy streams with the rest (round 1 copied each one on the compute stream right before its
    # DDIM iteration, which serialised 1/4 of the H2D bytes with the kernels).  The full-resolution logits boundary keeps the
    # two-buffer streaming form when T x 2 sets would not fit beside the working set (3.2 GB per tensor at B = 8).
    cost_bytes = inp["costs"][0].numel() * 4
    prefetch_costs = 2 * T_STEPS * cost_bytes < 24 * (1 << 30)
    if prefetch_costs:
        for d in sets:
            d["costs"] = [torch.empty_like(inp["costs"][0]) for _ in range(T_STEPS)]
    cost_bufs = [] if prefetch_costs else [torch.empty_like(inp["costs"][0]) for _ in range(2)]
    copy_streams = [torch.cuda.Stream(device=dev) for _ in range(max(1, args.e2e_copy_streams))]
    ready = [[torch.cuda.Event() for _ in copy_streams] for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    gen = torch.Generator(device=dev)
    gen.manual_seed(99)
    shape_q = tuple(inp["step_noises"][0].shape)

    def stage_logits(i):
        # the cost tensor of step i arrives from the host right before that step consumes it; two device buffers
        # alternate (stream order guarantees step i-2 has consumed a buffer before it is overwritten)
        buf = cost_bufs[i % 2]
        buf.copy_(host["costs"][0], non_blocking=True)
        return buf

    def prefetch(i):
        d = sets[i % 2]
        jobs = [(d[k], host[k]) for k in names] + [(dd, ss) for k in lists for dd, ss in zip(d[k], host[k])]
        if prefetch_costs:
            jobs += [(dd, host["costs"][0]) for dd in d["costs"]]
        jobs.sort(key=lambda j: -j[0].numel())            # big tensors first, dealt round-robin to the copy streams
        for si, st in enumerate(copy_streams):
            with torch.cuda.stream(st):
                st.wait_event(consumed[i % 2])             # the step that last read this set has finished
                for dd, ss in jobs[si::len(copy_streams)]:
                    dd.copy_(ss, non_blocking=True)
                ready[i % 2][si].record(st)

    def draw_noise():
        # acv_ddim.py:310 (unused start noise), then per non-final step :354 randn_like(img) (fp32 at step 1, fp64 after),
        # :358 randint, :243 randn_like(asd) fp32 (only a dtype/shape template there), :360 rand_like fp64
        torch.randn(shape_q, generator=gen, device=dev)
        sn, rn_ = [], []
        for i in range(T_STEPS - 1):
            sn.append(torch.randn(shape_q, generator=gen, device=dev, dtype=torch.float32 if i == 0 else torch.float64))
            torch.randint(0, 1000, (1,), generator=gen, device=dev)
            torch.randn(shape_q, generator=gen, device=dev)
            rn_.append(torch.rand(shape_q, generator=gen, device=dev, dtype=torch.float64))
        return sn, rn_

    def compute(i):
        d = sets[i % 2]
        cur = torch.cuda.current_stream(dev)
        for e in ready[i % 2]:
            cur.wait_event(e)
        sn, rn_ = draw_noise() if device_noise else (d["step_noises"], d["renoises"])
        out = path(**{k: d[k] for k in names}, costs=d["costs"] if prefetch_costs else stage_logits, shifts=d["shifts"],
                   step_noises=sn, renoises=rn_)
        consumed[i % 2].record(cur)
        pred_host.copy_(out["pred"], non_blocking=True)

    def run(n):
        prefetch(0)
        for i in range(n):
            if i + 1 < n:
                prefetch(i + 1)
            compute(i)

    steps = max(2, min(args.steps, args.e2e_steps))
    if args.regress != "logits" or inp["costs"][0].dim() == 5:
        steps = max(steps, min(args.steps, 8))          # the fused-upsample step is 10x shorter: time more of them
    for e in consumed:
        e.record(torch.cuda.current_stream(dev))
    run(1)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    run(steps)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    if dist is not None:
        tt = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    if affinity0 is not None:
        try:
            os.sched_setaffinity(0, affinity0)
        except OSError:
            pass
    return {"value": round(B * world * steps / (ms / 1e3), 2), "unit": "pairs/s", "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h), "steps": steps, "ms_per_step": round(ms / steps, 3),
            "h2d_GBs_per_gpu": round(h2d * steps / 1e9 / (ms / 1e3), 1),
            "noise": "drawn on the device inside the timed region, reference order/dtypes (acv_ddim.py:310,354-360)"
                     if device_noise else "pre-drawn on the host and copied with the inputs",
            "copy_streams": len(copy_streams), "cost_tensors": "prefetched with the step's other inputs" if prefetch_costs
            else "streamed on the compute stream, one DDIM iteration ahead (2 device buffers)",
            "host_placement_rank0": placement,
            "note": f"pinned host inputs incl. the T=5 per-step {list(host['costs'][0].shape)} cost tensors; PCIe-bound; "
                    "H2D of step i+1 overlapped with the compute of step i (copy streams + events)"}


# ------------------------------------------------------------------------------------------------
def cpu_reference(steps: int, warmup: int, regress: str = "logits", keep_io: bool = False):
    """The reference's op sequence on the host CPU (oracle/torch_port.py), one pair (B=1) per step.  keep_io: also return
    the leg's inputs and its outputs, so that the GPU path can be checked against what this leg computed."""
    from oracle import dv_oracle as O
    from oracle import torch_port as P
    ncpu = os.cpu_count() or 1
    torch.set_num_threads(ncpu)
    g = torch.Generator().manual_seed(0)
    h, w, D = H // 4, W // 4, MAXDISP // 4
    rn = lambda *s, dt=torch.float32: torch.randn(*s, generator=g, dtype=dt)
    sched = O.Schedule()
    a = dict(feat_l=rn(1, C_GWC, h, w), feat_r=rn(1, C_GWC, h, w), cfeat_l=rn(1, C_CAT, h, w), cfeat_r=rn(1, C_CAT, h, w),
             att_logits=rn(1, 1, D, h, w),
             costs=([rn(1, MAXDISP, H, W) * 4.0] if regress == "logits" else [rn(1, 1, D, h, w) * 4.0]) * T_STEPS,
             used=torch.rand(1, H, W, generator=g) * 191.0,
             asd=None,

             shifts=[rn(1, D) * 0.1 for _ in range(T_STEPS)],
             step_noises=[rn(1, D, h, w, dt=torch.float32 if i == 0 else torch.float64) for i in range(T_STEPS - 1)],
             renoises=[torch.rand(1, D, h, w, generator=g, dtype=torch.float64) for _ in range(T_STEPS - 1)])
    pred0 = torch.rand(1, H, W, generator=g) * 191.0          # the origin model's disparity the sampler starts from
    a["asd"] = P.xstart_from_pred(pred0)
    up = None if regress == "logits" else (MAXDISP, H, W)
    with torch.no_grad():
        for _ in range(warmup):
            P.hot_path_pair(**a, sched=sched, upsample_to=up)
        ts = []
        for _ in range(steps):
            t0 = time.perf_counter()
            out = P.hot_path_pair(**a, sched=sched, upsample_to=up)
            ts.append(time.perf_counter() - t0)
    sec = sum(ts) / len(ts)
    res = _cpu_result(steps, sec, regress)
    return (res, a, pred0, out) if keep_io else res


def _cpu_result(steps, sec, regress):
    return {"value": round(1.0 / sec, 4), "unit": "pairs/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{steps} x 1 pair (B=1, all T=5 steps, {H}x{W} D=192{'' if regress == 'logits' else ', incl. F.upsample trilinear per step'}), "
                      f"torch CPU op-for-op port of the reference (oracle/torch_port.py), {sec:.2f} s/pair",
            "seconds_per_pair": round(sec, 3)}


def parity_against_cpu_leg(a, pred0, cpu_out, filter_mode, regress, dev):
    """The fused GPU path on the very inputs the cpu_baseline leg just ran (B = 1, full size), compared with that leg's
    outputs: SURVEY.md §8d "parity gates reported with every timing" (volume max rel error, disparity EPE in px)."""
    from diffuvolume_b200 import ops
    from diffuvolume_b200.pipeline import AcvHotPath
    cpu_pred, (cpu_gwc, _, cpu_mask) = cpu_out
    to = lambda t: t.to(dev)
    h, w = H // 4, W // 4
    disp_q = ops.downsample_bilinear(to(pred0), (h, w), clamp=(0, MAXDISP - 1), post_scale=0.25)
    path = AcvHotPath(filter_mode=filter_mode, regress_mode="logits" if regress == "logits" else "fused_upsample")
    out = path(to(a["feat_l"]), to(a["feat_r"]), to(a["cfeat_l"]), to(a["cfeat_r"]), to(a["att_logits"]), [to(a["costs"][0])],
               to(a["used"]), disp_q, [to(t) for t in a["shifts"]], [to(t) for t in a["step_noises"]],
               [to(t) for t in a["renoises"]], keep_volumes=True)
    err = (out["pred"].cpu() - cpu_pred).abs()
    gwc = out["gwc"].cpu()
    return {"epe_px": float(err.mean()), "max_err_px": float(err.max()),
            "gwc_volume_max_rel_err": float((gwc - cpu_gwc).abs().max() / cpu_gwc.abs().max()),
            "renewal_mask_agreement": float((out["mask"].cpu() == cpu_mask).float().mean()),
            "gates": {"epe_px": 0.01, "volume_max_rel_err": 1e-4},
            "checker": "the cpu_baseline leg's own outputs (oracle/torch_port.py on the host), same inputs, B=1"}


def gpu_aten_baseline(a, regress, dev):
    """The reference's op sequence for one pair (oracle/torch_port.py:hot_path_pair — what the reference's PyTorch code
    launches between its convolutions) on CUDA tensors on THIS GPU: the baseline a user of the reference has today.
    Stated baseline only; nothing in the product path touches it."""
    from oracle import dv_oracle as O
    from oracle import torch_port as P
    to = lambda v: [t.to(dev) for t in v] if isinstance(v, (list, tuple)) else (v.to(dev) if torch.is_tensor(v) else v)
    d = {k: to(v) for k, v in a.items()}
    up = None if regress == "logits" else (MAXDISP, H, W)
    sched = O.Schedule()
    with torch.no_grad():
        ms, _ = _aten_time(lambda: P.hot_path_pair(**d, sched=sched, upsample_to=up), steps=3)
    return {"value": round(1e3 / ms, 2), "unit": "pairs/s", "ms_per_pair": round(ms, 3), "batch": 1,
            "what": "oracle/torch_port.py:hot_path_pair (the reference's ATen op sequence) on CUDA tensors, same GPU, B=1, "
                    "CUDA-event timed"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_reference(steps=max(1, args.steps), warmup=max(0, min(args.warmup, 1)), regress=args.regress)
    v = cb["value"]
    emit({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "pairs/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(1e3 / v, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        # the arm's config IS the other arm's; what differs is how much of it one timed step covers (cpu_baseline.sample)
        "config": dict(workload_config(args, args.batch, max(1, args.gpus)),
                       sample="each timed step = 1 pair of this workload on the host CPU (bounded sample; the metric is a per-pair rate)"),
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--batch", type=int, default=8, help="stereo pairs per GPU per step")
    ap.add_argument("--filter", choices=["regenerate", "volume"], default="regenerate")
    ap.add_argument("--regress", choices=["logits", "fused"], default="logits",
                    help="logits: softmax/regression reads the full-res [B,192,H,W] logits (the metric's op boundary); "
                         "fused: trilinear x4 upsample fused in, input [B,1,48,h,w] (SURVEY.md 8f row f2)")
    ap.add_argument("--no-fused", action="store_true", help="skip the extra fused-upsample measurement")
    ap.add_argument("--size", default="540x960",
                    help="full-resolution HxW of the synthetic pairs (multiples of 4): 540x960 is BASELINE.json's metric; "
                         "384x1248 (KITTI, 376 padded) is the north_star's second resolution")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--e2e-noise", choices=["device", "host"], default="device",
                    help="device: DDIM noise drawn on the GPU inside the timed region (as the reference does); host: shipped over PCIe")
    ap.add_argument("--e2e-copy-streams", type=int, default=2)
    ap.add_argument("--no-numa-bind", action="store_true", help="do not bind the rank to its GPU's local CPUs before pinning")
    ap.add_argument("--no-graph", action="store_true", help="skip the CUDA-graph replay leg")
    ap.add_argument("--cpu-steps", type=int, default=3)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the host-buffer leg")
    ap.add_argument("--no-legs", action="store_true", help="skip the size_384x1248 / pcwnet / igev legs")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 2 s sustained leg")
    ap.add_argument("--sustained-seconds", type=float, default=2.2)
    args = ap.parse_args()
    protect_stdout()
    global H, W, METRIC
    H, W = (int(v) for v in args.size.lower().split("x"))
    if H % 4 or W % 4 or H <= 0 or W <= 0:
        raise SystemExit("--size HxW must be positive multiples of 4")
    if (H, W) != (540, 960):
        METRIC = f"volume+DDIM-filter pairs/s @{H}x{W} D=192"
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
