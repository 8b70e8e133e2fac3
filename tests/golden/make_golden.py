"""Mint the golden fixtures in tests/golden/*.npz by running the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference):

    python tests/golden/make_golden.py            # all three sub-projects
    python tests/golden/make_golden.py sceneflow  # one of: sceneflow | kitti12 | kitti15

Each sub-project is imported in its own subprocess with that sub-project's directory on
sys.path (SceneFlow/, KITTI12/ and KITTI15/ all use the same top-level package names, so they
cannot live in one interpreter).  Inputs are regenerated from `tests/synth.py` seeds, so the
fixtures hold only outputs (plus the seeds / shapes that define the inputs).

Shims applied to the reference while generating (nothing else is touched):
  * torch.Tensor.cuda -> identity (the reference hard-codes .cuda(); there is no GPU here);
  * in the DDIM trace, the 3-D conv stack (dres0..dres3, classif2) is replaced by a cheap
    deterministic stand-in (mean over channels * 2 + a per-step bias) — the convolutions are
    out of scope and could not travel to the GPU box anyway — and
    torch.randn_like / torch.rand_like return seeded synthetic noise so that the oracle and the
    CUDA path can be fed the very same tensors (SURVEY.md §8a "RNG contract").
"""
from __future__ import annotations

import importlib.util
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF = Path(os.environ.get("DV_REFERENCE", "/root/reference"))
sys.path.insert(0, str(HERE.parent))
import synth  # noqa: E402


def _load(path: Path, name: str):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _t(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a))


def _sample(a: np.ndarray, n: int = 4096) -> np.ndarray:
    flat = a.reshape(-1)
    idx = np.linspace(0, flat.size - 1, min(n, flat.size)).astype(np.int64)
    return flat[idx]


# warp cases: name -> (shape [B,C,H,W], seed, disparity amplitude)
WARP_CASES = {"small": ((2, 6, 9, 40), 61, 30.0), "odd": ((1, 3, 7, 13), 63, 8.0), "wide": ((1, 4, 5, 156), 65, 60.0)}

# volume cases shared by the sub-projects: name -> (B, C, G, D, H, W, seed)
GWC_CASES = {
    "w39": (2, 32, 4, 6, 12, 39, 11),        # PCWNet 1/32 scale width (W % 4 != 0, HW % 4 == 0)
    "cpg12": (1, 24, 2, 12, 8, 20, 12),       # IGEV-style 12 channels per group
    "tiny_d_gt_w": (1, 8, 4, 9, 3, 7, 13),    # D > W: planes d >= W stay zero; HW % 4 != 0
    "d48": (1, 64, 8, 48, 4, 56, 14),         # ACV-style cpg=8, D=48
    "w78": (1, 16, 2, 12, 6, 78, 15),         # PCWNet 1/16 scale width
}
CONCAT_CASES = {
    "c12": (2, 12, 6, 6, 20, 21),             # (B, C, D, H, W, seed)
    "c4_d48": (1, 4, 48, 4, 56, 22),
    "tiny": (1, 3, 9, 3, 7, 23),
}


def trace_inputs(B, D, h, w, H, W):
    """Synthetic DDIM-trace inputs shared by make_golden.py and the tests: per-step logit biases
    peaked at a per-pixel disparity (so the regressed disparity has low uncertainty) and a `used`
    map that agrees with it on the left half of the image only — the renewal mask ends up mixed."""
    dstar = 6.0 + 30.0 * synth.uniform((B, 1, 1, h, w), 75, dtype=np.float32)
    dvals = np.arange(D, dtype=np.float32).reshape(1, 1, D, 1, 1)
    bias = []
    for i in range(5):
        peak = -2.0 * (dvals - (dstar + np.float32(0.3 * i))) ** 2
        bias.append((peak + synth.normal((B, 1, D, h, w), 80 + i) * np.float32(0.25)).astype(np.float32))
    up = np.repeat(np.repeat(dstar[:, 0, 0], 4, axis=1), 4, axis=2) * np.float32(4) + np.float32(1.5)
    used = up + (synth.uniform((B, H, W), 73, dtype=np.float32) - np.float32(0.5)) * np.float32(0.8)
    used[:, :, W // 2:] += np.float32(25)
    return bias, used.astype(np.float32)


def _volume_goldens(sub, out: dict, prefix: str, mask_left_expected: bool):
    import torch
    for name, (B, C, G, D, H, W, seed) in GWC_CASES.items():
        ref, tgt = synth.normal((B, C, H, W), seed), synth.normal((B, C, H, W), seed + 1000)
        v = sub.build_gwc_volume(_t(ref), _t(tgt), D, G).numpy()
        out[f"{prefix}.gwc.{name}"] = v
        if name == "cpg12":
            out[f"{prefix}.groupwise.{name}"] = sub.groupwise_correlation(_t(ref), _t(tgt), G).numpy()
    for name, (B, C, D, H, W, seed) in CONCAT_CASES.items():
        ref, tgt = synth.normal((B, C, H, W), seed), synth.normal((B, C, H, W), seed + 1000)
        v = sub.build_concat_volume(_t(ref), _t(tgt), D).numpy()
        out[f"{prefix}.concat.{name}"] = v
        # which variant is this file?  (left half zero-masked for x < d or not)
        masked = bool((v[:, :C, D - 1, :, : min(D - 1, W)] == 0).all())
        assert masked == mask_left_expected, (prefix, name, masked)
    # regression: softmax(N(0,1)*k) -> disparity_regression
    for k in (1, 10):
        cost = synth.normal((1, 192, 16, 32), 31) * np.float32(k)
        prob = torch.softmax(_t(cost), dim=1)
        out[f"{prefix}.regress.k{k}"] = sub.disparity_regression(prob, 192).numpy()


def gen_sceneflow():
    import torch
    import torch.nn as nn
    import torch.nn.functional as F
    sys.path.insert(0, str(REF / "SceneFlow"))
    torch.Tensor.cuda = lambda self, *a, **k: self
    import models.submodule as sub          # the file ACVNet actually uses
    import models.acv_ddim as acv_ddim
    top = _load(REF / "SceneFlow" / "submodule.py", "sf_top_submodule")  # orphan copy (variant T)
    out = {}
    _volume_goldens(sub, out, "sf", mask_left_expected=False)
    _volume_goldens(top, out, "sftop", mask_left_expected=True)
    # two-sided correlation volume (SceneFlow/submodule.py:172-186)
    ref, tgt = synth.normal((1, 32, 8, 64), 41), synth.normal((1, 32, 8, 64), 1041)
    out["sftop.corr2.m24"] = top.build_corrleation_volume(_t(ref), _t(tgt), 24, 1).numpy()
    ref, tgt = synth.normal((2, 8, 3, 7), 42), synth.normal((2, 8, 3, 7), 1042)
    out["sftop.corr2.tiny_m9"] = top.build_corrleation_volume(_t(ref), _t(tgt), 9, 2).numpy()

    # ACV attention volume (acv_ddim.py:390)
    B, C, D, h, w = 1, 4, 48, 4, 56
    cl, cr = synth.normal((B, C, h, w), 51), synth.normal((B, C, h, w), 1051)
    att = synth.normal((B, 1, D, h, w), 52) * np.float32(3)
    concat = sub.build_concat_volume(_t(cl), _t(cr), D)
    out["sf.acv_volume"] = (F.softmax(_t(att), dim=2) * concat).numpy()

    # ---- schedule constants + DDIM trace from the real ACVNet_DDIM ------------------------
    torch.manual_seed(0)
    net = acv_ddim.ACVNet_DDIM(192, False, False).eval()
    for name in ("betas", "alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
                 "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod"):
        out[f"sched.{name}"] = getattr(net, name).numpy()
    for S in (5, 3, 2):
        times = torch.linspace(-1, 999, steps=S + 1)
        times = list(reversed(times.int().tolist()))
        out[f"sched.time_pairs.S{S}"] = np.array(list(zip(times[:-1], times[1:])), dtype=np.int64)

    # q_sample / predict_noise_from_start as methods (acv_ddim.py:241-252)
    x0 = synth.uniform((2, 48, 4, 8), 61, dtype=np.float32) * 2 - 1
    nz = synth.normal((2, 48, 4, 8), 62)
    for t in (999, 599, 0):
        tt = torch.full((1,), t, dtype=torch.long)
        qs = net.q_sample(_t(x0), tt, _t(nz))
        out[f"sf.q_sample.t{t}"] = qs.numpy()
        out[f"sf.pred_noise.t{t}"] = net.predict_noise_from_start(qs, tt, _t(x0)).numpy()
        out[f"sf.pred_noise_f32.t{t}"] = net.predict_noise_from_start(_t(nz), tt, _t(x0)).numpy()

    # trace
    B, H, W = 2, 16, 32
    h, w, D, Cc = H // 4, W // 4, 48, 32
    cl, cr = synth.normal((B, Cc, h, w), 71), synth.normal((B, Cc, h, w), 1071)
    att = synth.normal((B, 1, D, h, w), 72) * np.float32(2)
    ac_volume = F.softmax(_t(att), dim=2) * sub.build_concat_volume(_t(cl), _t(cr), D)
    bias, used = trace_inputs(B, D, h, w, H, W)
    disp_q = synth.uniform((B, h, w), 74, dtype=np.float32) * np.float32(47.75)
    out["trace.disp_q"] = disp_q
    # asd = x_start of the initial disparity: use the reference's own inline code (acv_ddim.py:403-419)
    b = B
    disp = _t(disp_q).unsqueeze(1)
    disp_volume = torch.zeros([b, 48, h, w], dtype=torch.float32)
    real = torch.floor(disp).long()
    mask = real == 47
    coff = real - disp + 1
    disp_volume = disp_volume.view(b, 48, -1).scatter_(1, real.view(b, 1, -1), coff.view(b, 1, -1)).reshape(b, 48, h, w)
    disp_volume = disp_volume.view(b, 48, -1).scatter_(1, torch.clamp(real + 1, 0, 47).view(b, 1, -1),
                                                       (1 - coff).view(b, 1, -1)).reshape(b, 48, h, w)
    fuzhi = torch.zeros([b, 48, h, w], dtype=torch.float32)
    fuzhi[:, -1, :, :] = 1
    asd = torch.where(mask == True, fuzhi, disp_volume)  # noqa: E712
    asd = (asd * 2 - 1) * net.scale
    out["trace.asd"] = asd.numpy()

    calls = {"n": 0}

    class Stand0(nn.Module):
        def forward(self, v):
            return v.mean(1, keepdim=True) * 2.0

    class Zero(nn.Module):
        def forward(self, v):
            return torch.zeros_like(v)

    class Bias(nn.Module):
        def forward(self, v):
            o = v + _t(bias[calls["n"]])
            calls["n"] += 1
            return o

    net.dres0, net.dres1, net.dres2, net.dres3, net.classif2 = Stand0(), Zero(), nn.Identity(), nn.Identity(), Bias()

    rec = {"img": [], "eps": [], "x0": [], "disp": []}
    orig_mp = net.model_predictions

    def mp(volume, noise, t):
        rec["img"].append(noise.detach().clone())
        r = orig_mp(volume, noise, t)
        rec["eps"].append(r[0].clone()); rec["x0"].append(r[1].clone()); rec["disp"].append(r[2].clone())
        return r
    net.model_predictions = mp

    seeds = {"randn_like": [], "rand_like": []}
    k = {"n": 0}
    o_randn_like, o_rand_like = torch.randn_like, torch.rand_like

    def randn_like(x, **kw):
        seed = 2000 + k["n"]; k["n"] += 1
        seeds["randn_like"].append((seed, 1 if x.dtype == torch.float64 else 0))
        return _t(synth.normal(tuple(x.shape), seed, dtype=np.float64)).to(x.dtype)

    def rand_like(x, **kw):
        seed = 2000 + k["n"]; k["n"] += 1
        seeds["rand_like"].append((seed, 1 if x.dtype == torch.float64 else 0))
        return _t(synth.uniform(tuple(x.shape), seed, dtype=np.float64)).to(x.dtype)

    torch.randn_like, torch.rand_like = randn_like, rand_like
    try:
        with torch.no_grad():
            pred, final = net.ddim_sample(ac_volume, _t(used), asd)
    finally:
        torch.randn_like, torch.rand_like = o_randn_like, o_rand_like
    out["trace.pred"] = pred.numpy()
    out["trace.final"] = final.numpy()
    for key, lst in rec.items():
        for i, v in enumerate(lst):
            out[f"trace.{key}.{i}"] = v.numpy()
    out["trace.randn_like_seeds"] = np.array(seeds["randn_like"], dtype=np.int64)
    out["trace.rand_like_seeds"] = np.array(seeds["rand_like"], dtype=np.int64)
    # DynamicHead shift per sampled timestep (head.py:74-77): time_embedding(0, t) == shift
    with torch.no_grad():
        for t in (999, 799, 599, 399, 199):
            tc = torch.full((B,), t, dtype=torch.long)
            out[f"trace.shift.t{t}"] = net.time_embedding(torch.zeros(B, 48, 1, 1), tc).reshape(B, 48).numpy()
    out["trace.shape"] = np.array([B, Cc, D, h, w, H, W], dtype=np.int64)

    # ---- big-shape checksums (BASELINE config 2 at B=1): sums + strided samples only -------
    B, C, G, D, H, W = 1, 320, 40, 48, 135, 240
    ref, tgt = synth.normal((B, C, H, W), 91), synth.normal((B, C, H, W), 1091)
    torch.set_num_threads(os.cpu_count() or 1)
    v = sub.build_gwc_volume(_t(ref), _t(tgt), D, G).numpy()
    out["big.gwc.sum"] = np.array([v.sum(dtype=np.float64), np.abs(v).sum(dtype=np.float64), np.abs(v).max()])
    out["big.gwc.sample"] = _sample(v)
    cost = synth.normal((1, 192, 135, 240), 92) * np.float32(4)
    prob = torch.softmax(_t(cost), dim=1)
    d = sub.disparity_regression(prob, 192)
    out["big.regress.disp"] = d.numpy().astype(np.float32)
    disp_values = torch.arange(0, 192, dtype=d.dtype).view(1, 192, 1, 1)
    out["big.regress.unc"] = torch.sum(torch.abs(d.unsqueeze(1) - disp_values) * prob, dim=1).numpy()
    np.savez_compressed(HERE / "sceneflow.npz", **out)
    print("sceneflow:", len(out), "arrays")


def _pcw_sampler_trace(out):
    """Trace of the REFERENCE's PWCNet_ddim.ddim_sample / model_predictions (KITTI12/models/pwcnet_ddim.py:466-602): its
    unmodified sampler methods bound onto tests/pcw_mock.py:MockPCW (conv stacks replaced by cheap stand-ins); the
    reference's own warp / build_corrleation_volume / disparity_regression run underneath.  torch.randn / randn_like
    return seeded synthetic noise so that the CUDA path can be fed the same tensors."""
    import torch
    import models.pwcnet_ddim as M
    from pcw_mock import PCW_TRACE, MockPCW, pcw_trace_inputs
    sys.path.insert(0, str(HERE.parent.parent))
    from oracle import dv_oracle as O

    inp = pcw_trace_inputs("cpu")
    net = MockPCW(O.Schedule(), inp["shifts"])
    for name in ("q_sample", "predict_noise_from_start", "model_predictions", "ddim_sample"):
        setattr(MockPCW, name, getattr(M.PWCNet_ddim, name))
    # asd = x_start volume of the quarter-res initial disparity: the reference's inline code (pwcnet_ddim.py:738-754)
    b, h, w = inp["gt_q"].shape
    dn = inp["gt_q"].reshape(b, 1, 1, h, w)
    dv = torch.zeros([b, 48, h, w], dtype=torch.float32)
    real = torch.floor(dn).long()
    mask = real == 47
    coff = real - dn + 1
    dv = dv.view(b, 48, -1).scatter_(1, real.view(b, 1, -1), coff.view(b, 1, -1)).reshape(b, 48, h, w)
    dv = dv.view(b, 48, -1).scatter_(1, torch.clamp(real + 1, 0, 47).view(b, 1, -1), (1 - coff).view(b, 1, -1)).reshape(b, 48, h, w)
    fuzhi = torch.zeros([b, 48, h, w], dtype=torch.float32)
    fuzhi[:, -1] = 1
    asd = net.scale * (torch.where(mask.squeeze(1) == True, fuzhi, dv) * 2 - 1.)  # noqa: E712
    asd = torch.clamp(asd, min=-net.scale, max=net.scale)
    out["pcw.asd"] = asd.numpy()

    rec = {"img": [], "eps": [], "x0": [], "disp": [], "prob": []}
    orig_mp = net.model_predictions

    def mp(volume, img, t, fl, fr):
        rec["img"].append(img.detach().clone())
        r = orig_mp(volume, img, t, fl, fr)
        for key, v in zip(("eps", "x0", "disp", "prob"), r):
            rec[key].append(v.detach().clone())
        return r
    net.model_predictions = mp
    k = {"n": 0}
    seeds = []
    o_randn_like, o_randn = torch.randn_like, torch.randn

    def randn_like(x, **kw):
        seed = 5100 + k["n"]; k["n"] += 1
        seeds.append((seed, 1 if x.dtype == torch.float64 else 0))
        return _t(synth.normal(tuple(x.shape), seed, dtype=np.float64)).to(x.dtype)

    def randn(*shape, **kw):
        shape = tuple(shape[0]) if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)) else tuple(shape)
        return _t(synth.normal(shape, 5000))
    torch.randn_like, torch.randn = randn_like, randn
    try:
        with torch.no_grad():
            final, prob = net.ddim_sample(inp["volume"], inp["used"], asd, inp["fl"], inp["fr"])
    finally:
        torch.randn_like, torch.randn = o_randn_like, o_randn
    out["pcw.final"] = final.numpy()
    for key, lst in rec.items():
        for i, v in enumerate(lst):
            out[f"pcw.{key}.{i}"] = _sample(v.numpy()) if key == "prob" else v.numpy()
    out["pcw.randn_like_seeds"] = np.array(seeds, dtype=np.int64)
    print("pcw trace: steps", len(rec["disp"]), "randn_like draws", k["n"],
          "final range", float(final.min()), float(final.max()))


def gen_kitti12():
    import torch
    sys.path.insert(0, str(REF / "KITTI12"))
    torch.Tensor.cuda = lambda self, *a, **k: self
    import models.submodule as sub
    out = {}
    _volume_goldens(sub, out, "k12", mask_left_expected=True)
    ref, tgt = synth.normal((1, 32, 8, 64), 41), synth.normal((1, 32, 8, 64), 1041)
    out["k12.corr2.m24"] = sub.build_corrleation_volume(_t(ref), _t(tgt), 24, 1).numpy()
    ref, tgt = synth.normal((1, 32, 6, 80), 43), synth.normal((1, 32, 6, 80), 1043)
    out["k12.corr2.w80_m24"] = sub.build_corrleation_volume(_t(ref), _t(tgt), 24, 1).numpy()
    # warp (KITTI12/models/submodule.py:137-176): x.get_device() is -1 on the CPU, which torch.arange rejects
    torch.Tensor.get_device = lambda self: self.device
    for key, (shape, seed, amp) in WARP_CASES.items():
        x = synth.normal(shape, seed)
        disp = synth.uniform((shape[0], 1, shape[2], shape[3]), seed + 1, dtype=np.float32) * np.float32(amp) - np.float32(3)
        out["k12.warp." + key] = sub.warp(_t(x), _t(disp)).detach().numpy()
    _pcw_sampler_trace(out)
    np.savez_compressed(HERE / "kitti12.npz", **out)
    print("kitti12:", len(out), "arrays")


def _igev_sampler_trace(out, GeoDdim):
    """Trace of the REFERENCE's IGEVStereo_ddim.ddim_sample / model_predictions (KITTI15/core/igev_stereo_ddim.py:226-359).
    The class itself cannot be constructed (timm / opt_einsum / pretrained download), so its unmodified sampler methods
    are bound onto tests/igev_mock.py:MockIGEV; `timm` and `opt_einsum` are stubbed only to make the module importable."""
    import types
    import torch
    for name in ("timm", "opt_einsum"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if not hasattr(sys.modules["opt_einsum"], "contract"):
        sys.modules["opt_einsum"].contract = lambda *a, **k: None
    import core.igev_stereo_ddim as M
    from igev_mock import IGEV_TRACE, MockIGEV, igev_trace_inputs
    sys.path.insert(0, str(HERE.parent.parent))
    from oracle import dv_oracle as O

    inp = igev_trace_inputs("cpu")
    net = MockIGEV(O.Schedule(), inp["shifts"])
    for name in ("q_sample", "predict_noise_from_start", "model_predictions", "ddim_sample"):
        setattr(MockIGEV, name, getattr(M.IGEVStereo_ddim, name))
    geo_fn = GeoDdim(inp["f1"], inp["f2"], inp["geo"], radius=4, num_levels=2)
    # asd = x_start of the (quarter-res) ground-truth / origin disparity: the reference's inline code (:403-420)
    b, _, h, w = inp["gt_q"].shape
    tc = torch.clamp(inp["gt_q"], 0, 48 - 1)
    dv = torch.zeros([b, 48, h, w], dtype=torch.float32)
    real = torch.floor(tc).long()
    mask = real == 47
    coff = real - tc + 1
    dv = dv.view(b, 48, -1).scatter_(1, real.view(b, 1, -1), coff.view(b, 1, -1)).reshape(b, 48, h, w)
    dv = dv.view(b, 48, -1).scatter_(1, torch.clamp(real + 1, 0, 47).view(b, 1, -1), (1 - coff).view(b, 1, -1)).reshape(b, 48, h, w)
    fuzhi = torch.zeros([b, 48, h, w], dtype=torch.float32)
    fuzhi[:, -1] = 1
    asd = (torch.where(mask == True, fuzhi, dv) * 2 - 1) * net.scale  # noqa: E712
    out["igev.asd"] = asd.numpy()

    rec = {"img": [], "eps": [], "x0": [], "disp": [], "coords1": []}
    orig_mp = net.model_predictions

    def mp(*a):
        rec["img"].append(a[7].detach().clone())
        r = orig_mp(*a)
        for key, v in zip(("eps", "x0", "disp", "coords1"), r):
            rec[key].append(v.detach().clone())
        return r
    net.model_predictions = mp
    k = {"n": 0}
    seeds = []
    o_randn_like = torch.randn_like

    def randn_like(x, **kw):
        seed = 3000 + k["n"]; k["n"] += 1
        seeds.append((seed, 1 if x.dtype == torch.float64 else 0))
        return _t(synth.normal(tuple(x.shape), seed, dtype=np.float64)).to(x.dtype)
    torch.randn_like = randn_like
    try:
        with torch.no_grad():
            pred = net.ddim_sample(inp["init_disp"], inp["init_disp"], None, IGEV_TRACE["iters"], [], [], geo_fn, inp["used"],
                                   asd, None)
    finally:
        torch.randn_like = o_randn_like
    out["igev.pred"] = pred.numpy()
    for key, lst in rec.items():
        for i, v in enumerate(lst):
            out[f"igev.{key}.{i}"] = v.numpy()
    out["igev.randn_like_seeds"] = np.array(seeds, dtype=np.int64)


def gen_kitti15():
    import torch
    sys.path.insert(0, str(REF / "KITTI15"))
    torch.Tensor.cuda = lambda self, *a, **k: self
    sub = _load(REF / "KITTI15" / "core" / "submodule.py", "k15_submodule")
    from core.geometry import Combined_Geo_Encoding_Volume as GeoPlain
    from core.geometry_ddim import Combined_Geo_Encoding_Volume as GeoDdim
    head = _load(REF / "KITTI15" / "core" / "head.py", "k15_head")
    out = {}
    _volume_goldens(sub, out, "k15", mask_left_expected=False)
    cost = synth.normal((2, 48, 6, 40), 33) * np.float32(3)
    out["k15.regress.keepdim"] = sub.disparity_regression(torch.softmax(_t(cost), dim=1), 48).numpy()

    # geometry: [B, C, h, w] feature maps, geo volume [B, 8, 48, h, w]
    B, Cf, h, w, Cg, D = 2, 16, 6, 40, 8, 48
    f1, f2 = synth.normal((B, Cf, h, w), 101), synth.normal((B, Cf, h, w), 102)
    geo = synth.normal((B, Cg, D, h, w), 103)
    disp = synth.uniform((B, 1, h, w), 104, dtype=np.float32) * np.float32(50) - np.float32(2)   # some taps out of range
    coords = np.broadcast_to(np.arange(w, dtype=np.float32).reshape(1, 1, 1, w), (B, 1, h, w)).copy()
    noisy = synth.uniform((B, D, h, w), 105, dtype=np.float32)
    out["k15.corr"] = GeoDdim.corr(_t(f1), _t(f2)).numpy()
    gp = GeoPlain(_t(f1), _t(f2), _t(geo), num_levels=2, radius=4)
    out["k15.geo.plain"] = gp(_t(disp), _t(coords)).numpy()
    gd = GeoDdim(_t(f1), _t(f2), _t(geo), num_levels=2, radius=4)
    out["k15.geo.ddim"] = gd(_t(disp), _t(coords), _t(noisy)).numpy()
    out["k15.geo.pyr1"] = gd.geo_volume_pyramid[1].numpy()
    out["k15.corr.pyr1"] = gd.init_corr_pyramid[1].numpy()
    # IGEV DynamicHead: [B,180] shift linearly interpolated to D (core/head.py:74-83)
    torch.manual_seed(0)
    dh = head.DynamicHead(d_model=180) if "d_model" in head.DynamicHead.__init__.__code__.co_varnames else head.DynamicHead(180)
    with torch.no_grad():
        tc = torch.full((B,), 499, dtype=torch.long)
        z = torch.zeros(B, D, 1, 1)
        out["k15.head.shift48"] = dh(z, tc).reshape(B, D).numpy()
        out["k15.head.raw180"] = dh.block_time_mlp(dh.time_mlp(tc)).numpy()
    _igev_sampler_trace(out, GeoDdim)
    np.savez_compressed(HERE / "kitti15.npz", **out)
    print("kitti15:", len(out), "arrays")


def gen_f4():
    """context_upsample (KITTI15/core/submodule.py:241-253) with its autograd gradients, and the
    ACVNet patch convolutions (SceneFlow/models/acv_ddim.py:181-188,377-381: depth-wise (1,3,3) Conv3d chain, dilations
    1 then 1/2/3, then the channel concat) run through torch's own nn.Conv3d with seeded weights."""
    import torch
    sub = _load(REF / "KITTI15" / "core" / "submodule.py", "k15_submodule_f4")
    out = {}
    for name, (B, h, w) in {"a": (2, 6, 10), "b": (1, 24, 78)}.items():
        low = synth.uniform((B, 1, h, w), 301, dtype=np.float32) * np.float32(190)
        wts = synth.normal((B, 9, 4 * h, 4 * w), 302)
        wts = np.exp(wts) / np.exp(wts).sum(1, keepdims=True)            # softmax over the 9 taps, as spx_pred is
        lt, wt = _t(low).requires_grad_(True), _t(wts.astype(np.float32)).requires_grad_(True)
        res = sub.context_upsample(lt, wt)
        out[f"f4.ctxup.{name}"] = res.detach().numpy()
        gout = synth.normal((B, 4 * h, 4 * w), 303)
        res.backward(_t(gout))
        out[f"f4.ctxup.{name}.glow"] = lt.grad.numpy()
        out[f"f4.ctxup.{name}.gw"] = _sample(wt.grad.numpy()) if name == "b" else wt.grad.numpy()
    # ACV patch chain: the module definitions of acv_ddim.py:181-188 with seeded weights
    torch.manual_seed(0)
    import torch.nn as nn
    patch = nn.Conv3d(40, 40, kernel_size=(1, 3, 3), stride=1, dilation=1, groups=40, padding=(0, 1, 1), bias=False)
    l1 = nn.Conv3d(8, 8, kernel_size=(1, 3, 3), stride=1, dilation=1, groups=8, padding=(0, 1, 1), bias=False)
    l2 = nn.Conv3d(16, 16, kernel_size=(1, 3, 3), stride=1, dilation=2, groups=16, padding=(0, 2, 2), bias=False)
    l3 = nn.Conv3d(16, 16, kernel_size=(1, 3, 3), stride=1, dilation=3, groups=16, padding=(0, 3, 3), bias=False)
    for name, (B, D, H, W) in {"a": (1, 3, 9, 14), "b": (1, 4, 27, 60)}.items():
        vol = synth.normal((B, 40, D, H, W), 311)
        with torch.no_grad():
            g = patch(_t(vol))                                           # acv_ddim.py:377
            res = torch.cat((l1(g[:, :8]), l2(g[:, 8:24]), l3(g[:, 24:40])), dim=1)   # :378-381
        out[f"f4.patch.{name}.first"] = g.numpy() if name == "a" else _sample(g.numpy())
        out[f"f4.patch.{name}"] = res.numpy() if name == "a" else _sample(res.numpy())
    out["f4.patch.w_patch"] = patch.weight.detach().numpy().reshape(40, 9)
    out["f4.patch.w_l"] = np.concatenate([m.weight.detach().numpy().reshape(-1, 9) for m in (l1, l2, l3)], 0)
    np.savez_compressed(HERE / "f4.npz", **out)
    print("f4:", len(out), "arrays")


def gen_tier3():
    """Forward-level fixtures: the reference's UNMODIFIED ACVNet_DDIM.forward (eval branch with its own ddim_sample /
    model_predictions, and training branch with autograd) and ACVNet.forward, run on the real classes whose out-of-scope
    sub-networks were swapped for the light stand-ins of tests/acv_standin.py (same seeded weights on the GPU box)."""
    import warnings

    import torch
    sys.path.insert(0, str(REF / "SceneFlow"))
    torch.Tensor.cuda = lambda self, *a, **k: self
    warnings.simplefilter("ignore")
    import models.acv as acv
    import models.acv_ddim as acv_ddim
    from acv_standin import SeededNoise, graft, t3_inputs
    out = {}
    left, right, used, disp_q, mask_gt = (_t(a) for a in t3_inputs())

    def draws(rng):
        return np.array(["|".join(map(str, d)) for d in rng.log])

    net = graft(acv_ddim.ACVNet_DDIM(192, False, False)).eval()
    with SeededNoise() as rng, torch.no_grad():
        out["t3.eval.pred"] = net(left, right, used, disp_q, None)[0].numpy()
    out["t3.eval.draws"] = draws(rng)
    with SeededNoise() as rng, torch.no_grad():
        out["t3.eval_mask.pred"] = net(left, right, used, disp_q, mask_gt)[0].numpy()

    def train_pass(model, args, tag):
        model.train()
        model.zero_grad()
        with SeededNoise() as rng:
            preds = model(*args)
        loss = 0
        for i, p in enumerate(preds):
            out[f"t3.{tag}.pred{i}"] = p.detach().numpy()
            loss = loss + (p * _t(synth.normal(tuple(p.shape), 9700 + i))).sum() / p.numel()
        loss.backward()
        out[f"t3.{tag}.draws"] = draws(rng)
        for name, p in model.named_parameters():
            if p.grad is not None:
                out[f"t3.{tag}.grad.{name}"] = p.grad.numpy().copy()
        out[f"t3.{tag}.n_preds"] = np.array(len(preds))

    train_pass(net, (left, right, None, disp_q, None), "train")
    train_pass(net, (left, right, None, disp_q, mask_gt), "train_mask")
    # ACVNet (acv.py:167-247): eval, training, frozen-attention training, attention-only training
    for tag, (attn_only, freeze) in {"acv": (False, False), "acv_freeze": (False, True), "acv_attn": (True, False)}.items():
        net2 = graft(acv.ACVNet(192, attn_only, freeze)).eval()
        with torch.no_grad():
            out[f"t3.{tag}.eval.pred"] = net2(left, right)[0].numpy()
        train_pass(net2, (left, right), f"{tag}.train")
    np.savez_compressed(HERE / "tier3.npz", **out)
    print("tier3:", len(out), "arrays")


if __name__ == "__main__":
    which = sys.argv[1:] or ["sceneflow", "kitti12", "kitti15", "f4", "tier3"]
    if len(which) == 1 and os.environ.get("DV_GOLDEN_CHILD") == "1":
        {"sceneflow": gen_sceneflow, "kitti12": gen_kitti12, "kitti15": gen_kitti15, "f4": gen_f4,
         "tier3": gen_tier3}[which[0]]()
    else:
        for name in which:
            env = dict(os.environ, DV_GOLDEN_CHILD="1")
            subprocess.run([sys.executable, str(Path(__file__).resolve()), name], check=True, env=env)
        for f in sorted(HERE.glob("*.npz")):
            print(f.name, f.stat().st_size // 1024, "KiB")
