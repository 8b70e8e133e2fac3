"""CPU oracle for the DiffuVolume cost-volume hot path — TEST INFRASTRUCTURE, NOT PRODUCT.

A plain-numpy restatement of the reference's algorithm (iSEE-Laboratory/DiffuVolume; file:line
citations are relative to the reference root).  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` leg may import this module, and only as the
checker or the timed CPU baseline; nothing under `diffuvolume_b200/` imports it.

Pinning: the reference ships no tests or golden vectors for this path.
The oracle is therefore pinned against outputs of the reference itself: `tests/golden/make_golden.py`
imports the reference's own modules from /root/reference (possible only in the authoring
container), runs them on seeded inputs and commits the results as small fixtures;
`tests/test_oracle_golden.py` checks every function here against those fixtures.

dtype discipline: the reference mixes float32 tensors with float64 schedule buffers; every
promotion PyTorch performs is spelled out here with explicit numpy casts (numpy's own scalar
promotion rules are never relied on).
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

f32 = np.float32
f64 = np.float64


# =============================================================================================
# a1-a5: volumes
# =============================================================================================
def groupwise_correlation(fea1: np.ndarray, fea2: np.ndarray, num_groups: int) -> np.ndarray:
    """SceneFlow/models/submodule.py:209-215: (fea1*fea2).view(B,G,cpg,H,W).mean(2)."""
    B, C, H, W = fea1.shape
    assert C % num_groups == 0
    cpg = C // num_groups
    prod = (fea1 * fea2).reshape(B, num_groups, cpg, H, W)
    cost = prod.mean(axis=2, dtype=fea1.dtype)
    assert cost.shape == (B, num_groups, H, W)
    return cost


def build_gwc_volume(ref: np.ndarray, tgt: np.ndarray, maxdisp: int, num_groups: int) -> np.ndarray:
    """SceneFlow/models/submodule.py:228-238 (= KITTI12 :109-119, KITTI15 :159-169)."""
    B, C, H, W = ref.shape
    vol = np.zeros((B, num_groups, maxdisp, H, W), dtype=ref.dtype)
    for i in range(maxdisp):
        if i > 0:
            if i < W:  # empty slices for i >= W: the plane stays zero
                vol[:, :, i, :, i:] = groupwise_correlation(ref[:, :, :, i:], tgt[:, :, :, :-i], num_groups)
        else:
            vol[:, :, i, :, :] = groupwise_correlation(ref, tgt, num_groups)
    return vol


def build_concat_volume(ref: np.ndarray, tgt: np.ndarray, maxdisp: int, mask_left: bool) -> np.ndarray:
    """Variant M (mask_left=False): SceneFlow/models/submodule.py:180-191, KITTI15/core/submodule.py:206-217.
    Variant T (mask_left=True): SceneFlow/submodule.py:137-148, KITTI12/models/submodule.py:86-97."""
    B, C, H, W = ref.shape
    vol = np.zeros((B, 2 * C, maxdisp, H, W), dtype=ref.dtype)
    for i in range(maxdisp):
        if i > 0:
            if mask_left:
                vol[:, :C, i, :, i:] = ref[:, :, :, i:]
            else:
                vol[:, :C, i, :, :] = ref
            if i < W:
                vol[:, C:, i, :, i:] = tgt[:, :, :, :-i]
        else:
            vol[:, :C, i] = ref
            vol[:, C:, i] = tgt
    return vol


def build_corrleation_volume(ref: np.ndarray, tgt: np.ndarray, maxdisp: int, num_groups: int) -> np.ndarray:
    """KITTI12/models/submodule.py:121-135 (= SceneFlow/submodule.py:172-186), including the
    negative-shift quirk: `[..., :-i]` with i < 0 selects the FIRST -i columns."""
    B, C, H, W = ref.shape
    vol = np.zeros((B, num_groups, 2 * maxdisp + 1, H, W), dtype=ref.dtype)
    for i in range(-maxdisp, maxdisp + 1):
        if i > 0:
            if i < W:
                vol[:, :, i + maxdisp, :, i:] = groupwise_correlation(ref[:, :, :, i:], tgt[:, :, :, :-i], num_groups)
        elif i < 0:
            vol[:, :, i + maxdisp, :, :-i] = groupwise_correlation(ref[:, :, :, :-i], tgt[:, :, :, i:], num_groups)
        else:
            vol[:, :, i + maxdisp] = groupwise_correlation(ref, tgt, num_groups)
    return vol


def softmax(x: np.ndarray, axis: int) -> np.ndarray:
    """F.softmax: exp(x - max) / sum, in the dtype of x."""
    m = x.max(axis=axis, keepdims=True)
    e = np.exp(x - m, dtype=x.dtype)
    return e / e.sum(axis=axis, keepdims=True, dtype=x.dtype)


def acv_attention_volume(att_logits: np.ndarray, concat: np.ndarray) -> np.ndarray:
    """SceneFlow/models/acv_ddim.py:390 (acv.py:203): F.softmax(att_weights, dim=2) * concat_volume."""
    return softmax(att_logits, axis=2) * concat


# =============================================================================================
# a6: regression
# =============================================================================================
def disparity_regression(x: np.ndarray, maxdisp: int, keepdim: bool = False) -> np.ndarray:
    """SceneFlow/models/submodule.py:173-177 (keepdim=True: KITTI15/core/submodule.py:219-223)."""
    assert x.ndim == 4
    disp_values = np.arange(0, maxdisp, dtype=x.dtype).reshape(1, maxdisp, 1, 1)
    return np.sum(x * disp_values, axis=1, keepdims=keepdim, dtype=x.dtype)


def softmax_regress(cost: np.ndarray, maxdisp: Optional[int] = None) -> Tuple[np.ndarray, np.ndarray]:
    """F.softmax(cost, dim=1) then disparity_regression (acv_ddim.py:269-270). Returns (disp, prob)."""
    prob = softmax(cost, axis=1)
    return disparity_regression(prob, cost.shape[1] if maxdisp is None else maxdisp), prob


def uncertainty(disp: np.ndarray, prob: np.ndarray) -> np.ndarray:
    """SceneFlow/models/acv_ddim.py:324-329: sum_d |disp - d| * p[d]."""
    D = prob.shape[1]
    disp_values = np.arange(0, D, dtype=disp.dtype).reshape(1, D, 1, 1)
    difference = np.abs(disp[:, None] - disp_values)
    return np.sum(difference * prob, axis=1, dtype=disp.dtype)


def renewal_vote(disp: np.ndarray, used: np.ndarray, unc: Optional[np.ndarray], thr_dif: float,
                 thr_unc: float) -> np.ndarray:
    """acv_ddim.py:321-331 (pwcnet_ddim.py:560-570): (|disp-used| < thr_dif) * (unc < thr_unc), as float32."""
    m = np.abs(disp - used) < f32(thr_dif)
    if unc is not None:
        m = m & (unc < f32(thr_unc))
    return m.astype(f32)


def warp(x: np.ndarray, disp: np.ndarray) -> np.ndarray:
    """KITTI12/models/submodule.py:137-176 (= SceneFlow/submodule.py:188-227): warp the right features to the left view.
    x [B,C,H,W], disp [B,1,H,W].  The reference normalises the sampling grid as 2*v/(size-1)-1 (align_corners=True
    convention) but calls F.grid_sample with its default align_corners=False, so the sampled coordinate is
    ix = (x - disp) * W/(W-1) - 0.5, iy = y * H/(H-1) - 0.5 — reproduced here, zero padding, bilinear; the validity mask is
    the same sampling of a ones tensor, thresholded (< 0.999 -> 0, else 1)."""
    B, C, H, W = x.shape
    xx = np.broadcast_to(np.arange(W, dtype=f32).reshape(1, 1, W), (B, H, W))
    yy = np.broadcast_to(np.arange(H, dtype=f32).reshape(1, H, 1), (B, H, W))
    gx = (f32(2.0) * (xx - disp[:, 0]).astype(f32) / f32(max(W - 1, 1)) - f32(1.0)).astype(f32)
    gy = (f32(2.0) * yy / f32(max(H - 1, 1)) - f32(1.0)).astype(f32)
    ix = (((gx + f32(1)) * f32(W) - f32(1)) / f32(2)).astype(f32)
    iy = (((gy + f32(1)) * f32(H) - f32(1)) / f32(2)).astype(f32)
    x0, y0 = np.floor(ix), np.floor(iy)
    x1, y1 = x0 + 1, y0 + 1
    wts = [((x1 - ix) * (y1 - iy), x0, y0), ((ix - x0) * (y1 - iy), x1, y0),
           ((x1 - ix) * (iy - y0), x0, y1), ((ix - x0) * (iy - y0), x1, y1)]
    out = np.zeros_like(x, dtype=f32)
    mask = np.zeros((B, H, W), dtype=f32)
    bi = np.arange(B).reshape(B, 1, 1)
    for wgt, xi, yi in wts:
        ok = (xi >= 0) & (xi <= W - 1) & (yi >= 0) & (yi <= H - 1)
        xc, yc = np.clip(xi, 0, W - 1).astype(np.int64), np.clip(yi, 0, H - 1).astype(np.int64)
        wv = np.where(ok, wgt, f32(0)).astype(f32)
        vals = x[bi, :, yc, xc]                      # [B,H,W,C]
        out += (np.moveaxis(vals, -1, 1) * wv[:, None]).astype(f32)
        mask += wv
    mask = np.where(mask < f32(0.999), f32(0), f32(1)).astype(f32)
    return (out * mask[:, None]).astype(f32)


# =============================================================================================
# diffusion schedule (fp64) — a7, a8, a12 coefficients
# =============================================================================================
def cosine_beta_schedule(timesteps: int, s: float = 0.008) -> np.ndarray:
    """SceneFlow/models/acv_ddim.py:113-119 (float64)."""
    steps = timesteps + 1
    x = np.linspace(0, timesteps, steps, dtype=f64)
    ac = np.cos(((x / timesteps) + s) / (1 + s) * math.pi * 0.5) ** 2
    ac = ac / ac[0]
    betas = 1 - (ac[1:] / ac[:-1])
    return np.clip(betas, 0, 0.999)


class Schedule:
    """The registered float64 buffers of ACVNet_DDIM.__init__ (acv_ddim.py:131-157)."""

    def __init__(self, timesteps: int = 1000, sampling_timesteps: int = 5, eta: float = 1.0, scale: float = 1.0):
        betas = cosine_beta_schedule(timesteps)
        alphas = 1.0 - betas
        self.betas = betas
        self.alphas_cumprod = np.cumprod(alphas, axis=0)
        self.num_timesteps = int(betas.shape[0])
        self.sampling_timesteps = sampling_timesteps
        self.eta = eta
        self.scale = scale
        ac = self.alphas_cumprod
        self.sqrt_alphas_cumprod = np.sqrt(ac)
        self.sqrt_one_minus_alphas_cumprod = np.sqrt(1.0 - ac)
        self.sqrt_recip_alphas_cumprod = np.sqrt(1.0 / ac)
        self.sqrt_recipm1_alphas_cumprod = np.sqrt(1.0 / ac - 1)

    def time_pairs(self) -> List[Tuple[int, int]]:
        """acv_ddim.py:306-308: linspace(-1, T-1, S+1) (float32) -> int -> reversed -> pairs."""
        times = np.linspace(-1, self.num_timesteps - 1, self.sampling_timesteps + 1, dtype=f32)
        # torch.linspace (float32) computes start + step*i for the first half and end - step*(n-1-i)
        # for the second half; reproduce that so the int() truncation matches for any S.
        n = self.sampling_timesteps + 1
        start, end = f32(-1), f32(self.num_timesteps - 1)
        step = (end - start) / f32(n - 1)
        vals = []
        for i in range(n):
            if i < n // 2:
                vals.append(f32(start + step * f32(i)))
            else:
                vals.append(f32(end - step * f32(n - 1 - i)))
        times = [int(v) for v in vals]
        times = list(reversed(times))
        return list(zip(times[:-1], times[1:]))

    def ddim_coefficients(self, time: int, time_next: int) -> Tuple[float, float, float]:
        """acv_ddim.py:347-351: (sqrt(alpha_next), c, sigma) in float64."""
        alpha = self.alphas_cumprod[time]
        alpha_next = self.alphas_cumprod[time_next]
        sigma = self.eta * np.sqrt((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha))
        c = np.sqrt(1 - alpha_next - sigma ** 2)
        return float(np.sqrt(alpha_next)), float(c), float(sigma)


def q_sample(sched: Schedule, x_start: np.ndarray, t: int, noise: np.ndarray) -> np.ndarray:
    """acv_ddim.py:241-246: the extracted coefficients are float64 [1,1,1,1] tensors -> float64 result."""
    return sched.sqrt_alphas_cumprod[t] * x_start.astype(f64) + sched.sqrt_one_minus_alphas_cumprod[t] * noise.astype(f64)


def predict_noise_from_start(sched: Schedule, x_t: np.ndarray, t: int, x0: np.ndarray) -> np.ndarray:
    """acv_ddim.py:248-252 (float64)."""
    return (sched.sqrt_recip_alphas_cumprod[t] * x_t.astype(f64) - x0.astype(f64)) / sched.sqrt_recipm1_alphas_cumprod[t]


# =============================================================================================
# a9: filter
# =============================================================================================
def filter_factor(xt: np.ndarray, shift: Optional[np.ndarray], scale: float = 1.0) -> np.ndarray:
    """acv_ddim.py:256-258 with DynamicHead's `noisy + scale_shift` (head.py:74-77): dtype of xt."""
    dt = xt.dtype
    v = xt if shift is None else xt + shift.astype(dt)[:, :, None, None]
    s = dt.type(scale)
    v = np.clip(v, -s, s)
    return ((v / s) + dt.type(1)) / dt.type(2)


def volume_filter(vol: np.ndarray, xt: np.ndarray, shift: Optional[np.ndarray], scale: float = 1.0) -> np.ndarray:
    """acv_ddim.py:260: volume * noise.unsqueeze(1).float()."""
    return vol * filter_factor(xt, shift, scale).astype(f32)[:, None]


# =============================================================================================
# a10 / a11: down-sampling, 2-tap x_start, renewal mask
# =============================================================================================
def interpolate_bilinear(x: np.ndarray, size: Tuple[int, int]) -> np.ndarray:
    """F.interpolate(x, size=size, mode='bilinear') (align_corners=False) for [..., H, W] float32 maps
    (ATen area_pixel_compute_source_index / guard_index_and_lambda; blend along x, then y)."""
    H, W = x.shape[-2:]
    h, w = size

    def taps(n_in, n_out):
        scale = f32(n_in) / f32(n_out)
        src = scale * (np.arange(n_out, dtype=f32) + f32(0.5)) - f32(0.5)
        src = np.maximum(src, f32(0))
        i0 = np.minimum(np.floor(src).astype(np.int64), n_in - 1)
        i1 = np.minimum(i0 + 1, n_in - 1)
        l1 = np.clip(src - i0.astype(f32), f32(0), f32(1)).astype(f32)
        return i0, i1, (f32(1) - l1).astype(f32), l1

    y0, y1, ly0, ly1 = taps(H, h)
    x0, x1, lx0, lx1 = taps(W, w)
    x = x.astype(f32, copy=False)
    top = x[..., y0, :]
    bot = x[..., y1, :]
    r0 = lx0 * top[..., x0] + lx1 * top[..., x1]
    r1 = lx0 * bot[..., x0] + lx1 * bot[..., x1]
    return (ly0[:, None] * r0 + ly1[:, None] * r1).astype(f32)


def _linear_taps(n_in: int, n_out: int, align_corners: bool):
    if align_corners:
        scale = f32(n_in - 1) / f32(n_out - 1) if n_out > 1 else f32(0)
        src = scale * np.arange(n_out, dtype=f32)
    else:
        scale = f32(n_in) / f32(n_out)
        src = np.maximum(scale * (np.arange(n_out, dtype=f32) + f32(0.5)) - f32(0.5), f32(0))
    i0 = np.minimum(np.floor(src).astype(np.int64), n_in - 1)
    i1 = np.minimum(i0 + 1, n_in - 1)
    l1 = np.clip(src - i0.astype(f32), f32(0), f32(1)).astype(f32)
    return i0, i1, (f32(1) - l1).astype(f32), l1


def interpolate_trilinear(x: np.ndarray, size: Tuple[int, int, int], align_corners: bool = False) -> np.ndarray:
    """F.upsample(x, size, mode='trilinear') for [B,C,D,H,W] float32 (acv_ddim.py:267 align_corners=False;
    pwcnet_ddim.py:480 align_corners=True).  Separable blend W, then H, then D (ATen blends the 8 taps
    in one expression; the results agree to float32 rounding)."""
    x = x.astype(f32, copy=False)
    for axis, n_out in zip((4, 3, 2), (size[2], size[1], size[0])):
        i0, i1, l0, l1 = _linear_taps(x.shape[axis], n_out, align_corners)
        shp = [1] * 5
        shp[axis] = n_out
        x = (np.take(x, i0, axis=axis) * l0.reshape(shp) + np.take(x, i1, axis=axis) * l1.reshape(shp)).astype(f32)
    return x


def disp_to_quarter(pred: np.ndarray, maxdisp: int = 192) -> np.ndarray:
    """acv_ddim.py:272-274: clamp(pred, 0, maxdisp-1) -> bilinear to (H//4, W//4) -> /4.  pred [B,H,W]."""
    B, H, W = pred.shape
    d = np.clip(pred, f32(0), f32(maxdisp - 1))
    return interpolate_bilinear(d, (H // 4, W // 4)) / f32(4)


def xstart_from_disp(disp_q: np.ndarray, D: int = 48, scale: float = 1.0) -> np.ndarray:
    """acv_ddim.py:277-292: floor / two scatter_ / where(real == D-1, one_hot) / scale*(2x-1) / clamp.
    disp_q [B,h,w] float32 -> [B,D,h,w] float32."""
    B, h, w = disp_q.shape
    real = np.floor(disp_q).astype(np.int64)
    mask_num = real == D - 1
    coff = (real.astype(f32) - disp_q + f32(1)).astype(f32)
    vol = np.zeros((B, D, h, w), dtype=f32)
    bi, yi, xi = np.meshgrid(np.arange(B), np.arange(h), np.arange(w), indexing="ij")
    vol[bi, real, yi, xi] = coff
    vol[bi, np.clip(real + 1, 0, D - 1), yi, xi] = (f32(1) - coff).astype(f32)
    fuzhi = np.zeros((B, D, h, w), dtype=f32)
    fuzhi[:, -1] = 1
    x0 = np.where(mask_num[:, None], fuzhi, vol)
    x0 = f32(scale) * (x0 * f32(2) - f32(1))
    return np.clip(x0, f32(-scale), f32(scale)).astype(f32)


def update_mask(mask: np.ndarray, vote_full: np.ndarray) -> np.ndarray:
    """acv_ddim.py:333-338: mask = clamp(mask + bilinear_down(vote), 0, 1); mask [B,h,w], vote [B,H,W]."""
    h, w = mask.shape[-2:]
    return np.clip(mask + interpolate_bilinear(vote_full.astype(f32), (h, w)), f32(0), f32(1)).astype(f32)


# =============================================================================================
# a12: DDIM update
# =============================================================================================
def ddim_update(sched: Schedule, x0: np.ndarray, eps: np.ndarray, step_noise: np.ndarray, time: int,
                time_next: int) -> np.ndarray:
    """acv_ddim.py:344-356: img = x0*sqrt(alpha_next) + c*eps + sigma*noise.
    PyTorch dtype rules: x0 (fp32) times a 0-dim fp64 tensor stays fp32 (the scalar is cast to fp32);
    c*eps is fp64; sigma*noise has the dtype of noise; the sums promote to fp64 left to right."""
    san, c, sigma = sched.ddim_coefficients(time, time_next)
    t1 = (x0.astype(f32) * f32(san)).astype(f32)
    t2 = f64(c) * eps.astype(f64)
    if step_noise.dtype == f32:
        t3 = (f32(sigma) * step_noise).astype(f32).astype(f64)
    else:
        t3 = f64(sigma) * step_noise
    return (t1.astype(f64) + t2) + t3


def ensemble(maps: Sequence[np.ndarray], cof: Sequence[float]) -> np.ndarray:
    """acv_ddim.py:365-369: sum(cat(maps) * cof, dim=0) in float32."""
    acc = np.zeros_like(maps[0], dtype=f32)
    for m, c in zip(maps, cof):
        acc = acc + m.astype(f32) * f32(c)
    return acc


# =============================================================================================
# full ACV-style sampler trace (a7-a13) with an injected stand-in for the conv stack
# =============================================================================================
def ddim_sample_acv(sched: Schedule, volume: np.ndarray, used: np.ndarray, asd: np.ndarray,
                    shift_fn: Callable[[int], np.ndarray], cost_fn: Callable[[np.ndarray, int], np.ndarray],
                    step_noises: Sequence[np.ndarray], renoises: Sequence[np.ndarray],
                    cof: Sequence[float] = (0.5, 0.0, 0.0, 0.0, 0.2, 0.3), maxdisp: int = 192,
                    thr_dif: float = 1.0, thr_unc: float = 3.0, trace: Optional[dict] = None):
    """ACVNet_DDIM.ddim_sample (acv_ddim.py:298-370) + model_predictions (:254-296).

    `cost_fn(filtered_volume, step)` stands in for dres0..classif2 + trilinear upsample and must return
    the [B,maxdisp,H,W] logits; `shift_fn(t)` returns DynamicHead's [B,D] shift for time t.
    step_noises[i] = randn_like(img) of step i; renoises[i] = rand_like(...) (fp64 uniform) of step i.
    Returns (final_prediction [B,H,W], list of per-step disparities)."""
    B, C, D, h, w = volume.shape
    img = asd
    final = [used.astype(f32)]
    mask = np.zeros((B, h, w), dtype=f32)
    pairs = sched.time_pairs()
    for i, (time, time_next) in enumerate(pairs):
        shift = shift_fn(time)
        n = filter_factor(img, shift, sched.scale)
        vol_f = volume * n.astype(f32)[:, None]
        cost = cost_fn(vol_f, i)
        disp, prob = softmax_regress(cost, maxdisp)
        dq = disp_to_quarter(disp, maxdisp)
        x0 = xstart_from_disp(dq, D, sched.scale)
        eps = predict_noise_from_start(sched, n, time, x0)
        final.append(disp)
        unc = uncertainty(disp, prob)
        vote = renewal_vote(disp, used, unc, thr_dif, thr_unc)
        mask = update_mask(mask, vote)
        if trace is not None:
            trace.setdefault("disp", []).append(disp)
            trace.setdefault("x0", []).append(x0)
            trace.setdefault("eps", []).append(eps)
            trace.setdefault("mask", []).append(mask.copy())
            trace.setdefault("unc", []).append(unc)
        if time_next < 0:
            img = x0
            continue
        img = ddim_update(sched, x0, eps, step_noises[i], time, time_next)
        img = np.where(mask[:, None] == 0, renoises[i].astype(f64), img)
        if trace is not None:
            trace.setdefault("img", []).append(img)
    return ensemble(final, cof), final


# =============================================================================================
# a14 / a15: IGEV combined geometry encoding volume
# =============================================================================================
def corr1d_allpairs(fmap1: np.ndarray, fmap2: np.ndarray) -> np.ndarray:
    """KITTI15/core/geometry_ddim.py:72-80: einsum('aijk,aijh->ajkh') -> [B,H,W1,1,W2]."""
    B, D, H, W1 = fmap1.shape
    W2 = fmap2.shape[3]
    corr = np.einsum("aijk,aijh->ajkh", fmap1, fmap2, optimize=True).astype(fmap1.dtype)
    return corr.reshape(B, H, W1, 1, W2)


def avg_pool_w2(x: np.ndarray) -> np.ndarray:
    """F.avg_pool2d(x, [1,2], stride=[1,2]) on the last axis (floor)."""
    L = x.shape[-1] // 2
    return ((x[..., 0:2 * L:2] + x[..., 1:2 * L:2]) / x.dtype.type(2)).astype(x.dtype)


def bilinear_sampler_1d(img: np.ndarray, xcoord: np.ndarray) -> np.ndarray:
    """KITTI15/core/utils/utils.py:59-77 with H == 1: img [N,C,1,W], xcoord [N,T] pixel coordinates ->
    [N,C,T]; grid_sample bilinear, zero padding, align_corners=True."""
    N, C, _, W = img.shape
    xg = f32(2) * xcoord.astype(f32) / f32(W - 1) - f32(1)
    ix = ((xg + f32(1)) / f32(2)) * f32(W - 1)
    i0 = np.floor(ix)
    w0 = (i0 + f32(1) - ix).astype(f32)
    w1 = (ix - i0).astype(f32)
    i0 = i0.astype(np.int64)
    i1 = i0 + 1
    rows = img[:, :, 0, :]
    out = np.zeros((N, C, xcoord.shape[1]), dtype=f32)
    for idx, wgt in ((i0, w0), (i1, w1)):
        ok = (idx >= 0) & (idx < W)
        g = np.take_along_axis(rows, np.clip(idx, 0, W - 1)[:, None, :].repeat(C, axis=1), axis=2)
        out = out + np.where(ok[:, None, :], g * wgt[:, None, :], f32(0))
    return out.astype(f32)


class CombinedGeoEncodingVolume:
    """KITTI15/core/geometry_ddim.py:6-69 (and geometry.py:6-58 when `noisy` is None)."""

    def __init__(self, init_fmap1, init_fmap2, geo_volume, num_levels=2, radius=4):
        self.num_levels, self.radius = num_levels, radius
        init_corr = corr1d_allpairs(init_fmap1, init_fmap2)
        b, h, w, _, w2 = init_corr.shape
        b, c, d, h, w = geo_volume.shape
        self.channel = c
        geo = np.ascontiguousarray(geo_volume.transpose(0, 3, 4, 1, 2)).reshape(b * h * w, c, 1, d)
        corr = init_corr.reshape(b * h * w, 1, 1, w2)
        self.geo_volume_pyramid = [geo]
        self.init_corr_pyramid = [corr]
        for _ in range(num_levels - 1):
            geo = avg_pool_w2(geo)
            self.geo_volume_pyramid.append(geo)
        for _ in range(num_levels - 1):
            corr = avg_pool_w2(corr)
            self.init_corr_pyramid.append(corr)

    def __call__(self, disp, coords, noisy=None):
        r = self.radius
        b, _, h, w = disp.shape
        N = b * h * w
        noise = None
        if noisy is not None:
            # geometry_ddim.py:37: a raw reshape of the [B,D,h,w] buffer, no permute
            nz = np.ascontiguousarray(noisy).reshape(N, 1, 1, -1)
            noise = [nz]
            for _ in range(self.num_levels):
                nz = avg_pool_w2(nz)
                noise.append(nz)
        dx = np.linspace(-r, r, 2 * r + 1, dtype=f32).reshape(1, 2 * r + 1)
        dflat = disp.reshape(N, 1).astype(f32)
        cflat = coords.reshape(N, 1).astype(f32)
        out = []
        for i in range(self.num_levels):
            geo = self.geo_volume_pyramid[i]
            x0 = dx + dflat / f32(2 ** i)
            if noise is not None:
                geo = geo * noise[i]
            out.append(bilinear_sampler_1d(geo, x0).reshape(b, h, w, -1))
            init_x0 = cflat / f32(2 ** i) - dflat / f32(2 ** i) + dx
            out.append(bilinear_sampler_1d(self.init_corr_pyramid[i], init_x0).reshape(b, h, w, -1))
        res = np.concatenate(out, axis=-1)
        return np.ascontiguousarray(res.transpose(0, 3, 1, 2)).astype(f32)


# ------------------------------------------------------------------------------------------------
# f4: context_upsample  (KITTI15/core/submodule.py:241-253)
# ------------------------------------------------------------------------------------------------
def context_upsample(disp_low: np.ndarray, up_weights: np.ndarray) -> np.ndarray:
    """F.unfold(disp_low, 3, 1, 1) -> nearest x4 -> (* up_weights).sum(1).  disp_low [B,1,h,w], up_weights [B,9,4h,4w]
    -> [B,4h,4w].  Tap k = ky*3 + kx reads disp_low[y+ky-1, x+kx-1] (zero padding), F.unfold's channel order."""
    b, c, h, w = disp_low.shape
    assert c == 1 and up_weights.shape == (b, 9, 4 * h, 4 * w)
    pad = np.zeros((b, h + 2, w + 2), dtype=f32)
    pad[:, 1:-1, 1:-1] = disp_low[:, 0]
    out = np.zeros((b, 4 * h, 4 * w), dtype=f32)
    for ky in range(3):
        for kx in range(3):
            tap = pad[:, ky:ky + h, kx:kx + w]                          # [B,h,w]
            up = np.repeat(np.repeat(tap, 4, axis=1), 4, axis=2)        # nearest x4
            out = (out + up * up_weights[:, ky * 3 + kx].astype(f32)).astype(f32)
    return out


# ------------------------------------------------------------------------------------------------
# f4: ACVNet patch convolutions  (SceneFlow/models/acv_ddim.py:181-188,377-381)
# ------------------------------------------------------------------------------------------------
def depthwise_conv3x3(vol: np.ndarray, w: np.ndarray, dilation: int) -> np.ndarray:
    """nn.Conv3d(C, C, (1,3,3), groups=C, dilation=dilation, padding=(0,dilation,dilation), bias=False) on
    vol [B,C,D,H,W]; w is [C,9] (= weight[C,1,1,3,3] flattened).  Cross-correlation, zero padding."""
    B, C, D, H, W = vol.shape
    p = dilation
    pad = np.zeros((B, C, D, H + 2 * p, W + 2 * p), dtype=f32)
    pad[..., p:p + H, p:p + W] = vol
    out = np.zeros_like(vol, dtype=f32)
    for ky in range(3):
        for kx in range(3):
            tap = pad[..., ky * p:ky * p + H, kx * p:kx * p + W]
            out = (out + tap * w[:, ky * 3 + kx].reshape(1, C, 1, 1, 1).astype(f32)).astype(f32)
    return out


def acv_patch_volume(gwc: np.ndarray, w_patch: np.ndarray, w_l: np.ndarray, splits=(8, 16, 16), dils=(1, 2, 3)) -> np.ndarray:
    """g = patch(gwc); cat(patch_l1(g[:, :8]), patch_l2(g[:, 8:24]), patch_l3(g[:, 24:40]))  (acv_ddim.py:377-381)."""
    g = depthwise_conv3x3(gwc, w_patch, 1)
    outs, c = [], 0
    for n, dl in zip(splits, dils):
        outs.append(depthwise_conv3x3(g[:, c:c + n], w_l[c:c + n], dl))
        c += n
    return np.concatenate(outs, axis=1)
