"""Torch-facing functional ops over the C-ABI library (include/dv_b200.h).

PyTorch is plumbing here: it owns device memory (outputs are allocated with `torch.empty` on
the input's device, so the caching allocator and stream-ordered reuse are respected) and the
current stream.  Every op launches hand-written sm_100a kernels through ctypes; nothing in
this module computes with ATen, and there is no CPU path: CPU tensors raise.

Error conventions follow the reference: the conditions the reference guards
with a bare `assert` raise AssertionError here too; anything else that the C layer rejects
raises DvLibraryError with the dv_status name.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import DdimStepArgs, DvLibraryError, check


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _need_cuda(*tensors: Optional[torch.Tensor]) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise DvLibraryError(
                "diffuvolume_b200 ops run on CUDA (sm_100a) tensors only; got a CPU tensor "
                "(there is deliberately no CPU fallback)"
            )
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"Expected all tensors to be on the same device, but found {dev} and {t.device}")
    assert dev is not None
    return dev


_tile_counter_cache: dict = {}


def _tile_counters(dev: torch.device, family: int) -> int:
    """Device address of the {next, done} tile-counter pair a persistent kernel of `family` (0 = streaming volume
    producer, 1 = softmax-regression) draws from (include/dv_b200.h, `tile_counters`).  One zeroed pair per (device,
    stream, family): launches on one stream are ordered, so they can share a pair; other streams get their own.  While a
    CUDA graph is being captured every call gets a FRESH zeroed pair from the graph's private pool (a memset node in the
    graph), so replays of different graphs never share counters with each other or with eager launches."""
    if torch.cuda.is_current_stream_capturing():
        t = torch.zeros(2, dtype=torch.int32, device=dev)
        _tile_counter_cache.setdefault(("graph", dev.index), []).append(t)    # owned by the graph pool; keep the handle
        return t.data_ptr()
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    t = _tile_counter_cache.get(key)
    if t is None:
        t = torch.zeros(16, dtype=torch.int32, device=dev)
        _tile_counter_cache[key] = t
    return t.data_ptr() + 32 * family


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise DvLibraryError(f"{name}: expected float32, got {t.dtype}")
    return t if t.is_contiguous() else t.contiguous()


def _is_f64(t: torch.Tensor, name: str) -> int:
    if t.dtype == torch.float64:
        return 1
    if t.dtype == torch.float32:
        return 0
    raise DvLibraryError(f"{name}: expected float32 or float64, got {t.dtype}")


# --------------------------------------------------------------------------------------------
# volumes
# --------------------------------------------------------------------------------------------
def groupwise_correlation(fea1: torch.Tensor, fea2: torch.Tensor, num_groups: int) -> torch.Tensor:
    """a1 — SceneFlow/models/submodule.py:209-215."""
    B, Cc, H, W = fea1.shape
    assert Cc % num_groups == 0
    _need_cuda(fea1, fea2)
    fea1, fea2 = _f32c(fea1, "fea1"), _f32c(fea2, "fea2")
    if fea2.shape != fea1.shape:
        raise RuntimeError(f"The size of tensor a {tuple(fea1.shape)} must match the size of tensor b {tuple(fea2.shape)}")
    out = torch.empty((B, num_groups, H, W), dtype=torch.float32, device=fea1.device)
    with torch.cuda.device(fea1.device):
        check(_lib.lib().dv_groupwise_correlation_f32(_ptr(fea1), _ptr(fea2), _ptr(out), B, Cc, H, W, num_groups,
                                                      _stream(fea1)), "dv_groupwise_correlation_f32")
    assert out.shape == (B, num_groups, H, W)
    return out


def gwc_volume(ref: torch.Tensor, tgt: torch.Tensor, maxdisp: int, num_groups: int,
               out: Optional[torch.Tensor] = None, out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """a2 — build_gwc_volume, SceneFlow/models/submodule.py:228-238 -> [B,G,maxdisp,H,W].
    out_dtype=torch.bfloat16: fp32 features and accumulation, one round-to-nearest-even at the store (half the write
    traffic; 8 or 12 channels per group, H*W % 4 == 0)."""
    B, Cc, H, W = ref.shape
    assert Cc % num_groups == 0
    _need_cuda(ref, tgt)
    ref, tgt = _f32c(ref, "refimg_fea"), _f32c(tgt, "targetimg_fea")
    if tgt.shape != ref.shape:
        raise RuntimeError(f"The size of tensor a {tuple(ref.shape)} must match the size of tensor b {tuple(tgt.shape)}")
    if out_dtype not in (torch.float32, torch.bfloat16):
        raise DvLibraryError(f"gwc_volume: out_dtype must be float32 or bfloat16, got {out_dtype}")
    if out is None:
        out = torch.empty((B, num_groups, maxdisp, H, W), dtype=out_dtype, device=ref.device)
    else:
        assert out.shape == (B, num_groups, maxdisp, H, W) and out.dtype == out_dtype and out.is_contiguous()
    if out.numel() == 0:
        return out
    if out_dtype == torch.bfloat16:
        with torch.cuda.device(ref.device):
            check(_lib.lib().dv_gwc_volume_bf16(_ptr(ref), _ptr(tgt), _ptr(out), B, Cc, H, W, maxdisp, num_groups,
                                                _stream(ref)), "dv_gwc_volume_bf16")
        return out
    with torch.cuda.device(ref.device):
        check(_lib.lib().dv_gwc_volume_f32(_ptr(ref), _ptr(tgt), _ptr(out), B, Cc, H, W, maxdisp, num_groups,
                                           _stream(ref)), "dv_gwc_volume_f32")
    return out


def concat_volume(ref: torch.Tensor, tgt: torch.Tensor, maxdisp: int, *, mask_left: bool,
                  att_logits: Optional[torch.Tensor] = None, xt: Optional[torch.Tensor] = None,
                  shift: Optional[torch.Tensor] = None, scale: float = 1.0,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """a3 (+a4 +a9) — build_concat_volume with the optional fused factors.

    mask_left=False: variant M (SceneFlow/models/submodule.py:180-191, KITTI15/core/submodule.py:206-217);
    mask_left=True : variant T (SceneFlow/submodule.py:137-148, KITTI12/models/submodule.py:86-97).
    att_logits [B,1,D,H,W]: multiply by softmax over D (acv_ddim.py:390).
    xt [B,D,H,W] (+ shift [B,D]): multiply by the DDIM filter factor (acv_ddim.py:254-260).
    """
    B, Cc, H, W = ref.shape
    _need_cuda(ref, tgt, att_logits, xt, shift)
    ref, tgt = _f32c(ref, "refimg_fea"), _f32c(tgt, "targetimg_fea")
    if tgt.shape != ref.shape:
        raise RuntimeError(f"The size of tensor a {tuple(ref.shape)} must match the size of tensor b {tuple(tgt.shape)}")
    if att_logits is not None:
        att_logits = _f32c(att_logits, "att_logits")
        if att_logits.numel() != B * maxdisp * H * W:
            raise RuntimeError(f"att_logits has shape {tuple(att_logits.shape)}, expected [B,1,{maxdisp},{H},{W}]")
    xt_f64 = 0
    if xt is not None:
        xt_f64 = _is_f64(xt, "xt")
        xt = xt.contiguous()
        if xt.numel() != B * maxdisp * H * W:
            raise RuntimeError(f"xt has shape {tuple(xt.shape)}, expected [B,{maxdisp},{H},{W}]")
        if shift is not None:
            shift = _f32c(shift.reshape(B, maxdisp), "shift")
    if out is None:
        out = torch.empty((B, 2 * Cc, maxdisp, H, W), dtype=torch.float32, device=ref.device)
    else:
        assert out.shape == (B, 2 * Cc, maxdisp, H, W) and out.dtype == torch.float32 and out.is_contiguous()
    if out.numel() == 0:
        return out
    with torch.cuda.device(ref.device):
        check(_lib.lib().dv_concat_volume_f32(_ptr(ref), _ptr(tgt), _ptr(out), B, Cc, H, W, maxdisp, int(mask_left),
                                              _ptr(att_logits), _ptr(xt), xt_f64, _ptr(shift), float(scale),
                                              _stream(ref)), "dv_concat_volume_f32")
    return out


def att_softmax(att_logits: torch.Tensor) -> torch.Tensor:
    """a4 factor — softmax over D of the attention logits [B,1,D,H,W] -> fp32 weights [B,D,H,W]
    (F.softmax(att_weights, dim=2), acv_ddim.py:390)."""
    _need_cuda(att_logits)
    att_logits = _f32c(att_logits, "att_logits")
    if att_logits.dim() == 5:
        assert att_logits.shape[1] == 1
        B, _, D, H, W = att_logits.shape
    else:
        B, D, H, W = att_logits.shape
    out = torch.empty((B, D, H, W), dtype=torch.float32, device=att_logits.device)
    if out.numel():
        with torch.cuda.device(out.device):
            check(_lib.lib().dv_att_softmax_f32(_ptr(att_logits), _ptr(out), B, D, H, W, _stream(out)), "dv_att_softmax_f32")
    return out


def filter_factor(xt: torch.Tensor, shift: Optional[torch.Tensor] = None, scale: float = 1.0,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """a9 factor — n = float(((clamp(xt + shift[b,d], -s, s)/s)+1)/2) as an fp32 map [B,D,H,W] (acv_ddim.py:256-258)."""
    _need_cuda(xt, shift)
    B, D, H, W = xt.shape
    f64 = _is_f64(xt, "xt")
    xt = xt.contiguous()
    if shift is not None:
        shift = _f32c(shift.reshape(B, D), "shift")
    if out is None:
        out = torch.empty((B, D, H, W), dtype=torch.float32, device=xt.device)
    if out.numel():
        with torch.cuda.device(out.device):
            check(_lib.lib().dv_filter_factor_f32(_ptr(xt), f64, _ptr(shift), float(scale), _ptr(out), B, D, H, W,
                                                  _stream(out)), "dv_filter_factor_f32")
    return out


def filter_factor_pair(xt: torch.Tensor, shift: Optional[torch.Tensor] = None, scale: float = 1.0):
    """a9 factor in both precisions: (n in the dtype of xt — the tensor predict_noise_from_start consumes —, n as fp32 —
    the tensor that multiplies the volume / the IGEV geometry lookup)."""
    _need_cuda(xt, shift)
    B, D, H, W = xt.shape
    f64 = _is_f64(xt, "xt")
    xt = xt.contiguous()
    if shift is not None:
        shift = _f32c(shift.reshape(B, D), "shift")
    n_native = torch.empty_like(xt)
    n32 = torch.empty((B, D, H, W), dtype=torch.float32, device=xt.device) if f64 else None
    if xt.numel():
        with torch.cuda.device(xt.device):
            check(_lib.lib().dv_filter_factor(_ptr(xt), f64, _ptr(shift), float(scale), _ptr(n32), _ptr(n_native), B, D, H, W,
                                              _stream(xt)), "dv_filter_factor")
    return n_native, (n32 if f64 else n_native)


def concat_volume_weighted(ref: torch.Tensor, tgt: torch.Tensor, maxdisp: int, *, mask_left: bool,
                           att_weights: Optional[torch.Tensor] = None, n: Optional[torch.Tensor] = None,
                           out: Optional[torch.Tensor] = None, out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """a3 (+a4 +a9) with PRECOMPUTED fp32 factor maps (`att_softmax`, `filter_factor` / ddim_step's n_next):
    out = (concat * att_weights) * n — the same values as `concat_volume(att_logits=..., xt=...)`.
    Needs H*W % 4 == 0 (every reference shape); raises DvLibraryError(DV_ERR_MISALIGNED) otherwise."""
    B, Cc, H, W = ref.shape
    _need_cuda(ref, tgt, att_weights, n)
    ref, tgt = _f32c(ref, "refimg_fea"), _f32c(tgt, "targetimg_fea")
    if tgt.shape != ref.shape:
        raise RuntimeError(f"The size of tensor a {tuple(ref.shape)} must match the size of tensor b {tuple(tgt.shape)}")
    for name, t in (("att_weights", att_weights), ("n", n)):
        if t is not None and (t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != B * maxdisp * H * W):
            raise RuntimeError(f"{name} must be a contiguous float32 [B,{maxdisp},{H},{W}] map")
    if out_dtype not in (torch.float32, torch.bfloat16):
        raise DvLibraryError(f"concat_volume_weighted: out_dtype must be float32 or bfloat16, got {out_dtype}")
    if out is None:
        out = torch.empty((B, 2 * Cc, maxdisp, H, W), dtype=out_dtype, device=ref.device)
    else:
        assert out.shape == (B, 2 * Cc, maxdisp, H, W) and out.dtype == out_dtype and out.is_contiguous()
    if out.numel() == 0:
        return out
    if out_dtype == torch.bfloat16:
        # fp32 operands and products, one round-to-nearest-even at the store (half the write traffic of the filter pass)
        with torch.cuda.device(ref.device):
            check(_lib.lib().dv_concat_volume_weighted_bf16(_ptr(ref), _ptr(tgt), _ptr(out), B, Cc, H, W, maxdisp,
                                                            int(mask_left), _ptr(att_weights), _ptr(n),
                                                            _tile_counters(ref.device, 0), _stream(ref)),
                  "dv_concat_volume_weighted_bf16")
        return out
    with torch.cuda.device(ref.device):
        check(_lib.lib().dv_concat_volume_weighted_f32(_ptr(ref), _ptr(tgt), _ptr(out), B, Cc, H, W, maxdisp,
                                                       int(mask_left), _ptr(att_weights), _ptr(n),
                                                       _tile_counters(ref.device, 0), _stream(ref)),
              "dv_concat_volume_weighted_f32")
    return out


def volume_filter(vol: torch.Tensor, xt: torch.Tensor, shift: Optional[torch.Tensor] = None, scale: float = 1.0, *,
                  out: Optional[torch.Tensor] = None, return_n: bool = False):
    """a9 — `volume * ((clamp(xt + shift, -s, s)/s + 1)/2).unsqueeze(1).float()` (acv_ddim.py:254-260)."""
    B, Cc, D, H, W = vol.shape
    _need_cuda(vol, xt, shift)
    vol = _f32c(vol, "volume")
    xt_f64 = _is_f64(xt, "noise")
    xt = xt.contiguous()
    if tuple(xt.shape) != (B, D, H, W):
        raise RuntimeError(f"noise has shape {tuple(xt.shape)}, expected {(B, D, H, W)}")
    if shift is not None:
        shift = _f32c(shift.reshape(B, D), "shift")
    if out is None:
        out = torch.empty_like(vol)
    n_out = torch.empty_like(xt) if return_n else None
    if vol.numel():
        with torch.cuda.device(vol.device):
            check(_lib.lib().dv_volume_filter_f32(_ptr(vol), _ptr(out), B, Cc, D, H, W, _ptr(xt), xt_f64, _ptr(shift),
                                                  float(scale), _ptr(n_out), _stream(vol)), "dv_volume_filter_f32")
    return (out, n_out) if return_n else out


def corr_volume_2sided(ref: torch.Tensor, tgt: torch.Tensor, maxdisp: int, num_groups: int,
                       out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """a5 — build_corrleation_volume, KITTI12/models/submodule.py:121-135 -> [B,G,2*maxdisp+1,H,W].
    `out` (optional): a float32 [B, G*(2*maxdisp+1), H, W] block, contiguous per sample — e.g. a channel slice of a concat
    buffer; a batch-strided `out` is filled one sample per call (the C-ABI writes contiguous volumes)."""
    B, Cc, H, W = ref.shape
    assert Cc % num_groups == 0
    _need_cuda(ref, tgt, out)
    ref, tgt = _f32c(ref, "refimg_fea"), _f32c(tgt, "targetimg_fea")
    if tgt.shape != ref.shape:
        raise RuntimeError(f"The size of tensor a {tuple(ref.shape)} must match the size of tensor b {tuple(tgt.shape)}")
    S = num_groups * (2 * maxdisp + 1)
    if out is None:
        res = torch.empty((B, num_groups, 2 * maxdisp + 1, H, W), dtype=torch.float32, device=ref.device)
        calls = [(ref, tgt, res.data_ptr(), B)]
    else:
        res = out
        ptr, bstride = _slice_view(out, B, S, H, W, "out")
        if B == 1 or bstride == S * H * W:
            calls = [(ref, tgt, ptr, B)]
        elif num_groups == 1 and bstride % (H * W) == 0:
            # a channel slice of a contiguous [B, planes, H, W] buffer: one launch set, batch stride = planes per sample
            with torch.cuda.device(ref.device):
                check(_lib.lib().dv_corr_volume_2sided_into_f32(_ptr(ref), _ptr(tgt), ptr, bstride // (H * W), 0, B, Cc, H, W,
                                                                maxdisp, _stream(ref)), "dv_corr_volume_2sided_into_f32")
            return res
        else:
            calls = [(ref[b:b + 1], tgt[b:b + 1], ptr + 4 * b * bstride, 1) for b in range(B)]
    with torch.cuda.device(ref.device):
        for r_, t_, p_, b_ in calls:
            check(_lib.lib().dv_corr_volume_2sided_f32(_ptr(r_), _ptr(t_), p_, b_, Cc, H, W, maxdisp, num_groups,
                                                       _stream(ref)), "dv_corr_volume_2sided_f32")
    return res


# --------------------------------------------------------------------------------------------
# regression
# --------------------------------------------------------------------------------------------
def softmax_regress(cost: torch.Tensor, *, return_prob: bool = False, used: Optional[torch.Tensor] = None,
                    want_unc: bool = False, vote_thresholds: Optional[Sequence[float]] = None,
                    ens_acc: Optional[torch.Tensor] = None, ens_coef: float = 0.0, ens_init: bool = False):
    """a6 (+a11 +a13) — softmax over dim 1 then disparity regression, one read of `cost`.

    Returns a dict with 'disp' [B,H,W] and, when requested, 'prob' [B,D,H,W], 'unc' [B,H,W],
    'vote' [B,H,W] (needs `used` and `vote_thresholds=(thr_dif, thr_unc)`).
    """
    assert len(cost.shape) == 4
    B, D, H, W = cost.shape
    _need_cuda(cost, used, ens_acc)
    cost = _f32c(cost, "cost")
    dev = cost.device
    res = {"disp": torch.empty((B, H, W), dtype=torch.float32, device=dev)}
    if return_prob:
        res["prob"] = torch.empty_like(cost)
    if want_unc:
        res["unc"] = torch.empty((B, H, W), dtype=torch.float32, device=dev)
    thr_dif = thr_unc = 0.0
    if vote_thresholds is not None:
        if used is None:
            raise DvLibraryError("vote needs `used`")
        used = _f32c(used.reshape(B, H, W), "used")
        thr_dif, thr_unc = float(vote_thresholds[0]), float(vote_thresholds[1])
        res["vote"] = torch.empty((B, H, W), dtype=torch.float32, device=dev)
    if ens_acc is not None:
        assert ens_acc.dtype == torch.float32 and ens_acc.is_contiguous() and ens_acc.numel() == B * H * W
    with torch.cuda.device(dev):
        check(_lib.lib().dv_softmax_regress_f32(_ptr(cost), B, D, H, W, _ptr(res["disp"]), _ptr(res.get("prob")),
                                                _ptr(used), _ptr(res.get("unc")), _ptr(res.get("vote")), thr_dif,
                                                thr_unc, _ptr(ens_acc), float(ens_coef), int(ens_init),
                                                _tile_counters(dev, 1), _stream(cost)), "dv_softmax_regress_f32")
    return res


def upsample_softmax_regress(cost_q: torch.Tensor, size: Sequence[int], *, align_corners: bool = False,
                             used: Optional[torch.Tensor] = None, want_unc: bool = False,
                             vote_thresholds: Optional[Sequence[float]] = None,
                             ens_acc: Optional[torch.Tensor] = None, ens_coef: float = 0.0, ens_init: bool = False):
    """f2 — `F.upsample(cost_q, size, mode='trilinear')` + softmax over D + disparity regression (+ uncertainty, vote,
    ensemble accumulate) without materialising the [B,D,H,W] logits (acv_ddim.py:267-270, pwcnet_ddim.py:480-484).
    cost_q is [B,1,Dq,h,w] or [B,Dq,h,w]; size = (D, H, W).  Returns the same dict as `softmax_regress` (no 'prob')."""
    if cost_q.dim() == 5:
        assert cost_q.shape[1] == 1
        cost_q = cost_q[:, 0]
    assert cost_q.dim() == 4 and len(size) == 3
    B, Dq, h, w = cost_q.shape
    D, H, W = (int(v) for v in size)
    _need_cuda(cost_q, used, ens_acc)
    cost_q = _f32c(cost_q, "cost_q")
    dev = cost_q.device
    res = {"disp": torch.empty((B, H, W), dtype=torch.float32, device=dev)}
    if want_unc:
        res["unc"] = torch.empty((B, H, W), dtype=torch.float32, device=dev)
    thr_dif = thr_unc = 0.0
    if vote_thresholds is not None:
        if used is None:
            raise DvLibraryError("vote needs `used`")
        used = _f32c(used.reshape(B, H, W), "used")
        thr_dif, thr_unc = float(vote_thresholds[0]), float(vote_thresholds[1])
        res["vote"] = torch.empty((B, H, W), dtype=torch.float32, device=dev)
    if ens_acc is not None:
        assert ens_acc.dtype == torch.float32 and ens_acc.is_contiguous() and ens_acc.numel() == B * H * W
    with torch.cuda.device(dev):
        check(_lib.lib().dv_upsample_softmax_regress_f32(_ptr(cost_q), B, Dq, h, w, D, H, W, int(align_corners),
                                                         _ptr(res["disp"]), _ptr(used), _ptr(res.get("unc")),
                                                         _ptr(res.get("vote")), thr_dif, thr_unc, _ptr(ens_acc),
                                                         float(ens_coef), int(ens_init), _stream(cost_q)),
              "dv_upsample_softmax_regress_f32")
    return res


def uncertainty_vote(disp: torch.Tensor, prob: torch.Tensor, used: Optional[torch.Tensor], thr_dif: float,
                     thr_unc: float, return_unc: bool = False):
    """a11 — sum_d |disp - d| * prob[d] and the renewal vote for a disparity that is not the regression of
    `prob` (pwcnet_ddim.py:553-570).  Returns vote [B,H,W] (and unc when return_unc)."""
    B, D, H, W = prob.shape
    _need_cuda(disp, prob, used)
    prob = _f32c(prob, "prob")
    disp = _f32c(disp.reshape(B, H, W), "disp")
    if used is not None:
        used = _f32c(used.reshape(B, H, W), "used")
    vote = torch.empty((B, H, W), dtype=torch.float32, device=prob.device)
    unc = torch.empty((B, H, W), dtype=torch.float32, device=prob.device) if return_unc else None
    with torch.cuda.device(prob.device):
        check(_lib.lib().dv_uncertainty_vote_f32(_ptr(prob), _ptr(disp), _ptr(used), B, D, H, W, float(thr_dif),
                                                 float(thr_unc), _ptr(unc), _ptr(vote), _stream(prob)),
              "dv_uncertainty_vote_f32")
    return (vote, unc) if return_unc else vote


def softmax_uncertainty_vote(disp: torch.Tensor, cost: torch.Tensor, used: Optional[torch.Tensor], thr_dif: float,
                             thr_unc: float, return_unc: bool = False):
    """a11 from the logits — `uncertainty_vote(disp, softmax(cost, 1), ...)` without the probability volume: the softmax is
    recomputed in registers from one read of `cost` (pwcnet_ddim.py:483 + :553-570)."""
    B, D, H, W = cost.shape
    _need_cuda(disp, cost, used)
    cost = _f32c(cost, "cost")
    disp = _f32c(disp.reshape(B, H, W), "disp")
    if used is not None:
        used = _f32c(used.reshape(B, H, W), "used")
    vote = torch.empty((B, H, W), dtype=torch.float32, device=cost.device)
    unc = torch.empty((B, H, W), dtype=torch.float32, device=cost.device) if return_unc else None
    with torch.cuda.device(cost.device):
        check(_lib.lib().dv_softmax_uncertainty_vote_f32(_ptr(cost), _ptr(disp), _ptr(used), B, D, H, W, float(thr_dif),
                                                         float(thr_unc), _ptr(unc), _ptr(vote),
                                                         _tile_counters(cost.device, 1), _stream(cost)),
              "dv_softmax_uncertainty_vote_f32")
    return (vote, unc) if return_unc else vote


def disparity_regression(x: torch.Tensor, maxdisp: int, keepdim: bool = False) -> torch.Tensor:
    """a6 — SceneFlow/models/submodule.py:173-177 (keepdim=True: KITTI15/core/submodule.py:219-223)."""
    assert len(x.shape) == 4
    B, D, H, W = x.shape
    if D != maxdisp:
        raise RuntimeError(
            f"The size of tensor a ({D}) must match the size of tensor b ({maxdisp}) at non-singleton dimension 1")
    _need_cuda(x)
    x = _f32c(x, "x")
    out = torch.empty((B, 1, H, W) if keepdim else (B, H, W), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.lib().dv_disparity_regression_f32(_ptr(x), _ptr(out), B, D, H, W, _stream(x)),
              "dv_disparity_regression_f32")
    return out


# --------------------------------------------------------------------------------------------
# DDIM elementwise pieces
# --------------------------------------------------------------------------------------------
def q_sample(x_start: torch.Tensor, noise: torch.Tensor, sqrt_ac: float, sqrt_1m_ac: float) -> torch.Tensor:
    """a7 — acv_ddim.py:241-246; result is float64 (the schedule buffers are float64)."""
    _need_cuda(x_start, noise)
    x_start, noise = x_start.contiguous(), noise.contiguous()
    if x_start.shape != noise.shape:
        raise RuntimeError("x_start and noise must have the same shape")
    out = torch.empty(x_start.shape, dtype=torch.float64, device=x_start.device)
    with torch.cuda.device(x_start.device):
        check(_lib.lib().dv_q_sample(_ptr(x_start), _is_f64(x_start, "x_start"), _ptr(noise), _is_f64(noise, "noise"),
                                     float(sqrt_ac), float(sqrt_1m_ac), _ptr(out), out.numel(), _stream(out)),
              "dv_q_sample")
    return out


def predict_noise_from_start(x_t: torch.Tensor, x0: torch.Tensor, sqrt_recip: float, sqrt_recipm1: float) -> torch.Tensor:
    """a8 — acv_ddim.py:248-252; float64 result."""
    _need_cuda(x_t, x0)
    x_t, x0 = x_t.contiguous(), x0.contiguous()
    if x_t.shape != x0.shape:
        raise RuntimeError("x_t and x0 must have the same shape")
    out = torch.empty(x_t.shape, dtype=torch.float64, device=x_t.device)
    with torch.cuda.device(x_t.device):
        check(_lib.lib().dv_predict_noise_from_start(_ptr(x_t), _is_f64(x_t, "x_t"), _ptr(x0), _is_f64(x0, "x0"),
                                                     float(sqrt_recip), float(sqrt_recipm1), _ptr(out), out.numel(),
                                                     _stream(out)), "dv_predict_noise_from_start")
    return out


def xstart_from_disp(disp_q: torch.Tensor, D: int = 48, scale: float = 1.0) -> torch.Tensor:
    """a10 — quarter-res disparity [B,h,w] (already /4) -> 2-tap x_start volume [B,D,h,w] in [-s, s]."""
    _need_cuda(disp_q)
    disp_q = _f32c(disp_q, "disp")
    B, h, w = disp_q.shape[0], disp_q.shape[-2], disp_q.shape[-1]
    assert disp_q.numel() == B * h * w
    out = torch.empty((B, D, h, w), dtype=torch.float32, device=disp_q.device)
    with torch.cuda.device(disp_q.device):
        check(_lib.lib().dv_xstart_from_disp_f32(_ptr(disp_q), _ptr(out), B, D, h, w, float(scale), _stream(out)),
              "dv_xstart_from_disp_f32")
    return out


def downsample_bilinear(x: torch.Tensor, size, clamp=None, post_scale: float = 1.0) -> torch.Tensor:
    """F.interpolate(clamp(x), size=size, mode='bilinear') * post_scale for [B,H,W] maps."""
    _need_cuda(x)
    x = _f32c(x, "x")
    B, H, W = x.shape[0], x.shape[-2], x.shape[-1]
    assert x.numel() == B * H * W
    h, w = int(size[0]), int(size[1])
    lo, hi = (1.0, 0.0) if clamp is None else (float(clamp[0]), float(clamp[1]))
    out = torch.empty((B, h, w), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        check(_lib.lib().dv_downsample_bilinear_f32(_ptr(x), _ptr(out), B, H, W, h, w, lo, hi, float(post_scale),
                                                    _stream(out)), "dv_downsample_bilinear_f32")
    return out


def ensemble(maps: Sequence[torch.Tensor], cof: Sequence[float]) -> torch.Tensor:
    """a13 — torch.sum(cat(maps) * cof, dim=0) (acv_ddim.py:365-369)."""
    assert len(maps) == len(cof) and 0 < len(maps) <= 8
    _need_cuda(*maps)
    maps = [_f32c(m, "map") for m in maps]
    n = maps[0].numel()
    assert all(m.numel() == n for m in maps)
    out = torch.empty_like(maps[0])
    ptrs = (C.c_void_p * len(maps))(*[m.data_ptr() for m in maps])
    cofs = (C.c_float * len(maps))(*[float(c) for c in cof])
    with torch.cuda.device(out.device):
        check(_lib.lib().dv_ensemble_f32(ptrs, cofs, len(maps), _ptr(out), n, _stream(out)), "dv_ensemble_f32")
    return out


def select_close(a: torch.Tensor, b: torch.Tensor, thr: float) -> torch.Tensor:
    """torch.where(torch.abs(a - b) < thr, a, b) — IGEV's fallback to the initial disparity (igev_stereo_ddim.py:323-325)."""
    _need_cuda(a, b)
    a, b = _f32c(a, "a"), _f32c(b, "b")
    assert a.numel() == b.numel()
    out = torch.empty_like(a)
    if out.numel():
        with torch.cuda.device(out.device):
            check(_lib.lib().dv_select_close_f32(_ptr(a), _ptr(b), float(thr), _ptr(out), out.numel(), _stream(out)),
                  "dv_select_close_f32")
    return out


def ddim_step(*, disp: torch.Tensor, xt: torch.Tensor, shift: Optional[torch.Tensor], scale: float,
              sqrt_recip: float, sqrt_recipm1: float, last_step: bool,
              disp_clamp_hi: float = 191.0, coords0: Optional[torch.Tensor] = None,
              vote: Optional[torch.Tensor] = None, used: Optional[torch.Tensor] = None, vote_thr_dif: float = 0.0,
              mask: Optional[torch.Tensor] = None,
              sqrt_alpha_next: float = 0.0, c: float = 0.0, sigma: float = 0.0,
              step_noise: Optional[torch.Tensor] = None,
              renoise: Optional[torch.Tensor] = None,
              asd: Optional[torch.Tensor] = None, q_noise: Optional[torch.Tensor] = None,
              sqrt_ac: float = 0.0, sqrt_1m_ac: float = 0.0, want_asd_out: bool = False,
              want_eps: bool = False, shift_next: Optional[torch.Tensor] = None, want_n_next: bool = False):
    """a8+a10+a11+a12 — one fused DDIM sampler step (see dv_ddim_step_args in include/dv_b200.h).

    Returns dict(x0=[B,D,h,w] fp32, x_next=[B,D,h,w] fp64 (fp32 when last_step), eps=fp64 or None,
    asd_out=fp64 or None, n_next=fp32 filter factor of the NEXT step when want_n_next).  `mask` is updated in place.
    """
    _need_cuda(disp, xt, shift, coords0, vote, used, mask, step_noise, renoise, asd, q_noise)
    B, D, h, w = xt.shape
    disp = _f32c(disp, "disp")
    H, W = disp.shape[-2], disp.shape[-1]
    assert disp.numel() == B * H * W
    xt = xt.contiguous()
    dev = xt.device
    a = DdimStepArgs()
    a.B, a.D, a.h, a.w, a.H, a.W = B, D, h, w, H, W
    a.disp = _ptr(disp)
    a.disp_clamp_hi = float(disp_clamp_hi)
    keep = [disp, xt]
    if coords0 is not None:
        coords0 = _f32c(coords0, "coords0")
        assert coords0.numel() == B * h * w
        keep.append(coords0)
    a.coords0 = _ptr(coords0)
    a.xt = _ptr(xt)
    a.xt_is_f64 = _is_f64(xt, "xt")
    if shift is not None:
        shift = _f32c(shift.reshape(B, D), "shift")
        keep.append(shift)
    a.shift = _ptr(shift)
    a.scale = float(scale)
    if vote is not None:
        vote = _f32c(vote, "vote")
        assert vote.numel() == B * H * W
        keep.append(vote)
    if used is not None:
        used = _f32c(used, "used")
        assert used.numel() == B * H * W
        keep.append(used)
    a.vote, a.used, a.vote_thr_dif = _ptr(vote), _ptr(used), float(vote_thr_dif)
    if mask is not None:
        assert mask.dtype == torch.float32 and mask.is_contiguous() and mask.numel() == B * h * w
    a.mask = _ptr(mask)
    a.sqrt_recip, a.sqrt_recipm1 = float(sqrt_recip), float(sqrt_recipm1)
    a.last_step = int(last_step)
    a.sqrt_alpha_next, a.c, a.sigma = float(sqrt_alpha_next), float(c), float(sigma)
    if step_noise is not None:
        if step_noise.dtype != xt.dtype:
            raise DvLibraryError("step_noise must have the dtype of xt (it is randn_like(img) in the reference)")
        step_noise = step_noise.contiguous()
        assert step_noise.shape == xt.shape
        keep.append(step_noise)
    a.step_noise = _ptr(step_noise)
    a.renoise_mode = 0
    asd_out = None
    if not last_step:
        if renoise is not None:
            if renoise.dtype != torch.float64:
                raise DvLibraryError("renoise must be float64 (rand_like of a float64 tensor in the reference)")
            renoise = renoise.contiguous()
            assert renoise.shape == xt.shape
            keep.append(renoise)
            a.renoise_mode = 1
            a.renoise = _ptr(renoise)
        elif asd is not None:
            assert q_noise is not None
            asd, q_noise = asd.contiguous(), q_noise.contiguous()
            assert asd.shape == xt.shape and q_noise.shape == xt.shape
            keep += [asd, q_noise]
            a.renoise_mode = 2
            a.asd, a.asd_is_f64 = _ptr(asd), _is_f64(asd, "asd")
            a.q_noise, a.q_noise_is_f64 = _ptr(q_noise), _is_f64(q_noise, "q_noise")
            a.sqrt_ac, a.sqrt_1m_ac = float(sqrt_ac), float(sqrt_1m_ac)
            if want_asd_out:
                asd_out = torch.empty(xt.shape, dtype=torch.float64, device=dev)
            a.asd_out = _ptr(asd_out)
    x0 = torch.empty((B, D, h, w), dtype=torch.float32, device=dev)
    eps = torch.empty((B, D, h, w), dtype=torch.float64, device=dev) if want_eps else None
    x_next = torch.empty((B, D, h, w), dtype=torch.float32 if last_step else torch.float64, device=dev)
    a.x0_out, a.eps_out, a.x_next = _ptr(x0), _ptr(eps), _ptr(x_next)
    n_next = None
    if want_n_next and not last_step:
        if shift_next is not None:
            shift_next = _f32c(shift_next.reshape(B, D), "shift_next")
            keep.append(shift_next)
        n_next = torch.empty((B, D, h, w), dtype=torch.float32, device=dev)
        a.shift_next, a.n_next_out = _ptr(shift_next), _ptr(n_next)
    with torch.cuda.device(dev):
        check(_lib.lib().dv_ddim_step(C.byref(a), _stream(xt)), "dv_ddim_step")
    del keep
    return {"x0": x0, "x_next": x_next, "eps": eps, "asd_out": asd_out, "n_next": n_next}


def warp(x: torch.Tensor, disp: torch.Tensor) -> torch.Tensor:
    """f3 — warp(x, disp) of KITTI12/models/submodule.py:137-176: x [B,C,H,W] (right features), disp [B,1,H,W]."""
    B, C, H, W = x.shape
    _need_cuda(x, disp)
    x, disp = _f32c(x, "x"), _f32c(disp, "disp")
    if disp.numel() != B * H * W:
        raise RuntimeError(f"disp has shape {tuple(disp.shape)}, expected [B,1,{H},{W}]")
    out = torch.empty_like(x)
    if out.numel():
        with torch.cuda.device(x.device):
            check(_lib.lib().dv_warp_f32(_ptr(x), _ptr(disp), _ptr(out), B, C, H, W, _stream(x)), "dv_warp_f32")
    return out


def warp_bwd(grad_out: torch.Tensor, x: torch.Tensor, disp: torch.Tensor, need_x: bool = True, need_disp: bool = True):
    """f1 — gradients of warp(x, disp) w.r.t. the features and the disparity (the autograd of
    KITTI12/models/submodule.py:169-176).  Returns (grad_x or None, grad_disp [B,1,H,W] or None)."""
    B, C, H, W = x.shape
    _need_cuda(grad_out, x, disp)
    grad_out, x, disp = _f32c(grad_out, "grad_out"), _f32c(x, "x"), _f32c(disp, "disp")
    if tuple(grad_out.shape) != (B, C, H, W) or disp.numel() != B * H * W:
        raise RuntimeError(f"grad_out {tuple(grad_out.shape)} / disp {tuple(disp.shape)} do not match x {tuple(x.shape)}")
    if not (need_x or need_disp):
        return None, None
    gx = torch.empty_like(x) if need_x else None
    gd = torch.empty(B, 1, H, W, device=x.device, dtype=torch.float32) if need_disp else None
    if x.numel():
        with torch.cuda.device(x.device):
            check(_lib.lib().dv_warp_bwd_f32(_ptr(grad_out), _ptr(x), _ptr(disp), _ptr(gx) if need_x else None,
                                             _ptr(gd) if need_disp else None, B, C, H, W, _stream(x)), "dv_warp_bwd_f32")
    return gx, gd


def _slice_view(t: Optional[torch.Tensor], B: int, Cc: int, H: int, W: int, name: str):
    """(pointer, batch stride in floats) of an output that is a [B,Cc,H,W] float32 block, contiguous per sample — a whole
    tensor or a channel slice `buf[:, c0:c0+Cc]` of a contiguous concat buffer."""
    if t is None:
        return None, 0
    if t.dtype != torch.float32 or tuple(t.shape) != (B, Cc, H, W):
        raise RuntimeError(f"{name} must be float32 [{B},{Cc},{H},{W}], got {t.dtype} {tuple(t.shape)}")
    st = t.stride()
    if not (st[3] == 1 and st[2] == W and st[1] == H * W and (B == 1 or st[0] >= Cc * H * W)):
        raise RuntimeError(f"{name} must be contiguous within each sample (a channel slice of a contiguous buffer is fine)")
    return t.data_ptr(), (st[0] if B > 1 else Cc * H * W)


def refine_input_assemble(ref: torch.Tensor, src: torch.Tensor, disp: torch.Tensor, maxdisp: int, num_groups: int = 1, *,
                          corr_out: Optional[torch.Tensor] = None, diff_out: Optional[torch.Tensor] = None,
                          copy_out: Optional[torch.Tensor] = None):
    """PWCNet's refinement input without torch.cat (KITTI12/models/pwcnet_ddim.py:493-499): warp(src, disp), the +-maxdisp
    correlation volume of (ref, warped), `ref - warped` and a copy of `ref`, the last three written directly into
    caller-provided buffers that may be channel slices of one concat buffer.  Returns (warped, corr)."""
    B, Cc, H, W = ref.shape
    assert Cc % num_groups == 0
    _need_cuda(ref, src, disp, corr_out, diff_out, copy_out)
    ref, src, disp = _f32c(ref, "ref"), _f32c(src, "src"), _f32c(disp, "disp")
    if src.shape != ref.shape or disp.numel() != B * H * W:
        raise RuntimeError("refine_input_assemble: src must match ref and disp must be [B,1,H,W]")
    dp, ds = _slice_view(diff_out, B, Cc, H, W, "diff_out")
    yp, ys = _slice_view(copy_out, B, Cc, H, W, "copy_out")
    warped = torch.empty_like(src)
    if warped.numel():
        with torch.cuda.device(ref.device):
            check(_lib.lib().dv_warp_assemble_f32(_ptr(src), _ptr(disp), _ptr(ref), _ptr(warped), dp, ds, yp, ys, B, Cc, H, W,
                                                  _stream(ref)), "dv_warp_assemble_f32")
    corr = corr_volume_2sided(ref, warped, maxdisp, num_groups, out=corr_out)
    return warped, corr


# --------------------------------------------------------------------------------------------
# backward passes
# --------------------------------------------------------------------------------------------
def gwc_volume_bwd(grad_out: torch.Tensor, ref: torch.Tensor, tgt: torch.Tensor, num_groups: int, *,
                   need_ref: bool = True, need_tgt: bool = True, two_sided_maxdisp: Optional[int] = None):
    """Gradients of build_gwc_volume (or, with two_sided_maxdisp=m, build_corrleation_volume) w.r.t. (ref, tgt)."""
    B, C, H, W = ref.shape
    _need_cuda(grad_out, ref, tgt)
    grad_out, ref, tgt = _f32c(grad_out, "grad_out"), _f32c(ref, "ref"), _f32c(tgt, "tgt")
    gref = torch.empty_like(ref) if need_ref else None
    gtgt = torch.empty_like(tgt) if need_tgt else None
    if ref.numel() == 0 or not (need_ref or need_tgt):
        return gref, gtgt
    with torch.cuda.device(ref.device):
        if two_sided_maxdisp is None:
            D = grad_out.shape[2]
            assert tuple(grad_out.shape) == (B, num_groups, D, H, W)
            check(_lib.lib().dv_gwc_volume_bwd_f32(_ptr(grad_out), _ptr(ref), _ptr(tgt), _ptr(gref), _ptr(gtgt), B, C, H, W,
                                                   D, num_groups, _stream(ref)), "dv_gwc_volume_bwd_f32")
        else:
            m = int(two_sided_maxdisp)
            assert tuple(grad_out.shape) == (B, num_groups, 2 * m + 1, H, W)
            check(_lib.lib().dv_corr_volume_2sided_bwd_f32(_ptr(grad_out), _ptr(ref), _ptr(tgt), _ptr(gref), _ptr(gtgt), B,
                                                           C, H, W, m, num_groups, _stream(ref)),
                  "dv_corr_volume_2sided_bwd_f32")
    return gref, gtgt


def groupwise_correlation_bwd(grad_out: torch.Tensor, fea1: torch.Tensor, fea2: torch.Tensor, num_groups: int):
    B, C, H, W = fea1.shape
    _need_cuda(grad_out, fea1, fea2)
    grad_out, fea1, fea2 = _f32c(grad_out, "grad_out"), _f32c(fea1, "fea1"), _f32c(fea2, "fea2")
    g1, g2 = torch.empty_like(fea1), torch.empty_like(fea2)
    if fea1.numel():
        with torch.cuda.device(fea1.device):
            check(_lib.lib().dv_groupwise_correlation_bwd_f32(_ptr(grad_out), _ptr(fea1), _ptr(fea2), _ptr(g1), _ptr(g2), B, C,
                                                              H, W, num_groups, _stream(fea1)),
                  "dv_groupwise_correlation_bwd_f32")
    return g1, g2


def concat_volume_bwd(grad_out: torch.Tensor, mask_left: bool, *, need_ref: bool = True, need_tgt: bool = True):
    """Gradients of build_concat_volume w.r.t. (ref, tgt): grad_out is [B,2C,D,H,W]."""
    B, C2, D, H, W = grad_out.shape
    C = C2 // 2
    _need_cuda(grad_out)
    grad_out = _f32c(grad_out, "grad_out")
    gref = torch.empty((B, C, H, W), dtype=torch.float32, device=grad_out.device) if need_ref else None
    gtgt = torch.empty((B, C, H, W), dtype=torch.float32, device=grad_out.device) if need_tgt else None
    if grad_out.numel() and (need_ref or need_tgt):
        with torch.cuda.device(grad_out.device):
            check(_lib.lib().dv_concat_volume_bwd_f32(_ptr(grad_out), _ptr(gref), _ptr(gtgt), B, C, H, W, D, int(mask_left),
                                                      _stream(grad_out)), "dv_concat_volume_bwd_f32")
    return gref, gtgt


def disparity_regression_bwd(grad_out: torch.Tensor, maxdisp: int) -> torch.Tensor:
    """Gradient of disparity_regression w.r.t. x: [B,H,W] (or [B,1,H,W]) -> [B,maxdisp,H,W]."""
    _need_cuda(grad_out)
    grad_out = _f32c(grad_out, "grad_out")
    B, H, W = grad_out.shape[0], grad_out.shape[-2], grad_out.shape[-1]
    gx = torch.empty((B, maxdisp, H, W), dtype=torch.float32, device=grad_out.device)
    if gx.numel():
        with torch.cuda.device(gx.device):
            check(_lib.lib().dv_disparity_regression_bwd_f32(_ptr(grad_out), _ptr(gx), B, maxdisp, H, W, _stream(gx)),
                  "dv_disparity_regression_bwd_f32")
    return gx


def softmax_regress_bwd(cost: torch.Tensor, grad_disp: torch.Tensor) -> torch.Tensor:
    """Gradient of disparity_regression(F.softmax(cost, 1)) w.r.t. cost [B,D,H,W]; grad_disp is [B,H,W] (or [B,1,H,W])."""
    B, D, H, W = cost.shape
    _need_cuda(cost, grad_disp)
    cost, grad_disp = _f32c(cost, "cost"), _f32c(grad_disp, "grad_disp")
    assert grad_disp.numel() == B * H * W
    gx = torch.empty_like(cost)
    if gx.numel():
        with torch.cuda.device(gx.device):
            check(_lib.lib().dv_softmax_regress_bwd_f32(_ptr(cost), _ptr(grad_disp), _ptr(gx), B, D, H, W, _stream(gx)),
                  "dv_softmax_regress_bwd_f32")
    return gx


def acv_volume_bwd(grad_out: torch.Tensor, cl: torch.Tensor, cr: torch.Tensor, att_weights: Optional[torch.Tensor],
                   n: Optional[torch.Tensor], *, mask_left: bool = False, need_cl: bool = True, need_cr: bool = True,
                   need_att: bool = True):
    """Gradients of out = (concat(cl, cr) * att_weights) * n (`concat_volume_weighted`) w.r.t. cl, cr and the attention
    LOGITS (att_weights = softmax over D of them).  Returns (grad_cl, grad_cr, grad_att [B,D,H,W]); entries not needed
    are None."""
    B, C2, D, H, W = grad_out.shape
    Cc = C2 // 2
    _need_cuda(grad_out, cl, cr, att_weights, n)
    grad_out, cl, cr = _f32c(grad_out, "grad_out"), _f32c(cl, "cl"), _f32c(cr, "cr")
    assert tuple(cl.shape) == (B, Cc, H, W) and tuple(cr.shape) == (B, Cc, H, W)
    for name, t in (("att_weights", att_weights), ("n", n)):
        if t is not None and (t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != B * D * H * W):
            raise RuntimeError(f"{name} must be a contiguous float32 [B,{D},{H},{W}] map")
    need_att = need_att and att_weights is not None
    gcl = torch.empty_like(cl) if need_cl else None
    gcr = torch.empty_like(cr) if need_cr else None
    gatt = torch.empty((B, D, H, W), dtype=torch.float32, device=grad_out.device) if need_att else None
    if grad_out.numel() and (need_cl or need_cr or need_att):
        with torch.cuda.device(grad_out.device):
            check(_lib.lib().dv_acv_volume_bwd_f32(_ptr(grad_out), _ptr(cl), _ptr(cr), _ptr(att_weights), _ptr(n), _ptr(gcl),
                                                   _ptr(gcr), _ptr(gatt), B, Cc, H, W, D, int(mask_left), _stream(grad_out)),
                  "dv_acv_volume_bwd_f32")
    return gcl, gcr, gatt


# --------------------------------------------------------------------------------------------
# IGEV geometry
# --------------------------------------------------------------------------------------------
def corr1d_allpairs(fmap1: torch.Tensor, fmap2: torch.Tensor, return_pooled: bool = False):
    """a14 — Combined_Geo_Encoding_Volume.corr (geometry_ddim.py:72-80) -> [B,H,W1,1,W2]; with return_pooled also level 1
    of the correlation pyramid ([B,H,W1,1,W2//2], avg_pool2d([1,2]), geometry_ddim.py:27-30) from the same launch."""
    B, D, H, W1 = fmap1.shape
    _, _, _, W2 = fmap2.shape
    _need_cuda(fmap1, fmap2)
    fmap1, fmap2 = _f32c(fmap1, "fmap1"), _f32c(fmap2, "fmap2")
    out = torch.empty((B, H, W1, 1, W2), dtype=torch.float32, device=fmap1.device)
    with torch.cuda.device(fmap1.device):
        if return_pooled:
            pooled = torch.empty((B, H, W1, 1, W2 // 2), dtype=torch.float32, device=fmap1.device)
            check(_lib.lib().dv_corr1d_allpairs_pooled_f32(_ptr(fmap1), _ptr(fmap2), _ptr(out), _ptr(pooled), B, D, H, W1, W2,
                                                           _stream(out)), "dv_corr1d_allpairs_pooled_f32")
            return out, pooled
        check(_lib.lib().dv_corr1d_allpairs_f32(_ptr(fmap1), _ptr(fmap2), _ptr(out), B, D, H, W1, W2, _stream(out)),
              "dv_corr1d_allpairs_f32")
    return out


def geo_permute(geo: torch.Tensor) -> torch.Tensor:
    """geo [B,C,D,h,w] -> [B*h*w, C, 1, D] (geometry_ddim.py:19)."""
    B, Cc, D, h, w = geo.shape
    _need_cuda(geo)
    geo = _f32c(geo, "geo_volume")
    out = torch.empty((B * h * w, Cc, 1, D), dtype=torch.float32, device=geo.device)
    with torch.cuda.device(geo.device):
        check(_lib.lib().dv_geo_permute_f32(_ptr(geo), _ptr(out), B, Cc, D, h, w, _stream(out)), "dv_geo_permute_f32")
    return out


def avgpool_w2(rows: torch.Tensor) -> torch.Tensor:
    """F.avg_pool2d(rows, [1,2], stride=[1,2]) for [N, C, 1, L] rows (geometry_ddim.py:24-30)."""
    _need_cuda(rows)
    rows = _f32c(rows, "rows")
    L = rows.shape[-1]
    N = rows.numel() // L
    out = torch.empty((*rows.shape[:-1], L // 2), dtype=torch.float32, device=rows.device)
    with torch.cuda.device(rows.device):
        check(_lib.lib().dv_avgpool_w2_f32(_ptr(rows), _ptr(out), N, L, _stream(out)), "dv_avgpool_w2_f32")
    return out


def geo_lookup(geo_pyr: Sequence[torch.Tensor], corr_pyr: Sequence[torch.Tensor], disp: torch.Tensor,
               coords: torch.Tensor, noisy: Optional[torch.Tensor], radius: int) -> torch.Tensor:
    """a15 — Combined_Geo_Encoding_Volume.__call__ (geometry_ddim.py:33-69) -> [B, L*(C+1)*(2r+1), h, w]."""
    b, _, h, w = disp.shape
    _need_cuda(disp, coords, noisy, *geo_pyr, *corr_pyr)
    levels = len(geo_pyr)
    Cc, D = geo_pyr[0].shape[1], geo_pyr[0].shape[-1]
    W2 = corr_pyr[0].shape[-1]
    disp, coords = _f32c(disp, "disp"), _f32c(coords, "coords")
    assert coords.numel() == b * h * w
    if noisy is not None:
        noisy = _f32c(noisy, "noisy")
        assert noisy.numel() == b * h * w * D
    geo_pyr = [_f32c(g, "geo_pyramid") for g in geo_pyr]
    corr_pyr = [_f32c(c_, "corr_pyramid") for c_ in corr_pyr]
    out = torch.empty((b, levels * (Cc + 1) * (2 * radius + 1), h, w), dtype=torch.float32, device=disp.device)
    gp = (C.c_void_p * levels)(*[g.data_ptr() for g in geo_pyr])
    cp = (C.c_void_p * levels)(*[c_.data_ptr() for c_ in corr_pyr])
    with torch.cuda.device(disp.device):
        check(_lib.lib().dv_geo_lookup_f32(gp, cp, _ptr(noisy), _ptr(disp), _ptr(coords), _ptr(out), b, Cc, D, h, w,
                                           W2, levels, radius, _stream(out)), "dv_geo_lookup_f32")
    return out


def geo_pack(geo: torch.Tensor, num_levels: int = 2):
    """a14 — geo [B,C,D,h,w] -> hypothesis-major pyramid [[B*h*w, D>>l, C] for l < num_levels]: the reference's
    permute(0,3,4,1,2) (geometry_ddim.py:19) and its avg_pool2d chain (:24-26) in one pass over the volume."""
    B, Cc, D, h, w = geo.shape
    _need_cuda(geo)
    geo = _f32c(geo, "geo_volume")
    rows = [torch.empty((B * h * w, D >> l, Cc), dtype=torch.float32, device=geo.device) for l in range(num_levels)]
    rp = (C.c_void_p * num_levels)(*[r.data_ptr() for r in rows])
    with torch.cuda.device(geo.device):
        check(_lib.lib().dv_geo_pack_f32(_ptr(geo), rp, B, Cc, D, h, w, num_levels, _stream(geo)), "dv_geo_pack_f32")
    return rows


def geo_lookup_packed(geo_pyr: Sequence[torch.Tensor], corr_pyr: Sequence[torch.Tensor], disp: torch.Tensor,
                      coords: torch.Tensor, noisy: Optional[torch.Tensor], radius: int) -> torch.Tensor:
    """a15 on the packed pyramid of `geo_pack` (geo_pyr[l] is [N, D>>l, C]); same result as `geo_lookup`."""
    b, _, h, w = disp.shape
    _need_cuda(disp, coords, noisy, *geo_pyr, *corr_pyr)
    levels = len(geo_pyr)
    D, Cc = geo_pyr[0].shape[1], geo_pyr[0].shape[2]
    W2 = corr_pyr[0].shape[-1]
    disp, coords = _f32c(disp, "disp"), _f32c(coords, "coords")
    assert coords.numel() == b * h * w
    if noisy is not None:
        noisy = _f32c(noisy, "noisy")
        assert noisy.numel() == b * h * w * D
    geo_pyr = [_f32c(g, "geo_pyramid") for g in geo_pyr]
    corr_pyr = [_f32c(c_, "corr_pyramid") for c_ in corr_pyr]
    out = torch.empty((b, levels * (Cc + 1) * (2 * radius + 1), h, w), dtype=torch.float32, device=disp.device)
    gp = (C.c_void_p * levels)(*[g.data_ptr() for g in geo_pyr])
    cp = (C.c_void_p * levels)(*[c_.data_ptr() for c_ in corr_pyr])
    with torch.cuda.device(disp.device):
        check(_lib.lib().dv_geo_lookup_packed_f32(gp, cp, _ptr(noisy), _ptr(disp), _ptr(coords), _ptr(out), b, Cc, D, h,
                                                  w, W2, levels, radius, _stream(out)), "dv_geo_lookup_packed_f32")
    return out


def geo_filter_packed(geo_pyr: Sequence[torch.Tensor], noisy: torch.Tensor, out: Optional[Sequence[torch.Tensor]] = None):
    """a9 (IGEV) — geo_l * noise_l on the packed pyramid (geometry_ddim.py:37-43,56), every level in one launch.
    `noisy` is the caller's [B,D,h,w] tensor read as raw [N, D] rows (the reference's reshape without a permute)."""
    levels = len(geo_pyr)
    N, D, Cc = geo_pyr[0].shape
    _need_cuda(noisy, *geo_pyr)
    noisy = _f32c(noisy, "noisy")
    assert noisy.numel() == N * D
    geo_pyr = [_f32c(g, "geo_pyramid") for g in geo_pyr]
    if out is None:
        out = [torch.empty_like(g) for g in geo_pyr]
    ip = (C.c_void_p * levels)(*[g.data_ptr() for g in geo_pyr])
    op = (C.c_void_p * levels)(*[g.data_ptr() for g in out])
    with torch.cuda.device(noisy.device):
        check(_lib.lib().dv_geo_filter_packed_f32(ip, _ptr(noisy), op, N, Cc, D, levels, _stream(noisy)),
              "dv_geo_filter_packed_f32")
    return list(out)


# --------------------------------------------------------------------------------------------
# f4: context_upsample (KITTI15/core/submodule.py:241-253)
# --------------------------------------------------------------------------------------------
def context_upsample(disp_low: torch.Tensor, up_weights: torch.Tensor) -> torch.Tensor:
    """disp_low [B,1,h,w], up_weights [B,9,4h,4w] -> [B,4h,4w]."""
    b, c, h, w = disp_low.shape
    _need_cuda(disp_low, up_weights)
    if c != 1 or tuple(up_weights.shape) != (b, 9, 4 * h, 4 * w):
        raise RuntimeError(f"context_upsample: disp_low {tuple(disp_low.shape)} / up_weights {tuple(up_weights.shape)}, "
                           f"expected [B,1,h,w] / [B,9,4h,4w]")
    disp_low, up_weights = _f32c(disp_low, "disp_low"), _f32c(up_weights, "up_weights")
    out = torch.empty((b, 4 * h, 4 * w), dtype=torch.float32, device=disp_low.device)
    if out.numel():
        with torch.cuda.device(out.device):
            check(_lib.lib().dv_context_upsample_f32(_ptr(disp_low), _ptr(up_weights), _ptr(out), b, h, w, _stream(out)),
                  "dv_context_upsample_f32")
    return out


def context_upsample_bwd(grad_out: torch.Tensor, disp_low: torch.Tensor, up_weights: torch.Tensor, *,
                         need_low: bool = True, need_weights: bool = True):
    b, _, h, w = disp_low.shape
    _need_cuda(grad_out, disp_low, up_weights)
    grad_out, disp_low, up_weights = _f32c(grad_out, "grad_out"), _f32c(disp_low, "disp_low"), _f32c(up_weights, "up_weights")
    glow = torch.empty_like(disp_low) if need_low else None
    gw = torch.empty_like(up_weights) if need_weights else None
    if grad_out.numel() and (need_low or need_weights):
        with torch.cuda.device(grad_out.device):
            check(_lib.lib().dv_context_upsample_bwd_f32(_ptr(grad_out), _ptr(disp_low), _ptr(up_weights), _ptr(glow), _ptr(gw),
                                                         b, h, w, _stream(grad_out)), "dv_context_upsample_bwd_f32")
    return glow, gw


# --------------------------------------------------------------------------------------------
# f4: ACVNet patch convolutions (SceneFlow/models/acv_ddim.py:181-188,377-381)
# --------------------------------------------------------------------------------------------
def depthwise3x3_chain(vol: torch.Tensor, w1: torch.Tensor, w2: Optional[torch.Tensor], dil1: int, dil2: int = 1, *,
                       channels: Optional[Sequence[int]] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Depth-wise (1,3,3) Conv3d (zero padding = dilation, no bias) on vol [B,C,D,H,W], optionally chained with a second
    one; w1, w2 are the Conv3d weights [C,1,1,3,3] (or [C,9]).  `channels=(c0, c1)` restricts the work to a slice and
    leaves the other channels of `out` untouched."""
    B, Cc, D, H, W = vol.shape
    _need_cuda(vol, w1, w2, out)
    vol = _f32c(vol, "vol")
    w1 = _f32c(w1.reshape(-1, 9), "w1")
    if w2 is not None:
        w2 = _f32c(w2.reshape(-1, 9), "w2")
    c0, c1 = (0, Cc) if channels is None else (int(channels[0]), int(channels[1]))
    if w1.shape[0] != Cc or (w2 is not None and w2.shape[0] != Cc):
        raise RuntimeError(f"depthwise3x3_chain: weights must be [C={Cc},9]")
    if out is None:
        out = torch.empty_like(vol)
    assert out.shape == vol.shape and out.dtype == torch.float32 and out.is_contiguous()
    if vol.numel():
        with torch.cuda.device(vol.device):
            check(_lib.lib().dv_depthwise3x3_chain_f32(_ptr(vol), _ptr(w1), _ptr(w2), _ptr(out), B, Cc, D, H, W, c0, c1,
                                                       int(dil1), int(dil2), _stream(vol)), "dv_depthwise3x3_chain_f32")
    return out


def acv_patch_volume(gwc: torch.Tensor, w_patch: torch.Tensor, w_l1: torch.Tensor, w_l2: torch.Tensor,
                     w_l3: torch.Tensor) -> torch.Tensor:
    """cat(patch_l1(g[:, :8]), patch_l2(g[:, 8:24]), patch_l3(g[:, 24:40])) with g = patch(gwc)
    (acv_ddim.py:377-381): three launches (one per dilation class), the volume read once and written once."""
    n1, n2, n3 = w_l1.shape[0], w_l2.shape[0], w_l3.shape[0]
    assert gwc.shape[1] == n1 + n2 + n3 == w_patch.shape[0]
    w2 = torch.cat([w_l1.reshape(n1, 9), w_l2.reshape(n2, 9), w_l3.reshape(n3, 9)], 0)
    out = torch.empty_like(_f32c(gwc, "gwc"))
    depthwise3x3_chain(gwc, w_patch, w2, 1, 1, channels=(0, n1), out=out)
    depthwise3x3_chain(gwc, w_patch, w2, 1, 2, channels=(n1, n1 + n2), out=out)
    depthwise3x3_chain(gwc, w_patch, w2, 1, 3, channels=(n1 + n2, n1 + n2 + n3), out=out)
    return out
