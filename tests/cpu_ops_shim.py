"""TEST INFRASTRUCTURE — not a product path, never imported by diffuvolume_b200.

CPU stand-ins for the functions of `diffuvolume_b200.ops` that the tier-2 / tier-3 drop-ins call, each one a thin call
into the numpy oracle (oracle/dv_oracle.py).  tests/test_tier3_reference_cpu.py swaps them in for the CUDA ops so that
the HOST LOGIC of the drop-ins — which reference modules are called, in which order, with which tensors, which RNG draws
are made — can be run against the REAL, unmodified reference models (`/root/reference/SceneFlow/models/acv_ddim.py`,
`acv.py`) in the authoring container, where there is no GPU.  The CUDA kernels themselves are checked against the same
oracle by the `-m gpu` tests; together the two pin "reference model with install()" == "reference model" end to end.
"""
from __future__ import annotations

import numpy as np
import torch

from oracle import dv_oracle as O


def _np(t):
    return None if t is None else t.detach().cpu().numpy()


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


def gwc_volume(ref, tgt, maxdisp, num_groups, out=None, out_dtype=torch.float32):
    return _t(O.build_gwc_volume(_np(ref), _np(tgt), maxdisp, num_groups))


def gwc_volume_bwd(*a, **k):
    raise NotImplementedError("forward-only shim")


def concat_volume(ref, tgt, maxdisp, *, mask_left, att_logits=None, xt=None, shift=None, scale=1.0, out=None):
    v = O.build_concat_volume(_np(ref), _np(tgt), maxdisp, mask_left)
    if att_logits is not None:
        v = O.acv_attention_volume(_np(att_logits).reshape(v.shape[0], 1, maxdisp, *v.shape[-2:]), v)
    if xt is not None:
        v = O.volume_filter(v, _np(xt), _np(shift), scale)
    return _t(v)


def att_softmax(att_logits):
    a = _np(att_logits)
    if a.ndim == 5:
        a = a[:, 0]
    return _t(O.softmax(a, 1).astype(np.float32))


def filter_factor(xt, shift=None, scale=1.0, out=None):
    return _t(O.filter_factor(_np(xt), _np(shift), scale).astype(np.float32))


def filter_factor_pair(xt, shift=None, scale=1.0):
    n = O.filter_factor(_np(xt), _np(shift), scale)
    return _t(n), _t(n.astype(np.float32))


def concat_volume_weighted(ref, tgt, maxdisp, *, mask_left, att_weights=None, n=None, out=None, out_dtype=torch.float32):
    v = O.build_concat_volume(_np(ref), _np(tgt), maxdisp, mask_left)
    if att_weights is not None:
        v = (_np(att_weights)[:, None] * v).astype(np.float32)      # softmax(att) * concat (acv_ddim.py:390)
    if n is not None:
        v = (v * _np(n)[:, None]).astype(np.float32)                 # volume * noise.unsqueeze(1).float() (:260)
    return _t(v)


def volume_filter(vol, xt, shift=None, scale=1.0, *, out=None, return_n=False):
    n = O.filter_factor(_np(xt), _np(shift), scale)
    v = _t((_np(vol) * n.astype(np.float32)[:, None]).astype(np.float32))
    return (v, _t(n)) if return_n else v


def _regress(cost, used, vote_thresholds, ens_acc, ens_coef, return_prob=False):
    disp, prob = O.softmax_regress(cost)
    res = {"disp": _t(disp)}
    if return_prob:
        res["prob"] = _t(prob)
    if vote_thresholds is not None:
        unc = O.uncertainty(disp, prob)
        res["vote"] = _t(O.renewal_vote(disp, _np(used).reshape(disp.shape), unc, vote_thresholds[0], vote_thresholds[1]))
    if ens_acc is not None:
        ens_acc.add_(_t((disp * np.float32(ens_coef)).astype(np.float32)).reshape(ens_acc.shape))
    return res


def softmax_regress(cost, *, return_prob=False, used=None, want_unc=False, vote_thresholds=None, ens_acc=None,
                    ens_coef=0.0, ens_init=False):
    return _regress(_np(cost), used, vote_thresholds, ens_acc, ens_coef, return_prob)


def upsample_softmax_regress(cost_q, size, *, align_corners=False, used=None, want_unc=False, vote_thresholds=None,
                             ens_acc=None, ens_coef=0.0, ens_init=False):
    c = _np(cost_q)
    if c.ndim == 4:
        c = c[:, None]
    up = O.interpolate_trilinear(c, tuple(int(v) for v in size), align_corners)[:, 0]
    return _regress(up, used, vote_thresholds, ens_acc, ens_coef)


def q_sample(x_start, noise, sqrt_ac, sqrt_1m_ac):
    return _t(np.float64(sqrt_ac) * _np(x_start).astype(np.float64) + np.float64(sqrt_1m_ac) * _np(noise).astype(np.float64))


def predict_noise_from_start(x_t, x0, sqrt_recip, sqrt_recipm1):
    return _t((np.float64(sqrt_recip) * _np(x_t).astype(np.float64) - _np(x0).astype(np.float64)) / np.float64(sqrt_recipm1))


def xstart_from_disp(disp_q, D=48, scale=1.0):
    d = _np(disp_q)
    return _t(O.xstart_from_disp(d.reshape(d.shape[0], d.shape[-2], d.shape[-1]), D, scale))


def downsample_bilinear(x, size, clamp=None, post_scale=1.0):
    a = _np(x)
    if clamp is not None:
        a = np.clip(a, np.float32(clamp[0]), np.float32(clamp[1]))
    return _t((O.interpolate_bilinear(a, tuple(size)) * np.float32(post_scale)).astype(np.float32))


def ensemble(maps, cof):
    return _t(O.ensemble([_np(m) for m in maps], cof))


def ddim_step(*, disp, xt, shift, scale, sqrt_recip, sqrt_recipm1, last_step, disp_clamp_hi=191.0, coords0=None, vote=None,
              used=None, vote_thr_dif=0.0, mask=None, sqrt_alpha_next=0.0, c=0.0, sigma=0.0, step_noise=None, renoise=None,
              asd=None, q_noise=None, sqrt_ac=0.0, sqrt_1m_ac=0.0, want_asd_out=False, want_eps=False, shift_next=None,
              want_n_next=False):
    """The fused DDIM step (include/dv_b200.h: dv_ddim_step) restated with the oracle's pieces — ACV / PCW flavours."""
    assert coords0 is None, "IGEV flavour not needed by the CPU host-logic tests"
    f32, f64 = np.float32, np.float64
    d = _np(disp)
    B, D, h, w = xt.shape
    d = d.reshape(B, d.shape[-2], d.shape[-1])
    dq = (O.interpolate_bilinear(np.clip(d, f32(0), f32(disp_clamp_hi)), (h, w)) / f32(4)).astype(f32)
    x0 = O.xstart_from_disp(dq, D, scale)
    n = O.filter_factor(_np(xt), _np(shift), scale)
    eps = (f64(sqrt_recip) * n.astype(f64) - x0.astype(f64)) / f64(sqrt_recipm1)
    if vote is not None and mask is not None:
        mask.copy_(_t(O.update_mask(_np(mask), _np(vote).reshape(d.shape))))
    out = {"x0": _t(x0), "eps": _t(eps) if want_eps else None, "asd_out": None, "n_next": None}
    if last_step:
        out["x_next"] = _t(x0)
        return out
    t1 = (x0 * f32(sqrt_alpha_next)).astype(f32)
    sn = _np(step_noise)
    t3 = (f32(sigma) * sn).astype(f32).astype(f64) if sn.dtype == f32 else f64(sigma) * sn
    img = (t1.astype(f64) + f64(c) * eps) + t3
    m = _np(mask)[:, None] == 0
    if renoise is not None:
        img = np.where(m, _np(renoise).astype(f64), img)
    elif asd is not None:
        q = f64(sqrt_ac) * _np(asd).astype(f64) + f64(sqrt_1m_ac) * _np(q_noise).astype(f64)
        img = np.where(m, q, img)
        if want_asd_out:
            out["asd_out"] = _t(q)
    out["x_next"] = _t(img)
    if want_n_next:
        out["n_next"] = _t(O.filter_factor(img, _np(shift_next), scale).astype(f32))
    return out


def uncertainty_vote(disp, prob, used, thr_dif, thr_unc, return_unc=False):
    d = _np(disp).reshape(prob.shape[0], prob.shape[2], prob.shape[3])
    unc = O.uncertainty(d, _np(prob))
    vote = _t(O.renewal_vote(d, _np(used).reshape(d.shape), unc, thr_dif, thr_unc))
    return (vote, _t(unc)) if return_unc else vote


def softmax_uncertainty_vote(disp, cost, used, thr_dif, thr_unc, return_unc=False):
    _, prob = O.softmax_regress(_np(cost), cost.shape[1])
    return uncertainty_vote(disp, _t(prob), used, thr_dif, thr_unc, return_unc)


def warp(x, disp):
    return _t(O.warp(_np(x), _np(disp)))


def corr_volume_2sided(ref, tgt, maxdisp, num_groups):
    return _t(O.build_corrleation_volume(_np(ref), _np(tgt), maxdisp, num_groups))


def acv_patch_volume(*a, **k):
    raise AssertionError("the fused patch chain is CUDA-only; the CPU host-logic test must take the module path")


def disparity_regression(x, maxdisp, keepdim=False):
    return _t(O.disparity_regression(_np(x), maxdisp, keepdim))


def refine_input_assemble(ref, src, disp, maxdisp, num_groups=1, *, corr_out=None, diff_out=None, copy_out=None):
    w = _t(O.warp(_np(src), _np(disp)))
    c = _t(O.build_corrleation_volume(_np(ref), _np(w), maxdisp, num_groups)).reshape(ref.shape[0], -1, *ref.shape[-2:])
    if corr_out is not None:
        corr_out.copy_(c)
    if diff_out is not None:
        diff_out.copy_(ref - w)
    if copy_out is not None:
        copy_out.copy_(ref)
    return w, (corr_out if corr_out is not None else c)
