"""Parity of the CUDA path (through the C-ABI, via diffuvolume_b200.ops) against the oracle and the
reference-minted golden fixtures.  Runs on the B200 box: `pytest -m gpu`.

Tolerances (BASELINE.json north_star): volumes max|a-b|/max|ref| <= 1e-4 in fp32 (we hold 1e-5),
copies and zero regions bit-exact, disparities within 0.01 px (we hold 1e-3).
"""
import numpy as np
import pytest
import torch

import synth
from conftest import rel_max_err
from golden.make_golden import CONCAT_CASES, GWC_CASES, WARP_CASES, trace_inputs
from oracle import dv_oracle as O

pytestmark = pytest.mark.gpu
VOL_TOL = 1e-5


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module")
def ops():
    from diffuvolume_b200 import ops as _ops
    return _ops


# ------------------------------------------------------------------------------------------------
# a1 / a2
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("case", list(GWC_CASES))
def test_gwc_volume_golden(ops, golden, case):
    B, C, G, D, H, W, seed = GWC_CASES[case]
    ref, tgt = synth.normal((B, C, H, W), seed), synth.normal((B, C, H, W), seed + 1000)
    got = host(ops.gwc_volume(cu(ref), cu(tgt), D, G))
    want = golden[f"sf.gwc.{case}"]
    assert rel_max_err(got, want) < VOL_TOL
    for d in range(1, D):  # x < d is exactly +0.0
        z = got[:, :, d, :, : min(d, W)]
        assert (z == 0).all() and not np.signbit(z).any()


@pytest.mark.parametrize("shape", [
    (2, 320, 40, 48, 27, 240),    # ACV quarter-res width, 1/5 of the height
    (1, 320, 40, 24, 48, 156),    # PCWNet 1/8 scale
    (2, 320, 40, 12, 24, 78),     # PCWNet 1/16 (row stride not 16-byte aligned)
    (2, 320, 40, 6, 12, 39),      # PCWNet 1/32
    (1, 96, 8, 48, 24, 312),      # IGEV: 12 channels per group
    (1, 64, 4, 48, 20, 64),       # 16 channels per group
    (1, 16, 4, 48, 9, 60),        # 4 channels per group, H*W % 256 != 0 tail span
    (1, 10, 2, 7, 5, 13),         # cpg = 5, H*W % 4 != 0: shape-agnostic kernel
    (1, 8, 1, 100, 8, 64),        # D > 48: chunk loop, D > W
    (2, 16, 2, 48, 4, 8),         # H*W < D: several channel windows of (b=0, g=0) start before the tensor (clipped copies)
    (1, 64, 2, 48, 2, 8),         # same on the K-chunked kernel (32 channels per group)
])
def test_gwc_volume_vs_oracle(ops, shape):
    B, C, G, D, H, W = shape
    ref, tgt = synth.normal((B, C, H, W), 5), synth.normal((B, C, H, W), 6)
    got = host(ops.gwc_volume(cu(ref), cu(tgt), D, G))
    want = O.build_gwc_volume(ref, tgt, D, G)
    assert rel_max_err(got, want) < VOL_TOL
    assert np.array_equal(got == 0, want == 0) or rel_max_err(got, want) < VOL_TOL


def test_gwc_nonfinite_features_do_not_leak_into_zero_region(ops):
    B, C, G, D, H, W = 1, 8, 1, 12, 4, 64
    ref, tgt = synth.normal((B, C, H, W), 7), synth.normal((B, C, H, W), 8)
    ref[0, 0, 2, 3] = np.inf
    tgt[0, 1, 1, 60] = np.nan
    got = host(ops.gwc_volume(cu(ref), cu(tgt), D, G))
    for d in range(1, D):
        assert (got[:, :, d, :, :d] == 0).all()


def test_groupwise_correlation(ops, golden):
    B, C, G, D, H, W, seed = GWC_CASES["cpg12"]
    ref, tgt = synth.normal((B, C, H, W), seed), synth.normal((B, C, H, W), seed + 1000)
    got = host(ops.groupwise_correlation(cu(ref), cu(tgt), G))
    assert rel_max_err(got, golden["sf.groupwise.cpg12"]) < VOL_TOL
    with pytest.raises(AssertionError):
        ops.groupwise_correlation(cu(ref), cu(tgt), 5)


# ------------------------------------------------------------------------------------------------
# a3 / a4 / a9
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("prefix,mask_left", [("sf", False), ("k12", True)])
@pytest.mark.parametrize("case", list(CONCAT_CASES))
def test_concat_volume_bit_exact(ops, golden, prefix, mask_left, case):
    B, C, D, H, W, seed = CONCAT_CASES[case]
    ref, tgt = synth.normal((B, C, H, W), seed), synth.normal((B, C, H, W), seed + 1000)
    got = host(ops.concat_volume(cu(ref), cu(tgt), D, mask_left=mask_left))
    assert np.array_equal(got, golden[f"{prefix}.concat.{case}"])


@pytest.mark.parametrize("mask_left", [False, True])
@pytest.mark.parametrize("shape", [(2, 32, 48, 27, 240), (1, 12, 24, 48, 156), (2, 12, 6, 12, 39), (1, 12, 12, 24, 78),
                                   (1, 5, 9, 5, 13)])
def test_concat_volume_vs_oracle(ops, shape, mask_left):
    B, C, D, H, W = shape
    ref, tgt = synth.normal((B, C, H, W), 9), synth.normal((B, C, H, W), 10)
    got = host(ops.concat_volume(cu(ref), cu(tgt), D, mask_left=mask_left))
    assert np.array_equal(got, O.build_concat_volume(ref, tgt, D, mask_left))


def test_acv_attention_volume_golden(ops, golden):
    B, C, D, h, w = 1, 4, 48, 4, 56
    cl, cr = synth.normal((B, C, h, w), 51), synth.normal((B, C, h, w), 1051)
    att = synth.normal((B, 1, D, h, w), 52) * np.float32(3)
    got = host(ops.concat_volume(cu(cl), cu(cr), D, mask_left=False, att_logits=cu(att)))
    assert rel_max_err(got, golden["sf.acv_volume"]) < VOL_TOL


@pytest.mark.parametrize("xt_dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(2, 32, 48, 27, 240), (1, 4, 48, 5, 13)])
def test_fused_concat_att_filter_vs_oracle(ops, shape, xt_dtype):
    B, C, D, h, w = shape
    cl, cr = synth.normal((B, C, h, w), 11), synth.normal((B, C, h, w), 12)
    att = synth.normal((B, 1, D, h, w), 13) * np.float32(2)
    xt = (synth.normal((B, D, h, w), 14, dtype=np.float64) * 0.8).astype(xt_dtype)
    shift = synth.normal((B, D), 15) * np.float32(0.2)
    want = O.volume_filter(O.acv_attention_volume(att, O.build_concat_volume(cl, cr, D, False)), xt, shift, 1.0)
    got = host(ops.concat_volume(cu(cl), cu(cr), D, mask_left=False, att_logits=cu(att), xt=cu(xt), shift=cu(shift)))
    assert rel_max_err(got, want) < VOL_TOL
    # the op-boundary variant: volume in, filtered volume out (+ n itself)
    vol = cu(O.acv_attention_volume(att, O.build_concat_volume(cl, cr, D, False)))
    got2, n = ops.volume_filter(vol, cu(xt), cu(shift), 1.0, return_n=True)
    assert rel_max_err(host(got2), want) < 1e-6
    assert n.dtype == (torch.float64 if xt_dtype == np.float64 else torch.float32)
    np.testing.assert_allclose(host(n), O.filter_factor(xt, shift, 1.0), rtol=1e-6 if xt_dtype == np.float32 else 1e-14)


@pytest.mark.parametrize("xt_dtype", [np.float32, np.float64])
@pytest.mark.parametrize("mask_left", [False, True])
@pytest.mark.parametrize("shape", [(2, 32, 48, 27, 240), (1, 5, 48, 6, 14), (1, 12, 24, 48, 156), (2, 3, 7, 4, 9)])
def test_weighted_producer_vs_oracle_and_fused_kernel(ops, shape, mask_left, xt_dtype):
    """dv_att_softmax_f32 + dv_filter_factor_f32 + dv_concat_volume_weighted_f32 (the DDIM loop's producer)
    against the oracle, and BIT-identical to the single-kernel fused path (same roundings, same order)."""
    B, C, D, h, w = shape
    cl, cr = synth.normal((B, C, h, w), 21), synth.normal((B, C, h, w), 22)
    att = synth.normal((B, 1, D, h, w), 23) * np.float32(2)
    xt = (synth.normal((B, D, h, w), 24, dtype=np.float64) * 0.8).astype(xt_dtype)
    shift = synth.normal((B, D), 25) * np.float32(0.2)
    att_w = ops.att_softmax(cu(att))
    np.testing.assert_allclose(host(att_w), O.softmax(att[:, 0], axis=1), rtol=2e-6, atol=1e-9)
    n = ops.filter_factor(cu(xt), cu(shift), 1.0)
    np.testing.assert_allclose(host(n), O.filter_factor(xt, shift, 1.0).astype(np.float32), rtol=1e-6, atol=1e-7)
    concat = O.build_concat_volume(cl, cr, D, mask_left)
    # plain, att only, n only, both
    got = host(ops.concat_volume_weighted(cu(cl), cu(cr), D, mask_left=mask_left))
    np.testing.assert_array_equal(got, concat)
    got = host(ops.concat_volume_weighted(cu(cl), cu(cr), D, mask_left=mask_left, att_weights=att_w))
    assert rel_max_err(got, O.acv_attention_volume(att, concat)) < VOL_TOL
    np.testing.assert_array_equal(got, host(ops.concat_volume(cu(cl), cu(cr), D, mask_left=mask_left, att_logits=cu(att))))
    got = host(ops.concat_volume_weighted(cu(cl), cu(cr), D, mask_left=mask_left, n=n))
    assert rel_max_err(got, O.volume_filter(concat, xt, shift, 1.0)) < VOL_TOL
    want = O.volume_filter(O.acv_attention_volume(att, concat), xt, shift, 1.0)
    got = host(ops.concat_volume_weighted(cu(cl), cu(cr), D, mask_left=mask_left, att_weights=att_w, n=n))
    assert rel_max_err(got, want) < VOL_TOL
    fused = host(ops.concat_volume(cu(cl), cu(cr), D, mask_left=mask_left, att_logits=cu(att), xt=cu(xt), shift=cu(shift)))
    np.testing.assert_array_equal(got, fused)


def test_weighted_producer_rejects_unaligned_plane(ops):
    cl = cu(synth.normal((1, 2, 3, 7), 1))
    with pytest.raises(RuntimeError, match="MISALIGNED"):
        ops.concat_volume_weighted(cl, cl, 4, mask_left=False)


# ------------------------------------------------------------------------------------------------
# a5
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("key,shape,m,G,seed", [
    ("k12.corr2.m24", (1, 32, 8, 64), 24, 1, 41),
    ("k12.corr2.w80_m24", (1, 32, 6, 80), 24, 1, 43),
    ("sftop.corr2.tiny_m9", (2, 8, 3, 7), 9, 2, 42),
])
def test_corr_volume_2sided_golden(ops, golden, key, shape, m, G, seed):
    ref, tgt = synth.normal(shape, seed), synth.normal(shape, seed + 1000)
    got = host(ops.corr_volume_2sided(cu(ref), cu(tgt), m, G))
    want = golden[key]
    assert rel_max_err(got, want) < VOL_TOL
    assert np.array_equal(got == 0, want == 0)


def test_corr_volume_2sided_fullres_vs_oracle(ops):
    shape = (1, 32, 48, 1248)
    ref, tgt = synth.normal(shape, 16), synth.normal(shape, 17)
    got = host(ops.corr_volume_2sided(cu(ref), cu(tgt), 24, 1))
    assert rel_max_err(got, O.build_corrleation_volume(ref, tgt, 24, 1)) < VOL_TOL


# ------------------------------------------------------------------------------------------------
# a6 / a11 / a13
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("k", [1, 10])
def test_softmax_regress_golden(ops, golden, k):
    cost = synth.normal((1, 192, 16, 32), 31) * np.float32(k)
    r = ops.softmax_regress(cu(cost))
    assert np.abs(host(r["disp"]) - golden[f"sf.regress.k{k}"]).max() < 1e-3
    prob = O.softmax(cost, 1)
    got = host(ops.disparity_regression(cu(prob), 192))
    assert np.abs(got - golden[f"sf.regress.k{k}"]).max() < 1e-3
    got4 = host(ops.disparity_regression(cu(prob), 192, keepdim=True))
    assert got4.shape == (1, 1, 16, 32)


@pytest.mark.parametrize("shape", [(2, 192, 36, 64), (1, 48, 24, 312), (1, 192, 5, 13), (2, 96, 8, 20), (1, 300, 6, 16)])
def test_softmax_regress_full_vs_oracle(ops, shape):
    B, D, H, W = shape
    cost = synth.normal(shape, 18) * np.float32(4)
    disp_o, prob_o = O.softmax_regress(cost, D)
    unc_o = O.uncertainty(disp_o, prob_o)
    used = (disp_o + (synth.uniform((B, H, W), 19, dtype=np.float32) - np.float32(0.5)) * np.float32(4)).astype(np.float32)
    thr_unc = float(np.median(unc_o))
    ens = cu(synth.normal((B, H, W), 20))
    ens0 = host(ens).copy()
    r = ops.softmax_regress(cu(cost), return_prob=True, used=cu(used), want_unc=True, vote_thresholds=(1.0, thr_unc),
                            ens_acc=ens, ens_coef=0.3, ens_init=False)
    assert np.abs(host(r["disp"]) - disp_o).max() < 1e-3
    assert rel_max_err(host(r["prob"]), prob_o) < VOL_TOL
    assert np.abs(host(r["unc"]) - unc_o).max() < 1e-3
    vote_o = O.renewal_vote(disp_o, used, unc_o, 1.0, thr_unc)
    near = (np.abs(np.abs(disp_o - used) - 1.0) < 1e-3) | (np.abs(unc_o - thr_unc) < 1e-3)
    assert np.array_equal(host(r["vote"])[~near], vote_o[~near])
    np.testing.assert_allclose(host(ens), ens0 + np.float32(0.3) * disp_o, atol=1e-3)


@pytest.mark.parametrize("case", [
    # (B, Dq, h, w, D, H, W, align_corners)
    (2, 48, 9, 20, 192, 36, 80, False),      # the ACVNet call: x4 in every axis -> register fast path
    (1, 48, 7, 13, 192, 28, 52, False),      # fast path, ragged CTA tiles
    (1, 48, 9, 20, 192, 36, 80, True),       # PCWNet: align_corners=True (pwcnet_ddim.py:480) -> table path
    (1, 12, 5, 9, 48, 20, 36, False),        # other Dq
    (1, 48, 6, 10, 192, 22, 39, False),      # sizes that are not multiples (F.upsample accepts any size)
    (2, 7, 3, 5, 19, 11, 17, True),
])
def test_upsample_softmax_regress_vs_oracle(ops, case):
    B, Dq, h, w, D, H, W, ac = case
    cq = synth.normal((B, 1, Dq, h, w), 91) * np.float32(5)
    used = synth.uniform((B, H, W), 92, dtype=np.float32) * np.float32(D - 1)
    cost = O.interpolate_trilinear(cq, (D, H, W), align_corners=ac)[:, 0]
    disp, prob = O.softmax_regress(cost)
    unc = O.uncertainty(disp, prob)
    ens0 = synth.normal((B, H, W), 93)
    ens = cu(ens0.copy())
    r = ops.upsample_softmax_regress(cu(cq), (D, H, W), align_corners=ac, used=cu(used), want_unc=True,
                                     vote_thresholds=(1.0, 3.0), ens_acc=ens, ens_coef=0.3)
    assert np.abs(host(r["disp"]) - disp).max() < 1e-3
    assert np.abs(host(r["unc"]) - unc).max() < 1e-3
    vote = O.renewal_vote(disp, used, unc, 1.0, 3.0)
    borderline = (np.abs(np.abs(disp - used) - 1.0) < 2e-3) | (np.abs(unc - 3.0) < 2e-3)
    assert np.array_equal(host(r["vote"])[~borderline], vote[~borderline])
    assert np.abs(host(ens) - (ens0 + np.float32(0.3) * disp)).max() < 1e-3
    # disparity only, and agreement with the unfused CUDA path on the materialised logits
    r2 = ops.upsample_softmax_regress(cu(cq), (D, H, W), align_corners=ac)
    r3 = ops.softmax_regress(cu(cost))
    assert np.abs(host(r2["disp"]) - host(r3["disp"])).max() < 1e-3


def test_softmax_regress_big_golden(ops, golden):
    cost = synth.normal((1, 192, 135, 240), 92) * np.float32(4)
    r = ops.softmax_regress(cu(cost), want_unc=True)
    assert np.abs(host(r["disp"]) - golden["big.regress.disp"]).max() < 1e-3
    assert np.abs(host(r["unc"]) - golden["big.regress.unc"]).max() < 1e-3


def test_gwc_big_golden(ops, golden):
    B, C, G, D, H, W = 1, 320, 40, 48, 135, 240
    ref, tgt = synth.normal((B, C, H, W), 91), synth.normal((B, C, H, W), 1091)
    v = ops.gwc_volume(cu(ref), cu(tgt), D, G)
    s = golden["big.gwc.sum"]
    assert abs(float(v.abs().sum(dtype=torch.float64)) - s[1]) < 1e-6 * s[1]
    flat = host(v).reshape(-1)
    idx = np.linspace(0, flat.size - 1, 4096).astype(np.int64)
    assert np.abs(flat[idx] - golden["big.gwc.sample"]).max() < VOL_TOL * s[2]


# ------------------------------------------------------------------------------------------------
# a7 / a8 / a10 / a12 / a13
# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("t", [999, 599, 0])
def test_q_sample_and_pred_noise_golden(ops, golden, t):
    s = O.Schedule()
    x0 = synth.uniform((2, 48, 4, 8), 61, dtype=np.float32) * 2 - 1
    nz = synth.normal((2, 48, 4, 8), 62)
    qs = ops.q_sample(cu(x0), cu(nz), s.sqrt_alphas_cumprod[t], s.sqrt_one_minus_alphas_cumprod[t])
    assert qs.dtype == torch.float64
    np.testing.assert_allclose(host(qs), golden[f"sf.q_sample.t{t}"], rtol=1e-13, atol=1e-15)
    pn = ops.predict_noise_from_start(qs, cu(x0), s.sqrt_recip_alphas_cumprod[t], s.sqrt_recipm1_alphas_cumprod[t])
    np.testing.assert_allclose(host(pn), golden[f"sf.pred_noise.t{t}"], rtol=1e-11, atol=1e-11)
    pn32 = ops.predict_noise_from_start(cu(nz), cu(x0), s.sqrt_recip_alphas_cumprod[t], s.sqrt_recipm1_alphas_cumprod[t])
    np.testing.assert_allclose(host(pn32), golden[f"sf.pred_noise_f32.t{t}"], rtol=1e-11, atol=1e-11)


def test_xstart_golden_and_known_answers(ops, golden):
    got = host(ops.xstart_from_disp(cu(golden["trace.disp_q"]), 48, 1.0))
    assert np.array_equal(got, golden["trace.asd"])
    dq = np.array([0.0, 0.25, 3.0, 46.5, 47.0, 47.75], dtype=np.float32).reshape(1, 1, 6)
    vol = (host(ops.xstart_from_disp(cu(dq), 48, 1.0))[0, :, 0, :] + 1) / 2
    want = [{0: 1.0}, {0: 0.75, 1: 0.25}, {3: 1.0}, {46: 0.5, 47: 0.5}, {47: 1.0}, {47: 1.0}]
    for j, w in enumerate(want):
        ref = np.zeros(48, dtype=np.float32)
        for kk, v in w.items():
            ref[kk] = v
        np.testing.assert_allclose(vol[:, j], ref, atol=1e-6)


@pytest.mark.parametrize("HW", [(16, 32), (18, 30), (135 * 4, 64)])
def test_downsample_bilinear_vs_oracle(ops, HW):
    H, W = HW
    x = synth.normal((2, H, W), 21) * np.float32(60) + np.float32(90)
    got = host(ops.downsample_bilinear(cu(x), (H // 4, W // 4), clamp=(0, 191), post_scale=0.25))
    want = O.disp_to_quarter(x, 192)
    np.testing.assert_allclose(got, want, atol=2e-5)


def test_ensemble(ops):
    maps = [synth.normal((2, 16, 32), 22 + i) * np.float32(50) for i in range(6)]
    cof = [0.5, 0.0, 0.0, 0.0, 0.2, 0.3]
    got = host(ops.ensemble([cu(m) for m in maps], cof))
    np.testing.assert_array_equal(got, O.ensemble(maps, cof))


def test_ddim_trace_golden(ops, golden):
    """The reference's ddim_sample trace, replayed through the CUDA kernels step by step: fused
    producer (concat+ACV+filter), softmax-regress (+vote +ensemble) and the fused DDIM step.
    The stand-in for the conv stack (mean over channels * 2 + bias, trilinear x4) runs in numpy."""
    B, Cc, D, h, w, H, W = (int(v) for v in golden["trace.shape"])
    sched = O.Schedule()
    cl, cr = synth.normal((B, Cc, h, w), 71), synth.normal((B, Cc, h, w), 1071)
    att = synth.normal((B, 1, D, h, w), 72) * np.float32(2)
    bias, used = trace_inputs(B, D, h, w, H, W)
    rn, ru = golden["trace.randn_like_seeds"], golden["trace.rand_like_seeds"]
    img = ops.xstart_from_disp(cu(golden["trace.disp_q"]), D, 1.0)
    mask = torch.zeros((B, h, w), dtype=torch.float32, device="cuda")
    ens = torch.empty((B, H, W), dtype=torch.float32, device="cuda")
    cof = (0.5, 0.0, 0.0, 0.0, 0.2, 0.3)
    used_c = cu(used)
    ens.copy_(used_c * cof[0])
    pairs = sched.time_pairs()
    n_next = None
    for i, (time, time_next) in enumerate(pairs):
        np.testing.assert_allclose(host(img), golden[f"trace.img.{i}"], atol=1e-3)
        assert host(img).dtype == golden[f"trace.img.{i}"].dtype
        shift = cu(golden[f"trace.shift.t{time}"])
        vol_f = ops.concat_volume(cu(cl), cu(cr), D, mask_left=False, att_logits=cu(att), xt=img, shift=shift)
        if n_next is not None:   # the factor emitted by the previous fused step drives the weighted producer
            vol_w = ops.concat_volume_weighted(cu(cl), cu(cr), D, mask_left=False, att_weights=ops.att_softmax(cu(att)), n=n_next)
            assert torch.equal(vol_w, vol_f)
        c = host(vol_f).mean(axis=1, keepdims=True, dtype=np.float32) * np.float32(2.0)
        cost = O.interpolate_trilinear(c + bias[i], (192, H, W))[:, 0]
        r = ops.softmax_regress(cu(cost), used=used_c, vote_thresholds=(1.0, 3.0), ens_acc=ens, ens_coef=cof[i + 1])
        assert np.abs(host(r["disp"]) - golden[f"trace.disp.{i}"]).max() < 1e-3
        last = time_next < 0
        kw = {}
        if not last:
            san, c_, sigma = sched.ddim_coefficients(time, time_next)
            seed, is64 = rn[2 * i]
            sn = synth.normal((B, D, h, w), int(seed), dtype=np.float64).astype(np.float64 if is64 else np.float32)
            kw = dict(sqrt_alpha_next=san, c=c_, sigma=sigma, step_noise=cu(sn),
                      renoise=cu(synth.uniform((B, D, h, w), int(ru[i][0]), dtype=np.float64)),
                      shift_next=cu(golden[f"trace.shift.t{pairs[i + 1][0]}"]), want_n_next=True)
        st = ops.ddim_step(disp=r["disp"], xt=img, shift=shift, scale=1.0,
                           sqrt_recip=sched.sqrt_recip_alphas_cumprod[time],
                           sqrt_recipm1=sched.sqrt_recipm1_alphas_cumprod[time], last_step=last,
                           vote=r["vote"], mask=mask, want_eps=True, **kw)
        np.testing.assert_allclose(host(st["x0"]), golden[f"trace.x0.{i}"], atol=2e-4)
        np.testing.assert_allclose(host(st["eps"]), golden[f"trace.eps.{i}"], rtol=1e-6, atol=1e-3)
        img = st["x_next"]
        n_next = st["n_next"]
    assert 0.2 < float((mask == 0).float().mean()) < 0.8
    assert np.abs(host(ens) - golden["trace.pred"]).max() < 1e-3


# ------------------------------------------------------------------------------------------------
# a14 / a15
# ------------------------------------------------------------------------------------------------
def test_corr1d_and_geo_lookup_golden(ops, golden):
    B, Cf, h, w, Cg, D = 2, 16, 6, 40, 8, 48
    f1, f2 = synth.normal((B, Cf, h, w), 101), synth.normal((B, Cf, h, w), 102)
    geo = synth.normal((B, Cg, D, h, w), 103)
    disp = synth.uniform((B, 1, h, w), 104, dtype=np.float32) * np.float32(50) - np.float32(2)
    coords = np.broadcast_to(np.arange(w, dtype=np.float32).reshape(1, 1, 1, w), (B, 1, h, w)).copy()
    noisy = synth.uniform((B, D, h, w), 105, dtype=np.float32)
    corr = ops.corr1d_allpairs(cu(f1), cu(f2))
    assert corr.shape == (B, h, w, 1, w)
    assert rel_max_err(host(corr), golden["k15.corr"]) < VOL_TOL
    g0 = ops.geo_permute(cu(geo))
    g1 = ops.avgpool_w2(g0)
    c0 = corr.reshape(B * h * w, 1, 1, w)
    c1 = ops.avgpool_w2(c0)
    assert rel_max_err(host(g1), golden["k15.geo.pyr1"]) < 1e-6
    assert rel_max_err(host(c1), golden["k15.corr.pyr1"]) < VOL_TOL
    plain = ops.geo_lookup([g0, g1], [c0, c1], cu(disp), cu(coords), None, 4)
    assert plain.shape == (B, 162, h, w)
    assert rel_max_err(host(plain), golden["k15.geo.plain"]) < VOL_TOL
    ddim = ops.geo_lookup([g0, g1], [c0, c1], cu(disp), cu(coords), cu(noisy), 4)
    assert rel_max_err(host(ddim), golden["k15.geo.ddim"]) < VOL_TOL


def test_geo_lookup_igev_shape_vs_oracle(ops):
    B, Cf, h, w, Cg, D = 1, 96, 12, 312, 8, 48
    f1, f2 = synth.normal((B, Cf, h, w), 111), synth.normal((B, Cf, h, w), 112)
    geo = synth.normal((B, Cg, D, h, w), 113)
    disp = synth.uniform((B, 1, h, w), 114, dtype=np.float32) * np.float32(47)
    coords = np.broadcast_to(np.arange(w, dtype=np.float32).reshape(1, 1, 1, w), (B, 1, h, w)).copy()
    noisy = synth.uniform((B, D, h, w), 115, dtype=np.float32)
    vol = O.CombinedGeoEncodingVolume(f1, f2, geo, 2, 4)
    corr = ops.corr1d_allpairs(cu(f1), cu(f2))
    assert rel_max_err(host(corr), vol.init_corr_pyramid[0].reshape(corr.shape)) < VOL_TOL
    g0 = ops.geo_permute(cu(geo)); g1 = ops.avgpool_w2(g0)
    c0 = corr.reshape(B * h * w, 1, 1, w); c1 = ops.avgpool_w2(c0)
    got = ops.geo_lookup([g0, g1], [c0, c1], cu(disp), cu(coords), cu(noisy), 4)
    assert rel_max_err(host(got), vol(disp, coords, noisy)) < VOL_TOL


@pytest.mark.parametrize("shape", [(1, 8, 48, 12, 312), (2, 8, 48, 5, 13), (1, 4, 24, 6, 40), (1, 5, 17, 3, 11), (1, 16, 48, 4, 37)])
@pytest.mark.parametrize("levels", [1, 2, 3])
def test_geo_pack_and_packed_lookup_match_reference_layout(ops, shape, levels):
    """The hypothesis-major pyramid is the reference pyramid transposed (bit-exact), and the packed lookup gives the
    bit-identical result of the reference-layout lookup and the oracle's, with and without the noise multiply."""
    B, Cg, D, h, w = shape
    if (D >> (levels - 1)) < 2:
        pytest.skip("pyramid too deep for D")
    W2 = w + 3
    geo = synth.normal((B, Cg, D, h, w), 201)
    corr0 = synth.normal((B * h * w, 1, 1, W2), 202)
    disp = synth.uniform((B, 1, h, w), 203, dtype=np.float32) * np.float32(D + 6) - np.float32(3)
    coords = np.broadcast_to(np.arange(w, dtype=np.float32).reshape(1, 1, 1, w), (B, 1, h, w)).copy()
    noisy = synth.uniform((B, D, h, w), 204, dtype=np.float32)
    if (W2 >> (levels - 1)) < 2:
        pytest.skip("pyramid too deep for W2")
    ref_pyr = [ops.geo_permute(cu(geo))]
    corr_pyr = [cu(corr0)]
    for _ in range(levels - 1):
        ref_pyr.append(ops.avgpool_w2(ref_pyr[-1]))
        corr_pyr.append(ops.avgpool_w2(corr_pyr[-1]))
    packed = ops.geo_pack(cu(geo), levels)
    for l in range(levels):
        assert packed[l].shape == (B * h * w, D >> l, Cg)
        assert torch.equal(packed[l], ref_pyr[l][:, :, 0, :].transpose(1, 2).contiguous())
    for nz in (None, cu(noisy)):
        want = ops.geo_lookup(ref_pyr, corr_pyr, cu(disp), cu(coords), nz, 4)
        got = ops.geo_lookup_packed(packed, corr_pyr, cu(disp), cu(coords), nz, 4)
        assert torch.equal(got, want)
    # the DDIM filter taken once on the packed pyramid, then a noise-free lookup: the same bits
    filtered = ops.geo_filter_packed(packed, cu(noisy))
    assert torch.equal(ops.geo_lookup_packed(filtered, corr_pyr, cu(disp), cu(coords), None, 4), want)
    for r in (1, 3):   # run-time radius path
        assert torch.equal(ops.geo_lookup_packed(filtered, corr_pyr, cu(disp), cu(coords), None, r),
                           ops.geo_lookup(ref_pyr, corr_pyr, cu(disp), cu(coords), cu(noisy), r))
    if levels == 2:
        vol = O.CombinedGeoEncodingVolume.__new__(O.CombinedGeoEncodingVolume)
        vol.num_levels, vol.radius = 2, 4
        vol.geo_volume_pyramid = [host(g) for g in ref_pyr]
        vol.init_corr_pyramid = [host(c) for c in corr_pyr]
        assert rel_max_err(host(got), vol(disp, coords, noisy)) < VOL_TOL


@pytest.mark.parametrize("shape", [(1, 96, 12, 312), (2, 16, 3, 37), (1, 8, 2, 64), (1, 5, 1, 2)])
def test_corr1d_allpairs_pooled_epilogue(ops, shape):
    """Level 1 of the correlation pyramid written from the accumulators == avg_pool2d([1,2]) of the stored level 0
    (bit-exact), and level 0 is unchanged by the extra output."""
    B, C, H, W = shape
    f1, f2 = synth.normal(shape, 431), synth.normal(shape, 432)
    plain = ops.corr1d_allpairs(cu(f1), cu(f2))
    out, pooled = ops.corr1d_allpairs(cu(f1), cu(f2), return_pooled=True)
    assert torch.equal(out, plain) and pooled.shape == (B, H, W, 1, W // 2)
    assert torch.equal(pooled.reshape(B * H * W, 1, 1, W // 2), ops.avgpool_w2(out.reshape(B * H * W, 1, 1, W)))
    assert rel_max_err(host(out), O.corr1d_allpairs(f1, f2).reshape(out.shape)) < VOL_TOL


@pytest.mark.parametrize("shape", [(1, 96, 4, 312), (2, 40, 3, 136), (1, 32, 2, 320), (1, 8, 1, 8), (1, 104, 2, 200),
                                   (3, 96, 5, 312)])
def test_corr1d_allpairs_tcgen05_vs_fp64(ops, shape):
    """a14 on tcgen05 / TMEM (csrc/allpairs_tcgen05.cu): 3xTF32 against a float64 einsum — 1e-5 of the range (the
    north_star bound is 1e-4; plain TF32 would sit at ~1e-3) — with partial 128-row / 160-column tiles, a partial
    32-channel chunk, several persistent rounds per CTA, and the pooled level written from the same accumulators."""
    B, C, H, W = shape
    f1, f2 = synth.normal(shape, 441), synth.normal(shape, 442)
    out, pooled = ops.corr1d_allpairs(cu(f1), cu(f2), return_pooled=True)
    want = np.einsum("aijk,aijh->ajkh", f1.astype(np.float64), f2.astype(np.float64))
    got = host(out).reshape(want.shape).astype(np.float64)
    assert np.abs(got - want).max() / np.abs(want).max() < 1e-5
    wp = 0.5 * (want[..., 0::2] + want[..., 1::2])
    gp = host(pooled).reshape(wp.shape).astype(np.float64)
    assert np.abs(gp - wp).max() / np.abs(wp).max() < 1e-5
    again = ops.corr1d_allpairs(cu(f1), cu(f2))
    assert torch.equal(again, out)                       # deterministic, and independent of the pooled output


def test_kitti15_geo_class_matches_oracle(ops):
    """The drop-in class (packed pyramid inside) against the oracle class, both call conventions; the reference-layout
    attribute is still available."""
    from diffuvolume_b200 import kitti15
    B, Cf, h, w, Cg, D = 1, 32, 6, 44, 8, 48
    f1, f2 = synth.normal((B, Cf, h, w), 211), synth.normal((B, Cf, h, w), 212)
    geo = synth.normal((B, Cg, D, h, w), 213)
    disp = synth.uniform((B, 1, h, w), 214, dtype=np.float32) * np.float32(47)
    coords = np.broadcast_to(np.arange(w, dtype=np.float32).reshape(1, 1, 1, w), (B, 1, h, w)).copy()
    noisy = synth.uniform((B, D, h, w), 215, dtype=np.float32)
    vol = O.CombinedGeoEncodingVolume(f1, f2, geo, 2, 4)
    fn = kitti15.Combined_Geo_Encoding_Volume(cu(f1), cu(f2), cu(geo), num_levels=2, radius=4)
    assert rel_max_err(host(fn(cu(disp), cu(coords))), vol(disp, coords)) < VOL_TOL
    assert rel_max_err(host(fn(cu(disp), cu(coords), cu(noisy))), vol(disp, coords, noisy)) < VOL_TOL
    # the per-step filter cache: same tensor object -> cached product; in-place update or a new tensor -> recomputed
    nz = cu(noisy)
    first = fn(cu(disp), cu(coords), nz)
    assert torch.equal(fn(cu(disp), cu(coords), nz), first)
    nz.mul_(0.5)
    assert rel_max_err(host(fn(cu(disp), cu(coords), nz)), vol(disp, coords, noisy * np.float32(0.5))) < VOL_TOL
    assert rel_max_err(host(fn(cu(disp), cu(coords), cu(noisy))), vol(disp, coords, noisy)) < VOL_TOL
    assert [tuple(g.shape) for g in fn.geo_volume_pyramid] == [(B * h * w, Cg, 1, D), (B * h * w, Cg, 1, D // 2)]
    assert rel_max_err(host(fn.geo_volume_pyramid[1]), vol.geo_volume_pyramid[1].reshape(B * h * w, Cg, 1, D // 2)) < 1e-6


@pytest.mark.parametrize("shape", [(2, 64, 8, 48, 27, 240), (1, 96, 8, 48, 12, 312), (1, 320, 40, 24, 48, 156)])
def test_gwc_volume_bf16_is_the_rounded_fp32_volume(ops, shape):
    """bf16 output = the fp32 kernel's result rounded to nearest-even once (bit-exact), hence within 2^-8 relative of
    the oracle per element; the zero region stays exactly zero."""
    B, C, G, D, H, W = shape
    ref, tgt = synth.normal((B, C, H, W), 401), synth.normal((B, C, H, W), 402)
    f32 = ops.gwc_volume(cu(ref), cu(tgt), D, G)
    b16 = ops.gwc_volume(cu(ref), cu(tgt), D, G, out_dtype=torch.bfloat16)
    assert b16.dtype == torch.bfloat16 and b16.shape == f32.shape
    assert torch.equal(b16, f32.to(torch.bfloat16))
    want = O.build_gwc_volume(ref, tgt, D, G)
    got = b16.float().cpu().numpy()
    # stated bf16 tolerance: 2^-8 relative per element (+ the fp32 kernel's own 1e-6 * max floor near zero crossings)
    assert np.all(np.abs(got - want) <= 2.0 ** -8 * np.abs(want) + 1e-6 * np.abs(want).max())
    for d in range(1, min(D, W)):
        assert not got[:, :, d, :, :d].any()


@pytest.mark.parametrize("mask_left", [False, True])
@pytest.mark.parametrize("shape", [(2, 32, 48, 27, 240), (1, 12, 24, 48, 156)])
def test_weighted_producer_bf16_is_the_rounded_fp32_volume(ops, shape, mask_left):
    B, C, D, H, W = shape
    ref, tgt = synth.normal((B, C, H, W), 411), synth.normal((B, C, H, W), 412)
    att = ops.att_softmax(cu(synth.normal((B, 1, D, H, W), 413)))
    n = ops.filter_factor(cu(synth.normal((B, D, H, W), 414, dtype=np.float64)), None, 1.0)
    for aw, nn in ((None, None), (att, None), (att, n)):
        f32 = ops.concat_volume_weighted(cu(ref), cu(tgt), D, mask_left=mask_left, att_weights=aw, n=nn)
        b16 = ops.concat_volume_weighted(cu(ref), cu(tgt), D, mask_left=mask_left, att_weights=aw, n=nn,
                                         out_dtype=torch.bfloat16)
        assert b16.dtype == torch.bfloat16 and torch.equal(b16, f32.to(torch.bfloat16))


def test_bf16_volumes_reject_unsupported_layouts(ops):
    ref = cu(synth.normal((1, 16, 6, 20), 421))          # 4 channels per group: not a reference configuration
    with pytest.raises(Exception):
        ops.gwc_volume(ref, ref, 6, 4, out_dtype=torch.bfloat16)
    with pytest.raises(Exception):
        ops.gwc_volume(ref, ref, 6, 2, out_dtype=torch.float16)


# ------------------------------------------------------------------------------------------------
# BASELINE.json's full sizes: size-independent properties (the oracle cannot run these in seconds)
# ------------------------------------------------------------------------------------------------
def test_full_size_batch_independence_and_linearity(ops):
    """configs[1] at B = 8 (gwc C=320 G=40 D=48 @135x240, concat/ACV/filter C=32, softmax-regress over [B,192,540,960]):
    every op is per-sample, so sample b of the batched launch must be bit-identical to a B = 1 launch of that sample
    (different grids, tile schedules and batch strides), and the linear ops must scale exactly by powers of two."""
    g = torch.Generator(device="cuda"); g.manual_seed(77)
    B, h, w, D = 8, 135, 240, 48
    rn = lambda *sh: torch.randn(*sh, generator=g, device="cuda")
    fl, fr = rn(B, 320, h, w), rn(B, 320, h, w)
    vol = ops.gwc_volume(fl, fr, D, 40)
    for b in (0, 5, 7):
        assert torch.equal(vol[b:b + 1], ops.gwc_volume(fl[b:b + 1].contiguous(), fr[b:b + 1].contiguous(), D, 40))
    assert torch.equal(ops.gwc_volume(fl * 2, fr * 4, D, 40), vol * 8)          # bilinear, exact in binary fp
    for d in (1, 17, 47):
        assert not vol[:, :, d, :, :d].any()                                    # the zero region is exact zeros
    del vol
    cl, cr, att = rn(B, 32, h, w), rn(B, 32, h, w), rn(B, 1, D, h, w)
    xt = torch.randn(B, D, h, w, generator=g, device="cuda", dtype=torch.float64)
    aw, n = ops.att_softmax(att), ops.filter_factor(xt, None, 1.0)
    filt = ops.concat_volume_weighted(cl, cr, D, mask_left=False, att_weights=aw, n=n)
    for b in (0, 3, 7):
        one = ops.concat_volume_weighted(cl[b:b + 1].contiguous(), cr[b:b + 1].contiguous(), D, mask_left=False,
                                         att_weights=aw[b:b + 1].contiguous(), n=n[b:b + 1].contiguous())
        assert torch.equal(filt[b:b + 1], one)
    # the regenerated filtered volume == filter multiply applied to the materialised ACV volume (same rounding order)
    ac = ops.concat_volume(cl, cr, D, mask_left=False, att_logits=att)
    assert torch.equal(ops.volume_filter(ac, xt, None), filt)
    del ac, filt
    cost = rn(B, 192, 540, 960) * 4.0
    r = ops.softmax_regress(cost, want_unc=True)
    for b in (0, 6):
        rb = ops.softmax_regress(cost[b:b + 1].contiguous(), want_unc=True)
        assert torch.equal(r["disp"][b:b + 1], rb["disp"]) and torch.equal(r["unc"][b:b + 1], rb["unc"])
    assert float(r["disp"].min()) >= 0.0 and float(r["disp"].max()) <= 191.0     # an expectation over d in [0, 191]
    # softmax is shift-invariant: adding a per-pixel constant changes nothing beyond rounding
    r2 = ops.softmax_regress(cost + rn(B, 1, 540, 960))
    assert float((r2["disp"] - r["disp"]).abs().max()) < 1e-2


# ------------------------------------------------------------------------------------------------
# loud failure on CPU tensors (no fallback)
# ------------------------------------------------------------------------------------------------
# ------------------------------------------------------------------------------------------------
# f3: warp
# ------------------------------------------------------------------------------------------------
def _warp_inputs(shape, seed, amp):
    x = synth.normal(shape, seed)
    disp = synth.uniform((shape[0], 1, shape[2], shape[3]), seed + 1, dtype=np.float32) * np.float32(amp) - np.float32(3)
    return x, disp


@pytest.mark.parametrize("case", list(WARP_CASES))
def test_warp_golden(ops, golden, case):
    shape, seed, amp = WARP_CASES[case]
    x, disp = _warp_inputs(shape, seed, amp)
    got = host(ops.warp(cu(x), cu(disp)))
    want = golden["k12.warp." + case]
    assert np.array_equal(got == 0, want == 0)          # the validity mask agrees everywhere
    assert np.abs(got - want).max() < 5e-5


def test_warp_fullres_vs_oracle_and_autograd(ops):
    x, disp = _warp_inputs((1, 32, 24, 1248), 67, 150.0)
    got = host(ops.warp(cu(x), cu(disp)))
    want = O.warp(x, disp)
    assert np.array_equal(got == 0, want == 0)
    assert np.abs(got - want).max() < 2e-4               # |x| ~ 4 sigma, ix up to 1248: fp32 coordinate rounding
    # the autograd Function is kernel-backed (dv_warp_bwd_f32; parity in tests/test_gpu_backward.py)
    from diffuvolume_b200 import functional as Fn
    xs, ds = _warp_inputs((1, 4, 6, 20), 69, 10.0)
    xt, dt = cu(xs).requires_grad_(True), cu(ds).requires_grad_(True)
    out = Fn.warp(xt, dt)
    out.sum().backward()
    assert xt.grad is not None and dt.grad is not None and torch.isfinite(xt.grad).all() and torch.isfinite(dt.grad).all()
    assert np.abs(host(out) - O.warp(xs, ds)).max() < 1e-5


def test_cpu_tensor_raises(ops):
    from diffuvolume_b200._lib import DvLibraryError
    x = torch.zeros(1, 8, 4, 8)
    with pytest.raises(DvLibraryError):
        ops.gwc_volume(x, x, 4, 2)


def test_concurrent_host_threads_and_streams(ops):
    """nn.DataParallel drives the ops from one host thread per replica (SceneFlow/main.py:67).  Four threads, each on its
    own CUDA stream, hammer the kernels that keep device-side tile counters (the streaming concat producer, the TMA
    softmax-regression) plus gwc; every result must equal the single-threaded one."""
    import threading
    g = torch.Generator(device="cuda"); g.manual_seed(91)
    rn = lambda *s: torch.randn(*s, generator=g, device="cuda")
    B, h, w, D = 2, 27, 240, 48
    jobs = []
    for _ in range(4):
        fl, fr, cl, cr = rn(B, 64, h, w), rn(B, 64, h, w), rn(B, 16, h, w), rn(B, 16, h, w)
        att, cost = ops.att_softmax(rn(B, 1, D, h, w)), rn(B, 192, 4 * 9, 4 * 60) * 4.0
        want = (ops.gwc_volume(fl, fr, D, 8), ops.concat_volume_weighted(cl, cr, D, mask_left=False, att_weights=att),
                ops.softmax_regress(cost)["disp"])
        jobs.append(((fl, fr, cl, cr, att, cost), want))
    torch.cuda.synchronize()
    errors = []

    def worker(args, want):
        try:
            fl, fr, cl, cr, att, cost = args
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                for _ in range(25):
                    got = (ops.gwc_volume(fl, fr, D, 8), ops.concat_volume_weighted(cl, cr, D, mask_left=False, att_weights=att),
                           ops.softmax_regress(cost)["disp"])
                    for a, b in zip(got, want):
                        if not torch.equal(a, b):
                            errors.append("mismatch")
            st.synchronize()
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=j) for j in jobs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors[:3]


@pytest.mark.parametrize("shape,m,G", [((2, 32, 12, 64), 24, 1), ((1, 16, 7, 52), 24, 2), ((1, 8, 5, 20), 9, 1), ((3, 32, 40, 156), 24, 1),
                                       ((1, 6, 3, 7), 3, 1)])
def test_refine_input_assemble_fills_a_concat_buffer(ops, shape, m, G):
    """PWCNet's refinement input without torch.cat (pwcnet_ddim.py:493-499): `left - warp(right)`, the copy of `left` and
    the +-m volume land directly in channel slices of one concat buffer (batch-strided outputs), bit-identical to
    warp -> subtraction -> build_corrleation_volume -> cat; untouched channels stay untouched."""
    B, C, H, W = shape
    fl, fr = synth.normal(shape, 451), synth.normal(shape, 452)
    disp = (synth.uniform((B, 1, H, W), 453, dtype=np.float32) * np.float32(min(W, 40)) - np.float32(2)).astype(np.float32)
    warped = ops.warp(cu(fr), cu(disp))
    want_corr = ops.corr_volume_2sided(cu(fl), warped, m, G).reshape(B, G * (2 * m + 1), H, W)
    S = G * (2 * m + 1)
    buf = torch.full((B, 2 * C + 5 + S, H, W), float("nan"), device="cuda")
    w2, corr = ops.refine_input_assemble(cu(fl), cu(fr), cu(disp), m, G, corr_out=buf[:, 2 * C + 5:], diff_out=buf[:, :C],
                                         copy_out=buf[:, C:2 * C])
    assert corr.data_ptr() == buf[:, 2 * C + 5:].data_ptr() and torch.equal(w2, warped)
    want = torch.cat((cu(fl) - warped, cu(fl), buf[:, 2 * C:2 * C + 5], want_corr), 1)
    assert torch.equal(torch.nan_to_num(buf, nan=-7.0), torch.nan_to_num(want, nan=-7.0))
    assert torch.isnan(buf[:, 2 * C:2 * C + 5]).all()
    ref = O.build_corrleation_volume(fl, O.warp(fr, disp), m, G).reshape(B, S, H, W)
    assert rel_max_err(host(buf[:, 2 * C + 5:].contiguous()), ref) < 1e-4
