"""The numpy oracle against fixtures minted from the reference itself (tests/golden/make_golden.py).

This is what pins the oracle: every function of oracle/dv_oracle.py is compared with the output
of the reference's own code on the same seeded inputs.  CPU only.
"""
import numpy as np
import pytest

import synth
from conftest import rel_max_err
from golden.make_golden import CONCAT_CASES, GWC_CASES, WARP_CASES, trace_inputs
from oracle import dv_oracle as O

TOL = 1e-6  # fp32 restatement vs fp32 reference (different summation order only)


@pytest.mark.parametrize("prefix", ["sf", "sftop", "k12", "k15"])
@pytest.mark.parametrize("case", list(GWC_CASES))
def test_gwc_volume(golden, prefix, case):
    B, C, G, D, H, W, seed = GWC_CASES[case]
    ref, tgt = synth.normal((B, C, H, W), seed), synth.normal((B, C, H, W), seed + 1000)
    got = O.build_gwc_volume(ref, tgt, D, G)
    want = golden[f"{prefix}.gwc.{case}"]
    assert rel_max_err(got, want) < TOL
    # zeros for x < d are exact
    for d in range(1, D):
        assert (got[:, :, d, :, : min(d, W)] == 0).all()
        assert (want[:, :, d, :, : min(d, W)] == 0).all()


def test_groupwise_correlation(golden):
    B, C, G, D, H, W, seed = GWC_CASES["cpg12"]
    ref, tgt = synth.normal((B, C, H, W), seed), synth.normal((B, C, H, W), seed + 1000)
    assert rel_max_err(O.groupwise_correlation(ref, tgt, G), golden["sf.groupwise.cpg12"]) < TOL
    with pytest.raises(AssertionError):
        O.groupwise_correlation(ref, tgt, 5)


@pytest.mark.parametrize("prefix,mask_left", [("sf", False), ("k15", False), ("sftop", True), ("k12", True)])
@pytest.mark.parametrize("case", list(CONCAT_CASES))
def test_concat_volume_bit_exact(golden, prefix, mask_left, case):
    B, C, D, H, W, seed = CONCAT_CASES[case]
    ref, tgt = synth.normal((B, C, H, W), seed), synth.normal((B, C, H, W), seed + 1000)
    got = O.build_concat_volume(ref, tgt, D, mask_left)
    assert np.array_equal(got, golden[f"{prefix}.concat.{case}"])


@pytest.mark.parametrize("key,shape,m,G,seed", [
    ("sftop.corr2.m24", (1, 32, 8, 64), 24, 1, 41),
    ("k12.corr2.m24", (1, 32, 8, 64), 24, 1, 41),
    ("k12.corr2.w80_m24", (1, 32, 6, 80), 24, 1, 43),
    ("sftop.corr2.tiny_m9", (2, 8, 3, 7), 9, 2, 42),
])
def test_corr_volume_2sided(golden, key, shape, m, G, seed):
    ref, tgt = synth.normal(shape, seed), synth.normal(shape, seed + 1000)
    got = O.build_corrleation_volume(ref, tgt, m, G)
    want = golden[key]
    assert rel_max_err(got, want) < TOL
    assert np.array_equal(got == 0, want == 0)  # the negative-shift quirk: identical support


def test_acv_attention_volume(golden):
    B, C, D, h, w = 1, 4, 48, 4, 56
    cl, cr = synth.normal((B, C, h, w), 51), synth.normal((B, C, h, w), 1051)
    att = synth.normal((B, 1, D, h, w), 52) * np.float32(3)
    got = O.acv_attention_volume(att, O.build_concat_volume(cl, cr, D, False))
    assert rel_max_err(got, golden["sf.acv_volume"]) < TOL


@pytest.mark.parametrize("prefix", ["sf", "sftop", "k12", "k15"])
@pytest.mark.parametrize("k", [1, 10])
def test_softmax_regression(golden, prefix, k):
    cost = synth.normal((1, 192, 16, 32), 31) * np.float32(k)
    disp, _ = O.softmax_regress(cost, 192)
    want = golden[f"{prefix}.regress.k{k}"]
    if want.ndim == 4:  # KITTI15: keepdim=True
        want = want[:, 0]
    assert np.abs(disp - want).max() < 2e-4  # px; values up to 191


def test_regression_keepdim(golden):
    cost = synth.normal((2, 48, 6, 40), 33) * np.float32(3)
    got = O.disparity_regression(O.softmax(cost, 1), 48, keepdim=True)
    assert got.shape == (2, 1, 6, 40)
    assert np.abs(got - golden["k15.regress.keepdim"]).max() < 1e-4


def test_schedule_constants(golden):
    s = O.Schedule()
    for name in ("betas", "alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
                 "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod"):
        np.testing.assert_allclose(getattr(s, name), golden[f"sched.{name}"], rtol=1e-12, atol=0)
    # SURVEY.md §8c G6 known answers
    np.testing.assert_allclose(s.alphas_cumprod[[999, 799, 599, 399, 199]],
                               [2.4287669e-9, 0.0940456127, 0.3408096398, 0.6474782111, 0.8987059206], rtol=1e-7)
    for S in (5, 3, 2):
        pairs = O.Schedule(sampling_timesteps=S).time_pairs()
        assert np.array_equal(np.array(pairs), golden[f"sched.time_pairs.S{S}"])
    assert O.Schedule(sampling_timesteps=5).time_pairs() == [(999, 799), (799, 599), (599, 399), (399, 199), (199, -1)]
    assert O.Schedule(sampling_timesteps=3).time_pairs() == [(999, 665), (665, 332), (332, -1)]
    # SURVEY.md §8a a12 coefficients
    san, c, sigma = s.ddim_coefficients(999, 799)
    assert abs(sigma - 0.95182) < 1e-5 and abs(c - 1.456e-4) < 1e-6


@pytest.mark.parametrize("t", [999, 599, 0])
def test_q_sample_and_pred_noise(golden, t):
    s = O.Schedule()
    x0 = synth.uniform((2, 48, 4, 8), 61, dtype=np.float32) * 2 - 1
    nz = synth.normal((2, 48, 4, 8), 62)
    qs = O.q_sample(s, x0, t, nz)
    assert qs.dtype == np.float64
    np.testing.assert_allclose(qs, golden[f"sf.q_sample.t{t}"], rtol=1e-13, atol=1e-15)
    pn = O.predict_noise_from_start(s, golden[f"sf.q_sample.t{t}"], t, x0)
    np.testing.assert_allclose(pn, golden[f"sf.pred_noise.t{t}"], rtol=1e-12, atol=1e-12)
    pn32 = O.predict_noise_from_start(s, nz, t, x0)
    np.testing.assert_allclose(pn32, golden[f"sf.pred_noise_f32.t{t}"], rtol=1e-12, atol=1e-12)


def test_xstart_known_answers(golden):
    # SURVEY.md §8c G5 [probed on the reference]
    dq = np.array([0.0, 0.25, 3.0, 46.5, 47.0, 47.75], dtype=np.float32).reshape(1, 1, 6)
    x0 = O.xstart_from_disp(dq, 48, 1.0)[0, :, 0, :]
    vol = (x0 + 1) / 2
    want = [{0: 1.0}, {0: 0.75, 1: 0.25}, {3: 1.0}, {46: 0.5, 47: 0.5}, {47: 1.0}, {47: 1.0}]
    for j, w in enumerate(want):
        ref = np.zeros(48, dtype=np.float32)
        for k, v in w.items():
            ref[k] = v
        np.testing.assert_allclose(vol[:, j], ref, atol=1e-6)
    # and against the reference's inline code (acv_ddim.py:403-419)
    got = O.xstart_from_disp(golden["trace.disp_q"], 48, 1.0)
    assert np.array_equal(got, golden["trace.asd"])


def test_ddim_trace(golden):
    """Full ACVNet_DDIM.ddim_sample trace (reference code, stand-in conv stack, injected noise)."""
    B, Cc, D, h, w, H, W = (int(v) for v in golden["trace.shape"])
    sched = O.Schedule()
    cl, cr = synth.normal((B, Cc, h, w), 71), synth.normal((B, Cc, h, w), 1071)
    att = synth.normal((B, 1, D, h, w), 72) * np.float32(2)
    volume = O.acv_attention_volume(att, O.build_concat_volume(cl, cr, D, False))
    bias, used = trace_inputs(B, D, h, w, H, W)
    asd = O.xstart_from_disp(golden["trace.disp_q"], D, 1.0)

    def cost_fn(vol_f, i):
        c = vol_f.mean(axis=1, keepdims=True, dtype=np.float32) * np.float32(2.0)
        return O.interpolate_trilinear(c + bias[i], (192, H, W))[:, 0]

    rn, ru = golden["trace.randn_like_seeds"], golden["trace.rand_like_seeds"]
    # per non-final step the reference draws randn_like(img), randn_like(asd) [unused], rand_like(asdd)
    step_noises, renoises = [], []
    for i in range(4):
        seed, is64 = rn[2 * i]
        step_noises.append(synth.normal((B, D, h, w), int(seed), dtype=np.float64).astype(np.float64 if is64 else np.float32))
        seed_u, is64_u = ru[i]
        assert is64_u == 1
        renoises.append(synth.uniform((B, D, h, w), int(seed_u), dtype=np.float64))
    assert [int(x[1]) for x in rn[0::2]] == [0, 1, 1, 1]   # fp32 noise on step 1, fp64 afterwards
    trace = {}
    pred, final = O.ddim_sample_acv(sched, volume, used, asd, lambda t: golden[f"trace.shift.t{t}"], cost_fn,
                                    step_noises, renoises, trace=trace)
    for i in range(5):
        assert np.abs(trace["disp"][i] - golden[f"trace.disp.{i}"]).max() < 5e-4, i
        np.testing.assert_allclose(trace["x0"][i], golden[f"trace.x0.{i}"], atol=2e-4)
        np.testing.assert_allclose(trace["eps"][i], golden[f"trace.eps.{i}"], rtol=1e-6, atol=1e-3)
        if i > 0:
            got_img = trace["img"][i - 1]
            assert got_img.dtype == np.float64 and golden[f"trace.img.{i}"].dtype == np.float64
            np.testing.assert_allclose(got_img, golden[f"trace.img.{i}"], atol=1e-3)
    assert golden["trace.img.0"].dtype == np.float32
    # the renewal mask is genuinely mixed in this trace (both DDIM-update and re-noise paths are hit)
    frac = [(m == 0).mean() for m in trace["mask"]]
    assert 0.2 < frac[-1] < 0.8, frac
    assert np.abs(pred - golden["trace.pred"]).max() < 5e-4
    assert np.abs(np.stack(final) - golden["trace.final"]).max() < 5e-4


def test_corr1d_and_geo_lookup(golden):
    B, Cf, h, w, Cg, D = 2, 16, 6, 40, 8, 48
    f1, f2 = synth.normal((B, Cf, h, w), 101), synth.normal((B, Cf, h, w), 102)
    geo = synth.normal((B, Cg, D, h, w), 103)
    disp = synth.uniform((B, 1, h, w), 104, dtype=np.float32) * np.float32(50) - np.float32(2)
    coords = np.broadcast_to(np.arange(w, dtype=np.float32).reshape(1, 1, 1, w), (B, 1, h, w)).copy()
    noisy = synth.uniform((B, D, h, w), 105, dtype=np.float32)
    assert rel_max_err(O.corr1d_allpairs(f1, f2), golden["k15.corr"]) < TOL
    vol = O.CombinedGeoEncodingVolume(f1, f2, geo, 2, 4)
    assert rel_max_err(vol.geo_volume_pyramid[1], golden["k15.geo.pyr1"]) < TOL
    assert rel_max_err(vol.init_corr_pyramid[1], golden["k15.corr.pyr1"]) < TOL
    got_plain = vol(disp, coords)
    assert got_plain.shape == (B, 162, h, w)
    assert rel_max_err(got_plain, golden["k15.geo.plain"]) < 1e-5
    got_ddim = vol(disp, coords, noisy)
    assert rel_max_err(got_ddim, golden["k15.geo.ddim"]) < 1e-5


def test_big_shape_checksums(golden):
    """BASELINE config-2 shapes (B=1): the oracle against checksums of the reference's output."""
    B, C, G, D, H, W = 1, 320, 40, 48, 135, 240
    ref, tgt = synth.normal((B, C, H, W), 91), synth.normal((B, C, H, W), 1091)
    v = O.build_gwc_volume(ref, tgt, D, G)
    s = golden["big.gwc.sum"]
    assert abs(v.sum(dtype=np.float64) - s[0]) < 1e-3 * s[1] * 1e-3
    assert abs(np.abs(v).sum(dtype=np.float64) - s[1]) < 1e-6 * s[1]
    flat = v.reshape(-1)
    idx = np.linspace(0, flat.size - 1, 4096).astype(np.int64)
    assert np.abs(flat[idx] - golden["big.gwc.sample"]).max() < 1e-6 * s[2] + 1e-6
    cost = synth.normal((1, 192, 135, 240), 92) * np.float32(4)
    disp, prob = O.softmax_regress(cost, 192)
    assert np.abs(disp - golden["big.regress.disp"]).max() < 5e-4
    assert np.abs(O.uncertainty(disp, prob) - golden["big.regress.unc"]).max() < 5e-4


@pytest.mark.parametrize("align_corners", [False, True])
@pytest.mark.parametrize("shape,size", [((1, 1, 48, 9, 20), (192, 36, 80)), ((2, 1, 7, 3, 5), (19, 11, 17)),
                                        ((1, 1, 12, 6, 10), (48, 22, 39))])
def test_oracle_trilinear_matches_torch(shape, size, align_corners):
    """F.upsample(..., mode='trilinear') lives in PyTorch (torch 2.11 here; the reference pins 2.0 — README.md:25):
    the oracle's restatement of upsample_trilinear3d is pinned against torch itself on CPU (acv_ddim.py:267,
    pwcnet_ddim.py:480)."""
    import torch
    import torch.nn.functional as F
    x = synth.normal(shape, 77) * np.float32(3)
    want = F.interpolate(torch.from_numpy(x), size=size, mode="trilinear", align_corners=align_corners).numpy()
    got = O.interpolate_trilinear(x, size, align_corners=align_corners)
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("case", list(WARP_CASES))
def test_oracle_warp_golden(golden, case):
    """O.warp against the reference's own warp() (KITTI12/models/submodule.py:137-176) run on the CPU."""
    shape, seed, amp = WARP_CASES[case]
    x = synth.normal(shape, seed)
    disp = synth.uniform((shape[0], 1, shape[2], shape[3]), seed + 1, dtype=np.float32) * np.float32(amp) - np.float32(3)
    got, want = O.warp(x, disp), golden["k12.warp." + case]
    assert np.array_equal(got == 0, want == 0)
    assert np.abs(got - want).max() < 5e-5
