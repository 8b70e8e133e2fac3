"""Seeded random-shape sweep of every volume op against the oracle: ragged widths (W % 4 != 0), tiny planes, D > W,
odd channel counts per group, batch 1-3 — the shapes that decide between the 128-bit / TMA kernels and their
shape-agnostic fallbacks.  Tolerances as in tests/test_gpu_parity.py (copies bit-exact, volumes 1e-5 of max, 0.01 px)."""
import numpy as np
import pytest
import torch

import synth
from oracle import dv_oracle as O

pytestmark = pytest.mark.gpu
VOL_TOL = 1e-5


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def rel(a, b):
    return float(np.abs(a - b).max() / max(float(np.abs(b).max()), 1e-30))


def _cases(n, seed):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        B = int(rng.integers(1, 4))
        H = int(rng.integers(1, 13))
        W = int(rng.choice([3, 5, 8, 13, 20, 39, 64, 78, 130]))
        D = int(rng.choice([1, 2, 6, 9, 12, 24, 25, 48]))
        cpg = int(rng.choice([1, 2, 3, 4, 8, 12, 16, 32]))
        G = int(rng.integers(1, 4))
        out.append((B, cpg * G, G, D, H, W, seed * 100 + i))
    return out


@pytest.mark.parametrize("case", _cases(24, 7))
def test_fuzz_gwc_corr2_and_backward(case):
    from diffuvolume_b200 import ops
    B, C, G, D, H, W, seed = case
    ref, tgt = synth.normal((B, C, H, W), seed), synth.normal((B, C, H, W), seed + 1)
    want = O.build_gwc_volume(ref, tgt, D, G)
    got = ops.gwc_volume(cu(ref), cu(tgt), D, G).cpu().numpy()
    assert got.shape == want.shape and rel(got, want) < VOL_TOL
    assert np.array_equal(got == 0, want == 0) or rel(got, want) < VOL_TOL       # the zero region is exact zeros
    for d in range(1, min(D, W)):
        assert not got[:, :, d, :, :d].any()
    m = min(D, 6)
    want2 = O.build_corrleation_volume(ref, tgt, m, G)
    got2 = ops.corr_volume_2sided(cu(ref), cu(tgt), m, G).cpu().numpy()
    assert got2.shape == want2.shape and rel(got2, want2) < VOL_TOL
    # backward of gwc against torch autograd of the reference op sequence (float64 on the CPU)
    go = synth.normal(want.shape, seed + 2)
    rt, tt = torch.from_numpy(ref).double().requires_grad_(True), torch.from_numpy(tgt).double().requires_grad_(True)
    vol = rt.new_zeros([B, G, D, H, W])
    for d in range(min(D, W)):
        prod = (rt[:, :, :, d:] * tt[:, :, :, :W - d]) if d > 0 else rt * tt
        vol[:, :, d, :, d:] = prod.view(B, G, C // G, H, W - d).mean(dim=2)
    vol.backward(torch.from_numpy(go).double())
    gr, gt = ops.gwc_volume_bwd(cu(go), cu(ref), cu(tgt), G)
    assert rel(gr.cpu().numpy(), rt.grad.numpy()) < 1e-5 and rel(gt.cpu().numpy(), tt.grad.numpy()) < 1e-5


@pytest.mark.parametrize("case", _cases(16, 11))
def test_fuzz_concat_filter_and_regression(case):
    from diffuvolume_b200 import ops
    B, C, G, D, H, W, seed = case
    Cc = max(1, C // 4)
    ref, tgt = synth.normal((B, Cc, H, W), seed), synth.normal((B, Cc, H, W), seed + 1)
    att = synth.normal((B, 1, D, H, W), seed + 2)
    xt = synth.normal((B, D, H, W), seed + 3, dtype=np.float64) * 0.8
    shift = synth.normal((B, D), seed + 4) * np.float32(0.2)
    for mask_left in (False, True):
        plain = ops.concat_volume(cu(ref), cu(tgt), D, mask_left=mask_left).cpu().numpy()
        assert np.array_equal(plain, O.build_concat_volume(ref, tgt, D, mask_left))              # a copy: bit-exact
        want = O.volume_filter(O.acv_attention_volume(att, O.build_concat_volume(ref, tgt, D, mask_left)), xt, shift, 1.0)
        fused = ops.concat_volume(cu(ref), cu(tgt), D, mask_left=mask_left, att_logits=cu(att), xt=cu(xt), shift=cu(shift))
        assert rel(fused.cpu().numpy(), want) < VOL_TOL
        two = ops.volume_filter(ops.concat_volume(cu(ref), cu(tgt), D, mask_left=mask_left, att_logits=cu(att)), cu(xt), cu(shift))
        assert rel(two.cpu().numpy(), want) < VOL_TOL
    # softmax + regression over D (+ uncertainty), any D / plane size
    cost = synth.normal((B, D, H, W), seed + 5) * np.float32(3)
    p = O.softmax(cost, 1)
    want_d = O.disparity_regression(p, D)
    r = ops.softmax_regress(cu(cost), return_prob=True, want_unc=True)
    assert np.abs(r["disp"].cpu().numpy() - want_d).max() < 1e-3
    assert rel(r["prob"].cpu().numpy(), p) < 1e-5
    assert np.abs(r["unc"].cpu().numpy() - O.uncertainty(want_d, p)).max() < 1e-3


@pytest.mark.parametrize("case", _cases(12, 13))
def test_fuzz_warp_context_upsample_patch(case):
    from diffuvolume_b200 import ops
    B, C, G, D, H, W, seed = case
    H, W = max(H, 2), max(W, 2)
    x = synth.normal((B, C, H, W), seed)
    disp = synth.uniform((B, 1, H, W), seed + 1, dtype=np.float32) * np.float32(W) - np.float32(2)
    assert rel(ops.warp(cu(x), cu(disp)).cpu().numpy(), O.warp(x, disp)) < 1e-5
    low = synth.uniform((B, 1, H, W), seed + 2, dtype=np.float32) * np.float32(100)
    wts = synth.uniform((B, 9, 4 * H, 4 * W), seed + 3, dtype=np.float32)
    assert np.array_equal(ops.context_upsample(cu(low), cu(wts)).cpu().numpy(), O.context_upsample(low, wts))
    vol = synth.normal((B, 40, min(D, 3), H, W), seed + 4)
    wp, wl = synth.normal((40, 9), seed + 5), synth.normal((40, 9), seed + 6)
    got = ops.acv_patch_volume(cu(vol), cu(wp), cu(wl[:8]), cu(wl[8:24]), cu(wl[24:])).cpu().numpy()
    assert rel(got, O.acv_patch_volume(vol, wp, wl)) < 1e-5


@pytest.mark.parametrize("case", [(1, 8, 48, 3, 17, 21), (2, 4, 24, 2, 9, 22), (1, 3, 10, 4, 6, 23), (2, 8, 16, 1, 33, 24)])
def test_fuzz_geo_pyramid_lookup(case):
    from diffuvolume_b200 import kitti15
    B, Cg, D, h, w, seed = case
    f1, f2 = synth.normal((B, 12, h, w), seed), synth.normal((B, 12, h, w), seed + 1)
    geo = synth.normal((B, Cg, D, h, w), seed + 2)
    disp = synth.uniform((B, 1, h, w), seed + 3, dtype=np.float32) * np.float32(D + 8) - np.float32(4)
    coords = np.broadcast_to(np.arange(w, dtype=np.float32).reshape(1, 1, 1, w), (B, 1, h, w)).copy()
    noisy = synth.uniform((B, D, h, w), seed + 4, dtype=np.float32)
    for radius in (4, 2):
        vol = O.CombinedGeoEncodingVolume(f1, f2, geo, 2, radius)
        fn = kitti15.Combined_Geo_Encoding_Volume(cu(f1), cu(f2), cu(geo), num_levels=2, radius=radius)
        assert rel(fn(cu(disp), cu(coords)).cpu().numpy(), vol(disp, coords)) < VOL_TOL
        assert rel(fn(cu(disp), cu(coords), cu(noisy)).cpu().numpy(), vol(disp, coords, noisy)) < VOL_TOL
