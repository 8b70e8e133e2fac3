"""The torch-CPU baseline port (oracle/torch_port.py) against the numpy oracle — it is the timed CPU
baseline of bench.py, so it has to compute the same thing."""
import numpy as np
import torch

import synth
from conftest import rel_max_err
from oracle import dv_oracle as O
from oracle import torch_port as P

t = lambda a: torch.from_numpy(np.ascontiguousarray(a))


def test_volumes():
    ref, tgt = synth.normal((2, 16, 6, 20), 1), synth.normal((2, 16, 6, 20), 2)
    assert rel_max_err(P.gwc_volume(t(ref), t(tgt), 12, 4).numpy(), O.build_gwc_volume(ref, tgt, 12, 4)) < 1e-6
    for ml in (False, True):
        assert np.array_equal(P.concat_volume(t(ref), t(tgt), 12, ml).numpy(), O.build_concat_volume(ref, tgt, 12, ml))
    att = synth.normal((2, 1, 12, 6, 20), 3)
    cat = O.build_concat_volume(ref, tgt, 12, False)
    assert rel_max_err(P.acv_volume(t(att), t(cat)).numpy(), O.acv_attention_volume(att, cat)) < 1e-6


def test_regression_vote_xstart():
    cost = synth.normal((2, 192, 16, 32), 4) * np.float32(4)
    disp, prob = P.softmax_regress(t(cost))
    d_o, p_o = O.softmax_regress(cost, 192)
    assert np.abs(disp.numpy() - d_o).max() < 1e-3
    used = d_o + np.float32(0.5)
    vote = P.renewal_vote(disp, t(used), prob, 1.0, 1e9).numpy()
    assert vote.mean() > 0.99
    x0 = P.xstart_from_pred(disp).numpy()
    np.testing.assert_allclose(x0, O.xstart_from_disp(O.disp_to_quarter(d_o, 192), 48, 1.0), atol=2e-3)


def test_hot_path_pair_matches_oracle_trace():
    B, H, W, C, G, Cc, D = 1, 32, 64, 32, 4, 8, 48
    h, w = H // 4, W // 4
    fl, fr = synth.normal((B, C, h, w), 1), synth.normal((B, C, h, w), 2)
    cl, cr = synth.normal((B, Cc, h, w), 3), synth.normal((B, Cc, h, w), 4)
    att = synth.normal((B, 1, D, h, w), 5)
    costs = [synth.normal((B, 192, H, W), 10 + i) * np.float32(4) for i in range(5)]
    used = synth.uniform((B, H, W), 6, dtype=np.float32) * np.float32(191)
    disp_q = synth.uniform((B, h, w), 7, dtype=np.float32) * np.float32(47.75)
    shifts = [synth.normal((B, D), 20 + i) * np.float32(0.1) for i in range(5)]
    sn = [synth.normal((B, D, h, w), 30 + i, dtype=np.float64).astype(np.float32 if i == 0 else np.float64) for i in range(4)]
    rz = [synth.uniform((B, D, h, w), 40 + i, dtype=np.float64) for i in range(4)]
    sched = O.Schedule()
    asd = O.xstart_from_disp(disp_q, D, 1.0)
    vol = O.acv_attention_volume(att, O.build_concat_volume(cl, cr, D, False))
    pred_o, _ = O.ddim_sample_acv(sched, vol, used, asd, lambda tt: shifts[[999, 799, 599, 399, 199].index(tt)],
                                  lambda v, i: costs[i], sn, rz)
    pred_p, (gwc, _, _) = P.hot_path_pair(t(fl), t(fr), t(cl), t(cr), t(att), [t(c) for c in costs], t(used), t(asd),
                                       [t(s) for s in shifts], [t(s) for s in sn], [t(r) for r in rz], sched, G=G)
    assert np.abs(pred_p.numpy() - pred_o).max() < 1e-3
    assert rel_max_err(gwc.numpy(), O.build_gwc_volume(fl, fr, D, G)) < 1e-6


def test_pcw_and_igev_pieces_match_the_oracle():
    """warp, the two-sided correlation volume, the geometry class and context_upsample of the port (the gpu_aten_baseline /
    parity checker of bench.py's pcwnet and igev legs) against the numpy oracle, which is pinned to the reference fixtures."""
    x, d = synth.normal((2, 6, 9, 40), 61), (synth.uniform((2, 1, 9, 40), 62, dtype=np.float32) * np.float32(30))
    assert np.abs(P.warp(t(x), t(d)).numpy() - O.warp(x, d)).max() < 1e-5
    ref, tgt = synth.normal((1, 8, 3, 20), 63), synth.normal((1, 8, 3, 20), 64)
    assert rel_max_err(P.corr_volume_2sided(t(ref), t(tgt), 6, 2).numpy(), O.build_corrleation_volume(ref, tgt, 6, 2)) < 1e-6
    B, C, D, h, w, Cf = 1, 8, 48, 4, 24, 16
    f1, f2, geo = synth.normal((B, Cf, h, w), 65), synth.normal((B, Cf, h, w), 66), synth.normal((B, C, D, h, w), 67)
    disp = synth.uniform((B, 1, h, w), 68, dtype=np.float32) * np.float32(20)
    coords = np.broadcast_to(np.arange(w, dtype=np.float32).reshape(1, 1, 1, w), (B, 1, h, w)).copy()
    noisy = synth.uniform((B, D, h, w), 69, dtype=np.float32)
    got = P.GeoEncodingVolume(t(f1), t(f2), t(geo))(t(disp), t(coords), t(noisy)).numpy()
    want = O.CombinedGeoEncodingVolume(f1, f2, geo)(disp, coords, noisy)
    assert got.shape == want.shape and rel_max_err(got, want) < 1e-5
    low, wts = synth.normal((1, 1, 5, 7), 70), synth.uniform((1, 9, 20, 28), 71, dtype=np.float32)
    assert np.abs(P.context_upsample(t(low), t(wts)).numpy() - O.context_upsample(low, wts)).max() < 1e-5


def test_pcw_and_igev_pairs_run_and_are_deterministic():
    sched3, sched2 = O.Schedule(sampling_timesteps=3), O.Schedule(sampling_timesteps=2)
    B, H, W, D = 1, 32, 64, 48
    h, w = H // 4, W // 4
    n = lambda s, seed, dt=np.float32: t(synth.normal(s, seed, dtype=np.float64).astype(dt))
    scales = [(n((B, 16, H // s, W // s), 80 + s), n((B, 16, H // s, W // s), 90 + s), n((B, 4, H // s, W // s), 100 + s),
               n((B, 4, H // s, W // s), 110 + s), D * 4 // s) for s in (4, 8)]
    args = dict(scales=scales, combine=n((B, 8, D, h, w), 120), costs=[n((B, 192, H, W), 121) * 4], used=t(synth.uniform((B, H, W), 122, dtype=np.float32) * np.float32(190)),
                feat_l_full=n((B, 8, H, W), 123), feat_r_full=n((B, 8, H, W), 124), start=n((B, D, h, w), 125),
                asd=t(O.xstart_from_disp(synth.uniform((B, h, w), 126, dtype=np.float32) * np.float32(47), D, 1.0)),
                shifts=[n((B, D), 127 + i) * 0.1 for i in range(3)],
                step_noises=[n((B, D, h, w), 130 + i, np.float32 if i == 0 else np.float64) for i in range(2)],
                q_noises=[n((B, D, h, w), 140 + i) for i in range(2)])
    p1, extra = P.pcw_hot_path_pair(**args, sched=sched3, G=4)
    p2, _ = P.pcw_hot_path_pair(**args, sched=sched3, G=4)
    assert p1.shape == (B, H, W) and torch.equal(p1, p2) and extra[0].dtype == torch.float32 and extra[3].shape == (B, 49, H, W)
    coords = torch.arange(w, dtype=torch.float32).view(1, 1, 1, w).expand(B, 1, h, w).contiguous()
    ig = dict(fmap_l=n((B, 24, h, w), 150), fmap_r=n((B, 24, h, w), 151), geo=n((B, 8, D, h, w), 152), cost48=n((B, D, h, w), 153),
              up_weights=torch.softmax(n((B, 9, H, W), 154), 1), coords=coords, used=t(synth.uniform((B, H, W), 155, dtype=np.float32) * np.float32(47)),
              start=n((B, D, h, w), 156), asd=args["asd"], shifts=[n((B, D), 157 + i) * 0.1 for i in range(2)],
              step_noises=[n((B, D, h, w), 160)], q_noises=[n((B, D, h, w), 161)])
    q1, ex = P.igev_hot_path_pair(**ig, sched=sched2, iters=3, G=2)
    assert q1.shape == (B, H, W) and ex[2].shape == (B, 162, h, w) and torch.isfinite(q1).all()
