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
