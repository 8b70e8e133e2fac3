"""Tier-2 drop-ins (diffuvolume_b200.sampler) bound onto a stand-in for the reference's ACVNet_DDIM and
replayed against the trace the REAL ACVNet_DDIM.ddim_sample produced (tests/golden/make_golden.py): same
stand-in conv stack, same injected noise.  Runs on the GPU box."""
import numpy as np
import pytest
import torch
import torch.nn as nn

import synth
from golden.make_golden import trace_inputs
from oracle import dv_oracle as O

pytestmark = pytest.mark.gpu


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


class ShiftTable(nn.Module):
    """DynamicHead stand-in: noisy + shift[t] (head.py:74-77), shifts taken from the golden fixture."""

    def __init__(self, golden):
        super().__init__()
        self.table = {t: cu(golden[f"trace.shift.t{t}"]) for t in (999, 799, 599, 399, 199)}

    def forward(self, noisy, t):
        s = self.table[int(t.reshape(-1)[0].item())]
        return noisy + s[:, :, None, None]


class Stand0(nn.Module):
    def forward(self, v):
        return v.mean(1, keepdim=True) * 2.0


class Zero(nn.Module):
    def forward(self, v):
        return torch.zeros_like(v)


class Bias(nn.Module):
    def __init__(self, bias):
        super().__init__()
        self.bias, self.n = [cu(b) for b in bias], 0

    def forward(self, v):
        o = v + self.bias[self.n]
        self.n += 1
        return o


class MockACVNetDDIM(nn.Module):
    """Attribute-compatible stand-in for SceneFlow/models/acv_ddim.py:ACVNet_DDIM (sampler-relevant part)."""

    def __init__(self, golden, bias):
        super().__init__()
        s = O.Schedule()
        self.maxdisp, self.scale = 192, 1.0
        self.num_timesteps, self.sampling_timesteps, self.ddim_sampling_eta = 1000, 5, 1
        self.renewal, self.use_ensemble = True, True
        for name in ("alphas_cumprod", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod",
                     "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod"):
            self.register_buffer(name, torch.from_numpy(getattr(s, name)))
        self.time_embedding = ShiftTable(golden)
        self.dres0, self.dres1, self.dres2, self.dres3, self.classif2 = Stand0(), Zero(), nn.Identity(), nn.Identity(), Bias(bias)


@pytest.fixture()
def bound(golden):
    from diffuvolume_b200 import ops, sampler
    B, Cc, D, h, w, H, W = (int(v) for v in golden["trace.shape"])
    bias, used = trace_inputs(B, D, h, w, H, W)
    net = MockACVNetDDIM(golden, bias).cuda()
    for name, fn in (("q_sample", sampler.q_sample), ("predict_noise_from_start", sampler.predict_noise_from_start),
                     ("model_predictions", sampler.acv_model_predictions), ("ddim_sample", sampler.acv_ddim_sample)):
        setattr(MockACVNetDDIM, name, fn)
    cl, cr = synth.normal((B, Cc, h, w), 71), synth.normal((B, Cc, h, w), 1071)
    att = synth.normal((B, 1, D, h, w), 72) * np.float32(2)
    volume = ops.concat_volume(cu(cl), cu(cr), D, mask_left=False, att_logits=cu(att))
    asd = ops.xstart_from_disp(cu(golden["trace.disp_q"]), D, 1.0)
    return net, volume, cu(used), asd


def test_methods_q_sample_and_pred_noise(bound, golden):
    net = bound[0]
    x0 = synth.uniform((2, 48, 4, 8), 61, dtype=np.float32) * 2 - 1
    nz = synth.normal((2, 48, 4, 8), 62)
    for t in (999, 599, 0):
        tt = torch.full((1,), t, dtype=torch.long, device="cuda")
        qs = net.q_sample(cu(x0), tt, cu(nz))
        np.testing.assert_allclose(qs.cpu().numpy(), golden[f"sf.q_sample.t{t}"], rtol=1e-13, atol=1e-15)
        pn = net.predict_noise_from_start(qs, tt, cu(x0))
        np.testing.assert_allclose(pn.cpu().numpy(), golden[f"sf.pred_noise.t{t}"], rtol=1e-11, atol=1e-11)


def test_model_predictions_first_step(bound, golden):
    net, volume, used, asd = bound
    t = torch.full((volume.shape[0],), 999, dtype=torch.long, device="cuda")
    pred_noise, x_start, pred, prob = net.model_predictions(volume, asd, t)
    assert pred_noise.dtype == torch.float64 and x_start.dtype == torch.float32 and prob.shape[1] == 192
    assert np.abs(pred.cpu().numpy() - golden["trace.disp.0"]).max() < 1e-3
    np.testing.assert_allclose(x_start.cpu().numpy(), golden["trace.x0.0"], atol=2e-4)
    np.testing.assert_allclose(pred_noise.cpu().numpy(), golden["trace.eps.0"], rtol=1e-6, atol=1e-3)
    np.testing.assert_allclose(prob.sum(1).cpu().numpy(), 1.0, atol=1e-5)


def test_ddim_sample_replays_the_reference_trace(bound, golden, monkeypatch):
    net, volume, used, asd = bound
    k = {"n": 0}

    def randn_like(x, **kw):
        seed = 2000 + k["n"]; k["n"] += 1
        return cu(synth.normal(tuple(x.shape), seed, dtype=np.float64)).to(kw.get("dtype", x.dtype))

    def rand_like(x, **kw):
        seed = 2000 + k["n"]; k["n"] += 1
        return cu(synth.uniform(tuple(x.shape), seed, dtype=np.float64)).to(kw.get("dtype", x.dtype))

    monkeypatch.setattr(torch, "randn_like", randn_like)
    monkeypatch.setattr(torch, "rand_like", rand_like)
    pred, final = net.ddim_sample(volume, used, asd)
    assert k["n"] == 12                                    # 3 draws per non-final step, like the reference
    assert final.shape == (6, *used.shape)
    assert np.abs(final.cpu().numpy() - golden["trace.final"]).max() < 1e-3
    assert np.abs(pred.cpu().numpy() - golden["trace.pred"]).max() < 1e-3


def test_uncertainty_vote_matches_oracle():
    from diffuvolume_b200 import ops
    cost = synth.normal((2, 192, 12, 20), 5) * np.float32(4)
    disp_o, prob_o = O.softmax_regress(cost, 192)
    refined = disp_o + synth.normal((2, 12, 20), 6) * np.float32(0.7)
    used = refined + (synth.uniform((2, 12, 20), 7, dtype=np.float32) - np.float32(0.5)) * np.float32(3)
    unc_o = O.uncertainty(refined, prob_o)
    thr = float(np.median(unc_o))
    vote, unc = ops.uncertainty_vote(cu(refined), cu(prob_o), cu(used), 1.0, thr, return_unc=True)
    np.testing.assert_allclose(unc.cpu().numpy(), unc_o, atol=1e-3)
    want = O.renewal_vote(refined, used, unc_o, 1.0, thr)
    near = (np.abs(np.abs(refined - used) - 1.0) < 1e-3) | (np.abs(unc_o - thr) < 1e-3)
    assert np.array_equal(vote.cpu().numpy()[~near], want[~near])


@pytest.mark.parametrize("shape", [(2, 192, 12, 20), (1, 192, 7, 9), (2, 48, 6, 10), (1, 96, 5, 6), (1, 200, 3, 5)])
def test_softmax_uncertainty_vote_from_logits_matches_oracle(shape):
    """dv_softmax_uncertainty_vote_f32: the uncertainty / vote of an EXTERNAL disparity against softmax(cost) taken straight
    from the logits (pwcnet_ddim.py:483 + :553-570) — every kernel path (tensor-map, register, generic D > 192, odd H*W),
    with and without `used`, and equal to the two-pass form through a materialised probability volume."""
    from diffuvolume_b200 import ops
    B, D, H, W = shape
    cost = synth.normal(shape, 15) * np.float32(4)
    disp_o, prob_o = O.softmax_regress(cost, D)
    refined = disp_o + synth.normal((B, H, W), 16) * np.float32(0.7)
    used = refined + (synth.uniform((B, H, W), 17, dtype=np.float32) - np.float32(0.5)) * np.float32(3)
    unc_o = O.uncertainty(refined, prob_o)
    thr = float(np.median(unc_o))
    vote, unc = ops.softmax_uncertainty_vote(cu(refined), cu(cost), cu(used), 1.0, thr, return_unc=True)
    np.testing.assert_allclose(unc.cpu().numpy(), unc_o, atol=1e-3)
    want = O.renewal_vote(refined, used, unc_o, 1.0, thr)
    near = (np.abs(np.abs(refined - used) - 1.0) < 1e-3) | (np.abs(unc_o - thr) < 1e-3)
    assert np.array_equal(vote.cpu().numpy()[~near], want[~near])
    # without `used` the vote is the uncertainty test alone
    v2 = ops.softmax_uncertainty_vote(cu(refined), cu(cost), None, 1.0, thr).cpu().numpy()
    near2 = np.abs(unc_o - thr) < 1e-3
    assert np.array_equal(v2[~near2], (unc_o < thr).astype(np.float32)[~near2])
    # two-pass form: softmax_regress(return_prob) + uncertainty_vote
    prob = ops.softmax_regress(cu(cost), return_prob=True)["prob"]
    _, unc2 = ops.uncertainty_vote(cu(refined), prob, cu(used), 1.0, thr, return_unc=True)
    assert float((unc - unc2).abs().max()) < 1e-4
    # the regression's own outputs are untouched by the external-disparity plumbing
    r = ops.softmax_regress(cu(cost), used=cu(used), vote_thresholds=(1.0, thr), want_unc=True)
    np.testing.assert_allclose(r["disp"].cpu().numpy(), disp_o, atol=2e-4)
    np.testing.assert_allclose(r["unc"].cpu().numpy(), O.uncertainty(disp_o, prob_o), atol=1e-3)


@pytest.mark.gpu
def test_hot_path_cuda_graph_replay_matches_eager():
    """The whole ACV hot path captured in a CUDA graph: replays are bit-identical to eager calls, also after the
    inputs were updated in place, and issue no new launches from the host."""
    import torch
    from diffuvolume_b200 import _lib
    from diffuvolume_b200.pipeline import AcvHotPath
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(5)
    B, H, W, D = 2, 64, 128, 48
    h, w = H // 4, W // 4
    rn = lambda *s, dt=torch.float32: torch.randn(*s, generator=g, device=dev, dtype=dt)
    ru = lambda *s, dt=torch.float32: torch.rand(*s, generator=g, device=dev, dtype=dt)
    inp = dict(feat_l=rn(B, 64, h, w), feat_r=rn(B, 64, h, w), cfeat_l=rn(B, 8, h, w), cfeat_r=rn(B, 8, h, w),
               att_logits=rn(B, 1, D, h, w), costs=[rn(B, 192, H, W) * 4.0], used=ru(B, H, W) * 191.0,
               disp_q=ru(B, h, w) * 47.75, shifts=[rn(B, D) * 0.1 for _ in range(5)],
               step_noises=[rn(B, D, h, w, dt=torch.float32 if i == 0 else torch.float64) for i in range(4)],
               renoises=[ru(B, D, h, w, dt=torch.float64) for _ in range(4)])
    path = AcvHotPath(num_groups=8)
    want = {k: v.clone() for k, v in path(**inp).items()}
    replay, out = path.graphed(**inp)
    n0 = _lib.launch_count()
    replay()
    torch.cuda.synchronize()
    assert _lib.launch_count() == n0                       # nothing was issued through the C-ABI: the graph ran
    for k in want:
        assert torch.equal(out[k], want[k]), k
    inp["feat_l"].mul_(0.5); inp["used"].add_(1.0); inp["costs"][0].mul_(-1.0)
    replay()
    torch.cuda.synchronize()
    got = {k: v.clone() for k, v in out.items()}
    want2 = path(**inp)
    for k in want2:
        assert torch.equal(got[k], want2[k]), k
    assert not torch.equal(got["pred"], want["pred"])


@pytest.mark.gpu
@pytest.mark.parametrize("filter_mode,regress_mode", [("regenerate", "logits"), ("volume", "logits"), ("regenerate", "fused_upsample")])
def test_full_size_hot_path_vs_reference_op_sequence_on_cuda(filter_mode, regress_mode):
    """BASELINE.json's headline size (540x960, D=192, C=320/G=40, T=5), one pair: the fused hot path against the
    reference's own op sequence (oracle/torch_port.py — pinned on the CPU against the oracle and the reference-minted
    fixtures) executed on CUDA tensors, which is what the reference does on a GPU box.  North-star gates: volume
    max relative error <= 1e-4, disparity EPE <= 0.01 px; the renewal mask may differ only where a vote sits on a threshold."""
    import torch
    from diffuvolume_b200 import ops
    from diffuvolume_b200.pipeline import AcvHotPath
    from oracle import torch_port as P
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev); g.manual_seed(2024)
    B, H, W, D = 1, 540, 960, 48
    h, w = H // 4, W // 4
    rn = lambda *s, dt=torch.float32: torch.randn(*s, generator=g, device=dev, dtype=dt)
    ru = lambda *s, dt=torch.float32: torch.rand(*s, generator=g, device=dev, dtype=dt)
    fl, fr, cl, cr = rn(B, 320, h, w), rn(B, 320, h, w), rn(B, 32, h, w), rn(B, 32, h, w)
    att = rn(B, 1, D, h, w)
    fused = regress_mode == "fused_upsample"
    costs = [(rn(B, 1, D, h, w) if fused else rn(B, 192, H, W)) * 4.0 for _ in range(5)]
    used = ru(B, H, W) * 191.0
    disp_q = ru(B, h, w) * 47.75
    shifts = [rn(B, D) * 0.1 for _ in range(5)]
    sn = [rn(B, D, h, w, dt=torch.float32 if i == 0 else torch.float64) for i in range(4)]
    rz = [ru(B, D, h, w, dt=torch.float64) for _ in range(4)]
    path = AcvHotPath(filter_mode=filter_mode, regress_mode=regress_mode)
    out = path(fl, fr, cl, cr, att, costs, used, disp_q, shifts, sn, rz, keep_volumes=True)
    asd = ops.xstart_from_disp(disp_q, D, 1.0)
    want_pred, (want_gwc, want_img, want_mask) = P.hot_path_pair(fl, fr, cl, cr, att, costs, used, asd, shifts, sn, rz,
                                                                 sched=O.Schedule(), upsample_to=(192, H, W) if fused else None)
    assert float((out["gwc"] - want_gwc).abs().max() / want_gwc.abs().max()) < 1e-4
    want_ac = P.acv_volume(att, P.concat_volume(cl, cr, D, mask_left=False))
    assert float((out["ac"] - want_ac).abs().max() / want_ac.abs().max()) < 1e-4
    epe = float((out["pred"] - want_pred).abs().mean())
    assert epe < 0.01 and float((out["pred"] - want_pred).abs().max()) < 0.05, epe
    assert float((out["mask"] == want_mask).float().mean()) > 0.999
    close = (out["x_last"].double() - want_img.double()).abs() < 5e-3        # 2-tap x_start: discontinuous at integer crossings
    assert float(close.float().mean()) > 0.999
