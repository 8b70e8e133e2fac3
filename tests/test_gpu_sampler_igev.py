"""IGEV tier-2 drop-ins (diffuvolume_b200.sampler.igev_*) bound onto tests/igev_mock.py:MockIGEV and replayed against the
trace that the REFERENCE's own IGEVStereo_ddim.model_predictions / ddim_sample produced on the same mock
(tests/golden/make_golden.py:_igev_sampler_trace): same stand-in GRU block, same injected noise, the reference's
Combined_Geo_Encoding_Volume replaced by the CUDA lookup.  Runs on the GPU box."""
import numpy as np
import pytest
import torch

import synth
from igev_mock import IGEV_TRACE, MockIGEV, igev_trace_inputs
from oracle import dv_oracle as O

pytestmark = pytest.mark.gpu


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.fixture()
def bound(golden):
    from diffuvolume_b200 import sampler
    from diffuvolume_b200.kitti15 import Combined_Geo_Encoding_Volume
    inp = igev_trace_inputs("cuda")
    net = MockIGEV(O.Schedule(), inp["shifts"]).cuda()
    for name, fn in (("q_sample", sampler.q_sample), ("predict_noise_from_start", sampler.predict_noise_from_start),
                     ("model_predictions", sampler.igev_model_predictions), ("ddim_sample", sampler.igev_ddim_sample)):
        setattr(MockIGEV, name, fn)
    geo_fn = Combined_Geo_Encoding_Volume(inp["f1"], inp["f2"], inp["geo"], radius=4, num_levels=2)
    return net, geo_fn, inp, cu(golden["igev.asd"])


def test_igev_model_predictions_replays_reference_steps(bound, golden):
    net, geo_fn, inp, asd = bound
    for i, t in enumerate(IGEV_TRACE["times"]):
        img = cu(golden[f"igev.img.{i}"])
        tc = torch.full((IGEV_TRACE["B"],), t, dtype=torch.long, device="cuda")
        coords1_in = inp["init_disp"] if i == 0 else cu(golden[f"igev.coords1.{i - 1}"])
        eps, x0, disp, coords1 = net.model_predictions(inp["init_disp"], coords1_in, None, IGEV_TRACE["iters"], [], [], geo_fn,
                                                       img, tc, None)
        assert eps.dtype == torch.float64 and x0.dtype == torch.float32
        assert np.abs(disp.cpu().numpy() - golden[f"igev.disp.{i}"]).max() < 2e-3
        assert np.abs(coords1.cpu().numpy() - golden[f"igev.coords1.{i}"]).max() < 1e-3
        # the 2-tap x_start is discontinuous where the quarter-res disparity crosses an integer: compare away from there
        x0_ref = golden[f"igev.x0.{i}"]
        close = np.abs(x0.cpu().numpy() - x0_ref) < 5e-3
        assert close.mean() > 0.999
        eps_ref = golden[f"igev.eps.{i}"]
        ok = np.abs(eps.cpu().numpy() - eps_ref) < 1e-6 * np.abs(eps_ref) + 0.2    # sqrt_recip(999) = 2e4 amplifies x0's 1e-5
        assert (ok | ~close).mean() > 0.999


def test_igev_ddim_sample_replays_the_reference_trace(bound, golden, monkeypatch):
    net, geo_fn, inp, asd = bound
    k = {"n": 0}

    def randn_like(x, **kw):
        seed = 3000 + k["n"]; k["n"] += 1
        return cu(synth.normal(tuple(x.shape), seed, dtype=np.float64)).to(kw.get("dtype", x.dtype))

    monkeypatch.setattr(torch, "randn_like", randn_like)
    pred = net.ddim_sample(inp["init_disp"], inp["init_disp"], None, IGEV_TRACE["iters"], [], [], geo_fn, inp["used"], asd, None)
    assert k["n"] == len(golden["igev.randn_like_seeds"]) == 3      # start state + (randn_like(img), randn_like(asd)) once
    want = golden["igev.pred"]
    assert tuple(pred.shape) == tuple(want.shape)
    err = np.abs(pred.cpu().numpy() - want)
    # the fallback |disp - used| < 3 is a hard switch: a pixel within 2e-3 of the threshold may take the other branch
    assert (err < 2e-3).mean() > 0.999 and err.max() < 3.0


def test_filter_factor_pair_matches_oracle():
    from diffuvolume_b200 import ops
    xt = synth.normal((2, 48, 5, 12), 71, dtype=np.float64) * 0.8
    shift = synth.normal((2, 48), 72) * np.float32(0.2)
    n, n32 = ops.filter_factor_pair(cu(xt), cu(shift), 1.0)
    want = O.filter_factor(xt, shift, 1.0)
    assert n.dtype == torch.float64 and n32.dtype == torch.float32
    np.testing.assert_allclose(n.cpu().numpy(), want, rtol=1e-14, atol=1e-15)
    np.testing.assert_allclose(n32.cpu().numpy(), want.astype(np.float32), rtol=1e-6)
    n, n32 = ops.filter_factor_pair(cu(xt.astype(np.float32)), cu(shift), 1.0)
    assert n.dtype == torch.float32 and n32 is n
