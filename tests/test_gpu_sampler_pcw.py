"""PCWNet tier-2 drop-ins (diffuvolume_b200.sampler: q_sample, predict_noise_from_start, pcw_model_predictions,
pcw_ddim_sample), bound by diffuvolume_b200.install — the same table it applies to the reference's PWCNet_ddim — onto
tests/pcw_mock.py:MockPCW and replayed against the trace that the REFERENCE's own PWCNet_ddim.model_predictions /
ddim_sample (KITTI12/models/pwcnet_ddim.py:466-602) produced on the same mock (tests/golden/make_golden.py:
_pcw_sampler_trace): same stand-in conv modules, same injected noise; filter, softmax-regression, warp, +-24 correlation
volume, x_start, pred_noise, uncertainty vote, DDIM update and cumulative re-noising run on the CUDA kernels."""
import numpy as np
import pytest
import torch

import synth
from oracle import dv_oracle as O
from pcw_mock import PCW_TRACE, MockPCW, pcw_trace_inputs

pytestmark = pytest.mark.gpu


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _sample(a, n=4096):
    flat = a.reshape(-1)
    return flat[np.linspace(0, flat.size - 1, min(n, flat.size)).astype(np.int64)]


@pytest.fixture(scope="module")
def golden():
    from pathlib import Path
    return np.load(Path(__file__).resolve().parent / "golden" / "kitti12.npz")


@pytest.fixture()
def bound(golden):
    import types

    from diffuvolume_b200 import install as dvi
    inp = pcw_trace_inputs("cuda")
    net = MockPCW(O.Schedule(), inp["shifts"]).cuda()
    done = dvi.install("kitti12", modules={"models.pwcnet_ddim": types.SimpleNamespace(PWCNet_ddim=MockPCW)})
    assert "models.pwcnet_ddim.PWCNet_ddim.model_predictions" in done and "models.pwcnet_ddim.PWCNet_ddim.ddim_sample" in done
    yield net, inp, cu(golden["pcw.asd"])
    dvi.uninstall()


def test_pcw_model_predictions_replays_reference_steps(bound, golden):
    """Each step from the reference's own input state: filter -> softmax/regression -> warp -> +-24 correlation ->
    x_start -> pred_noise, against what the reference's model_predictions returned."""
    net, inp, asd = bound
    for i, t in enumerate(PCW_TRACE["times"]):
        img = cu(golden[f"pcw.img.{i}"])
        tc = torch.full((PCW_TRACE["B"],), t, dtype=torch.long, device="cuda")
        eps, x0, disp, prob = net.model_predictions(inp["volume"], img, tc, inp["fl"], inp["fr"])
        assert np.abs(disp.cpu().numpy() - golden[f"pcw.disp.{i}"]).max() < 1e-2            # disparity: 0.01 px
        assert np.abs(_sample(prob.cpu().numpy()) - golden[f"pcw.prob.{i}"]).max() < 1e-5
        # the 2-tap x_start is discontinuous where the quarter-res disparity crosses an integer: compare away from there
        close = np.abs(x0.cpu().numpy() - golden[f"pcw.x0.{i}"]) < 5e-3
        assert close.mean() > 0.995
        eps_ref = golden[f"pcw.eps.{i}"]
        ok = np.abs(eps.cpu().numpy() - eps_ref) < 1e-6 * np.abs(eps_ref) + 0.2      # sqrt_recip(999) = 2e4 amplifies 1e-5
        assert (ok | ~close).mean() > 0.995
        assert eps.dtype == torch.float64 and x0.dtype == torch.float32


def test_pcw_ddim_sample_replays_the_reference_trace(bound, golden, monkeypatch):
    net, inp, asd = bound
    k = {"n": 0}
    seen = {"img": []}

    def randn_like(x, **kw):
        seed = 5100 + k["n"]; k["n"] += 1
        return cu(synth.normal(tuple(x.shape), seed, dtype=np.float64)).to(kw.get("dtype", x.dtype))

    def randn(*shape, **kw):
        shape = tuple(shape[0]) if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)) else tuple(shape)
        return cu(synth.normal(shape, 5000))

    orig_mp = MockPCW.model_predictions

    def mp(self, volume, img, t, fl, fr):
        seen["img"].append(img.detach().clone())
        return orig_mp(self, volume, img, t, fl, fr)

    monkeypatch.setattr(torch, "randn_like", randn_like)
    monkeypatch.setattr(torch, "randn", randn)
    monkeypatch.setattr(MockPCW, "model_predictions", mp)
    final, prob = net.ddim_sample(inp["volume"], inp["used"], asd, inp["fl"], inp["fr"])
    assert k["n"] == len(golden["pcw.randn_like_seeds"]) == 4         # (randn_like(img), randn_like(asd)) x 2 steps
    assert len(seen["img"]) == 3
    # the sampler state entering every step: DDIM update, renewal mask and cumulative re-noising of the previous step
    for i in range(3):
        got, want = seen["img"][i].double().cpu().numpy(), golden[f"pcw.img.{i}"].astype(np.float64)
        close = np.abs(got - want) < 1e-4 + 1e-6 * np.abs(want)
        assert close.mean() > 0.99, (i, close.mean())      # a pixel on a vote threshold may take the other branch
    want = golden["pcw.final"]
    assert tuple(final.shape) == tuple(want.shape)
    err = np.abs(final.cpu().numpy() - want)
    assert (err < 1e-2).mean() > 0.99 and err.max() < 1.0
    assert np.abs(_sample(prob.cpu().numpy()) - golden["pcw.prob.2"]).max() < 1e-3


def test_pcw_ddim_sample_probability_free_steps_equal_the_materialised_form(bound, golden, monkeypatch):
    """With our own model_predictions bound, pcw_ddim_sample writes pred3_volume on the LAST step only and takes the other
    steps' uncertainty from the logits (dv_softmax_uncertainty_vote_f32); a user-supplied model_predictions (here: a plain
    wrapper, which is what the trace test above installs) makes it fall back to the reference's materialised form.  Same
    injected noise -> the two forms must agree: same returned probability volume, same ensemble up to vote-threshold ties."""
    from diffuvolume_b200 import _lib
    net, inp, asd = bound

    def run(wrapped):
        k = {"n": 0}

        def randn_like(x, **kw):
            seed = 5100 + k["n"]; k["n"] += 1
            return cu(synth.normal(tuple(x.shape), seed, dtype=np.float64)).to(kw.get("dtype", x.dtype))

        def randn(*shape, **kw):
            shape = tuple(shape[0]) if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)) else tuple(shape)
            return cu(synth.normal(shape, 5000))

        with monkeypatch.context() as mp:
            mp.setattr(torch, "randn_like", randn_like)
            mp.setattr(torch, "randn", randn)
            if wrapped:
                orig = MockPCW.model_predictions
                mp.setattr(MockPCW, "model_predictions", lambda self, *a: orig(self, *a))
            n0 = _lib.launch_count()
            final, prob = net.ddim_sample(inp["volume"], inp["used"], asd.clone(), inp["fl"], inp["fr"])
            return final, prob, _lib.launch_count() - n0

    f_ours, p_ours, n_ours = run(False)
    f_user, p_user, n_user = run(True)
    assert n_ours == n_user                       # same number of launches: the vote kernel just reads logits instead of prob
    assert torch.equal(p_ours, p_user)            # the returned (last-step) probability volume is the same kernel output
    err = (f_ours - f_user).abs()
    assert float((err < 1e-3).float().mean()) > 0.995 and float(err.max()) < 1.0
    want = golden["pcw.final"]
    e2 = np.abs(f_ours.cpu().numpy() - want)
    assert (e2 < 1e-2).mean() > 0.99 and e2.max() < 1.0
