"""Tier-3 drop-ins (`forward`) and tier-2 sampler methods bound by `install()` onto the REAL, unmodified reference
models — SceneFlow/models/acv_ddim.py:ACVNet_DDIM and acv.py:ACVNet, random-init, full conv stacks — and compared with
the reference's own forward on the same inputs, weights and RNG stream (|dEPE| <= 0.01 px).

There is no GPU in the authoring container, so the CUDA ops are swapped for oracle-backed CPU stand-ins
(tests/cpu_ops_shim.py): what is under test here is the HOST side of the drop-in — module call order, tensor plumbing,
dtype promotions, RNG draws, train / eval branches, freeze / attention-only switches.  Skipped where the reference tree
is not mounted (the GPU box); the kernels behind the real `ops` are pinned to the same oracle by the `-m gpu` tests.
"""
import importlib
import os
import sys
import warnings
from pathlib import Path

import numpy as np
import pytest
import torch

REF = Path(os.environ.get("DV_REFERENCE", "/root/reference"))
pytestmark = pytest.mark.skipif(not REF.exists(), reason="reference tree not mounted")

PKGS = ("models", "core", "utils", "datasets")


@pytest.fixture()
def sceneflow(monkeypatch):
    """The reference's SceneFlow sub-project imported for real, `.cuda()` -> identity, ops -> oracle-backed stand-ins."""
    import cpu_ops_shim
    from diffuvolume_b200 import functional, install as dvi, sampler
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in PKGS}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, str(REF / "SceneFlow"))
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    monkeypatch.setattr(sampler, "ops", cpu_ops_shim)
    monkeypatch.setattr(functional, "ops", cpu_ops_shim)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            acv_ddim = importlib.import_module("models.acv_ddim")
            acv = importlib.import_module("models.acv")
        yield acv_ddim, acv, dvi
    finally:
        dvi.uninstall()
        sys.path.remove(str(REF / "SceneFlow"))
        for k in [k for k in sys.modules if k.split(".")[0] in PKGS]:
            del sys.modules[k]
        sys.modules.update(saved)


def _inputs(B=1, H=64, W=128, seed=3):
    g = torch.Generator().manual_seed(seed)
    left, right = torch.randn(B, 3, H, W, generator=g), torch.randn(B, 3, H, W, generator=g)
    used = torch.rand(B, H, W, generator=g) * 60.0
    disp_q = torch.rand(B, 1, H // 4, W // 4, generator=g) * 15.0
    return left, right, used, disp_q


def _run(fn, seed):
    torch.manual_seed(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return fn()


def test_acvnet_ddim_eval_forward_matches_the_reference_model(sceneflow):
    acv_ddim, _, dvi = sceneflow
    torch.manual_seed(0)
    net = acv_ddim.ACVNet_DDIM(192, False, False).eval()
    left, right, used, disp_q = _inputs()
    with torch.no_grad():
        ref = _run(lambda: net(left, right, used, disp_q, None), 11)
    rng_after_ref = torch.rand(1)
    done = dvi.install("sceneflow")
    assert "models.acv_ddim.ACVNet_DDIM.forward" in done and "models.acv_ddim.ACVNet_DDIM.ddim_sample" in done
    with torch.no_grad():
        ours = _run(lambda: net(left, right, used, disp_q, None), 11)
    rng_after_ours = torch.rand(1)
    assert isinstance(ours, list) and len(ours) == len(ref) == 1
    assert ours[0].shape == ref[0].shape and ours[0].dtype == ref[0].dtype
    err = (ours[0] - ref[0]).abs()
    assert float(err.mean()) <= 0.01, float(err.mean())
    assert float((err < 0.01).float().mean()) > 0.99
    assert torch.equal(rng_after_ref, rng_after_ours)               # the same number of RNG draws, of the same kinds


def test_acvnet_ddim_eval_forward_with_mask_gt_and_reference_ddim_sample(sceneflow):
    """mask_gt path of the initial x_start, and the non-regenerate path: tier 3 without tier 2 leaves the reference's own
    ddim_sample bound -> the drop-in must hand it a materialised ac_volume."""
    acv_ddim, _, dvi = sceneflow
    from diffuvolume_b200 import sampler
    torch.manual_seed(0)
    net = acv_ddim.ACVNet_DDIM(192, False, False).eval()
    left, right, used, disp_q = _inputs(seed=5)
    mask_gt = (torch.rand(1, 1, 16, 32, generator=torch.Generator().manual_seed(9)) > 0.3).float()
    with torch.no_grad():
        ref = _run(lambda: net(left, right, used, disp_q, mask_gt), 12)
    dvi.install("sceneflow", tier2=False)
    dvi._bind(acv_ddim.ACVNet_DDIM, "forward", sampler.acv_ddim_forward)      # forward only: sampler stays the reference's
    with torch.no_grad():
        ours = _run(lambda: net(left, right, used, disp_q, mask_gt), 12)
    err = (ours[0] - ref[0]).abs()
    assert float(err.mean()) <= 0.01, float(err.mean())


def test_acvnet_ddim_training_forward_matches_the_reference_model(sceneflow):
    acv_ddim, _, dvi = sceneflow
    torch.manual_seed(0)
    net = acv_ddim.ACVNet_DDIM(192, False, False).train()
    left, right, used, disp_q = _inputs(B=2, seed=7)
    with torch.no_grad():                                             # BatchNorm runs in batch-stat mode both times
        ref = _run(lambda: net(left, right, None, disp_q, None), 13)
        dvi.install("sceneflow")
        ours = _run(lambda: net(left, right, None, disp_q, None), 13)
    assert len(ours) == len(ref) == 4
    for a, b in zip(ours, ref):
        assert a.shape == b.shape
        assert float((a - b).abs().max()) < 2e-2 and float((a - b).abs().mean()) < 1e-3


@pytest.mark.parametrize("attn_only,freeze,train", [(False, False, False), (False, False, True), (True, False, True),
                                                    (True, False, False), (False, True, True)])
def test_acvnet_forward_matches_the_reference_model(sceneflow, attn_only, freeze, train):
    _, acv, dvi = sceneflow
    torch.manual_seed(0)
    net = acv.ACVNet(192, attn_only, freeze)
    net.train(train)
    left, right, _, _ = _inputs(B=2 if train else 1, seed=8)
    with torch.no_grad():
        ref = _run(lambda: net(left, right), 14)
        dvi.install("sceneflow")
        ours = _run(lambda: net(left, right), 14)
    assert len(ours) == len(ref)
    for a, b in zip(ours, ref):
        assert a.shape == b.shape
        assert float((a - b).abs().max()) < 2e-2 and float((a - b).abs().mean()) < 1e-3


@pytest.fixture()
def kitti12(monkeypatch):
    import cpu_ops_shim
    from diffuvolume_b200 import functional, install as dvi, sampler
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in PKGS}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, str(REF / "KITTI12"))
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    # warp (KITTI12/models/submodule.py:147): x.get_device() is -1 on the CPU, which torch.arange rejects
    monkeypatch.setattr(torch.Tensor, "get_device", lambda self: self.device)
    monkeypatch.setattr(sampler, "ops", cpu_ops_shim)
    monkeypatch.setattr(functional, "ops", cpu_ops_shim)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            mod = importlib.import_module("models.pwcnet_ddim")
        yield mod, dvi
    finally:
        dvi.uninstall()
        sys.path.remove(str(REF / "KITTI12"))
        for k in [k for k in sys.modules if k.split(".")[0] in PKGS]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_pwcnet_ddim_eval_forward_matches_the_reference_model(kitti12):
    """install('kitti12') on the real PWCNet_ddim: tier-1 names + q_sample / predict_noise_from_start /
    model_predictions / ddim_sample (pwcnet_ddim.py:453-602) against the unmodified model, same weights, inputs, RNG."""
    mod, dvi = kitti12
    torch.manual_seed(0)
    net = mod.PWCNet_ddimgc(192).eval()
    left, right, used, disp_q = _inputs(B=1, H=64, W=128, seed=21)
    with torch.no_grad():
        ref, ref3 = _run(lambda: net(left, right, used, disp_q, None), 15)
    rng_ref = torch.rand(1)
    done = dvi.install("kitti12")
    assert "models.pwcnet_ddim.PWCNet_ddim.model_predictions" in done
    with torch.no_grad():
        ours, ours3 = _run(lambda: net(left, right, used, disp_q, None), 15)
    rng_ours = torch.rand(1)
    assert ours[0].shape == ref[0].shape and ours3[0].shape == ref3[0].shape
    # a random-init refinenet3 is numerically degenerate (|disp_finetune| ~ 1e11 here): compare relative to the range;
    # the sampler state behind it (x_start, pred_noise, re-noised img, probability volume) agrees exactly
    err = (ours[0] - ref[0]).abs()
    assert float(err.max() / ref[0].abs().max()) < 1e-4
    assert float((ours3[0] - ref3[0]).abs().max()) < 1e-5              # the returned probability volume
    assert torch.equal(rng_ref, rng_ours)
