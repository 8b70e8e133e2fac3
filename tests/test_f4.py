"""context_upsample (KITTI15/core/submodule.py:241-253) and the ACVNet patch convolutions
(SceneFlow/models/acv_ddim.py:181-188,377-381).  CPU: the oracle against fixtures minted from the reference
(tests/golden/make_golden.py f4).  GPU: the kernels against the oracle, the fixtures and torch's own modules."""
from pathlib import Path

import numpy as np
import pytest
import torch

import synth
from oracle import dv_oracle as O

GOLD = Path(__file__).resolve().parent / "golden" / "f4.npz"
CTX_CASES = {"a": (2, 6, 10), "b": (1, 24, 78)}
PATCH_CASES = {"a": (1, 3, 9, 14), "b": (1, 4, 27, 60)}


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def _ctx_inputs(name):
    B, h, w = CTX_CASES[name]
    low = synth.uniform((B, 1, h, w), 301, dtype=np.float32) * np.float32(190)
    wts = synth.normal((B, 9, 4 * h, 4 * w), 302)
    wts = (np.exp(wts) / np.exp(wts).sum(1, keepdims=True)).astype(np.float32)
    return low, wts, synth.normal((B, 4 * h, 4 * w), 303)


def _sample(a, n=4096):
    flat = a.reshape(-1)
    return flat[np.linspace(0, flat.size - 1, min(n, flat.size)).astype(np.int64)]


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


# ---------------------------------------------------------------------------------------------- CPU: oracle pinning
@pytest.mark.parametrize("name", list(CTX_CASES))
def test_oracle_context_upsample_matches_reference(gold, name):
    low, wts, _ = _ctx_inputs(name)
    assert np.array_equal(O.context_upsample(low, wts), gold[f"f4.ctxup.{name}"])


@pytest.mark.parametrize("name", list(PATCH_CASES))
def test_oracle_patch_chain_matches_reference(gold, name):
    B, D, H, W = PATCH_CASES[name]
    vol = synth.normal((B, 40, D, H, W), 311)
    first = O.depthwise_conv3x3(vol, gold["f4.patch.w_patch"], 1)
    res = O.acv_patch_volume(vol, gold["f4.patch.w_patch"], gold["f4.patch.w_l"])
    if name == "a":
        assert rel(first, gold["f4.patch.a.first"]) < 1e-6 and rel(res, gold["f4.patch.a"]) < 1e-6
    else:
        assert rel(_sample(first), gold["f4.patch.b.first"]) < 1e-6 and rel(_sample(res), gold["f4.patch.b"]) < 1e-6


def test_oracle_patch_intermediate_is_zero_padded():
    """The second stencil pads ITS input with zeros: a 1x1 plane sees only the centre taps of both kernels."""
    vol = np.full((1, 2, 1, 1, 1), 3.0, np.float32)
    w1, w2 = synth.normal((2, 9), 1), synth.normal((2, 9), 2)
    got = O.depthwise_conv3x3(O.depthwise_conv3x3(vol, w1, 1), w2, 2)
    assert np.allclose(got[0, :, 0, 0, 0], 3.0 * w1[:, 4] * w2[:, 4], rtol=1e-6)


# ---------------------------------------------------------------------------------------------- GPU
cu = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CTX_CASES))
def test_context_upsample_golden_and_grads(gold, name):
    from diffuvolume_b200 import kitti15
    low, wts, gout = _ctx_inputs(name)
    lt, wt = cu(low).requires_grad_(True), cu(wts).requires_grad_(True)
    res = kitti15.context_upsample(lt, wt)
    assert res.shape == gold[f"f4.ctxup.{name}"].shape and res.dtype == torch.float32
    assert np.array_equal(res.detach().cpu().numpy(), gold[f"f4.ctxup.{name}"])      # same tap order: bit-exact
    res.backward(cu(gout))
    assert rel(lt.grad.cpu().numpy(), gold[f"f4.ctxup.{name}.glow"]) < 1e-5
    gw = wt.grad.cpu().numpy()
    assert rel(_sample(gw) if name == "b" else gw, gold[f"f4.ctxup.{name}.gw"]) < 1e-6


@pytest.mark.gpu
def test_context_upsample_igev_size_vs_oracle_and_half_weights():
    from diffuvolume_b200 import kitti15, ops
    B, h, w = 2, 96, 312
    low = synth.uniform((B, 1, h, w), 321, dtype=np.float32) * np.float32(190)
    wts = synth.uniform((B, 9, 4 * h, 4 * w), 322, dtype=np.float32)
    got = ops.context_upsample(cu(low), cu(wts))
    assert np.array_equal(got.cpu().numpy(), O.context_upsample(low, wts))
    # linearity in disp_low (size-independent property)
    got2 = ops.context_upsample(cu(low * np.float32(2)), cu(wts))
    assert torch.equal(got2, got * 2)
    # autocast call site (igev_stereo_ddim.py:209): fp16 weights, fp32 disparity -> fp32 result
    half = kitti15.context_upsample(cu(low), cu(wts).half())
    assert half.dtype == torch.float32
    assert rel(half.cpu().numpy(), O.context_upsample(low, wts.astype(np.float16).astype(np.float32))) < 1e-6
    with pytest.raises(RuntimeError):
        ops.context_upsample(cu(low), cu(wts[:, :8]))


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(PATCH_CASES))
def test_patch_chain_golden(gold, name):
    from diffuvolume_b200 import ops
    B, D, H, W = PATCH_CASES[name]
    vol = synth.normal((B, 40, D, H, W), 311)
    wl = gold["f4.patch.w_l"]
    res = ops.acv_patch_volume(cu(vol), cu(gold["f4.patch.w_patch"]), cu(wl[:8]), cu(wl[8:24]), cu(wl[24:])).cpu().numpy()
    first = ops.depthwise3x3_chain(cu(vol), cu(gold["f4.patch.w_patch"]), None, 1).cpu().numpy()
    if name == "a":
        assert rel(first, gold["f4.patch.a.first"]) < 1e-5 and rel(res, gold["f4.patch.a"]) < 1e-5
    else:
        assert rel(_sample(first), gold["f4.patch.b.first"]) < 1e-5 and rel(_sample(res), gold["f4.patch.b"]) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1, 40, 6, 135, 240), (2, 40, 2, 33, 65), (1, 40, 1, 1, 1), (1, 40, 2, 3, 130),
                                   # the row-streaming kernel: 12 / 16 columns per lane, a ragged last lane (244 = 30.5 x 8),
                                   # planes shorter than the prefetch depth and than the dilation
                                   (1, 40, 2, 96, 312), (1, 40, 1, 5, 400), (1, 40, 1, 1, 64), (1, 40, 1, 2, 244),
                                   (2, 40, 1, 7, 68), (1, 40, 1, 3, 512)])
def test_patch_chain_vs_oracle(shape):
    from diffuvolume_b200 import ops
    B, C, D, H, W = shape
    vol = synth.normal(shape, 331)
    wp, wl = synth.normal((40, 9), 332), synth.normal((40, 9), 333)
    want = O.acv_patch_volume(vol, wp, wl)
    got = ops.acv_patch_volume(cu(vol), cu(wp), cu(wl[:8]), cu(wl[8:24]), cu(wl[24:])).cpu().numpy()
    assert rel(got, want) < 1e-5
    # channel slices do not touch their neighbours
    out = torch.full(shape, 7.0, device="cuda")
    ops.depthwise3x3_chain(cu(vol), cu(wp), cu(wl), 1, 2, channels=(8, 24), out=out)
    assert torch.all(out[:, :8] == 7.0) and torch.all(out[:, 24:] == 7.0)
    assert rel(out[:, 8:24].cpu().numpy(), want[:, 8:24]) < 1e-5


class _AcvPatchStandIn(torch.nn.Module):
    """The four module definitions of acv_ddim.py:181-188 and the call chain of :377-381, verbatim in structure."""

    def __init__(self):
        super().__init__()
        nn = torch.nn
        self.patch = nn.Conv3d(40, 40, kernel_size=(1, 3, 3), stride=1, dilation=1, groups=40, padding=(0, 1, 1), bias=False)
        self.patch_l1 = nn.Conv3d(8, 8, kernel_size=(1, 3, 3), stride=1, dilation=1, groups=8, padding=(0, 1, 1), bias=False)
        self.patch_l2 = nn.Conv3d(16, 16, kernel_size=(1, 3, 3), stride=1, dilation=2, groups=16, padding=(0, 2, 2), bias=False)
        self.patch_l3 = nn.Conv3d(16, 16, kernel_size=(1, 3, 3), stride=1, dilation=3, groups=16, padding=(0, 3, 3), bias=False)

    def forward(self, gwc_volume):
        gwc_volume = self.patch(gwc_volume)
        patch_l1 = self.patch_l1(gwc_volume[:, :8])
        patch_l2 = self.patch_l2(gwc_volume[:, 8:24])
        patch_l3 = self.patch_l3(gwc_volume[:, 24:40])
        return torch.cat((patch_l1, patch_l2, patch_l3), dim=1)


@pytest.mark.gpu
def test_fuse_acv_patch_drop_in_matches_cudnn_and_keeps_state_dict():
    from diffuvolume_b200 import _lib, install
    torch.manual_seed(0)
    model = _AcvPatchStandIn().cuda().eval()
    keys = sorted(model.state_dict())
    vol = torch.randn(1, 40, 12, 54, 96, device="cuda")
    torch.backends.cudnn.allow_tf32 = False
    with torch.no_grad():
        want = model(vol)
    assert install.fuse_acv_patch(model)
    try:
        assert sorted(model.state_dict()) == keys
        n0 = _lib.launch_count()
        with torch.no_grad():
            got = model(vol)
        assert _lib.launch_count() - n0 == 3                      # one launch per dilation class
        assert float((got - want).abs().max() / want.abs().max()) < 1e-5
        # nn.DataParallel replicas are shallow copies of the instance: the fused forward must use the REPLICA's parameters
        rep = model.patch._replicate_for_data_parallel()
        rep._parameters["weight"] = torch.zeros_like(model.patch.weight)
        with torch.no_grad():
            assert not rep(vol).any() and model.patch(vol).any()
        # with autograd on the original convolutions run (weights receive gradients)
        n0 = _lib.launch_count()
        model(vol).sum().backward()
        assert _lib.launch_count() == n0 and model.patch.weight.grad is not None
    finally:
        install.uninstall()
    with torch.no_grad():
        assert torch.equal(model(vol), want)
