"""Combined_Geo_Encoding_Volume under autograd: the training forward back-propagates through the lookup
into the geometry volume and both feature maps (KITTI15/core/igev_stereo_ddim.py:402,443; train_stereo.py:122).  The
differentiable path of diffuvolume_b200.kitti15 (matmul / avg_pool1d / gather) is compared with the reference classes —
values and gradients — where the reference tree is mounted; on the GPU it is compared with the forward-only kernels."""
import os
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

import synth

REF = Path(os.environ.get("DV_REFERENCE", "/root/reference"))


def _case(dev):
    B, C, D, h, w, Cf = 1, 8, 48, 6, 40, 16
    t = lambda a: torch.from_numpy(a).to(dev)
    leaves = lambda: [t(synth.normal((B, Cf, h, w), 801)).requires_grad_(True), t(synth.normal((B, Cf, h, w), 802)).requires_grad_(True),
                      t(synth.normal((B, C, D, h, w), 803)).requires_grad_(True)]
    disp = t(synth.uniform((B, 1, h, w), 804, dtype=np.float32) * np.float32(20))
    coords = torch.arange(w, dtype=torch.float32, device=dev).view(1, 1, 1, w).expand(B, 1, h, w).contiguous()
    noisy = t(synth.uniform((B, D, h, w), 805, dtype=np.float32))
    return leaves, disp, coords, noisy


@pytest.mark.skipif(not REF.exists(), reason="reference tree not mounted")
@pytest.mark.parametrize("use_noise", [True, False])
def test_autograd_path_matches_the_reference_class(use_noise):
    from diffuvolume_b200.kitti15 import Combined_Geo_Encoding_Volume as Ours
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("core", "models", "utils")}
    for k in saved:
        del sys.modules[k]
    sys.path.insert(0, str(REF / "KITTI15"))
    try:
        if use_noise:
            from core.geometry_ddim import Combined_Geo_Encoding_Volume as Ref
        else:
            from core.geometry import Combined_Geo_Encoding_Volume as Ref
        leaves, disp, coords, noisy = _case("cpu")
        a, b = leaves(), leaves()
        r, o = Ref(*a), Ours(*b)
        yr = r(disp, coords, noisy) if use_noise else r(disp, coords)
        yo = o(disp, coords, noisy if use_noise else None)
        assert yo.shape == yr.shape and yo.dtype == yr.dtype
        assert float((yr - yo).abs().max()) < 1e-4 * float(yr.abs().max())
        g = torch.from_numpy(synth.normal(tuple(yr.shape), 806))
        yr.backward(g)
        yo.backward(g)
        for x, y in zip(a, b):
            assert y.grad is not None
            assert float((x.grad - y.grad).abs().max()) < 1e-4 * float(x.grad.abs().max())
        assert [tuple(t.shape) for t in o.geo_volume_pyramid] == [tuple(t.shape) for t in r.geo_volume_pyramid]
        assert [tuple(t.shape) for t in o.init_corr_pyramid] == [tuple(t.shape) for t in r.init_corr_pyramid]
    finally:
        sys.path.remove(str(REF / "KITTI15"))
        for k in [k for k in sys.modules if k.split(".")[0] in ("core", "models", "utils")]:
            del sys.modules[k]
        sys.modules.update(saved)


def test_no_grad_construction_never_takes_the_autograd_path():
    """Without a GPU the kernel path raises (no CPU fallback) — which proves the no-grad branch does not silently run ATen."""
    from diffuvolume_b200._lib import DvLibraryError
    from diffuvolume_b200.kitti15 import Combined_Geo_Encoding_Volume as Ours
    leaves, disp, coords, noisy = _case("cpu")
    with torch.no_grad(), pytest.raises(DvLibraryError):
        Ours(*leaves())


@pytest.mark.gpu
@pytest.mark.parametrize("use_noise", [True, False])
def test_autograd_path_agrees_with_the_kernels_on_the_gpu(use_noise):
    from diffuvolume_b200.kitti15 import Combined_Geo_Encoding_Volume as Ours
    leaves, disp, coords, noisy = _case("cuda")
    n = noisy if use_noise else None
    with torch.no_grad():
        want = Ours(*[t.detach() for t in leaves()])(disp, coords, n)
    lv = leaves()
    vol = Ours(*lv)
    assert vol._autograd
    got = vol(disp, coords, n)
    assert got.requires_grad
    assert float((got - want).abs().max()) < 1e-4 * float(want.abs().max())
    got.sum().backward()
    assert all(t.grad is not None and torch.isfinite(t.grad).all() for t in lv)
