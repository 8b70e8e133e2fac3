"""Host-side logic that needs no GPU: schedule constants, the installer, error conventions, autograd
formulas of the drop-in functions, metric sharding."""
import sys
import types
from pathlib import Path

import numpy as np
import pytest
import torch

import synth
from oracle import dv_oracle as O
from oracle import torch_port as P

REF = Path("/root/reference")


def test_schedule_matches_oracle_and_golden(golden):
    from diffuvolume_b200.pipeline import DdimSchedule
    s, o = DdimSchedule(), O.Schedule()
    np.testing.assert_allclose(s.alphas_cumprod, golden["sched.alphas_cumprod"], rtol=1e-12)
    for S in (5, 3, 2):
        assert DdimSchedule(sampling_timesteps=S).time_pairs() == O.Schedule(sampling_timesteps=S).time_pairs()
        assert np.array_equal(np.array(DdimSchedule(sampling_timesteps=S).time_pairs()), golden[f"sched.time_pairs.S{S}"])
    for t, tn in s.time_pairs()[:-1]:
        np.testing.assert_allclose(s.update_coefficients(t, tn), o.ddim_coefficients(t, tn), rtol=1e-14)
        assert s.sqrt_recip(t) == pytest.approx(golden["sched.sqrt_recip_alphas_cumprod"][t], rel=1e-12)
        assert s.sqrt_recipm1(t) == pytest.approx(golden["sched.sqrt_recipm1_alphas_cumprod"][t], rel=1e-12)
        assert s.sqrt_ac(t) == pytest.approx(golden["sched.sqrt_alphas_cumprod"][t], rel=1e-12)
        assert s.sqrt_1m_ac(t) == pytest.approx(golden["sched.sqrt_one_minus_alphas_cumprod"][t], rel=1e-12)


def test_reference_error_conventions_on_the_mirrors():
    from diffuvolume_b200 import kitti12, kitti15, sceneflow
    from diffuvolume_b200._lib import DvLibraryError
    x = torch.zeros(1, 10, 4, 8)
    for m in (sceneflow, kitti12, kitti15):
        with pytest.raises(AssertionError):          # assert C % num_groups == 0 (submodule.py:211)
            m.build_gwc_volume(x, x, 4, 3)
        with pytest.raises(AssertionError):          # assert len(x.shape) == 4 (submodule.py:174)
            m.disparity_regression(torch.zeros(1, 4, 8), 4)
        with pytest.raises(DvLibraryError):          # CPU tensors: loud failure, no fallback
            m.build_gwc_volume(x, x, 4, 2)
        with pytest.raises(DvLibraryError):
            m.build_concat_volume(x, x, 4)
    with pytest.raises(AssertionError):
        kitti12.build_corrleation_volume(x, x, 2, 3)


def test_installer_with_stand_in_modules():
    import diffuvolume_b200.install as dvi
    from diffuvolume_b200 import sampler, sceneflow
    sub = types.ModuleType("models.submodule")
    sub.build_gwc_volume = lambda *a: "ref-gwc"
    sub.build_concat_volume = lambda *a: "ref-concat"
    sub.disparity_regression = lambda *a: "ref-reg"
    sub.groupwise_correlation = lambda *a: "ref-gc"
    sub.convbn = "untouched"
    consumer = types.ModuleType("models.acv_ddim")
    for k, v in vars(sub).items():            # what `from models.submodule import *` does
        if not k.startswith("__"):
            setattr(consumer, k, v)

    class ACVNet_DDIM:
        def q_sample(self, x, t, noise=None):
            return "ref"

        def predict_noise_from_start(self, x_t, t, x0):
            return "ref"

        def model_predictions(self, volume, noise, t):
            return "ref"

        def ddim_sample(self, volume, used, asd):
            return "ref"

    consumer.ACVNet_DDIM = ACVNet_DDIM
    mods = {"models.submodule": sub, "models.acv_ddim": consumer}
    done = dvi.install("sceneflow", modules=mods)
    assert "models.acv_ddim.build_gwc_volume" in done and "models.acv_ddim.ACVNet_DDIM.ddim_sample" in done
    assert consumer.build_gwc_volume is sceneflow.build_gwc_volume
    assert sub.disparity_regression is sceneflow.disparity_regression
    assert consumer.convbn == "untouched"
    assert ACVNet_DDIM.ddim_sample is sampler.acv_ddim_sample
    n = dvi.uninstall()
    assert n == len(done)
    assert consumer.build_gwc_volume() == "ref-gwc" and ACVNet_DDIM().ddim_sample(0, 0, 0) == "ref"
    with pytest.raises(ValueError):
        dvi.install("nope")


@pytest.mark.skipif(not (REF / "SceneFlow").exists(), reason="reference checkout not present (authoring container only)")
def test_installer_on_the_real_reference():
    """Import the reference's SceneFlow package, install, and check that its own consumer modules now hold
    our functions (calling them with CPU tensors then fails loudly — no silent fallback)."""
    import subprocess
    code = r'''
import sys, torch
sys.path.insert(0, "/root/reference/SceneFlow"); sys.path.insert(0, "%s")
import models.acv_ddim as M, models.acv as A
import diffuvolume_b200.install as dvi
from diffuvolume_b200 import sceneflow, sampler
from diffuvolume_b200._lib import DvLibraryError
done = dvi.install("sceneflow")
assert M.build_gwc_volume is sceneflow.build_gwc_volume and A.build_concat_volume is sceneflow.build_concat_volume
assert M.ACVNet_DDIM.ddim_sample is sampler.acv_ddim_sample
x = torch.zeros(1, 8, 4, 8)
try:
    M.build_gwc_volume(x, x, 4, 2)
    raise SystemExit("expected a loud failure on CPU tensors")
except DvLibraryError:
    pass
dvi.uninstall()
assert M.build_gwc_volume.__module__ == "models.submodule"
print("OK", len(done))
''' % str(Path(__file__).resolve().parent.parent)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0 and "OK" in r.stdout, r.stderr[-2000:]


def test_shard_range_covers_everything():
    from diffuvolume_b200.distributed import shard_range
    for n in (0, 1, 7, 8, 4370):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_fuse_acv_patch_is_transparent_on_cpu_and_restores_classes():
    """install.fuse_acv_patch re-classes the four Conv3d instances; off the GPU (or with autograd on) the original
    convolutions run, state_dict keys are untouched, and uninstall() restores the original classes."""
    import torch
    from diffuvolume_b200 import install

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            nn = torch.nn
            self.patch = nn.Conv3d(40, 40, kernel_size=(1, 3, 3), stride=1, dilation=1, groups=40, padding=(0, 1, 1), bias=False)
            self.patch_l1 = nn.Conv3d(8, 8, kernel_size=(1, 3, 3), stride=1, dilation=1, groups=8, padding=(0, 1, 1), bias=False)
            self.patch_l2 = nn.Conv3d(16, 16, kernel_size=(1, 3, 3), stride=1, dilation=2, groups=16, padding=(0, 2, 2), bias=False)
            self.patch_l3 = nn.Conv3d(16, 16, kernel_size=(1, 3, 3), stride=1, dilation=3, groups=16, padding=(0, 3, 3), bias=False)

        def forward(self, v):
            v = self.patch(v)
            return torch.cat((self.patch_l1(v[:, :8]), self.patch_l2(v[:, 8:24]), self.patch_l3(v[:, 24:40])), dim=1)

    torch.manual_seed(0)
    net = Net().eval()
    x = torch.randn(1, 40, 2, 6, 9)
    with torch.no_grad():
        want = net(x)
    keys, classes = sorted(net.state_dict()), [type(m) for m in (net.patch, net.patch_l1, net.patch_l2, net.patch_l3)]
    assert install.fuse_acv_patch(net)
    try:
        assert sorted(net.state_dict()) == keys
        assert isinstance(net.patch, torch.nn.Conv3d) and type(net.patch) is not classes[0]
        with torch.no_grad():
            assert torch.equal(net(x), want)                     # CPU tensors: the cuDNN/ATen path, no CUDA needed
    finally:
        install.uninstall()
    assert [type(m) for m in (net.patch, net.patch_l1, net.patch_l2, net.patch_l3)] == classes
    assert not install.fuse_acv_patch(torch.nn.Linear(2, 2))     # nothing to fuse


def test_cpulist_parsing_and_numa_helpers_are_safe_without_a_gpu():
    """Host placement helpers of distributed.py: the sysfs CPU-list grammar, and no exception (just None / empty info) on a
    box without the device."""
    from diffuvolume_b200.distributed import _parse_cpulist, bind_to_gpu_numa_node, gpu_numa_info
    assert _parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert _parse_cpulist("") == []
    info = gpu_numa_info(0)
    assert set(info) == {"pci", "numa_node", "local_cpus"}
    if not torch.cuda.is_available():
        assert info["pci"] is None and bind_to_gpu_numa_node(0) is None
