"""Bind the sm_100a ops into an imported copy of the reference so its scripts run unchanged.

The reference has no plugin/operator registry: consumers pull the hot-path functions in with
star-imports (`from models.submodule import *`, SceneFlow/models/acv_ddim.py:7; KITTI12/models/pwcnet_ddim.py:8;
`from core.submodule import *` + `from core.geometry_ddim import Combined_Geo_Encoding_Volume`,
KITTI15/core/igev_stereo_ddim.py:7-8).  A star-import copies bindings, so replacing the function in
`models.submodule` after `models.acv_ddim` was imported changes nothing; `install()` therefore
rebinds the names on every consumer module that is already in sys.modules (tier 1), and rebinds the
sampler methods on the reference's model classes (tier 2) so that the unnamed ops between the named
functions (softmax, the filter multiply, the DDIM arithmetic) are fused too.

    import models                      # the reference, cwd = SceneFlow/
    import diffuvolume_b200.install as dvi
    dvi.install("sceneflow")           # or "kitti12", "kitti15"
    ...                                # run test_sceneflow_ddim.py's code path unchanged
    dvi.uninstall()

Nothing is added to any nn.Module (no parameters, no buffers): checkpoints load unchanged.
"""
from __future__ import annotations

import sys
from typing import Dict, List, Tuple

from . import kitti12, kitti15, sampler, sceneflow

_TIER1 = {
    "sceneflow": (sceneflow, ["models.submodule", "models.acv", "models.acv_ddim", "models.temp", "models"]),
    "kitti12": (kitti12, ["models.submodule", "models.pwcnet", "models.pwcnet_ddim", "models"]),
    "kitti15": (kitti15, ["core.submodule", "core.geometry", "core.geometry_ddim", "core.igev_stereo",
                          "core.igev_stereo_ddim", "core.extractor"]),
}

_TIER2 = {
    "sceneflow": [("models.acv_ddim", "ACVNet_DDIM", {
        "q_sample": sampler.q_sample,
        "predict_noise_from_start": sampler.predict_noise_from_start,
        "model_predictions": sampler.acv_model_predictions,
        "ddim_sample": sampler.acv_ddim_sample,
    })],
    "kitti12": [("models.pwcnet_ddim", "PWCNet_ddim", {
        "q_sample": sampler.q_sample,
        "predict_noise_from_start": sampler.predict_noise_from_start,
        "model_predictions": sampler.pcw_model_predictions,
        "ddim_sample": sampler.pcw_ddim_sample,
    })],
    "kitti15": [("core.igev_stereo_ddim", "IGEVStereo_ddim", {
        "q_sample": sampler.q_sample,
        "predict_noise_from_start": sampler.predict_noise_from_start,
        "model_predictions": sampler.igev_model_predictions,
        "ddim_sample": sampler.igev_ddim_sample,
    })],
}

# tier 3: the models' `forward` (only there can the unnamed ops between the named functions be fused and the concat
# features be handed to the sampler — see sampler.py)
_TIER3 = {
    "sceneflow": [("models.acv_ddim", "ACVNet_DDIM", {"forward": sampler.acv_ddim_forward}),
                  ("models.acv", "ACVNet", {"forward": sampler.acvnet_forward})],
    "kitti12": [],
    "kitti15": [],
}

_undo: List[Tuple[object, str, object, bool]] = []   # (owner, name, original, existed)


def _bind(owner, name, value):
    existed = hasattr(owner, name) and (name in vars(owner))
    _undo.append((owner, name, getattr(owner, name, None), existed))
    setattr(owner, name, value)


def install(project: str, tier2: bool = True, modules: Dict[str, object] = None, tier3: bool = True) -> List[str]:
    """Rebind the hot-path names of `project` ('sceneflow' | 'kitti12' | 'kitti15').  `modules` overrides
    sys.modules (used by the tests with stand-in modules).  tier2: the sampler methods; tier3 (needs tier2): the models'
    `forward`.  Returns the list of 'module.name' rebound."""
    if project not in _TIER1:
        raise ValueError(f"unknown project {project!r}; expected one of {sorted(_TIER1)}")
    mods = sys.modules if modules is None else modules
    mirror, consumers = _TIER1[project]
    done = []
    for mname in consumers:
        mod = mods.get(mname)
        if mod is None:
            continue
        for name in mirror.__all__:
            if hasattr(mod, name):
                _bind(mod, name, getattr(mirror, name))
                done.append(f"{mname}.{name}")
    if tier2:
        for mname, cname, methods in _TIER2[project]:
            mod = mods.get(mname)
            cls = getattr(mod, cname, None) if mod is not None else None
            if cls is None:
                continue
            for name, fn in methods.items():
                _bind(cls, name, fn)
                done.append(f"{mname}.{cname}.{name}")
    if tier2 and tier3:
        for mname, cname, methods in _TIER3[project]:
            mod = mods.get(mname)
            cls = getattr(mod, cname, None) if mod is not None else None
            if cls is None:
                continue
            for name, fn in methods.items():
                _bind(cls, name, fn)
                done.append(f"{mname}.{cname}.{name}")
    return done


def uninstall() -> int:
    """Restore everything install() rebound (LIFO).  Returns the number of bindings restored."""
    n = 0
    while _undo:
        owner, name, orig, existed = _undo.pop()
        if existed:
            setattr(owner, name, orig)
        else:
            try:
                delattr(owner, name)
            except AttributeError:
                pass
        n += 1
    return n


# ------------------------------------------------------------------------------------------------
# ACVNet's patch convolutions
# ------------------------------------------------------------------------------------------------
def fuse_acv_patch(model) -> bool:
    """Fuse `patch` -> `patch_l1/l2/l3` -> `torch.cat` of an ACVNet / ACVNet_DDIM instance
    (SceneFlow/models/acv_ddim.py:181-188,377-381) into the depth-wise chain kernel WITHOUT touching the model's
    forward or its state_dict: the four Conv3d modules keep their parameters; `patch.forward` returns the finished
    patch volume and `patch_l*.forward` pass the slices of that tensor through, so the reference's own
    `torch.cat((patch_l1, patch_l2, patch_l3), dim=1)` reassembles it.  Only taken when autograd is off (inference);
    with gradients enabled the original cuDNN convolutions run.  The four instances are re-classed to dynamic
    subclasses of their own class (not given instance-level `forward` attributes), so `nn.DataParallel` replicas — which
    are shallow copies of the instance `__dict__` — run the fused forward with THEIR parameters on THEIR device.
    Returns False when `model` has no such modules."""
    import torch
    import torch.nn.functional as F

    from . import ops

    model = getattr(model, "module", model)          # nn.DataParallel
    names = ("patch", "patch_l1", "patch_l2", "patch_l3")
    if not all(hasattr(model, n) for n in names):
        return False
    patch, l1, l2, l3 = (getattr(model, n) for n in names)
    siblings = (l1, l2, l3)                          # the originals: their (tiny) weights are moved to x.device per call

    def conv(self, x):
        return F.conv3d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)

    class FusedPatchConv(type(patch)):
        def forward(self, x):
            if (torch.is_grad_enabled() and (x.requires_grad or self.weight.requires_grad)) or not x.is_cuda:
                return conv(self, x)
            w = [m.weight.detach().to(x.device) for m in siblings]
            out = ops.acv_patch_volume(x, self.weight.detach(), w[0], w[1], w[2]).to(x.dtype)
            out._dv_fused_patch = True
            return out

    def make_slice_class(base):
        class PatchSliceConv(base):
            def forward(self, x):
                src = x._base if x._base is not None else x
                if getattr(src, "_dv_fused_patch", False):
                    return x                              # already patch_l*(patch(gwc)[:, slice])
                return conv(self, x)
        return PatchSliceConv

    for m, cls in ((patch, FusedPatchConv), (l1, make_slice_class(type(l1))), (l2, make_slice_class(type(l2))),
                   (l3, make_slice_class(type(l3)))):
        _undo.append((m, "__class__", m.__class__, True))
        m.__class__ = cls
    return True
