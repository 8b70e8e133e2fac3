"""Drop-in contract: every name the mirrors export has the parameter list of the reference function / method it replaces.
Runs only where the reference tree is mounted (the authoring container); skipped on the GPU box, where /root/reference
does not exist.  Nothing is executed from the reference beyond importing the modules."""
import importlib.util
import inspect
import os
import sys
import types
from pathlib import Path

import pytest

REF = Path(os.environ.get("DV_REFERENCE", "/root/reference"))
pytestmark = pytest.mark.skipif(not REF.exists(), reason="reference tree not mounted")


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _params(fn, drop_self=False):
    ps = [(p.name, p.default if p.default is not inspect._empty else "<required>") for p in inspect.signature(fn).parameters.values()]
    return ps[1:] if drop_self else ps


@pytest.mark.parametrize("mirror_name,ref_path", [
    ("sceneflow", "SceneFlow/models/submodule.py"),
    ("kitti12", "KITTI12/models/submodule.py"),
    ("kitti12", "SceneFlow/submodule.py"),            # the orphan top-level copy the north_star names
    ("kitti15", "KITTI15/core/submodule.py"),
])
def test_function_mirrors_have_the_reference_signatures(mirror_name, ref_path):
    mirror = importlib.import_module(f"diffuvolume_b200.{mirror_name}")
    ref = _load(REF / ref_path, "ref_" + ref_path.replace("/", "_").replace(".", "_"))
    checked = 0
    for name in mirror.__all__:
        if not hasattr(ref, name) or inspect.isclass(getattr(mirror, name)):
            continue
        assert _params(getattr(mirror, name)) == _params(getattr(ref, name)), name
        checked += 1
    assert checked >= 4


def test_geometry_class_has_the_reference_interface():
    from diffuvolume_b200 import kitti15
    sys.path.insert(0, str(REF / "KITTI15"))
    try:
        plain = _load(REF / "KITTI15/core/geometry.py", "ref_k15_geometry")
        ddim = _load(REF / "KITTI15/core/geometry_ddim.py", "ref_k15_geometry_ddim")
    finally:
        sys.path.remove(str(REF / "KITTI15"))
    ours = kitti15.Combined_Geo_Encoding_Volume
    assert _params(ours.__init__, True) == _params(plain.Combined_Geo_Encoding_Volume.__init__, True)
    assert _params(ours.__init__, True) == _params(ddim.Combined_Geo_Encoding_Volume.__init__, True)
    assert [n for n, _ in _params(ours.__call__, True)][:2] == [n for n, _ in _params(plain.Combined_Geo_Encoding_Volume.__call__, True)]
    assert [n for n, _ in _params(ours.__call__, True)] == [n for n, _ in _params(ddim.Combined_Geo_Encoding_Volume.__call__, True)]
    assert _params(ours.corr) == _params(ddim.Combined_Geo_Encoding_Volume.corr)


@pytest.mark.parametrize("project,module,cls", [
    ("SceneFlow", "models.acv_ddim", "ACVNet_DDIM"),
    ("KITTI12", "models.pwcnet_ddim", "PWCNet_ddim"),
    ("KITTI15", "core.igev_stereo_ddim", "IGEVStereo_ddim"),
])
def test_sampler_methods_have_the_reference_signatures(project, module, cls):
    """Tier 2: the methods install() binds onto the reference's own classes.  Each sub-project is imported in a
    subprocess-free way by temporarily owning sys.path / sys.modules (they share top-level package names)."""
    from diffuvolume_b200 import install as dvi
    saved = {k: v for k, v in sys.modules.items() if k.split(".")[0] in ("models", "core", "utils", "datasets")}
    for k in saved:
        del sys.modules[k]
    for stub in ("timm", "opt_einsum"):
        sys.modules.setdefault(stub, types.ModuleType(stub))
    if not hasattr(sys.modules["opt_einsum"], "contract"):
        sys.modules["opt_einsum"].contract = lambda *a, **k: None
    sys.path.insert(0, str(REF / project))
    try:
        mod = importlib.import_module(module)
        ref_cls = getattr(mod, cls)
        key = {"SceneFlow": "sceneflow", "KITTI12": "kitti12", "KITTI15": "kitti15"}[project]
        (_, _, methods), = dvi._TIER2[key]
        for name, fn in methods.items():
            assert _params(fn, True) == _params(getattr(ref_cls, name), True), name
        # and install() really rebinds them on the imported reference (then restores)
        done = dvi.install(key)
        try:
            assert f"{module}.{cls}.ddim_sample" in done
            assert getattr(ref_cls, "ddim_sample") is methods["ddim_sample"]
            for name in importlib.import_module(f"diffuvolume_b200.{key}").__all__:
                if hasattr(mod, name):
                    assert getattr(mod, name) is getattr(importlib.import_module(f"diffuvolume_b200.{key}"), name), name
        finally:
            dvi.uninstall()
        assert getattr(ref_cls, "ddim_sample") is not methods["ddim_sample"]
    finally:
        sys.path.remove(str(REF / project))
        for k in [k for k in sys.modules if k.split(".")[0] in ("models", "core", "utils", "datasets")]:
            del sys.modules[k]
        sys.modules.update(saved)
