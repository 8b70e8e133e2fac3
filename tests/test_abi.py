"""The C-ABI: header, ctypes table and built library must agree (no compute calls — CPU only)."""
import ctypes
import re
import subprocess
from pathlib import Path

import pytest

from diffuvolume_b200 import _lib
from diffuvolume_b200.build import LIB, build_library

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "dv_b200.h").read_text()


def header_functions():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(dv_[a-z0-9_]+)\s*\(", body)))


@pytest.fixture(scope="module")
def lib():
    build_library()
    return _lib.lib()


def test_header_matches_ctypes_table():
    assert header_functions() == sorted(_lib.SIGNATURES)


def test_library_exports_every_declared_symbol(lib):
    out = subprocess.run(["nm", "-D", "--defined-only", str(LIB)], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (dv_[a-z0-9_]+)", out))
    assert set(header_functions()) <= exported
    for name in header_functions():
        assert getattr(lib, name) is not None


def test_introspection_calls(lib):
    assert lib.dv_built_for_sm() == 100
    assert lib.dv_version() >= 100
    assert lib.dv_status_string(0) == b"ok"
    assert lib.dv_status_string(6) == b"unsupported configuration"
    assert lib.dv_launch_count() >= 0


def test_library_is_sm100a_and_uses_the_tma_engine(lib):
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in sass
    assert "UBLKCP" in sass          # cp.async.bulk (TMA engine) in the gwc kernel
    assert "SYNCS" in sass           # mbarrier
    assert "UTMALDG" in sass         # tensor-map TMA loads (streaming producers)
    # tensor cores appear in exactly one op: the compute-leaning all-pairs correlation (a14) — the tcgen05 kernel
    # (UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld from tensor memory, UTMASTG = TMA tensor store) and its warp-level
    # 3xTF32 fallback (HMMA); every HBM-bound volume kernel stays on plain FFMA
    fn = None
    owners, tc5 = set(), set()
    for line in sass.splitlines():
        if "Function :" in line:
            fn = line.split("Function :")[1].strip()
        elif "UTCHMMA" in line:
            tc5.add(fn)
        elif "HMMA" in line:
            owners.add(fn)
    assert owners and all("corr1d_allpairs_mma" in o for o in owners), owners
    assert tc5 and all("corr1d_allpairs_tcgen05" in o for o in tc5), tc5
    assert "LDTM" in sass and "UTMASTG" in sass


def test_argument_validation_without_a_gpu(lib):
    # every entry point validates before launching: NULL pointers / bad shapes never reach CUDA
    assert lib.dv_gwc_volume_f32(None, None, None, 1, 8, 4, 8, 4, 2, None) == 5          # DV_ERR_NULL
    buf = ctypes.create_string_buffer(64)
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert lib.dv_gwc_volume_f32(p, p, p, 1, 7, 4, 8, 4, 2, None) == 1                   # C % G != 0
    assert lib.dv_gwc_volume_f32(p, p, p, 0, 8, 4, 8, 4, 2, None) == 1
    assert lib.dv_softmax_regress_f32(None, 1, 4, 2, 2, None, None, None, None, None, 0, 0, None, 0, 0, None, None) == 5
    assert lib.dv_volume_filter_f32(p, p, 1, 1, 1, 1, 1, p, 7, None, 1.0, None, None) == 2  # bad dtype flag
    assert lib.dv_ddim_step(None, None) == 5
    assert lib.dv_geo_lookup_f32(p, p, None, p, p, p, 1, 8, 48, 2, 2, 8, 9, 4, None) == 6   # too many levels


def test_ddim_args_struct_layout_matches_the_header(tmp_path):
    """Compile a C probe against include/dv_b200.h and compare sizeof/offsetof with the ctypes mirror."""
    fields = [f[0] for f in _lib.DdimStepArgs._fields_]
    src = '#include <stdio.h>\n#include <stddef.h>\n#include "dv_b200.h"\nint main(void){\n'
    src += 'printf("%zu\\n", sizeof(dv_ddim_step_args));\n'
    for f in fields:
        src += f'printf("{f} %zu\\n", offsetof(dv_ddim_step_args, {f}));\n'
    src += "return 0;}\n"
    c = tmp_path / "probe.c"
    c.write_text(src)
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-I", str(ROOT / "include"), str(c), "-o", str(exe)], check=True)
    lines = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split("\n")
    assert int(lines[0]) == ctypes.sizeof(_lib.DdimStepArgs)
    for ln in lines[1:]:
        if ln.strip():
            name, off = ln.split()
            assert getattr(_lib.DdimStepArgs, name).offset == int(off), name


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "nope.so")
    with pytest.raises(_lib.DvLibraryError, match="no CPU fallback"):
        _lib.lib()
