"""ctypes binding of the C-ABI shared library (include/dv_b200.h).

The library is the product: there is no CPU or PyTorch fallback.  `lib()` raises
`DvLibraryError` when `libdv_b200.so` has not been built (run `python -m diffuvolume_b200.build`
or `__graft_entry__.build()`), and every wrapper raises on a non-zero status.
"""
from __future__ import annotations

import ctypes as C
import threading
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "lib" / "libdv_b200.so"

STATUS = {
    0: "DV_OK",
    1: "DV_ERR_BAD_SHAPE",
    2: "DV_ERR_BAD_DTYPE",
    3: "DV_ERR_MISALIGNED",
    4: "DV_ERR_LAUNCH",
    5: "DV_ERR_NULL",
    6: "DV_ERR_UNSUPPORTED",
}


class DvLibraryError(RuntimeError):
    """libdv_b200.so is missing, cannot be loaded, or a call returned a non-zero dv_status."""


class DdimStepArgs(C.Structure):
    """Mirror of `dv_ddim_step_args` (include/dv_b200.h)."""

    _fields_ = [
        ("B", C.c_int64), ("D", C.c_int64), ("h", C.c_int64), ("w", C.c_int64),
        ("H", C.c_int64), ("W", C.c_int64),
        ("disp", C.c_void_p),
        ("disp_clamp_hi", C.c_float),
        ("coords0", C.c_void_p),
        ("xt", C.c_void_p),
        ("xt_is_f64", C.c_int),
        ("shift", C.c_void_p),
        ("scale", C.c_double),
        ("vote", C.c_void_p),
        ("used", C.c_void_p),
        ("vote_thr_dif", C.c_float),
        ("mask", C.c_void_p),
        ("sqrt_recip", C.c_double), ("sqrt_recipm1", C.c_double),
        ("last_step", C.c_int),
        ("sqrt_alpha_next", C.c_double), ("c", C.c_double), ("sigma", C.c_double),
        ("step_noise", C.c_void_p),
        ("renoise_mode", C.c_int),
        ("renoise", C.c_void_p),
        ("asd", C.c_void_p),
        ("asd_is_f64", C.c_int),
        ("q_noise", C.c_void_p),
        ("q_noise_is_f64", C.c_int),
        ("sqrt_ac", C.c_double), ("sqrt_1m_ac", C.c_double),
        ("asd_out", C.c_void_p),
        ("x0_out", C.c_void_p),
        ("eps_out", C.c_void_p),
        ("x_next", C.c_void_p),
        ("shift_next", C.c_void_p),
        ("n_next_out", C.c_void_p),
    ]


_P, _I64, _I, _F, _D = C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_double

# name -> (restype, argtypes); every function declared in include/dv_b200.h must be listed here
# (tests/test_abi.py checks the header, this table and the built library against each other).
SIGNATURES = {
    "dv_version": (_I, []),
    "dv_status_string": (C.c_char_p, [_I]),
    "dv_launch_count": (_I64, []),
    "dv_built_for_sm": (_I, []),
    "dv_groupwise_correlation_f32": (_I, [_P, _P, _P, _I64, _I64, _I64, _I64, _I64, _P]),
    "dv_gwc_volume_f32": (_I, [_P, _P, _P, _I64, _I64, _I64, _I64, _I64, _I64, _P]),
    "dv_gwc_volume_bf16": (_I, [_P, _P, _P, _I64, _I64, _I64, _I64, _I64, _I64, _P]),
    "dv_concat_volume_f32": (_I, [_P, _P, _P, _I64, _I64, _I64, _I64, _I64, _I, _P, _P, _I, _P, _D, _P]),
    "dv_att_softmax_f32": (_I, [_P, _P, _I64, _I64, _I64, _I64, _P]),
    "dv_filter_factor_f32": (_I, [_P, _I, _P, _D, _P, _I64, _I64, _I64, _I64, _P]),
    "dv_filter_factor": (_I, [_P, _I, _P, _D, _P, _P, _I64, _I64, _I64, _I64, _P]),
    "dv_concat_volume_weighted_f32": (_I, [_P, _P, _P, _I64, _I64, _I64, _I64, _I64, _I, _P, _P, _P, _P]),
    "dv_concat_volume_weighted_bf16": (_I, [_P, _P, _P, _I64, _I64, _I64, _I64, _I64, _I, _P, _P, _P, _P]),
    "dv_volume_filter_f32": (_I, [_P, _P, _I64, _I64, _I64, _I64, _I64, _P, _I, _P, _D, _P, _P]),
    "dv_corr_volume_2sided_f32": (_I, [_P, _P, _P, _I64, _I64, _I64, _I64, _I64, _I64, _P]),
    "dv_softmax_regress_f32": (_I, [_P, _I64, _I64, _I64, _I64, _P, _P, _P, _P, _P, _F, _F, _P, _F, _I, _P, _P]),
    "dv_upsample_softmax_regress_f32": (_I, [_P, _I64, _I64, _I64, _I64, _I64, _I64, _I64, _I, _P, _P, _P, _P, _F, _F, _P, _F, _I, _P]),
    "dv_uncertainty_vote_f32": (_I, [_P, _P, _P, _I64, _I64, _I64, _I64, _F, _F, _P, _P, _P]),
    "dv_softmax_uncertainty_vote_f32": (_I, [_P, _P, _P, _I64, _I64, _I64, _I64, _F, _F, _P, _P, _P, _P]),
    "dv_disparity_regression_f32": (_I, [_P, _P, _I64, _I64, _I64, _I64, _P]),
    "dv_q_sample": (_I, [_P, _I, _P, _I, _D, _D, _P, _I64, _P]),
    "dv_predict_noise_from_start": (_I, [_P, _I, _P, _I, _D, _D, _P, _I64, _P]),
    "dv_xstart_from_disp_f32": (_I, [_P, _P, _I64, _I64, _I64, _I64, _D, _P]),
    "dv_downsample_bilinear_f32": (_I, [_P, _P, _I64, _I64, _I64, _I64, _I64, _F, _F, _F, _P]),
    "dv_ddim_step": (_I, [C.POINTER(DdimStepArgs), _P]),
    "dv_corr1d_allpairs_f32": (_I, [_P, _P, _P, _I64, _I64, _I64, _I64, _I64, _P]),
    "dv_corr1d_allpairs_pooled_f32": (_I, [_P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _P]),
    "dv_geo_permute_f32": (_I, [_P, _P, _I64, _I64, _I64, _I64, _I64, _P]),
    "dv_avgpool_w2_f32": (_I, [_P, _P, _I64, _I64, _P]),
    "dv_geo_lookup_f32": (_I, [_P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _I64, _I, _I, _P]),
    "dv_geo_pack_f32": (_I, [_P, _P, _I64, _I64, _I64, _I64, _I64, _I, _P]),
    "dv_depthwise3x3_chain_f32": (_I, [_P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _I64, _I64, _I, _I, _P]),
    "dv_context_upsample_f32": (_I, [_P, _P, _P, _I64, _I64, _I64, _P]),
    "dv_context_upsample_bwd_f32": (_I, [_P, _P, _P, _P, _P, _I64, _I64, _I64, _P]),
    "dv_geo_filter_packed_f32": (_I, [_P, _P, _P, _I64, _I64, _I64, _I, _P]),
    "dv_geo_lookup_packed_f32": (_I, [_P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _I64, _I, _I, _P]),
    "dv_gwc_volume_bwd_f32": (_I, [_P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _I64, _P]),
    "dv_corr_volume_2sided_bwd_f32": (_I, [_P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _I64, _P]),
    "dv_groupwise_correlation_bwd_f32": (_I, [_P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _P]),
    "dv_concat_volume_bwd_f32": (_I, [_P, _P, _P, _I64, _I64, _I64, _I64, _I64, _I, _P]),
    "dv_disparity_regression_bwd_f32": (_I, [_P, _P, _I64, _I64, _I64, _I64, _P]),
    "dv_softmax_regress_bwd_f32": (_I, [_P, _P, _P, _I64, _I64, _I64, _I64, _P]),
    "dv_acv_volume_bwd_f32": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _I64, _I, _P]),
    "dv_warp_f32": (_I, [_P, _P, _P, _I64, _I64, _I64, _I64, _P]),
    "dv_warp_bwd_f32": (_I, [_P, _P, _P, _P, _P, _I64, _I64, _I64, _I64, _P]),
    "dv_corr_volume_2sided_into_f32": (_I, [_P, _P, _P, _I64, _I64, _I64, _I64, _I64, _I64, _I64, _P]),
    "dv_warp_assemble_f32": (_I, [_P, _P, _P, _P, _P, _I64, _P, _I64, _I64, _I64, _I64, _I64, _P]),
    "dv_select_close_f32": (_I, [_P, _P, _F, _P, _I64, _P]),
    "dv_ensemble_f32": (_I, [_P, _P, _I, _P, _I64, _P]),
}

_lock = threading.Lock()
_lib = None


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raise DvLibraryError if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not LIB_PATH.exists():
            raise DvLibraryError(
                f"{LIB_PATH} not found: the sm_100a CUDA library is not built. "
                "Run `python -m diffuvolume_b200.build` (needs nvcc). There is no CPU fallback."
            )
        try:
            handle = C.CDLL(str(LIB_PATH))
        except OSError as e:  # pragma: no cover - depends on the box
            raise DvLibraryError(f"cannot load {LIB_PATH}: {e}") from e
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(handle, name)
            except AttributeError as e:
                raise DvLibraryError(f"{LIB_PATH} does not export {name}; rebuild the library") from e
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        raise DvLibraryError(f"{what} failed: {STATUS.get(status, status)}")


def launch_count() -> int:
    return int(lib().dv_launch_count())
