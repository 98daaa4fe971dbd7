"""ctypes binding of ``libgss.so`` (the C ABI declared in ``include/gss_api.h``).

There is NO fallback: if the shared library is missing, or a call returns a
non-zero status, this module raises.  The CPU oracle under ``oracle/`` is test
infrastructure and is never imported from here.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_void_p, POINTER

_HERE = os.path.dirname(os.path.abspath(__file__))
# GSS_LIB: developer override (instrumented / tuning builds); the product path is the in-tree library
LIB_PATH = os.environ.get("GSS_LIB") or os.path.join(_HERE, "lib", "libgss.so")

GSS_OK, GSS_EINVAL, GSS_EUNSUPPORTED, GSS_ECUDA, GSS_ENOMEM = 0, -1, -2, -3, -4
FLAG_LOG, FLAG_EXP, FLAG_REVERSE = 1, 2, 4

# name -> (restype, argtypes); mirrors include/gss_api.h one to one
_P = c_void_p
SIGNATURES = {
    "gss_version": (c_int, []),
    "gss_last_error": (c_char_p, []),
    "gss_launch_count": (c_int64, []),
    "gss_supported_fft_sizes": (c_int, [POINTER(c_int), c_int]),
    "gss_frame_count": (c_int, [c_int64, c_int, c_int, POINTER(c_int64), POINTER(c_int64)]),
    "gss_stft_packed": (c_int, [_P, c_int64, c_int64, c_int64, c_int, c_int, c_int, c_float, _P, _P]),
    "gss_stft_packed_i16": (c_int, [_P, c_int64, c_int64, c_int64, c_int, c_int, c_int, c_float, _P, _P]),
    "gss_istft_packed": (c_int, [_P, c_int64, c_int64, c_int, c_int, c_int, c_float, _P, c_int64, _P]),
    "gss_mask_istft": (c_int, [_P, _P, c_int64, c_int, c_int64, c_int64, c_int, c_int, _P, c_int64, _P]),
    "gss_stft_packed_dual": (c_int, [_P, c_int64, c_int64, c_int64, c_int, c_int, c_float, _P, _P, _P]),
    "gss_mask_istft_feature": (c_int, [_P, _P, c_int64, c_int, c_int64, c_int, c_int, c_int, _P, c_int64, _P]),
    "gss_mask_istft_feature_ae": (c_int, [_P, _P, c_int64, c_int, c_int64, c_int, c_int, c_int, _P, c_int64, _P, _P]),
    "gss_metric_finalise": (c_int, [_P, _P, c_int64, c_int, c_int, ctypes.c_double, _P, _P]),
    "gss_resample_workspace_bytes": (ctypes.c_size_t, [c_int64, c_int64]),
    "gss_resample_f64": (c_int, [_P, c_int64, c_int64, c_int64, c_int64, _P, c_int64, _P, ctypes.c_size_t, _P]),
    "gss_gather_rows_i16": (c_int, [_P, _P, _P, _P, c_int64, c_int64, _P, _P]),
    "gss_apply_mask": (c_int, [_P, _P, c_int64, c_int, c_int64, c_int, _P, _P]),
    "gss_ola_norm_scale": (c_int, [_P, _P, c_int64, c_int64, c_int64, c_int64, c_int64, c_int, c_int, c_int, c_float, _P]),
    "gss_scale_packed": (c_int, [_P, _P, c_int64, c_int, c_float, c_float, _P]),
    "gss_to_log": (c_int, [_P, _P, c_int64, c_int, c_float, _P]),
    "gss_to_exp": (c_int, [_P, _P, c_int64, c_int, c_float, _P]),
    "gss_cross_snr": (c_int, [_P, _P, c_int64, c_int, c_int, c_int64, c_float, _P, _P]),
    "gss_ae_partial": (c_int, [_P, _P, c_int64, c_int, c_int64, _P, _P]),
    "gss_wav16_normalise": (c_int, [_P, c_int64, c_int64, c_int64, _P, _P, _P]),
    "gss_stft_packed_host": (c_int, [_P, c_int64, c_int64, c_int, c_int, c_int, c_float, _P]),
    "gss_istft_packed_host": (c_int, [_P, c_int64, c_int64, c_int, c_int, c_int, c_float, _P]),
    "gss_stft_h2d": (c_int, [_P, _P, c_int64, c_int64, c_int64, c_int, c_int, c_int, c_float, _P, c_int, _P]),
    "gss_mask_istft_d2h": (c_int, [_P, _P, c_int64, c_int, c_int64, c_int64, c_int, c_int, _P, _P, c_int64, c_int, _P]),
    "gss_stft_h2d_async": (c_int, [_P, _P, c_int64, c_int64, c_int64, c_int, c_int, c_int, c_float, _P, c_int, _P]),
    "gss_mask_istft_d2h_async": (c_int, [_P, _P, c_int64, c_int, c_int64, c_int64, c_int, c_int, _P, _P, c_int64, c_int, _P]),
    "gss_wait_host": (c_int, [_P]),
    "gss_stft_h2d_i16_async": (c_int, [_P, _P, _P, c_int64, c_int64, c_int64, c_int, c_int, c_int, c_float, _P, c_int, _P]),
    "gss_mask_istft_d2h_pcm16_async": (c_int, [_P, _P, c_int64, c_int, c_int64, c_int64, c_int, c_int, _P, _P, _P, _P, c_int64, c_int, _P]),
    "gss_mix_features": (c_int, [_P, _P, c_int64, c_int, c_int64, c_int, c_int, c_float, _P, _P, _P]),
    "gss_to_log_bwd": (c_int, [_P, _P, _P, c_int64, c_int, c_float, _P]),
    "gss_to_exp_bwd": (c_int, [_P, _P, _P, c_int64, c_int, c_float, _P]),
    "gss_apply_mask_bwd": (c_int, [_P, _P, _P, c_int64, c_int, c_int64, c_int, _P, _P, _P]),
}

# only in lib/libgss_experimental.so (include/gss_api.h under GSS_EXPERIMENTAL): process-wide kernel-family switches
EXPERIMENTAL_SIGNATURES = {
    "gss_set_path": (c_int, [c_int]),
    "gss_set_synth_variant": (c_int, [c_int]),
}
EXPERIMENTAL_LIB_PATH = os.path.join(_HERE, "lib", "libgss_experimental.so")

_lib = None
_product_lib = None


class GssError(RuntimeError):
    """A libgss call failed with GSS_ECUDA / GSS_ENOMEM."""


def lib():
    """Load (once) and return the ctypes handle.  Raises if the library is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -m gan_sass_tf_b200.build` "
                "(nvcc, sm_100a).  There is no CPU fallback.")
        h = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(h, name)          # AttributeError if the ABI and this table disagree
            fn.restype = res
            fn.argtypes = args
        _lib = h
    return _lib


def check(rc: int):
    """0 -> None; GSS_EINVAL/GSS_EUNSUPPORTED -> ValueError (what the reference's
    asserts / SciPy raise for bad shapes); anything else -> GssError."""
    if rc == GSS_OK:
        return
    msg = lib().gss_last_error().decode("utf-8", "replace")
    if rc in (GSS_EINVAL, GSS_EUNSUPPORTED):
        raise ValueError(f"libgss: {msg}")
    raise GssError(f"libgss ({rc}): {msg}")


def frame_count(n: int, N: int, H: int):
    """(T, nadd) of ``scipy.signal.stft(boundary='zeros', padded=True)`` (main.py:97)."""
    T, nadd = c_int64(0), c_int64(0)
    check(lib().gss_frame_count(n, N, H, ctypes.byref(T), ctypes.byref(nadd)))
    return T.value, nadd.value


def supported_fft_sizes():
    buf = (c_int * 16)()
    k = lib().gss_supported_fft_sizes(buf, 16)
    return tuple(buf[i] for i in range(min(k, 16)))


def load_experimental():
    """Switch this process to ``lib/libgss_experimental.so`` (a superset of the product ABI: the measured-slower
    synthesis variants, the team kernels at 256 / 512 and the switches below).  For the cross-check tests and the
    tuning tools; ``unload_experimental()`` goes back to the product library."""
    global _lib, _product_lib
    if not os.path.exists(EXPERIMENTAL_LIB_PATH):
        raise ImportError(f"{EXPERIMENTAL_LIB_PATH} is missing: build it with `python -m gan_sass_tf_b200.build`")
    if _product_lib is None:
        _product_lib = lib()
    h = ctypes.CDLL(EXPERIMENTAL_LIB_PATH)
    for name, (res, args) in {**SIGNATURES, **EXPERIMENTAL_SIGNATURES}.items():
        fn = getattr(h, name)
        fn.restype = res
        fn.argtypes = args
    _lib = h
    return h


def unload_experimental():
    global _lib
    if _product_lib is not None:
        _lib = _product_lib


def _experimental():
    h = lib()
    if not hasattr(h, "gss_set_path"):
        raise RuntimeError("kernel-family switches exist in lib/libgss_experimental.so only: call _native.load_experimental() first")
    return h


def set_path(path: int):
    """0 = automatic kernel selection, 1 = no register-exchange kernels (team kernels), 2 = per-frame kernels only.
    Experimental library only."""
    check(_experimental().gss_set_path(path))


def set_synth_variant(variant: int):
    """N = 512 fused synthesis: 0 = register-resident kernel, 1 = role-split CTAs, 2 = tensor-memory state.
    Experimental library only."""
    check(_experimental().gss_set_synth_variant(variant))


def launch_count() -> int:
    return int(lib().gss_launch_count())
