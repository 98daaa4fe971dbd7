"""Packed-feature layout helpers (host side), same names as the reference's
``app/utils.py:8-37``.

On the device the packing is fused into the STFT epilogue / iSTFT prologue
(``csrc/gss_stream.cuh``: split_pair / pack_pair); these NumPy helpers exist so
code written against the reference (``utils.spectrum_to_feature(Z)``) keeps
working for host arrays, e.g. when a spectrum comes from another tool.

Layout: ``feat[t, k] = Re X[k]`` (0 <= k < N/2), ``feat[t, N/2] = Re X[N/2]`` (the
Nyquist real part rides in the imaginary-DC slot), ``feat[t, N/2+k] = Im X[k]``.
"""
from __future__ import annotations

import numpy as np

from . import hparams


def spectrum_to_feature(freqs):
    """``[FFT_SIZE/2+1, LEN]`` complex -> ``[LEN, FFT_SIZE]`` real (utils.py:8-26)."""
    freqs = np.asarray(freqs)
    half = freqs.shape[0] - 1
    out = np.empty((freqs.shape[1], 2 * half), dtype=hparams.FLOATX)
    out[:, :half] = freqs.real[:half].T
    out[:, half:] = freqs.imag[:half].T
    out[:, half] = freqs.real[half]
    return out


def feature_to_spectrum(features):
    """reverse of :func:`spectrum_to_feature` (utils.py:29-37); width = ``hparams.FFT_SIZE``."""
    features = np.asarray(features)
    half = hparams.FFT_SIZE // 2
    assert features.shape[-1] == 2 * half, f"feature width {features.shape[-1]} != FFT_SIZE {2 * half}"
    cdt = np.complex128 if features.dtype == np.float64 else np.complex64
    Z = np.zeros((half + 1, features.shape[0]), dtype=cdt)
    Z[:half] = features[:, :half].T + 1j * features[:, half:].T
    Z[half] = features[:, half]
    Z[0] = features[:, 0]
    return Z
