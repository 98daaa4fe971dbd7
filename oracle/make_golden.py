"""Generate tests/golden/*.npz from the reference itself.  TEST INFRASTRUCTURE.

Run in the build container only (``/root/reference`` does not exist on the GPU
box):  ``python oracle/make_golden.py``.

What is executed:

* the reference's own ``app/utils.py`` (``spectrum_to_feature`` /
  ``feature_to_spectrum``), imported with ``nltk`` and ``tensorflow`` stubbed
  because neither is installed (utils.py:1, hparams.py:57);
* the reference's own ``app/ops.py`` functions ``to_log_signal``,
  ``to_exp_signal``, ``batch_snr``, ``batch_cross_snr`` (ops.py:162-251),
  imported with ``tensorflow`` replaced by a NumPy shim of exactly the
  ``tf.*`` calls those four functions make.  The reference's operator
  composition is therefore what produced the fixtures; only the 12 leaf
  kernels are NumPy instead of TF 1.x;
* ``scipy.signal.stft`` / ``istft`` called the way main.py:97 / :111 call them
  (scipy version recorded in the fixture).
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import scipy
import scipy.signal

REF = os.environ.get("GSS_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


class _Shape:
    def __init__(self, shp):
        self._s = list(shp)

    def as_list(self):
        return list(self._s)


class _T(np.ndarray):
    """ndarray with TF1's ``get_shape().as_list()``."""

    def get_shape(self):
        return _Shape(self.shape)


def _t(x):
    return np.asarray(x).view(_T)


def _tf_shim():
    tf = types.ModuleType("tensorflow")
    tf.square = lambda x: _t(np.square(x))
    tf.add = lambda a, b: _t(np.add(a, b))
    tf.split = lambda x, sizes, axis: [_t(p) for p in np.split(x, np.cumsum(sizes)[:-1], axis=axis)]
    tf.log1p = lambda x: _t(np.log1p(x))
    tf.log = lambda x: _t(np.log(x))
    tf.rsqrt = lambda x: _t(1.0 / np.sqrt(x))
    tf.sqrt = lambda x: _t(np.sqrt(x))
    tf.expm1 = lambda x: _t(np.expm1(x))
    tf.tile = lambda x, reps: _t(np.tile(x, reps))
    tf.reduce_mean = lambda x, axis=None: _t(np.mean(x, axis=None if axis is None else tuple(axis)))
    tf.expand_dims = lambda x, axis: _t(np.expand_dims(x, axis))
    contrib = types.ModuleType("tensorflow.contrib")
    tf.contrib = contrib
    return tf


def _import_reference():
    sys.modules["tensorflow"] = _tf_shim()
    sys.modules["nltk"] = types.ModuleType("nltk")
    sys.path.insert(0, REF)
    import app.hparams as hparams          # noqa: E402
    import app.utils as utils              # noqa: E402
    import app.ops as ops                  # noqa: E402
    return hparams, utils, ops


def speechish(rng, n, scale=0.05):
    """1-pole low-passed Gaussian noise, clipped to [-1,1]."""
    e = rng.normal(0.0, scale, n)
    y = np.empty(n)
    acc = 0.0
    for i in range(n):
        acc = 0.95 * acc + e[i]
        y[i] = acc
    return np.clip(y, -1, 1).astype(np.float32)


def main():
    hparams, utils, ops = _import_reference()
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20261018)
    meta = dict(scipy=scipy.__version__, numpy=np.__version__)

    # ---- STFT / pack / iSTFT through the reference's utils + scipy ----------
    cases = [  # (name, n, N, H)   H=None -> the reference's literal default call
        ("ref_default_256", 3000, 256, None),
        ("c2_512_128", 4000, 512, 128),
        ("ragged_512_128", 4797, 512, 128),       # nadd = 67
        ("c3_1024_256", 6000, 1024, 256),
        ("n256_h64", 2048, 256, 64),
        ("n2048_h512", 9000, 2048, 512),
        ("n4096_h1024", 12000, 4096, 1024),
        ("short_equal_N", 512, 512, 128),          # n == N edge
    ]
    store = {}
    for name, n, N, H in cases:
        x = speechish(rng, n)
        hparams.FFT_SIZE = N
        kw = {} if H is None else dict(noverlap=N - H)
        Z = scipy.signal.stft(x, nperseg=N, **kw)[2]            # main.py:97
        feat = utils.spectrum_to_feature(Z)                       # main.py:98
        Zb = utils.feature_to_spectrum(feat)                      # main.py:110
        y = scipy.signal.istft(Zb, nperseg=N, **kw)[1]            # main.py:111
        Z64 = scipy.signal.stft(x.astype(np.float64), nperseg=N, **kw)[2]
        store[name + "/x"] = x
        store[name + "/NH"] = np.array([N, N // 2 if H is None else H])
        store[name + "/feat"] = feat.astype(np.float32)
        hparams.FLOATX = 'float64'                               # yardstick: same code, f64
        store[name + "/feat64"] = utils.spectrum_to_feature(Z64)
        hparams.FLOATX = 'float32'
        store[name + "/istft"] = y.astype(np.float32)
    # int16 input path (process.py:94-98): int16 waveform -> complex64
    xi = (speechish(rng, 3000) * 20000).astype(np.int16)
    hparams.FFT_SIZE = 256
    Zi = scipy.signal.stft(xi, nperseg=256)[2]
    store["int16_256/x"] = xi
    store["int16_256/feat"] = utils.spectrum_to_feature(Zi).astype(np.float32)
    store["int16_256/dtype"] = np.array([str(Zi.dtype)])
    np.savez_compressed(os.path.join(OUT, "stft_istft.npz"), **store,
                        **{"meta/" + k: np.array([v]) for k, v in meta.items()})

    # ---- the four TF elementwise/metric ops, reference source under the shim --
    store = {}
    for N in (256, 512):
        hparams.FFT_SIZE = N
        f = (rng.normal(0, 1.0, (3, 7, N)) * rng.choice([1e-4, 1e-2, 1.0, 8.0], (3, 7, 1))).astype(np.float32)
        store[f"logexp_{N}/f"] = f
        store[f"logexp_{N}/to_log"] = np.asarray(ops.to_log_signal(_t(f)), dtype=np.float32)
        store[f"logexp_{N}/to_exp"] = np.asarray(ops.to_exp_signal(_t(f * 0.3)), dtype=np.float32)
        store[f"logexp_{N}/exp_of_log"] = np.asarray(
            ops.to_exp_signal(_t(np.asarray(ops.to_log_signal(_t(f)), dtype=np.float32))), dtype=np.float32)
    clear = rng.normal(0, 1, (4, 3, 9, 256)).astype(np.float32)
    noisy = (clear[:, [0, 1, 2, 0]] * 0.8 + rng.normal(0, 0.3, (4, 4, 9, 256))).astype(np.float32)
    store["snr/clear"] = clear
    store["snr/noisy"] = noisy
    store["snr/cross"] = np.asarray(ops.batch_cross_snr(_t(clear), _t(noisy)), dtype=np.float32)
    store["snr/batch"] = np.asarray(ops.batch_snr(_t(clear[:, 0]), _t(noisy[:, 0])), dtype=np.float32)
    np.savez_compressed(os.path.join(OUT, "tf_ops.npz"), **store)

    # ---- WAV normalise (main.py:112-116 restated inline: it writes a file) ----
    store = {}
    d = np.array([-1, 0, .5, 1], dtype=np.float32)
    dd = d.copy(); lo, hi = np.min(dd), np.max(dd); dd -= lo; dd *= (32767. / (hi - lo))
    store["wav16/in"] = d
    store["wav16/out"] = dd.astype(np.int16)
    np.savez_compressed(os.path.join(OUT, "wav16.npz"), **store)
    print("golden fixtures written to", os.path.normpath(OUT))


if __name__ == "__main__":
    main()
