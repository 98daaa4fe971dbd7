"""Executes the ctypes stub of INTEGRATION.md section 1 VERBATIM (the code block is read from the document) against
the oracle's SciPy calls - the binding a maintainer of the reference would add at main.py:97-98 / :110-111 - plus the
threading / multi-device behaviour of the host-buffer entry points."""
import ctypes
import os
import re
import threading

import numpy as np
import pytest

from oracle import ref_oracle as R

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REL_L2 = 1e-5


@pytest.fixture(scope="module")
def gss():
    import torch
    assert torch.cuda.is_available()
    from gan_sass_tf_b200 import _native
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"```python\n(# app/gss\.py.*?)```", text, re.S)
    assert m, "INTEGRATION.md: stub code block not found"
    code = m.group(1)
    assert 'ctypes.CDLL("libgss.so")' in code
    ns = {}
    real = ctypes.CDLL

    def cdll(name, *a, **k):             # the stub names the library as an installed one; the tree holds it under lib/
        return real(_native.LIB_PATH if name == "libgss.so" else name, *a, **k)
    ctypes.CDLL = cdll
    try:
        exec(compile(code, "INTEGRATION.md:app/gss.py", "exec"), ns)
    finally:
        ctypes.CDLL = real
    return type("gss", (), {k: staticmethod(v) if callable(v) else v for k, v in ns.items() if not k.startswith("__")})


@pytest.mark.parametrize("N,n,B", [(256, 48000, 2), (256, 4797, 3), (512, 6000, 1), (1024, 20000, 2)])
def test_stub_stft_and_istft_match_the_reference_calls(gss, N, n, B):
    rng = np.random.default_rng(N + n)
    x = (rng.standard_normal((B, n)) * 0.1).astype(np.float32)
    feat = gss.stft_feature(x, N)                                   # SciPy's default hop N/2, as main.py:97
    ref = np.stack([R.stft_feature_scipy(r, N) for r in x])         # scipy.signal.stft(...)[2] + spectrum_to_feature
    assert feat.shape == ref.shape and feat.dtype == np.float32
    assert R.rel_l2(feat, ref) < REL_L2
    wave = gss.istft_feature(feat)
    refw = np.stack([R.istft_feature_scipy(f) for f in ref])        # feature_to_spectrum + scipy.signal.istft
    assert wave.shape == refw.shape and R.rel_l2(wave, refw) < REL_L2
    assert R.snr_db(x, wave[:, :n]) >= 100.0
    # single clip, exactly the edits of INTEGRATION.md section 2
    f1 = gss.stft_feature(x[0], N)[0]
    assert np.array_equal(f1, feat[0])
    assert np.array_equal(gss.istft_feature(f1), wave[0])
    # fused log / exp flags
    lg = gss.stft_feature(x, N, log=True)
    assert R.rel_l2(lg, R.to_log_signal(ref.astype(np.float64))) < REL_L2
    ex = gss.istft_feature(lg, exp=True)
    refe = np.stack([R.istft_feature_np(R.to_exp_signal(f.astype(np.float64))) for f in R.to_log_signal(ref.astype(np.float64))])
    assert R.rel_l2(ex, refe) < REL_L2


def test_stub_error_mapping(gss):
    with pytest.raises(ValueError):
        gss.stft_feature(np.zeros(100, np.float32), 256)            # n < FFT_SIZE
    with pytest.raises(ValueError):
        gss.stft_feature(np.zeros(4000, np.float32), 300)           # not a power of two


def test_host_entry_points_from_several_threads(gss):
    """re-entrancy of the *_host calls (per-device workspace behind a mutex, private stream)"""
    rng = np.random.default_rng(0)
    xs = [(rng.standard_normal((2, 3000 + 500 * i)) * 0.1).astype(np.float32) for i in range(6)]
    out = [None] * len(xs)

    def work(i):
        out[i] = gss.istft_feature(gss.stft_feature(xs[i], 256))
    th = [threading.Thread(target=work, args=(i,)) for i in range(len(xs))]
    for t in th:
        t.start()
    for t in th:
        t.join()
    for i, x in enumerate(xs):
        assert R.snr_db(x, out[i][:, :x.shape[1]]) >= 100.0


def test_wait_host_from_another_thread():
    """gss_wait_host on a thread that did not enqueue must really wait (copy-stream state is per device)"""
    import torch
    from gan_sass_tf_b200.app.spectral import SpectralPipeline
    B, n, S, N, H = 64, 48000, 3, 512, 128
    pipe = SpectralPipeline(B, n, S, N, H, chunks=4, depth=2)
    g = torch.Generator().manual_seed(0)
    x = (torch.randn(B, n, generator=g) * 0.1)
    mask = torch.rand(B, S, pipe.T, N // 2, device="cuda")
    mask = mask / mask.sum(dim=1, keepdim=True)
    pipe.out_hs[0].zero_()
    pipe.analyse(x, slot=0, block=False)
    pipe.synthesise(mask, slot=0, block=False)
    got = {}

    def waiter():
        torch.cuda.set_device(pipe.device)
        got["buf"] = pipe.wait(0).clone()
    t = threading.Thread(target=waiter)
    t.start(); t.join()
    rec = got["buf"].reshape(B, S, -1).sum(dim=1)[:, :n]
    assert R.snr_db(x.numpy(), rec.numpy()) >= 100.0          # all zeros (no wait) would give 0 dB


def test_pipelines_on_two_devices_in_one_thread():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from gan_sass_tf_b200.app.spectral import SpectralPipeline
    B, n, S, N, H = 8, 16000, 2, 512, 128
    x = torch.randn(B, n) * 0.1
    for d in (0, 1, 0):
        with torch.cuda.device(d):
            pipe = SpectralPipeline(B, n, S, N, H, device=f"cuda:{d}", chunks=2, depth=1)
            mask = torch.full((B, S, pipe.T, N // 2), 0.5, device=f"cuda:{d}")
            pipe.analyse(x, block=False)
            w = pipe.synthesise(mask, block=True)
            assert R.snr_db(x.numpy(), w.reshape(B, S, -1).sum(dim=1)[:, :n].numpy()) >= 100.0
