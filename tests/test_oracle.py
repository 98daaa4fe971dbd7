"""Pins oracle/ref_oracle.py: against live SciPy (the reference's third-party
arithmetic), against the K1-K10 known answers of SURVEY.md section 4 and against
the fixtures oracle/make_golden.py produced by running the reference's own
app/utils.py and app/ops.py source."""
import numpy as np
import pytest
import scipy.signal

from oracle import ref_oracle as R


# K1 -----------------------------------------------------------------------
@pytest.mark.parametrize("n,N,H,T,nadd", [
    (48000, 256, 128, 376, 0), (48000, 512, 128, 376, 0), (46797, 512, 128, 367, 51),
    (64000, 512, 128, 501, 0), (960000, 1024, 256, 3751, 0), (48000, 4096, 1024, 48, 128),
    (48000, 256, 64, 751, 0), (48000, 1024, 256, 189, 128), (48000, 2048, 512, 95, 128),
])
def test_k1_frame_count(n, N, H, T, nadd):
    assert R.frame_count(n, N, H) == (T, nadd)
    Z = scipy.signal.stft(np.zeros(n, np.float32), nperseg=N, noverlap=N - H)[2]
    assert Z.shape == (N // 2 + 1, T)


# K2 / K10 -------------------------------------------------------------------
@pytest.mark.parametrize("n,N,H", [(3000, 256, 128), (4797, 512, 128), (6000, 1024, 256), (2048, 256, 64)])
def test_k2_np_equals_scipy_f64(n, N, H):
    x = np.random.default_rng(n).normal(size=n)
    a = R.stft_np(x, N, H, np.float64)
    b = R.stft_scipy(x, N, H)
    assert a.shape == b.shape
    assert np.max(np.abs(a - b)) < 1e-13
    assert abs(R.hann_periodic(N).sum() - N / 2) < 1e-9
    y = R.istft_np(a, N, H)
    yb = R.istft_scipy(b, N, H)
    assert y.shape == yb.shape == ((a.shape[-1] - 1) * H,)            # K4
    assert np.max(np.abs(y - yb)) < 1e-12


def test_k3_scaling():
    N = 512
    t = np.arange(8192)
    x = 3 * np.cos(2 * np.pi * 37 * t / N) + 0.5
    Z = R.stft_np(x, N, 128)
    mid = Z[:, Z.shape[1] // 2]
    assert abs(abs(mid[37]) - 1.5) < 1e-9
    assert abs(abs(mid[36]) - 0.75) < 1e-9 and abs(abs(mid[38]) - 0.75) < 1e-9
    assert abs(mid[0] - 0.5) < 1e-9 and abs(abs(mid[1]) - 0.25) < 1e-9


def test_k4_k5_roundtrip_and_length():
    x = np.random.default_rng(0).normal(size=46797)
    Z = R.stft_np(x, 512, 128)
    y = R.istft_np(Z, 512, 128)
    assert y.shape[0] == 46848
    assert R.snr_db(x, y[:46797]) > 280
    x32 = x.astype(np.float32)
    y32 = R.istft_scipy(R.stft_scipy(x32, 512, 128), 512, 128)
    assert y32.dtype == np.float32
    assert R.snr_db(x32, y32[:46797]) > 120


def test_k6_pack_unpack():
    Z = R.stft_scipy(np.random.default_rng(1).normal(size=3000).astype(np.float32), 256)
    f = R.spectrum_to_feature(Z)
    assert f.shape == (Z.shape[1], 256) and f.dtype == np.float32
    Zb = R.feature_to_spectrum(f)
    Zc = Z.copy(); Zc[0].imag = 0; Zc[-1].imag = 0
    assert np.array_equal(Zb, Zc)
    assert np.array_equal(f[:, 0], Z[0].real) and np.array_equal(f[:, 128], Z[128].real)
    assert np.array_equal(f[:, 5], Z[5].real) and np.array_equal(f[:, 128 + 5], Z[5].imag)


def test_k7_log_exp_not_inverse():
    f = np.random.default_rng(2).normal(size=(4, 256)).astype(np.float64)
    g = R.to_exp_signal(R.to_log_signal(f))
    m = np.sqrt(f[:, :128] ** 2 + f[:, 128:] ** 2)
    mg = np.sqrt(g[:, :128] ** 2 + g[:, 128:] ** 2)
    assert np.max(np.abs(mg - (np.sqrt(1 + m ** 2) - 1))) < 1e-4
    assert np.max(np.abs(mg - m)) > 0.1


def test_k8_dtype():
    assert R.stft_scipy(np.zeros(1000, np.int16), 256).dtype == np.complex64
    assert R.stft_scipy(np.zeros(1000, np.float32), 256).dtype == np.complex64
    assert R.stft_scipy(np.zeros(1000, np.float64), 256).dtype == np.complex128


def test_k9_wav16(golden):
    g = golden("wav16.npz")
    assert np.array_equal(R.wav16_normalise(g["wav16/in"]), g["wav16/out"])
    assert list(g["wav16/out"]) == [0, 16383, 24575, 32767]


def test_k10_torch_cross_oracle():
    import torch
    n, N, H = 4797, 512, 128
    T, nadd = R.frame_count(n, N, H)
    x = np.random.default_rng(3).normal(size=n)
    xt = torch.nn.functional.pad(torch.from_numpy(x), (0, nadd))
    w = torch.hann_window(N, periodic=True, dtype=torch.float64)
    Zt = torch.stft(xt, N, H, N, w, center=True, pad_mode="constant", normalized=False,
                    onesided=True, return_complex=True) / w.sum()
    assert np.max(np.abs(Zt.numpy() - R.stft_np(x, N, H))) < 1e-13


# golden fixtures made by the reference's own code ----------------------------
CASES = ["ref_default_256", "c2_512_128", "ragged_512_128", "c3_1024_256", "n256_h64",
         "n2048_h512", "n4096_h1024", "short_equal_N"]


@pytest.mark.parametrize("name", CASES)
def test_golden_stft_istft(golden, name):
    g = golden("stft_istft.npz")
    x = g[name + "/x"]
    N, H = (int(v) for v in g[name + "/NH"])
    feat64 = R.stft_feature_np(x, N, H, np.float64, np.float64)
    assert feat64.shape == g[name + "/feat"].shape
    assert np.max(np.abs(feat64 - g[name + "/feat64"])) < 1e-14
    assert R.rel_l2(g[name + "/feat"], feat64) < 2e-6             # reference's f32 path vs f64 yardstick
    y = R.istft_feature_np(g[name + "/feat"], H, np.float64)
    assert y.shape == g[name + "/istft"].shape
    assert R.rel_l2(g[name + "/istft"], y) < 2e-6
    # the literal scipy call chain reproduces the fixture bit for bit
    assert np.array_equal(R.stft_feature_scipy(x, N, H), g[name + "/feat"])
    assert np.array_equal(R.istft_feature_scipy(g[name + "/feat"], H).astype(np.float32), g[name + "/istft"])


def test_golden_int16(golden):
    g = golden("stft_istft.npz")
    assert g["int16_256/dtype"][0] == "complex64"
    f = R.stft_feature_scipy(g["int16_256/x"], 256)
    assert np.array_equal(f, g["int16_256/feat"])
    assert R.rel_l2(f, R.stft_feature_np(g["int16_256/x"], 256, None)) < 2e-6


@pytest.mark.parametrize("N", [256, 512])
def test_golden_log_exp(golden, N):
    g = golden("tf_ops.npz")
    f = g[f"logexp_{N}/f"]
    assert np.allclose(R.to_log_signal(f), g[f"logexp_{N}/to_log"], rtol=2e-6, atol=1e-9)
    assert np.allclose(R.to_exp_signal(f * np.float32(0.3)), g[f"logexp_{N}/to_exp"], rtol=2e-6, atol=1e-9)
    assert np.allclose(R.to_exp_signal(R.to_log_signal(f)), g[f"logexp_{N}/exp_of_log"], rtol=5e-6, atol=1e-9)


def test_golden_snr(golden):
    g = golden("tf_ops.npz")
    assert np.allclose(R.batch_cross_snr(g["snr/clear"], g["snr/noisy"]), g["snr/cross"], rtol=1e-5, atol=1e-5)
    assert np.allclose(R.batch_snr(g["snr/clear"][:, 0], g["snr/noisy"][:, 0]), g["snr/batch"], rtol=1e-5, atol=1e-5)
    assert g["snr/cross"].shape == (4, 3, 4)


def test_mask_linearity_and_identity():
    rng = np.random.default_rng(5)
    B, S, n, N, H = 2, 3, 3000, 256, 64
    x = rng.normal(size=(B, n))
    T, nadd = R.frame_count(n, N, H)
    ones = np.ones((B, S, T, N // 2))
    y = R.mask_istft_np(x, ones, N, H)
    assert y.shape == (B, S, n + nadd)
    assert np.max(np.abs(y[:, 0, :n] - x)) < 1e-12                  # all-pass mask = round trip
    m = rng.random((B, S, T, N // 2))
    m[:, 2] = 1.0 - m[:, 0] - m[:, 1]
    z = R.mask_istft_np(x, m, N, H)
    assert np.max(np.abs(z.sum(axis=1)[:, :n] - x)) < 1e-11          # masks summing to 1 -> sources sum to mix
    f = R.stft_feature_np(x, N, H, np.float64, np.float64)
    ym = R.apply_mask(f, m)
    assert ym.shape == (B * S, T, N)
    assert np.array_equal(ym[1 * S + 2], np.concatenate([m[1, 2], m[1, 2]], -1) * f[1])


def test_mix_ae_snr_metric():
    rng = np.random.default_rng(6)
    B, n_sig, T, N = 2, 3, 5, 256
    src = rng.normal(size=(B * n_sig, T, N))
    noise = rng.normal(0, 0.1, size=(B, T, N))
    mix = R.mix_features(src, B, n_sig, noise)
    assert np.allclose(mix[1], src[3] + src[4] + src[5] + noise[1])
    sep = np.concatenate([src.reshape(B, n_sig, T, N), noise[:, None]], 1).reshape(B * 4, T, N)
    assert R.autoencoder_loss(sep, mix, B, 4) < 1e-28
    assert R.snr_metric(src, sep, B, n_sig) > 60
    assert R.resample_pad_size(1000, 256) == 24 and R.resample_pad_size(1024, 256) == 0
