"""GPU parity of the rows either side of the transforms: feature-domain mixing (A10), the
backward kernels, demo-mode edges (A13/A14: resample, pad, WAV in/out), the toy
config C1 end to end and the device-resident waveform dataset (8f.2).  All through the C ABI."""
import os

import numpy as np
import pytest

from oracle import ref_oracle as R

pytestmark = pytest.mark.gpu
REL_L2 = 1e-5


@pytest.fixture(scope="module")
def T():
    import torch
    assert torch.cuda.is_available()
    return torch


@pytest.fixture(scope="module")
def ops():
    from gan_sass_tf_b200.app import ops as o
    from gan_sass_tf_b200 import _native
    _native.lib()
    return o


@pytest.fixture()
def hp():
    from gan_sass_tf_b200.app import hparams
    saved = {k: getattr(hparams, k) for k in ("FFT_SIZE", "HOP_SIZE", "SEPARATOR_TYPE", "DATASET_TYPE", "BATCH_SIZE", "MAX_N_SIGNAL")}
    yield hparams
    for k, v in saved.items():
        setattr(hparams, k, v)


def dev(T, a):
    return T.from_numpy(np.ascontiguousarray(a)).cuda()


# ---- A10 --------------------------------------------------------------------------
@pytest.mark.parametrize("B,n_sig,Tn,N", [(8, 3, 128, 256), (2, 2, 37, 512), (1, 4, 5, 64)])
def test_mix_signals_matches_oracle(T, ops, hp, B, n_sig, Tn, N):
    hp.FFT_SIZE = N
    rng = np.random.default_rng(B + Tn)
    src = rng.random((B * n_sig, Tn, N), dtype=np.float32)
    noise = (rng.standard_normal((B, Tn, N)) * 0.1).astype(np.float32)
    ref = R.mix_features(src.astype(np.float64), B, n_sig, noise.astype(np.float64))
    got = ops.mix_signals(dev(T, src), n_sig, noise=dev(T, noise))
    assert got.shape == (B, Tn, N)
    assert R.rel_l2(got.cpu().numpy(), ref) < 1e-6
    mix, mix_log = ops.mix_signals(dev(T, src), n_sig, noise=dev(T, noise), log=True)
    assert np.array_equal(mix.cpu().numpy(), got.cpu().numpy())
    assert R.rel_l2(mix_log.cpu().numpy(), R.to_log_signal(ref)) < REL_L2
    # no noise / device-drawn noise (reproducible with a generator, N(0, 0.1^2) like tf.random_normal(stddev=0.1))
    assert R.rel_l2(ops.mix_signals(dev(T, src), n_sig, noise=False).cpu().numpy(), R.mix_features(src, B, n_sig)) < 1e-6
    g1 = T.Generator(device="cuda").manual_seed(3)
    g2 = T.Generator(device="cuda").manual_seed(3)
    a = ops.mix_signals(dev(T, src), n_sig, generator=g1)
    b = ops.mix_signals(dev(T, src), n_sig, generator=g2)
    assert T.equal(a, b)
    drawn = (a - ops.mix_signals(dev(T, src), n_sig, noise=False)).cpu().numpy()
    assert abs(drawn.std() - 0.1) < 0.02 and abs(drawn.mean()) < 0.02


# ---- backward kernels -------------------------------------------------------------
def _torch_log(x, eps):
    h = x.shape[-1] // 2
    a2 = x[..., :h] ** 2 + x[..., h:] ** 2
    g = 0.5 * (a2).log1p() * (a2 + eps).rsqrt()
    return x * g.repeat(*([1] * (x.dim() - 1)), 2)


def _torch_exp(x, eps):
    h = x.shape[-1] // 2
    a = (x[..., :h] ** 2 + x[..., h:] ** 2 + eps).sqrt()
    g = a.expm1() / a
    return x * g.repeat(*([1] * (x.dim() - 1)), 2)


@pytest.mark.parametrize("which", ["log", "exp"])
@pytest.mark.parametrize("scale", [1e-3, 0.3, 3.0])
def test_log_exp_backward_matches_autograd(T, ops, hp, which, scale):
    hp.FFT_SIZE = N = 256
    g = T.Generator(device="cuda").manual_seed(11)
    x = (T.randn(3, 17, N, device="cuda", generator=g) * scale).requires_grad_(True)
    go = T.randn(3, 17, N, device="cuda", generator=g)
    y = (ops.to_log_signal if which == "log" else ops.to_exp_signal)(x)
    y.backward(go)
    xd = x.detach().double().requires_grad_(True)
    yr = (_torch_log if which == "log" else _torch_exp)(xd, hp.EPS)
    yr.backward(go.double())
    assert R.rel_l2(y.detach().cpu().numpy(), yr.detach().cpu().numpy()) < REL_L2
    assert R.rel_l2(x.grad.cpu().numpy(), xd.grad.cpu().numpy()) < 2e-5


def test_apply_mask_backward_matches_autograd(T, ops):
    B, S, Tn, N = 2, 3, 9, 128
    g = T.Generator(device="cuda").manual_seed(5)
    mix = T.randn(B, Tn, N, device="cuda", generator=g).requires_grad_(True)
    mask = T.rand(B, S, Tn, N // 2, device="cuda", generator=g).requires_grad_(True)
    go = T.randn(B * S, Tn, N, device="cuda", generator=g)
    ops.apply_mask(mix, mask).backward(go)
    md, kd = mix.detach().double().requires_grad_(True), mask.detach().double().requires_grad_(True)
    ref = (T.cat([kd, kd], dim=-1) * md[:, None]).reshape(B * S, Tn, N)
    ref.backward(go.double())
    assert R.rel_l2(mix.grad.cpu().numpy(), md.grad.cpu().numpy()) < 1e-6
    assert R.rel_l2(mask.grad.cpu().numpy(), kd.grad.cpu().numpy()) < 1e-6


# ---- adjoints of the transforms (optional part) -----------------------
def _t_window(T, N):
    return 0.5 - 0.5 * T.cos(2 * np.pi * T.arange(N, dtype=T.float64, device="cuda") / N)


def _torch_stft_feature(T, x, N, H):
    """float64, differentiable restatement of oracle.stft_feature_np (scipy.signal.stft defaults + packing)"""
    n = x.shape[-1]
    nadd = (-n % H) % N
    xp = T.nn.functional.pad(x, (N // 2, N // 2 + nadd))
    Z = T.fft.rfft(xp.unfold(-1, N, H) * _t_window(T, N), dim=-1) * (2.0 / N)
    return T.cat([Z.real[..., :N // 2], Z.real[..., N // 2:], Z.imag[..., 1:N // 2]], dim=-1)


def _torch_istft_feature(T, f, H):
    """float64, differentiable restatement of oracle.istft_feature_np (scipy.signal.istft defaults)"""
    N, Tn = f.shape[-1], f.shape[-2]
    zero = T.zeros_like(f[..., :1])
    Z = T.complex(T.cat([f[..., :N // 2], f[..., N // 2:N // 2 + 1]], -1), T.cat([zero, f[..., N // 2 + 1:], zero], -1))
    w = _t_window(T, N)
    y = T.fft.irfft(Z, n=N, dim=-1) * (N / 2) * w
    total = N + (Tn - 1) * H
    acc = T.zeros(f.shape[:-2] + (total,), dtype=T.float64, device=f.device)
    nrm = T.zeros(total, dtype=T.float64, device=f.device)
    for t in range(Tn):
        acc = acc + T.nn.functional.pad(y[..., t, :], (t * H, total - N - t * H))
        nrm[t * H:t * H + N] += w * w
    nrm = T.where(nrm > 1e-10, nrm, T.ones_like(nrm))
    return (acc / nrm)[..., N // 2:total - N // 2]


@pytest.mark.parametrize("N,H", [(512, 128), (256, 128), (256, 32), (1024, 256), (64, 16)])
@pytest.mark.parametrize("log", [False, True])
def test_stft_backward_matches_autograd(T, ops, hp, N, H, log):
    g = T.Generator(device="cuda").manual_seed(N + H)
    n = 5 * N + 3 * H + 7
    x = (T.randn(2, n, device="cuda", generator=g) * 0.2).requires_grad_(True)
    y = ops.stft(x, N, H, log=log)
    go = T.randn(y.shape, device="cuda", generator=g)
    y.backward(go)
    xd = x.detach().double().requires_grad_(True)
    yr = _torch_stft_feature(T, xd, N, H)
    if log:
        yr = _torch_log(yr, hp.EPS)
    yr.backward(go.double())
    assert R.rel_l2(y.detach().cpu().numpy(), yr.detach().cpu().numpy()) < REL_L2
    assert R.rel_l2(x.grad.cpu().numpy(), xd.grad.cpu().numpy()) < 2e-5


@pytest.mark.parametrize("N,H", [(512, 128), (256, 128), (256, 32), (1024, 256), (64, 16)])
@pytest.mark.parametrize("exp", [False, True])
def test_istft_backward_matches_autograd(T, ops, hp, N, H, exp):
    g = T.Generator(device="cuda").manual_seed(N + H + 1)
    Tn = 5 * N // H + 4
    f = (T.randn(3, Tn, N, device="cuda", generator=g) * 0.3).requires_grad_(True)
    y = ops.istft(f, H, exp=exp, length=(Tn - 1) * H - 5)          # the trimmed tail gets a zero gradient
    go = T.randn(y.shape, device="cuda", generator=g)
    y.backward(go)
    fd = f.detach().double().requires_grad_(True)
    yr = _torch_istft_feature(T, _torch_exp(fd, hp.EPS) if exp else fd, H)[..., :(Tn - 1) * H - 5]
    yr.backward(go.double())
    assert R.rel_l2(y.detach().cpu().numpy(), yr.detach().cpu().numpy()) < REL_L2
    assert R.rel_l2(f.grad.cpu().numpy(), fd.grad.cpu().numpy()) < 2e-5


@pytest.mark.parametrize("N,H", [(512, 128), (512, 256), (256, 64), (1024, 256)])
def test_mask_istft_backward_matches_autograd(T, ops, N, H):
    """a waveform-domain loss back-propagated to the masks (what trains a mask separator) and to the mixture"""
    g = T.Generator(device="cuda").manual_seed(N + H + 2)
    B, S, n = 2, 3, 6 * N + H + 3
    Tn, _ = R.frame_count(n, N, H)
    x = (T.randn(B, n, device="cuda", generator=g) * 0.2).requires_grad_(True)
    m = T.rand(B, S, Tn, N // 2, device="cuda", generator=g).requires_grad_(True)
    y = ops.mask_istft(x, m, N, H)
    go = T.randn(y.shape, device="cuda", generator=g)
    y.backward(go)
    xd, md = x.detach().double().requires_grad_(True), m.detach().double().requires_grad_(True)
    feat = _torch_stft_feature(T, xd, N, H)
    masked = (T.cat([md, md], dim=-1) * feat[:, None]).reshape(B * S, Tn, N)
    yr = _torch_istft_feature(T, masked, H)
    yr.backward(go.double())
    assert R.rel_l2(y.detach().cpu().numpy(), yr.detach().cpu().numpy()) < REL_L2
    assert R.rel_l2(m.grad.cpu().numpy(), md.grad.cpu().numpy()) < 2e-5
    assert R.rel_l2(x.grad.cpu().numpy(), xd.grad.cpu().numpy()) < 2e-5


def test_mask_separator_trains_through_native_kernels(T, ops, hp):
    """The plugin surface end to end with gradients: waveforms -> stft_log -> registered mask separator ->
    mask_istft -> waveform-domain loss -> Adam.  Two sources in disjoint bands are separable by a per-bin mask,
    so a few steps through the native forward and adjoint kernels must cut the loss substantially."""
    from gan_sass_tf_b200.app import modules  # noqa: F401  (registers 'toy-mask')
    hp.FFT_SIZE, hp.HOP_SIZE, hp.MAX_N_SIGNAL = 256, 64, 1
    N, H, S, B, n = 256, 64, 2, 8, 4096
    hp.SEPARATOR_TYPE = 'toy-mask'
    sep = hp.get_separator()(None, 'train_test_sep')
    g = T.Generator(device="cuda").manual_seed(21)
    t = T.arange(n, device="cuda", dtype=T.float32) / 16000.0
    lo = T.sin(2 * np.pi * 500.0 * t)[None] * (0.2 + 0.1 * T.rand(B, 1, device="cuda", generator=g))
    hi = T.sin(2 * np.pi * 5000.0 * t)[None] * (0.2 + 0.1 * T.rand(B, 1, device="cuda", generator=g))
    target = T.stack([lo, hi], dim=1).reshape(B * S, n)
    mix = lo + hi
    logf = ops.stft_log(mix, N, H)
    sep(logf)                                            # creates the lazily-built layers
    opt = T.optim.Adam(list(sep.p.parameters()), lr=3e-3)
    losses = []
    for _ in range(30):
        opt.zero_grad()
        masks = sep(logf)                                # [B, S, T, N/2], differentiable torch module
        est = ops.mask_istft(mix, masks, N, H)[:, :n]    # native forward, native adjoint
        loss = ((est - target) ** 2).mean()
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
    assert all(p.grad is not None and T.isfinite(p.grad).all() for p in sep.p.parameters())
    assert losses[-1] < 0.25 * losses[0], losses


# ---- demo-mode edges --------------------------------------------------------------
@pytest.mark.parametrize("n,num", [(1000, 363), (1001, 364), (44100, 16000), (8000, 16000), (8001, 16001), (999, 500), (500, 1001)])
def test_resample_matches_scipy(T, ops, n, num):
    import scipy.signal
    rng = np.random.default_rng(n)
    x = rng.standard_normal((2, n))
    ref = scipy.signal.resample(x, num, axis=-1)
    got = ops.resample(dev(T, x), num).cpu().numpy()
    assert got.shape == ref.shape and np.abs(got - ref).max() < 1e-12
    got32 = ops.resample(dev(T, x.astype(np.float32)), num).cpu().numpy()
    assert R.rel_l2(got32, ref) < 1e-6
    for L in (1, 255, 256, 257, 5000):
        assert ops.resample_pad_size(L, 256) == R.resample_pad_size(L, 256)


@pytest.mark.parametrize("rate,n,dtype", [(16000, 6000, np.int16), (8000, 3001, np.int16), (44100, 9000, np.int16),
                                          (16000, 5000, np.float32), (22050, 5000, np.float32)])
def test_load_save_wavfile(T, ops, hp, tmp_path, rate, n, dtype):
    import scipy.io.wavfile
    from gan_sass_tf_b200 import main as drv
    hp.FFT_SIZE, hp.HOP_SIZE = 256, None
    rng = np.random.default_rng(n + rate)
    x = (rng.standard_normal(n) * (3000 if dtype == np.int16 else 0.1)).astype(dtype)
    path = os.path.join(tmp_path, "in.wav")
    scipy.io.wavfile.write(path, rate, x)
    feat = drv.load_wavfile(path)
    ref = R.load_wave_features(x, rate, 256)
    assert tuple(feat.shape) == ref.shape                       # frame count bit-exact (K1, pad rule)
    assert R.rel_l2(feat.cpu().numpy(), ref) < REL_L2
    out = os.path.join(tmp_path, "out.wav")
    drv.save_wavfile(out, feat)
    sr, pcm = scipy.io.wavfile.read(out)
    ref_pcm = R.save_wave_pcm(ref, 256)
    assert sr == 16000 and pcm.dtype == np.int16 and pcm.shape == ref_pcm.shape
    assert np.max(np.abs(pcm.astype(np.int32) - ref_pcm.astype(np.int32))) <= 1     # +-1 LSB at the truncation boundary
    with pytest.raises(FileNotFoundError):
        drv.load_wavfile(None)


# ---- config C1: toy data, reference defaults, forward end to end -------------------
def test_c1_toy_forward(T, ops, hp):
    """BASELINE config C1: toy batch [24,128,256] -> mix + N(0,0.1) -> to_log -> toy separator
    -> to_exp -> iSTFT of the 32 outputs (16 256 samples each), against the oracle chain with the
    separator's weights evaluated in float64."""
    from gan_sass_tf_b200.app import modules
    from gan_sass_tf_b200.app.datasets.dataset import WhiteNoiseData
    hp.FFT_SIZE, hp.HOP_SIZE, hp.BATCH_SIZE, hp.MAX_N_SIGNAL = 256, None, 8, 3
    N, B, n_sig, S = 256, 8, 3, 4
    ds = WhiteNoiseData(seed=0)
    ds.install_and_load()
    src = next(iter(ds.epoch('train', B * n_sig)))[0]                      # [24,128,256] uniform [0,1)
    assert src.shape == (24, 128, 256) and src.dtype == np.float32
    noise = (np.random.default_rng(1).standard_normal((B, 128, N)) * 0.1).astype(np.float32)
    mix, mix_log = ops.mix_signals(dev(T, src), n_sig, noise=dev(T, noise), log=True)
    sep_mod = modules.ToySeparator(None, 'c1/separator')
    sep_log = sep_mod(mix_log)
    assert tuple(sep_log.shape) == (B * S, 128, N)
    sep = ops.to_exp_signal(sep_log)
    waves = ops.istft(sep, N // 2)
    assert tuple(waves.shape) == (B * S, 127 * 128)                        # 16 256 samples each
    # oracle
    W0 = sep_mod.p.layers['c1/separator/linear0'].weight.detach().double().cpu().numpy()
    b0 = sep_mod.p.layers['c1/separator/linear0'].bias.detach().double().cpu().numpy()
    W1 = sep_mod.p.layers['c1/separator/linear1'].weight.detach().double().cpu().numpy()
    b1 = sep_mod.p.layers['c1/separator/linear1'].bias.detach().double().cpu().numpy()
    rmix = R.mix_features(src.astype(np.float64), B, n_sig, noise.astype(np.float64))
    rlog = R.to_log_signal(rmix)
    h = rlog @ W0.T + b0
    h = np.maximum(h, hp.RELU_LEAKAGE * h)
    o = (h @ W1.T + b1).reshape(B, 128, S, N).transpose(0, 2, 1, 3).reshape(B * S, 128, N)
    rsep = R.to_exp_signal(o)
    rwav = R.istft_feature_np(rsep, N // 2)
    assert R.rel_l2(mix_log.cpu().numpy(), rlog) < REL_L2
    assert R.rel_l2(sep_log.detach().cpu().numpy(), o) < 1e-4            # fp32 matmuls of the stand-in
    assert R.rel_l2(waves.detach().cpu().numpy(), rwav) < 1e-4
    # metrics of main.py:353-361 / :446-457 on the same tensors
    assert abs(float(ops.ae_loss(sep.detach(), mix, S)) - R.autoencoder_loss(rsep, rmix, B, S)) < 1e-3 * R.autoencoder_loss(rsep, rmix, B, S)
    assert abs(float(ops.snr_metric(dev(T, src), sep.detach(), n_sig)) - R.snr_metric(src.astype(np.float64), rsep, B, n_sig)) < 1e-3


# ---- device-resident waveform dataset ---------------------------------------------
def test_waveform_dataset_features(T, ops, hp):
    from gan_sass_tf_b200.app.datasets.wave import WaveformData
    hp.FFT_SIZE, hp.HOP_SIZE = 512, 128
    rng = np.random.default_rng(4)
    waves = [(rng.standard_normal(int(n)) * 2000).astype(np.int16) for n in (3000, 5000, 2200, 4100, 3333, 600)]
    ds = WaveformData()
    ds.add_subset('train', waves)
    seen = 0
    for feats, frames in ds.epoch('train', 2):
        assert feats.dim() == 3 and feats.shape[0] == 2 and feats.shape[2] == 512
        seen += 1
    assert seen == 3
    feats, frames = next(iter(ds.epoch('train', 3)))
    order = np.argsort([len(w) for w in waves], kind="stable")[:3]
    n_max = max(max(len(waves[i]) for i in order), 512)
    for r, i in enumerate(order):
        x = np.zeros(n_max, np.int16)
        x[:len(waves[i])] = waves[i]
        assert R.rel_l2(feats[r].cpu().numpy(), R.stft_feature_np(x, 512, 128)) < REL_L2
        assert int(frames[r]) == R.frame_count(len(waves[i]), 512, 128)[0]
    # changing FFT_SIZE needs no re-install: the next epoch is transformed with the new size
    hp.FFT_SIZE, hp.HOP_SIZE = 256, None
    feats2, _ = next(iter(ds.epoch('train', 3)))
    assert feats2.shape[2] == 256


def test_demo_mode_writes_separated_files(T, ops, hp, tmp_path):
    """main.py:749-771 with one clip as one batch row (A13)."""
    import scipy.io.wavfile
    from gan_sass_tf_b200 import main as drv
    hp.FFT_SIZE, hp.HOP_SIZE, hp.SEPARATOR_TYPE, hp.DATASET_TYPE = 256, None, 'toy-mask', 'toy'
    x = (np.random.default_rng(2).standard_normal(20000) * 2500).astype(np.int16)
    path = os.path.join(tmp_path, "clip.wav")
    scipy.io.wavfile.write(path, 16000, x)
    drv.main(['-m', 'demo', '-if', path])
    Tn = R.frame_count(20000, 256, 128)[0]
    for i in range(1, hp.MAX_N_SIGNAL + 2):
        sr, pcm = scipy.io.wavfile.read(os.path.join(tmp_path, "clip_separated_%d.wav" % i))
        assert sr == 16000 and pcm.dtype == np.int16 and pcm.shape == ((Tn - 1) * 128,)
        assert pcm.min() == 0 and pcm.max() >= 32766
    res = drv.g_model.test(drv.g_dataset)
    assert set(res) == {'SNR', 'AE'} and np.isfinite(res['SNR']) and np.isfinite(res['AE'])


# ---- -m train: the spectral slice of main.py:568-624 through the native ops ---------------------------------
@pytest.mark.parametrize("sep_type", ["toy", "toy-mask"])
def test_train_mode_runs_and_loss_falls(T, ops, hp, tmp_path, monkeypatch, sep_type):
    from gan_sass_tf_b200 import main as drv
    monkeypatch.chdir(tmp_path)                                   # saves/<name>_e<k> land in the temporary directory
    hp.FFT_SIZE, hp.HOP_SIZE, hp.BATCH_SIZE, hp.MAX_N_SIGNAL = 256, None, 4, 2
    hp.SEPARATOR_TYPE, hp.DATASET_TYPE = sep_type, 'toy'
    drv.main(['-m', 'train', '-ne', '1', '-n', 'unit', '-o', os.path.join(tmp_path, 'final.pt')])
    assert os.path.exists(os.path.join(tmp_path, 'saves', 'unit_e1')) and os.path.exists(os.path.join(tmp_path, 'final.pt'))
    model = drv.g_model
    ds = drv.g_dataset
    reports = model.train(ds, 3, lr=2e-3, save_on_epoch=False, test_on_epoch=True, out=open(os.devnull, 'w'))
    assert len(reports) == 3 and all(np.isfinite(r['AE']) and np.isfinite(r['test_SNR']) for r in reports)
    if sep_type == 'toy':
        assert reports[-1]['AE'] < reports[0]['AE']               # Adam on the auto-encoder loss makes progress
    else:
        # softmax masks sum to one, so sum_s mask_s * mix - mix = 0: the auto-encoder term vanishes identically (the
        # reference's generator objective is ae - gan, main.py:481; its GAN term is out of scope) - a check of A7 + A12
        assert all(r['AE'] < 1e-12 and r['test_AE'] < 1e-12 for r in reports)
    # a checkpoint restores the weights bit for bit
    before = [p.detach().clone() for p in model.parameters()]
    model.save_params(os.path.join(tmp_path, 'ck.pt'))
    for p in model.parameters():
        p.data.add_(1.0)
    model.load_params(os.path.join(tmp_path, 'ck.pt'))
    assert all(T.equal(a, b) for a, b in zip(before, model.parameters()))


def test_load_wavfile_stereo_and_short(T, ops, hp, tmp_path):
    """a (nsamples, nchannels) WAV: the FIRST CHANNEL is used (upstream's indexing takes the first frame);
    a clip shorter than FFT_SIZE is rejected with a message that names the file."""
    import scipy.io.wavfile
    from gan_sass_tf_b200 import main as drv
    hp.FFT_SIZE, hp.HOP_SIZE = 256, None
    rng = np.random.default_rng(11)
    st = (rng.standard_normal((4000, 2)) * 3000).astype(np.int16)
    path = os.path.join(tmp_path, "stereo.wav")
    scipy.io.wavfile.write(path, 16000, st)
    feat = drv.load_wavfile(path)
    ref = R.load_wave_features(st[:, 0].copy(), 16000, 256)
    assert tuple(feat.shape) == ref.shape and R.rel_l2(feat.cpu().numpy(), ref) < REL_L2
    short = os.path.join(tmp_path, "short.wav")
    scipy.io.wavfile.write(short, 16000, st[:100, 0].copy())
    with pytest.raises(ValueError, match="fewer than FFT_SIZE"):
        drv.load_wavfile(short)


def test_waveform_dataset_from_directory(T, ops, hp, tmp_path):
    """file-backed corpus (TIMIT/process.py:89-110): 16 kHz WAVs under a directory, 'sa*' skipped, other rates rejected;
    the gathered batch equals the oracle's features of the zero-padded utterances; epochs shuffle differently."""
    import scipy.io.wavfile
    from gan_sass_tf_b200.app.datasets.wave import WaveformData
    hp.FFT_SIZE, hp.HOP_SIZE = 256, None
    rng = np.random.default_rng(5)
    lens = {"a1.wav": 3001, "b2.wav": 5000, "sa1.wav": 4000, "c3.wav": 2500, "d4.wav": 2600}
    os.makedirs(os.path.join(tmp_path, "dr1"))
    waves = {}
    for name, n in lens.items():
        waves[name] = (rng.standard_normal(n) * 2500).astype(np.int16)
        scipy.io.wavfile.write(os.path.join(tmp_path, "dr1", name), 16000, waves[name])
    ds = WaveformData()
    assert ds.add_directory('train', str(tmp_path)) == 4                          # sa1.wav skipped
    feats, frames = next(iter(ds.epoch('train', 4)))
    order = sorted((n for n in lens if not n.startswith("sa")), key=lambda k: lens[k])
    n_max = feats.shape[1]
    for r, name in enumerate(order):
        x = np.zeros(5000, np.int16)
        x[:lens[name]] = waves[name]
        ref = R.stft_feature_np(x, 256, 128)
        assert feats.shape[1] == ref.shape[0] and R.rel_l2(feats[r].cpu().numpy(), ref) < REL_L2
        assert int(frames[r]) == R.frame_count(lens[name], 256, 128)[0]
    o1 = [f.shape for f, _ in ds.epoch('train', 1, shuffle=True)]
    o2 = [f.shape for f, _ in ds.epoch('train', 1, shuffle=True)]
    o3 = [f.shape for f, _ in ds.epoch('train', 1, shuffle=True)]
    assert sorted(o1) == sorted(o2) and (o1 != o2 or o1 != o3)                    # same batches, a new order per epoch
    scipy.io.wavfile.write(os.path.join(tmp_path, "dr1", "e5.wav"), 8000, waves["a1.wav"])
    with pytest.raises(ValueError, match="Sampling rate"):
        WaveformData().add_directory('train', str(tmp_path))


def test_resample_long_clip_and_batch(T, ops):
    """the hand-written Bluestein resampler at demo sizes: 10 s of 44.1 kHz -> 16 kHz, and a small batch of rows"""
    import scipy.signal
    rng = np.random.default_rng(3)
    x = rng.standard_normal(441000)
    got = ops.resample(dev(T, x), 160000).cpu().numpy()
    assert np.abs(got - scipy.signal.resample(x, 160000)).max() < 1e-11
    xb = rng.standard_normal((3, 2, 1234))
    gotb = ops.resample(dev(T, xb), 777).cpu().numpy()
    assert gotb.shape == (3, 2, 777) and np.abs(gotb - scipy.signal.resample(xb, 777, axis=-1)).max() < 1e-12


def test_per_frame_path_takes_more_than_65535_rows(T, ops):
    """the per-frame kernels (FFT_SIZE 64 / 128) index rows through grid.y: long batches are split into several launches"""
    N, H, n, B = 64, 32, 96, 70000
    g = T.Generator(device="cuda").manual_seed(9)
    x = T.randn(B, n, device="cuda", generator=g) * 0.1
    f = ops.stft(x, N, H)
    y = ops.istft(f, H)
    assert T.equal(f[65535:65540], ops.stft(x[65535:65540].contiguous(), N, H))
    err = (y[:, :n] - x).double().pow(2).sum() / x.double().pow(2).sum()
    assert 10 * np.log10(1.0 / float(err)) >= 100.0
    m = T.rand(B, 2, f.shape[1], N // 2, device="cuda", generator=g)
    w = ops.mask_istft(x, m, N, H)
    ref = R.mask_istft_np(x[69999:].cpu().numpy(), m[69999:].cpu().numpy(), N, H).reshape(2, -1)
    assert R.rel_l2(w.reshape(B, 2, -1)[69999].cpu().numpy(), ref) < REL_L2
