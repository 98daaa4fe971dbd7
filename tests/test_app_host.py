"""CPU-side checks of the host logic around the path: registries, the toy generator (A16), the
driver's mode switch (main.py:736-781) and the oracle's restatement of the WAV edges."""
import numpy as np
import pytest

from oracle import ref_oracle as R


def test_registries_and_unknown_names():
    from gan_sass_tf_b200.app import hparams, modules, datasets  # noqa: F401
    assert {'toy', 'toy-mask', 'dc-v1'} <= set(hparams.separator_registry)
    assert {'toy', 'wave'} <= set(hparams.dataset_registry)
    old = hparams.SEPARATOR_TYPE
    try:
        hparams.SEPARATOR_TYPE = 'no-such-separator'
        with pytest.raises(KeyError):
            hparams.get_separator()
    finally:
        hparams.SEPARATOR_TYPE = old
    with pytest.raises(NotImplementedError):
        hparams.separator_registry['dc-v1'](None, 'x')          # stub upstream as well (modules.py:447-457)


def test_toy_dataset_contract():
    """dataset.py:58-71: 10 batches of rand(batch, 128, FFT_SIZE) + a dense-as-sparse text triple."""
    from gan_sass_tf_b200.app import hparams
    from gan_sass_tf_b200.app.datasets.dataset import WhiteNoiseData
    ds = WhiteNoiseData(seed=0)
    with pytest.raises(RuntimeError):
        next(iter(ds.epoch('train', 4)))
    ds.install_and_load()
    batches = list(ds.epoch('train', hparams.BATCH_SIZE * hparams.MAX_N_SIGNAL))
    assert len(batches) == 10
    sig, (ti, tv, ts) = batches[0]
    assert sig.shape == (24, 128, hparams.FFT_SIZE) and sig.dtype == np.float32
    assert 0.0 <= sig.min() and sig.max() < 1.0
    assert ti.shape == (24 * 64, 2) and tv.shape == (24 * 64,) and ts == (24, 64)
    assert tv.min() >= 0 and tv.max() < hparams.CHARSET_SIZE - 1
    again = next(iter(_loaded(WhiteNoiseData(seed=0)).epoch('train', 24)))
    assert np.array_equal(again[0], sig)                         # seeded: reproducible


def _loaded(ds):
    ds.install_and_load()
    return ds


def test_driver_mode_switch_without_gpu():
    from gan_sass_tf_b200 import main as drv
    with pytest.raises(ValueError, match='Unknown mode'):
        drv.main(['-m', 'bogus'])
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match='CUDA only'):       # the mode is wired (tests/test_gpu_app.py trains); no CPU fallback
            drv.main(['-m', 'train', '-ne', '1'])
    with pytest.raises(FileNotFoundError):
        drv.load_wavfile(None)


@pytest.mark.parametrize("rate,n", [(16000, 4000), (8000, 3001), (44100, 9000), (22050, 5000)])
def test_oracle_wav_edges(rate, n):
    """main.py:83-99 / :102-116: resample branch pads to a multiple of FFT_SIZE, the 16 kHz branch does not."""
    N = 256
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) * 3000).astype(np.int16)
    feat = R.load_wave_features(x, rate, N)
    if rate == 16000:
        L = n
    else:
        L = int(max(n * (16000 / rate), 1))
        L += R.resample_pad_size(L, N)
        assert L % N == 0
    assert feat.shape == (R.frame_count(L, N, N // 2)[0], N) and feat.dtype == np.float32
    pcm = R.save_wave_pcm(feat, N)
    assert pcm.dtype == np.int16 and pcm.min() == 0 and pcm.max() >= 32766
