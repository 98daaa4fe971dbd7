"""Device-resident waveform dataset.

The reference bakes ``FFT_SIZE`` into its pickles: ``install.sh`` runs
``scipy.signal.stft(nperseg=FFT_SIZE)`` over every TIMIT file once
(``TIMIT/process.py:80-161``) and training reads the packed features
(``app/datasets/timit.py:30-107``), so changing ``FFT_SIZE`` means re-running the
install.  Here the utterances stay as int16 PCM in HBM (2 bytes/sample, what the WAV
files hold) and every batch is transformed on the device when it is drawn
(``gss_stft_packed_i16``: the int16 -> float convert is fused into the load stage), so
``FFT_SIZE`` / ``HOP_SIZE`` are read at ``epoch()`` time like every other hyper-parameter.

Batches are length-bucketed (utterances sorted by length, a batch = neighbours) and
zero-padded to the longest member, the padding rule of ``timit.py:47-52``; the PCM scale
(``process.py:97`` feeds raw int16 sample values to SciPy) is kept.
"""
from __future__ import annotations

import numpy as np
import torch
from scipy.signal import lfilter

import glob
import os

from .. import hparams, ops
from ... import _native as _n
from .dataset import Dataset


@hparams.register_dataset('wave')
class WaveformData(Dataset):
    def __init__(self, device=None, seed=0):
        self.is_loaded = False
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.seed = seed
        self.subset = {}
        self.n_epochs = 0          # advances the shuffle seed: every epoch gets its own order

    def add_subset(self, name, waves):
        """``waves``: list of 1-D int16 (or float in [-1, 1), scaled to int16) arrays at 16 kHz."""
        pcm = []
        for w in waves:
            w = np.asarray(w)
            if w.dtype != np.int16:
                w = np.clip(np.round(w.astype(np.float64) * 32768.0), -32768, 32767).astype(np.int16)
            if w.ndim != 1 or w.size < 1:
                raise ValueError("WaveformData: utterances must be non-empty 1-D arrays")
            pcm.append(w)
        order = np.argsort([len(w) for w in pcm], kind="stable")
        lengths = np.asarray([len(pcm[i]) for i in order], dtype=np.int64)
        offsets = np.concatenate([[0], np.cumsum(lengths)])
        flat = torch.from_numpy(np.concatenate([pcm[i] for i in order])).to(self.device)
        self.subset[name] = (flat, offsets, lengths,
                             torch.from_numpy(offsets[:-1].astype(np.int64)).to(self.device), torch.from_numpy(lengths).to(self.device))
        self.is_loaded = True

    def add_directory(self, name, path, pattern="**/*.wav", skip_prefixes=("sa",)):
        """File-backed corpus: every 16 kHz mono WAV under ``path`` becomes an utterance of subset ``name`` - the
        per-file loop of ``TIMIT/process.py:89-110`` (which skips the dialect sentences ``sa*`` and raises on any other
        sampling rate, process.py:90-96) without its STFT: the waveforms are stored, the transform runs per batch."""
        import scipy.io.wavfile
        waves = []
        for fname in sorted(glob.glob(os.path.join(path, pattern), recursive=True)):
            if os.path.basename(fname).lower().startswith(tuple(skip_prefixes)):
                continue
            rate, w = scipy.io.wavfile.read(fname)
            if rate != hparams.SAMPLE_RATE:
                raise ValueError('Sampling rate of "%s" is %d, must be %d' % (fname, rate, hparams.SAMPLE_RATE))   # process.py:95-96
            if w.ndim != 1:
                w = w.reshape(w.shape[0], -1)[:, 0]
            waves.append(w)
        if not waves:
            raise FileNotFoundError('no WAV files matching "%s" under "%s"' % (pattern, path))
        self.add_subset(name, waves)
        return len(waves)

    def install_and_load(self):
        """Synthetic stand-in corpus (no TIMIT in this environment): 2 x 64 utterances of
        low-passed noise, 1-3 s, so that ``-m test`` / ``-m demo`` have something to draw."""
        rng = np.random.default_rng(self.seed)
        for name, count in (("train", 64), ("test", 64)):
            waves = []
            for _ in range(count):
                n = int(rng.integers(16000, 48001))
                x = rng.normal(0, 0.05, n)
                waves.append(np.clip(lfilter([1.0], [1.0, -0.95], x), -1, 1))      # 1-pole low-pass (SURVEY 8d C2 recipe)
            self.add_subset(name, waves)

    def epoch(self, subset, batch_size, shuffle=False):
        """Yields ``(signals [batch, T, FFT_SIZE] on the device, lengths [batch] in frames)`` with
        ``signals`` = packed STFT features of the zero-padded batch (``timit.py:47-52``)."""
        if not self.is_loaded:
            raise RuntimeError('Dataset is not loaded.')
        if subset not in self.subset:
            raise KeyError('Unknown subset "%s", valid options are %s' % (subset, list(self.subset.keys())))
        flat, offsets, lengths, offsets_d, lengths_d = self.subset[subset]
        N, H = hparams.FFT_SIZE, hparams.hop_size()
        nb = (len(lengths) + batch_size - 1) // batch_size
        order = np.random.default_rng([self.seed, self.n_epochs]).permutation(nb) if shuffle else np.arange(nb)
        self.n_epochs += 1
        for bi in order:
            lo = bi * batch_size
            idx = np.arange(lo, lo + batch_size) % len(lengths)          # the last batch wraps (timit.py:74 re-uses the tail)
            n_max = max(int(lengths[idx].max()), N)
            n_max += n_max % 2                                           # even rows: aligned 2-sample loads in the transform
            batch = torch.empty((batch_size, n_max), dtype=torch.int16, device=self.device)
            idx_d = torch.from_numpy(idx.astype(np.int64)).to(self.device, non_blocking=True)
            with torch.cuda.device(self.device):                         # one gather launch instead of a per-row copy loop
                _n.check(_n.lib().gss_gather_rows_i16(flat.data_ptr(), offsets_d.data_ptr(), lengths_d.data_ptr(), idx_d.data_ptr(),
                                                      batch_size, n_max, batch.data_ptr(), torch.cuda.current_stream().cuda_stream))
            feats = ops.stft(batch, N, H)
            frames = torch.as_tensor((lengths[idx] + ((-lengths[idx]) % H) % N) // H + 1)
            yield feats, frames
