"""Dataset base class and the toy generator: same contracts as the reference's
``app/datasets/dataset.py`` (base :8-42, ``WhiteNoiseData`` :45-76)."""
from __future__ import annotations

from itertools import product

import numpy as np

from .. import hparams


class Dataset(object):
    def __init__(self):
        self.is_loaded = False

    def epoch(self, subset, batch_size, shuffle=False):
        """Iterator over batches ``(signals, (text_indices, text_values, text_shape))``:
        ``signals`` rank-3 float32 ``[batch_size, time, FFT_SIZE]``, the texts a sparse triple
        (dataset.py:12-27)."""
        raise NotImplementedError()

    def install_and_load(self):
        raise NotImplementedError()

    def encode_from_str(arr):
        raise NotImplementedError()

    def decode_to_str(arr):
        raise NotImplementedError()


@hparams.register_dataset('toy')
class WhiteNoiseData(Dataset):
    """Always generates uniform noise: 10 batches of ``rand(batch, 128, FFT_SIZE)`` packed
    features and dummy texts of length 64 (dataset.py:45-76).  Upstream is unseeded; ``seed``
    makes the config-C1 runs reproducible, ``device`` draws the signals on the GPU (torch's
    generator, different stream than NumPy's) so no host copy is needed."""
    LENGTH = 128
    TEXT_LENGTH = 64
    N_BATCHES = 10

    def __init__(self, seed=None, device=None):
        self.is_loaded = False
        self.seed = seed
        self.device = device

    def epoch(self, subset, batch_size, shuffle=False):
        if not self.is_loaded:
            raise RuntimeError('Dataset is not loaded.')
        rng = np.random.default_rng(self.seed) if self.seed is not None else np.random
        gen = None
        if self.device is not None:
            import torch
            gen = torch.Generator(device=self.device)
            gen.manual_seed(0 if self.seed is None else int(self.seed))
        for _ in range(self.N_BATCHES):
            if gen is not None:
                import torch
                signal = torch.rand((batch_size, self.LENGTH, hparams.FFT_SIZE), dtype=torch.float32,
                                    device=self.device, generator=gen)
            elif self.seed is not None:
                signal = rng.random((batch_size, self.LENGTH, hparams.FFT_SIZE), dtype=np.float32)
            else:
                signal = np.random.rand(batch_size, self.LENGTH, hparams.FFT_SIZE).astype(hparams.FLOATX)
            text_indices = np.asarray(list(product(range(batch_size), range(self.TEXT_LENGTH))), dtype=hparams.INTX)
            if self.seed is not None:
                text_values = rng.integers(0, hparams.CHARSET_SIZE - 1, (batch_size, self.TEXT_LENGTH)).astype(hparams.INTX).ravel()
            else:
                text_values = np.asarray(np.random.randint(0, hparams.CHARSET_SIZE - 1, (batch_size, self.TEXT_LENGTH),
                                                           dtype=hparams.INTX).flat)
            yield signal, (text_indices, text_values, (batch_size, self.TEXT_LENGTH))

    def install_and_load(self):
        self.is_loaded = True
