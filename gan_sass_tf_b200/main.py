"""Command-line driver of the spectral path: the reference's ``main.py`` (flags :689-707, mode
switch :736-781) with the SciPy / TF pieces of the path replaced by the device ops.

    python -m gan_sass_tf_b200.main -m demo -if clip.wav      # load -> separator -> save  (main.py:749-771)
    python -m gan_sass_tf_b200.main -m test                   # SNR sweep over the test subset (main.py:652-668)

``load_wavfile`` / ``save_wavfile`` keep the reference's names and contracts (main.py:67-116).
Training the GAN (``-m train``, main.py:575-651) is out of scope of this repository:
the mode exists and raises ``NotImplementedError`` naming the boundary.
"""
from __future__ import annotations

import argparse
import os
from sys import stdout

import numpy as np
import scipy.io.wavfile
import torch

from .app import hparams, ops
from .app import modules  # noqa: F401  (registers the plugins)
from .app import datasets  # noqa: F401  (registers the datasets)

g_args = None
g_model = None
g_dataset = None


def load_wavfile(filename, device=None):
    """main.py:67-99: read a WAV file, take the first channel, resample to 16 kHz (+ the pad rule
    of main.py:93, applied only in the resample branch, as upstream), STFT, pack.
    Returns a float32 CUDA tensor ``[time, FFT_SIZE]``."""
    if filename is None:
        raise FileNotFoundError('WAV file not specified, please specify via --input-file argument.')
    smprate, data = scipy.io.wavfile.read(filename)
    if data.ndim != 1:
        print('Warning: WAV file is not of single channel, using the first channel')
        data = data[(0,) * (data.ndim - 1)]          # upstream indexing (main.py:88), kept verbatim in meaning
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    x = torch.from_numpy(np.ascontiguousarray(data)).to(dev)
    if smprate != hparams.SAMPLE_RATE:
        nsmp = x.shape[-1]
        new_nsmp = int(max(nsmp * (hparams.SAMPLE_RATE / smprate), 1))
        x = ops.resample(x.to(torch.float64), new_nsmp).to(torch.float32)     # SciPy resamples in float64
        x = torch.nn.functional.pad(x, (0, ops.resample_pad_size(x.shape[-1])))
    if x.dtype not in (torch.int16, torch.float32):
        x = x.to(torch.float32)
    return ops.stft(x.reshape(1, -1), hparams.FFT_SIZE, hparams.hop_size())[0]


def save_wavfile(filename, feature):
    """main.py:102-116: packed feature ``[time, FFT_SIZE]`` -> iSTFT -> shift/scale to int16 -> WAV."""
    f = feature if isinstance(feature, torch.Tensor) else torch.from_numpy(np.asarray(feature, dtype=np.float32)).cuda()
    wave = ops.istft(f.reshape(1, f.shape[-2], f.shape[-1]), hparams.hop_size())
    pcm = ops.wav16_normalise(wave)[0]
    scipy.io.wavfile.write(filename, hparams.SAMPLE_RATE, pcm.cpu().numpy())


class Model(object):
    """The inference slice of the reference's ``Model`` (main.py:161-565): mixture features ->
    ``to_log_signal`` -> separator plugin -> separated features (main.py:338-342).  A separator with
    ``EMITS_MASK`` returns masks, applied to the mixture spectrum."""
    def __init__(self, name='Model'):
        self.name = name
        self.separator = None

    def build(self):
        self.separator = hparams.get_separator()(self, 'separator')

    def infer(self, mixture_feature):
        """``[B,T,N]`` packed mixture features -> ``[B*(MAX_N_SIGNAL+1), T, N]`` separated features."""
        with torch.no_grad():
            log_mix = ops.to_log_signal(mixture_feature)
            out = self.separator(log_mix, s_dropout_keep=1.)
            if getattr(self.separator, 'EMITS_MASK', False):
                return ops.apply_mask(mixture_feature, out)
            return ops.to_exp_signal(out)

    def test(self, dataset):
        """main.py:652-668 reduced to the spectral metrics: mean best-output SNR and auto-encoder loss."""
        n_sig, rep, nb = hparams.MAX_N_SIGNAL, {'SNR': 0.0, 'AE': 0.0}, 0
        for data_pt in dataset.epoch('test', hparams.BATCH_SIZE * n_sig):
            src = data_pt[0]
            src = src if isinstance(src, torch.Tensor) else torch.from_numpy(np.asarray(src)).cuda()
            mix = ops.mix_signals(src, n_sig)
            sep = self.infer(mix)
            rep['SNR'] += float(ops.snr_metric(src, sep, n_sig))
            rep['AE'] += float(ops.ae_loss(sep, mix, n_sig + 1))
            nb += 1
        return {k: v / max(nb, 1) for k, v in rep.items()}


def main(argv=None):
    global g_args, g_model, g_dataset
    parser = argparse.ArgumentParser()
    parser.add_argument('-n', '--name', default='UnamedExperiment', help='name of experiment, affects checkpoint saves')
    parser.add_argument('-m', '--mode', default='train', help='Mode, "train", "test", "demo" or "interactive"')
    parser.add_argument('-i', '--input-pfile', help='path to input model parameter file')
    parser.add_argument('-o', '--output-pfile', help='path to output model parameters file')
    parser.add_argument('-ne', '--num-epoch', type=int, default=10, help='number of training epoch')
    parser.add_argument('--no-save-on-epoch', action='store_true', help="don't save parameter after each epoch")
    parser.add_argument('--no-test-on-epoch', action='store_true', help="don't sweep test set after training epoch")
    parser.add_argument('-if', '--input-file', help='input WAV file for "demo" mode')
    g_args = parser.parse_args(argv)

    stdout.write('Preparing dataset "%s" ... ' % hparams.DATASET_TYPE)
    stdout.flush()
    g_dataset = hparams.get_dataset()()
    g_dataset.install_and_load()
    stdout.write('done\n')
    print('Separator type: "%s"' % hparams.SEPARATOR_TYPE)
    g_model = Model(name=g_args.name)
    g_model.build()

    if g_args.mode == 'interactive':
        print('Now in interactive mode, you should run this with python -i')
        return
    elif g_args.mode == 'train':
        raise NotImplementedError('GAN training (main.py:575-651) is outside the spectral hot path; see DESIGN.md 6')
    elif g_args.mode == 'test':
        print(' '.join('%s=%s' % kv for kv in g_model.test(g_dataset).items()))
    elif g_args.mode == 'demo':
        if g_args.input_file is None:
            filename = 'demo.wav'
            for features in g_dataset.epoch('test', hparams.MAX_N_SIGNAL):
                break
            f0 = features[0] if isinstance(features[0], torch.Tensor) else torch.from_numpy(np.asarray(features[0])).cuda()
            save_wavfile(filename, f0[0] + f0[1])
            features = f0.sum(dim=0, keepdim=True)
        else:
            filename = g_args.input_file
            features = load_wavfile(g_args.input_file)[None]
        # one clip is one batch row: the reference tiles it to BATCH_SIZE and throws 7/8 of the work
        # away (main.py:763-767, "inefficient !")
        signals = g_model.infer(features)[:(hparams.MAX_N_SIGNAL + 1)]
        filename, fileext = os.path.splitext(filename)
        for i, s in enumerate(signals):
            save_wavfile(filename + ('_separated_%d' % (i + 1)) + fileext, s)
    else:
        raise ValueError('Unknown mode "%s"' % g_args.mode)


if __name__ == '__main__':
    main()
