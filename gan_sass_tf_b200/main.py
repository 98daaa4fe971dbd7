"""Command-line driver of the spectral path: the reference's ``main.py`` (flags :689-707, mode
switch :736-781) with the SciPy / TF pieces of the path replaced by the device ops.

    python -m gan_sass_tf_b200.main -m demo -if clip.wav      # load -> separator -> save  (main.py:749-771)
    python -m gan_sass_tf_b200.main -m test                   # SNR sweep over the test subset (main.py:652-668)

    python -m gan_sass_tf_b200.main -m train -ne 1            # epochs through the native ops (main.py:568-624)

``load_wavfile`` / ``save_wavfile`` keep the reference's names and contracts (main.py:67-116).
``-m train`` drives the separator plugin through the differentiable native ops (mix + to_log -> plugin ->
to_exp / mask -> auto-encoder loss, Adam, per-epoch test sweep and checkpoint, NaN rollback): the spectral
slice of main.py:568-624.  The GAN objective, the discriminator and the ASR branch (main.py:363-445,
:522-565) stay out of scope.
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
        # upstream indexes data[(0,)*(ndim-1)] (main.py:88), which for SciPy's (nsamples, nchannels) layout picks the
        # first sample FRAME, not the first channel; its warning text states the intent
        print('Warning: WAV file is not of single channel, using the first channel')
        data = data.reshape(data.shape[0], -1)[:, 0]
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    x = torch.from_numpy(np.ascontiguousarray(data)).to(dev)
    if smprate != hparams.SAMPLE_RATE:
        nsmp = x.shape[-1]
        new_nsmp = int(max(nsmp * (hparams.SAMPLE_RATE / smprate), 1))
        x = ops.resample(x.to(torch.float64), new_nsmp).to(torch.float32)     # SciPy resamples in float64
        x = torch.nn.functional.pad(x, (0, ops.resample_pad_size(x.shape[-1])))
    if x.dtype not in (torch.int16, torch.float32):
        x = x.to(torch.float32)
    if x.shape[-1] < hparams.FFT_SIZE:
        # SciPy would silently shrink nperseg to the clip length (a feature width the model cannot take); libgss
        # rejects n < FFT_SIZE.  Say so here instead of failing inside the transform.
        raise ValueError('WAV file "%s" holds %d samples after resampling, fewer than FFT_SIZE=%d' % (filename, x.shape[-1], hparams.FFT_SIZE))
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

    def forward_train(self, src, generator=None):
        """One differentiable pass of the spectral slice of the training graph (main.py:328-361): sources ->
        mix (+ noise) -> to_log -> separator -> to_exp (or masks on the mixture) -> auto-encoder loss; the SNR
        metric of main.py:446-457 rides along without gradients.  Returns ``(loss, snr)``."""
        n_sig = hparams.MAX_N_SIGNAL
        mix, log_mix = ops.mix_signals(src, n_sig, noise_stddev=0.1, generator=generator, log=True)
        out = self.separator(log_mix, s_dropout_keep=hparams.DROPOUT_KEEP_PROB if hasattr(hparams, 'DROPOUT_KEEP_PROB') else 1.)
        sep = ops.apply_mask(mix, out) if getattr(self.separator, 'EMITS_MASK', False) else ops.to_exp_signal(out)
        B, S = mix.shape[0], n_sig + 1
        loss = (sep.reshape(B, S, *mix.shape[1:]).sum(dim=1) - mix).pow(2).mean()      # main.py:353-361 (autograd form)
        with torch.no_grad():
            snr = ops.snr_metric(src, sep, n_sig)
        return loss, snr

    def parameters(self):
        return list(self.separator.p.parameters()) if hasattr(self.separator, 'p') else []

    def save_params(self, filename, step=None):
        """main.py:278-285: ``saves/<name>`` checkpoints of the trainable variables (torch.save instead of tf.train.Saver)"""
        os.makedirs(os.path.dirname(filename) or '.', exist_ok=True)
        torch.save({k: v.state_dict() for k, v in self.separator.p.layers.items()}, filename)

    def load_params(self, filename):
        """main.py:287-292"""
        state = torch.load(filename, map_location='cuda')
        for k, sd in state.items():
            if k in self.separator.p.layers:
                self.separator.p.layers[k].load_state_dict(sd)

    def train(self, dataset, n_epoch, lr=1e-3, save_on_epoch=True, test_on_epoch=True, seed=0, out=stdout):
        """The epoch loop of main.py:568-624 for the spectral slice: per batch one Adam step on the auto-encoder
        loss (the generator's ``ae`` term, main.py:481-484, without the GAN term), a ':' tick per batch, epoch
        means, per-epoch checkpoint ``saves/<name>_e<k>`` (main.py:600), NaN rollback to the previous one
        (main.py:587-596) and the test sweep (main.py:605-607).  Returns the per-epoch reports."""
        if not torch.cuda.is_available():
            raise RuntimeError('train: the spectral ops run on CUDA only (no CPU fallback)')
        n_sig = hparams.MAX_N_SIGNAL
        gen = torch.Generator(device='cuda').manual_seed(seed)
        opt = None
        reports = []
        for i_epoch in range(n_epoch):
            rep, nb = {'AE': 0.0, 'SNR': 0.0}, 0
            for data_pt in dataset.epoch('train', hparams.BATCH_SIZE * n_sig, shuffle=True):
                src = data_pt[0]
                src = src if isinstance(src, torch.Tensor) else torch.from_numpy(np.asarray(src)).cuda()
                loss, snr = self.forward_train(src, gen)
                if opt is None:                                   # the plugin creates its layers on first use
                    opt = torch.optim.Adam(self.parameters(), lr=lr)
                    loss, snr = self.forward_train(src, gen)
                opt.zero_grad(set_to_none=True)
                loss.backward()
                opt.step()
                rep['AE'] += float(loss.detach()); rep['SNR'] += float(snr); nb += 1
                out.write(':'); out.flush()
            rep = {k: v / max(nb, 1) for k, v in rep.items()}
            out.write('\n')
            ckpt = os.path.join('saves', '%s_e%d' % (self.name, i_epoch + 1))
            if any(v != v for v in rep.values()):                 # NaN: roll back (main.py:587-596)
                prev = os.path.join('saves', '%s_e%d' % (self.name, i_epoch))
                if i_epoch == 0 or not os.path.exists(prev):
                    raise FloatingPointError('NaN in the first epoch (epoch %d): nothing to roll back to' % (i_epoch + 1))
                print('Epoch %d has NaN, reloading epoch %d' % (i_epoch + 1, i_epoch))
                self.load_params(prev)
                opt = torch.optim.Adam(self.parameters(), lr=lr)
                continue
            if save_on_epoch:
                self.save_params(ckpt)
            print('Epoch %d/%d %s' % (i_epoch + 1, n_epoch, ' '.join('%s=%.6g' % kv for kv in rep.items())))
            if test_on_epoch:
                t = self.test(dataset)
                rep.update({'test_' + k: v for k, v in t.items()})
                print('  test ' + ' '.join('%s=%.6g' % kv for kv in t.items()))
            reports.append(rep)
        return reports

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
        if g_args.input_pfile is not None:
            g_model.load_params(g_args.input_pfile)
        g_model.train(g_dataset, g_args.num_epoch, save_on_epoch=not g_args.no_save_on_epoch,
                      test_on_epoch=not g_args.no_test_on_epoch)
        if g_args.output_pfile is not None:
            g_model.save_params(g_args.output_pfile)
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
