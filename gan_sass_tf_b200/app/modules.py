"""Plug-in sub-modules: same base classes, constructor and call contracts as the
reference's ``app/modules.py:10-108``, on ``torch.Tensor`` (CUDA, float32, last
dim ``FFT_SIZE``) instead of TF1 symbolic tensors.

    Separator.__call__(s_mixture[B,T,N], s_dropout_keep=1.)      -> [B*(n_sig+1), T, N]
    Discriminator.__call__(s_signals[B',T,N], s_texts=None, ...)  -> [B', 3]
    Recognizer.__call__(s_signals, s_dropout_keep=1.)             -> [B', T, CHARSET+1]

User extensions register exactly as in the reference's README:

    @hparams.register_separator('my-separator')
    class MySeparator(Separator): ...

The learnt models themselves (bi-LSTM / bi-GRU stacks, modules.py:194-444) are out
of scope; the toy MLPs (modules.py:111-191) are
kept as stand-ins so the spectral path can be driven end to end, plus one
mask-emitting separator, which the reference only has as a stub
(modules.py:447-457, ``NotImplementedError``).
"""
from __future__ import annotations

import zlib

import torch

from . import hparams


class ModelModule(object):
    """abstract sub-module of a model (modules.py:10-21)"""
    def __init__(self, model, name):
        pass

    def __call__(self, s_dropout_keep=1.):
        raise NotImplementedError()


class Separator(ModelModule):
    """separate signal and noise from the input mixture (modules.py:24-46)"""
    def __init__(self, model, name):
        self.name = name

    def __call__(self, s_mixture, s_dropout_keep=1.):
        raise NotImplementedError()


class Recognizer(ModelModule):
    """speech recognizer (modules.py:49-82); ``IS_CTC`` tells whether logits are returned"""
    IS_CTC = False

    def __init__(self, model, name):
        pass

    def __call__(self, s_signals, s_dropout_keep=1.):
        raise NotImplementedError()


class Discriminator(ModelModule):
    """audio (+ optional text) -> 3-way logits (modules.py:85-108)"""
    def __init__(self, model, name):
        self.name = name

    def __call__(self, s_signals, s_texts, s_dropout_keep=1.):
        raise NotImplementedError()


def _leaky(x):
    # ops.relu(x, RELU_LEAKAGE) of the reference: max(x, leak*x)
    return torch.nn.functional.leaky_relu(x, hparams.RELU_LEAKAGE)


class _Params:
    """lazily created, seeded nn.Linear layers (the reference creates variables at graph-build time)"""
    def __init__(self):
        self.layers = {}

    def linear(self, key, n_in, n_out, device, bias=True):
        if key not in self.layers:
            # a stable hash: Python salts str hashes per process, which made the "seeded" weights differ between runs and ranks
            g = torch.Generator().manual_seed((zlib.crc32(key.encode()) + 7919 * int(getattr(hparams, 'SEED', 0))) % (2 ** 31))
            lin = torch.nn.Linear(n_in, n_out, bias=bias)
            with torch.no_grad():
                lin.weight.copy_(torch.randn(n_out, n_in, generator=g) * (1.0 / n_in) ** 0.5)
                if bias:
                    lin.bias.zero_()
            self.layers[key] = lin.to(device)
        return self.layers[key]

    def parameters(self):
        for lin in self.layers.values():
            yield from lin.parameters()


@hparams.register_separator('toy')
class ToySeparator(Separator):
    """2-layer MLP emitting the separated log-features directly (modules.py:111-135)"""
    def __init__(self, model, name):
        self.name = name
        self.p = _Params()

    def __call__(self, s_signals, s_dropout_keep=1.):
        N, S = hparams.FFT_SIZE, hparams.MAX_N_SIGNAL + 1
        B = s_signals.shape[0]
        mid = _leaky(self.p.linear(self.name + '/linear0', N, 2 * N, s_signals.device)(s_signals))
        out = self.p.linear(self.name + '/linear1', 2 * N, N * S, s_signals.device)(mid)
        out = out.reshape(B, -1, S, N).transpose(1, 2)          # [B, S, T, N]
        return out.reshape(B * S, -1, N)                        # row b*S+s (modules.py:396-399)


@hparams.register_separator('toy-mask')
class ToyMaskSeparator(Separator):
    """Mask-emitting stand-in (no reference implementation: its only mask separator is the
    'dc-v1' stub, modules.py:447-457).  Returns soft masks ``[B, S, T, N/2]`` in (0,1) that
    sum to one over sources, to be applied with ``ops.mask_istft`` / ``ops.apply_mask``."""
    EMITS_MASK = True

    def __init__(self, model, name):
        self.name = name
        self.p = _Params()

    def __call__(self, s_signals, s_dropout_keep=1.):
        N, S = hparams.FFT_SIZE, hparams.MAX_N_SIGNAL + 1
        B = s_signals.shape[0]
        mid = _leaky(self.p.linear(self.name + '/linear0', N, N, s_signals.device)(s_signals))
        logits = self.p.linear(self.name + '/linear1', N, (N // 2) * S, s_signals.device)(mid)
        logits = logits.reshape(B, -1, S, N // 2).transpose(1, 2)
        return torch.softmax(logits, dim=1).contiguous()


@hparams.register_recognizer('toy')
class ToyRecognizer(Recognizer):
    """always outputs CTC logits with a 2-layer MLP (modules.py:138-165)"""
    IS_CTC = True

    def __init__(self, model, name):
        self.name = name
        self.p = _Params()

    def __call__(self, s_signals, s_dropout_keep=1.):
        N = s_signals.shape[-1]
        mid = _leaky(self.p.linear(self.name + '/linear0', N, 2 * N, s_signals.device)(s_signals))
        return self.p.linear(self.name + '/linear1', 2 * N, hparams.CHARSET_SIZE + 1, s_signals.device)(mid)


@hparams.register_discriminator('toy')
class ToyDiscriminator(Discriminator):
    """time-mean of the signal (+ text) -> MLP -> 3 logits (modules.py:168-191)"""
    def __init__(self, model, name):
        self.name = name
        self.p = _Params()

    def __call__(self, s_signals, s_texts=None, s_dropout_keep=1.):
        s_input = s_signals.mean(dim=-2)
        if s_texts is not None:
            s_input = torch.cat([s_input, s_texts.mean(dim=-2)], dim=-1)
        d = s_input.shape[-1]
        mid = _leaky(self.p.linear(self.name + f'/linear0_{d}', d, 2 * d, s_input.device)(s_input))
        return self.p.linear(self.name + f'/linear1_{d}', 2 * d, 3, s_input.device)(mid)


@hparams.register_separator('dc-v1')
class DeepClusterSeparator(Separator):
    """unimplemented upstream as well (modules.py:447-457)"""
    def __init__(self, model, name):
        raise NotImplementedError()
