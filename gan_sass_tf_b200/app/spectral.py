"""Host-buffer front end of the spectral path: what ``main.py``'s demo/test modes
and the dataset feature extraction call.

``SpectralPipeline`` owns the device workspaces (waveforms stay resident between
the analysis and the synthesis stage, so the mixture spectrum can be recomputed
instead of stored) and pinned host buffers.  Per batch:

    logfeat = pipe.analyse(wave_host)          # H2D + STFT + to_log   (main.py:97-98, :338)
    mask    = separator(logfeat)               # plugin, on the device  (main.py:340)
    waves   = pipe.synthesise(mask)            # mask + iSTFT + D2H     (main.py:110-111)

Copies and kernels are chunked and overlapped inside libgss
(``gss_stft_h2d`` / ``gss_mask_istft_d2h``).

With ``depth > 1`` the pipeline keeps ``depth`` sets of workspaces ("slots") and the
calls become non-blocking, so a loop over many batches (``main.py:749-771`` run over a
list of clips, or the dataset's feature extraction) keeps both directions of the host
link busy: the upload of batch k+1 overlaps the download of batch k.

    for k, wave in enumerate(batches):
        slot = k % pipe.depth
        if k >= pipe.depth:
            consume(pipe.wait(slot))                       # results of batch k - depth
        feat = pipe.analyse(wave, slot=slot, block=False)
        pipe.synthesise(separator(feat), slot=slot, block=False)
"""
from __future__ import annotations

import numpy as np
import torch

from . import hparams
from .. import _native as _n


class SpectralPipeline:
    def __init__(self, batch, n_samples, n_out, fft_size=None, hop=None, device=None, chunks=4, depth=1, pcm16=False):
        if not torch.cuda.is_available():
            raise RuntimeError("SpectralPipeline needs a CUDA device (no CPU fallback)")
        self.N = hparams.FFT_SIZE if fft_size is None else int(fft_size)
        self.H = (hparams.hop_size() if self.N == hparams.FFT_SIZE else self.N // 2) if hop is None else int(hop)
        self.B, self.n, self.S = int(batch), int(n_samples), int(n_out)
        self.T, self.nadd = _n.frame_count(self.n, self.N, self.H)
        self.L = (self.T - 1) * self.H
        self.chunks = max(1, min(int(chunks), self.B))
        self.depth = max(1, int(depth))
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        d = self.device
        mk = lambda shape, **kw: [torch.empty(shape, dtype=torch.float32, **kw) for _ in range(self.depth)]
        self.wave_ds = mk((self.B, self.n), device=d)
        self.feat_ds = mk((self.B, self.T, self.N), device=d)
        self.out_ds = mk((self.B * self.S, self.L), device=d)
        # pcm16: int16 PCM on the host link in both directions (the WAV sample format of main.py:83 / :116):
        # raw int16 sample values in, per-clip min/max-normalised int16 out (main.py:112-116)
        self.pcm16 = bool(pcm16)
        hdt = torch.int16 if self.pcm16 else torch.float32
        self.wave_hs = [torch.empty((self.B, self.n), dtype=hdt, pin_memory=True) for _ in range(self.depth)]
        self.out_hs = [torch.empty((self.B * self.S, self.L), dtype=hdt, pin_memory=True) for _ in range(self.depth)]
        if self.pcm16:
            mki = lambda shape: [torch.empty(shape, dtype=torch.int16, device=d) for _ in range(self.depth)]
            self.pcm_in_ds = mki((self.B, self.n))
            self.pcm_out_ds = mki((self.B * self.S, self.L))
            self.minmax_ds = mk((self.B * self.S, 2), device=d)
        # slot 0 under the names the blocking API has always used
        self.wave_d, self.feat_d, self.out_d = self.wave_ds[0], self.feat_ds[0], self.out_ds[0]
        self.wave_h, self.out_h = self.wave_hs[0], self.out_hs[0]

    # bytes crossing PCIe per batch
    @property
    def h2d_bytes(self):
        return self.wave_h.numel() * self.wave_h.element_size()

    @property
    def d2h_bytes(self):
        return self.out_h.numel() * self.out_h.element_size()

    def analyse(self, wave_host=None, log=True, slot=0, block=True):
        """``wave_host [B,n]`` float32 (numpy or CPU tensor; ``None`` = already in
        ``self.wave_hs[slot]``) -> log-compressed packed features ``[B,T,N]`` on the device.
        ``block=False`` only enqueues (the features are ordered on the current stream)."""
        wave_h = self.wave_hs[slot]
        if wave_host is not None:
            src = torch.from_numpy(wave_host) if isinstance(wave_host, np.ndarray) else wave_host
            assert tuple(src.shape) == (self.B, self.n), f"analyse: expected {(self.B, self.n)}, got {tuple(src.shape)}"
            assert (src.dtype == torch.int16) == self.pcm16, "analyse: int16 input needs SpectralPipeline(pcm16=True) and vice versa"
            wave_h.copy_(src)
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            if self.pcm16:
                _n.check(_n.lib().gss_stft_h2d_i16_async(wave_h.data_ptr(), self.pcm_in_ds[slot].data_ptr(),
                                                         self.wave_ds[slot].data_ptr(), self.B, self.n, self.n, self.N, self.H,
                                                         _n.FLAG_LOG if log else 0, hparams.EPS,
                                                         self.feat_ds[slot].data_ptr(), self.chunks, st))
                if block:
                    torch.cuda.current_stream().synchronize()
            else:
                fn = _n.lib().gss_stft_h2d if block else _n.lib().gss_stft_h2d_async
                _n.check(fn(wave_h.data_ptr(), self.wave_ds[slot].data_ptr(), self.B, self.n, self.n,
                            self.N, self.H, _n.FLAG_LOG if log else 0, hparams.EPS,
                            self.feat_ds[slot].data_ptr(), self.chunks, st))
        return self.feat_ds[slot]

    def synthesise(self, mask, slot=0, block=True):
        """``mask [B,S,T,N/2]`` on the device -> separated waveforms ``[B*S,(T-1)H]`` in
        pinned host memory (row ``b*S+s``).  With ``block=False`` the returned buffer is
        valid after ``wait(slot)``."""
        assert mask.is_cuda and mask.dtype == torch.float32 and mask.is_contiguous()
        assert tuple(mask.shape) == (self.B, self.S, self.T, self.N // 2), \
            f"synthesise: mask shape {tuple(mask.shape)} != {(self.B, self.S, self.T, self.N // 2)}"
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            if self.pcm16:
                _n.check(_n.lib().gss_mask_istft_d2h_pcm16_async(
                    self.wave_ds[slot].data_ptr(), mask.data_ptr(), self.B, self.S, self.n, self.n, self.N, self.H,
                    self.out_ds[slot].data_ptr(), self.minmax_ds[slot].data_ptr(), self.pcm_out_ds[slot].data_ptr(),
                    self.out_hs[slot].data_ptr(), self.L, self.chunks, st))
                if block:
                    self.wait(slot)
            else:
                fn = _n.lib().gss_mask_istft_d2h if block else _n.lib().gss_mask_istft_d2h_async
                _n.check(fn(self.wave_ds[slot].data_ptr(), mask.data_ptr(), self.B, self.S, self.n, self.n,
                            self.N, self.H, self.out_ds[slot].data_ptr(), self.out_hs[slot].data_ptr(), self.L,
                            self.chunks, st))
        return self.out_hs[slot]

    def wait(self, slot=None):
        """Block until the download of ``slot`` (``None``: every slot) has landed; returns the
        pinned host waveforms of that slot."""
        if slot is None:
            _n.check(_n.lib().gss_wait_host(None))
            return self.out_hs
        _n.check(_n.lib().gss_wait_host(self.out_hs[slot].data_ptr()))
        return self.out_hs[slot]
