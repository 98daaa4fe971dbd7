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
"""
from __future__ import annotations

import numpy as np
import torch

from . import hparams
from .. import _native as _n


class SpectralPipeline:
    def __init__(self, batch, n_samples, n_out, fft_size=None, hop=None, device=None, chunks=4):
        if not torch.cuda.is_available():
            raise RuntimeError("SpectralPipeline needs a CUDA device (no CPU fallback)")
        self.N = hparams.FFT_SIZE if fft_size is None else int(fft_size)
        self.H = (hparams.hop_size() if self.N == hparams.FFT_SIZE else self.N // 2) if hop is None else int(hop)
        self.B, self.n, self.S = int(batch), int(n_samples), int(n_out)
        self.T, self.nadd = _n.frame_count(self.n, self.N, self.H)
        self.L = (self.T - 1) * self.H
        self.chunks = max(1, min(int(chunks), self.B))
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        d = self.device
        self.wave_d = torch.empty((self.B, self.n), dtype=torch.float32, device=d)
        self.feat_d = torch.empty((self.B, self.T, self.N), dtype=torch.float32, device=d)
        self.out_d = torch.empty((self.B * self.S, self.L), dtype=torch.float32, device=d)
        self.wave_h = torch.empty((self.B, self.n), dtype=torch.float32, pin_memory=True)
        self.out_h = torch.empty((self.B * self.S, self.L), dtype=torch.float32, pin_memory=True)

    # bytes crossing PCIe per batch
    @property
    def h2d_bytes(self):
        return self.wave_h.numel() * 4

    @property
    def d2h_bytes(self):
        return self.out_h.numel() * 4

    def analyse(self, wave_host=None, log=True):
        """``wave_host [B,n]`` float32 (numpy or CPU tensor; ``None`` = already in
        ``self.wave_h``) -> log-compressed packed features ``[B,T,N]`` on the device."""
        if wave_host is not None:
            src = torch.from_numpy(wave_host) if isinstance(wave_host, np.ndarray) else wave_host
            assert tuple(src.shape) == (self.B, self.n), f"analyse: expected {(self.B, self.n)}, got {tuple(src.shape)}"
            self.wave_h.copy_(src)
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            _n.check(_n.lib().gss_stft_h2d(self.wave_h.data_ptr(), self.wave_d.data_ptr(), self.B, self.n, self.n,
                                           self.N, self.H, _n.FLAG_LOG if log else 0, hparams.EPS,
                                           self.feat_d.data_ptr(), self.chunks, st))
        return self.feat_d

    def synthesise(self, mask):
        """``mask [B,S,T,N/2]`` on the device -> separated waveforms ``[B*S,(T-1)H]`` in
        pinned host memory (row ``b*S+s``)."""
        assert mask.is_cuda and mask.dtype == torch.float32 and mask.is_contiguous()
        assert tuple(mask.shape) == (self.B, self.S, self.T, self.N // 2), \
            f"synthesise: mask shape {tuple(mask.shape)} != {(self.B, self.S, self.T, self.N // 2)}"
        with torch.cuda.device(self.device):
            st = torch.cuda.current_stream().cuda_stream
            _n.check(_n.lib().gss_mask_istft_d2h(self.wave_d.data_ptr(), mask.data_ptr(), self.B, self.S, self.n, self.n,
                                                 self.N, self.H, self.out_d.data_ptr(), self.out_h.data_ptr(), self.L,
                                                 self.chunks, st))
        return self.out_h
