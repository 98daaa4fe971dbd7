"""Spectral ops on CUDA tensors - the drop-in for the reference's ``app/ops.py``.

Same style as the reference: free functions, tensor in / tensor out, last axis =
``hparams.FFT_SIZE`` read at call time, any leading rank, shape errors by
``assert`` (ops.py:26-31, :176, :205-206).  Kept names / argument order:
``to_log_signal`` (ops.py:228-238), ``to_exp_signal`` (:241-251), ``batch_snr``
(:162-189), ``batch_cross_snr`` (:191-225).  New ops replace the host-side SciPy
calls: ``stft`` / ``stft_log`` (main.py:97-98 + :338), ``istft`` (main.py:110-111),
``apply_mask`` and the fused ``mask_istft``.

Every function enqueues hand-written sm_100a kernels (``csrc/``) on the current
torch CUDA stream through the C ABI; PyTorch only owns the buffers.  CPU tensors
are rejected - there is no fallback path.
"""
from __future__ import annotations

import torch

from . import hparams
from .. import _native as _n


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _dev(t: torch.Tensor, name: str) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(f"{name}: tensor is on {t.device}; gan_sass_tf_b200 ops run on CUDA only (no CPU fallback)")
    return t


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    _dev(t, name)
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    return t.contiguous()


def _nh(fft_size, hop):
    """(N, H): explicit arguments win; else ``hparams.FFT_SIZE`` and, for that size,
    ``hparams.hop_size()``; any other size defaults to SciPy's ``N // 2``."""
    N = hparams.FFT_SIZE if fft_size is None else int(fft_size)
    if hop is not None:
        return N, int(hop)
    return N, (hparams.hop_size() if N == hparams.FFT_SIZE else N // 2)


# ---------------------------------------------------------------------------
# transforms
# ---------------------------------------------------------------------------
def stft(wave, fft_size=None, hop=None, log=False):
    """``scipy.signal.stft(x, nperseg=N)[2]`` + ``utils.spectrum_to_feature``
    (main.py:97-98) for a batch: ``wave [..., n]`` (float32 or int16) ->
    packed feature ``[..., T, N]`` float32.  ``log=True`` fuses ``to_log_signal``.
    Differentiable in ``wave`` (adjoint = scaled overlap-add iSTFT, ``_stft_adjoint``)."""
    _dev(wave, "stft")
    N, H = _nh(fft_size, hop)
    if torch.is_grad_enabled() and wave.requires_grad:
        return _Stft.apply(wave, N, H, bool(log))
    return _stft_fwd(wave, N, H, log)


def _stft_fwd(wave, N, H, log=False):
    if wave.dtype == torch.int16:
        w, fn = wave.contiguous(), _n.lib().gss_stft_packed_i16
    else:
        w, fn = _f32c(wave, "stft"), _n.lib().gss_stft_packed
    lead, n = w.shape[:-1], w.shape[-1]
    assert n >= 1, "stft: empty waveform"
    T, _ = _n.frame_count(n, N, H)
    B = 1
    for d in lead:
        B *= d
    feat = torch.empty(lead + (T, N), dtype=torch.float32, device=w.device)
    with torch.cuda.device(w.device):
        _n.check(fn(w.data_ptr(), B, n, n, N, H, _n.FLAG_LOG if log else 0, hparams.EPS, feat.data_ptr(), _stream()))
    return feat


def stft_log(wave, fft_size=None, hop=None):
    """STFT with ``to_log_signal`` fused into the epilogue: what the separator consumes (main.py:338)."""
    return stft(wave, fft_size, hop, log=True)


def istft(feature, hop=None, length=None, exp=False):
    """``utils.feature_to_spectrum`` + ``scipy.signal.istft(Z, nperseg=N)``
    (main.py:110-111): ``feature [..., T, N]`` -> ``[..., (T-1)*H]``.  ``exp=True``
    applies ``to_exp_signal`` first (main.py:342).  ``length`` trims the tail
    (SciPy itself returns the ``nadd`` padding samples, K4).  Differentiable in ``feature``
    (adjoint = scaled STFT of the normalised gradient, ``_istft_adjoint``)."""
    _dev(feature, "istft")
    assert feature.dim() >= 2, "istft: feature must be [..., T, N]"
    N, H = _nh(feature.shape[-1], hop)
    if torch.is_grad_enabled() and feature.requires_grad:
        out = _Istft.apply(feature, H, bool(exp))
    else:
        out = _istft_fwd(feature, H, exp)
    return out if length is None else out[..., :length]


def _istft_fwd(feature, H, exp=False):
    f = _f32c(feature, "istft")
    T, N = f.shape[-2], f.shape[-1]
    lead = f.shape[:-2]
    R = 1
    for d in lead:
        R *= d
    L = (T - 1) * H
    out = torch.empty(lead + (L,), dtype=torch.float32, device=f.device)
    with torch.cuda.device(f.device):
        _n.check(_n.lib().gss_istft_packed(f.data_ptr(), R, T, N, H, _n.FLAG_EXP if exp else 0, hparams.EPS,
                                           out.data_ptr(), L, _stream()))
    return out


# adjoints of the two transforms: each is the OTHER transform's kernel between two scalings
def _scale_packed(f, c_all, c_edge):
    N = f.shape[-1]
    out = torch.empty_like(f)
    with torch.cuda.device(f.device):
        _n.check(_n.lib().gss_scale_packed(f.data_ptr(), out.data_ptr(), f.numel() // N, N, c_all, c_edge, _stream()))
    return out


def _ola_norm_scale(x, length, T, N, H, inverse, scale):
    """x [R, ld] (first ``length`` samples of a row used) -> [R, length] times or over the overlap-add weight"""
    R, ld = x.shape
    out = torch.empty((R, length), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _n.check(_n.lib().gss_ola_norm_scale(x.data_ptr(), out.data_ptr(), R, length, ld, length, T, N, H,
                                             1 if inverse else 0, scale, _stream()))
    return out


def _istft_adjoint(g, T, N, H):
    """g = dL/d(wave) ``[R, (T-1)H]`` -> dL/d(feature) ``[R, T, N]``: (N/2) STFT(g / norm), DC and Nyquist halved"""
    g = _f32c(g, "istft.backward")
    L = (T - 1) * H
    assert g.dim() == 2 and g.shape[1] == L
    u = _ola_norm_scale(g, L, T, N, H, True, 1.0)
    G = _stft_fwd(u, N, H)
    assert G.shape[-2] == T
    return _scale_packed(G, 0.5 * N, 0.5)


def _stft_adjoint(g, n, N, H):
    """g = dL/d(feature) ``[B, T, N]`` -> dL/d(wave) ``[B, n]``: (4/N) norm * iSTFT(g with interior bins halved)"""
    g = _f32c(g, "stft.backward")
    T = g.shape[-2]
    y = _istft_fwd(_scale_packed(g, 0.5, 2.0), H)               # [B, (T-1)H], already divided by norm
    return _ola_norm_scale(y, n, T, N, H, False, 4.0 / N)


def _logexp_bwd_n(x, gout, fn_name, N):
    g = _f32c(gout, fn_name)
    gin = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _n.check(getattr(_n.lib(), fn_name)(x.data_ptr(), g.data_ptr(), gin.data_ptr(), x.numel() // N, N, hparams.EPS, _stream()))
    return gin


class _Stft(torch.autograd.Function):
    @staticmethod
    def forward(ctx, wave, N, H, log):
        w = _f32c(wave, "stft")
        ctx.save_for_backward(w)
        ctx.cfg = (N, H, log)
        return _stft_fwd(w, N, H, log)

    @staticmethod
    def backward(ctx, gout):
        (w,) = ctx.saved_tensors
        N, H, log = ctx.cfg
        n = w.shape[-1]
        g = _f32c(gout, "stft.backward").reshape(-1, gout.shape[-2], N)
        if log:     # chain through to_log_signal: needs the uncompressed features, recomputed (one STFT launch)
            raw = _stft_fwd(w.reshape(-1, n), N, H, False)
            g = _logexp_bwd_n(raw, g, "gss_to_log_bwd", N)
        return _stft_adjoint(g, n, N, H).reshape(w.shape), None, None, None


class _Istft(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feature, H, exp):
        f = _f32c(feature, "istft")
        ctx.save_for_backward(f if exp else f.new_empty(0))
        ctx.cfg = (H, exp, f.shape)
        return _istft_fwd(f, H, exp)

    @staticmethod
    def backward(ctx, gout):
        (f,) = ctx.saved_tensors
        H, exp, shape = ctx.cfg
        T, N = shape[-2], shape[-1]
        g = _istft_adjoint(gout.reshape(-1, gout.shape[-1]), T, N, H)
        if exp:
            g = _logexp_bwd_n(f.reshape(-1, T, N), g, "gss_to_exp_bwd", N)
        return g.reshape(shape), None, None


class _ApplyMask(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mix_feature, mask):
        x, m = _f32c(mix_feature, "apply_mask"), _f32c(mask, "apply_mask")
        ctx.save_for_backward(x, m)
        return _apply_mask_fwd(x, m)

    @staticmethod
    def backward(ctx, gout):
        x, m = ctx.saved_tensors
        g = _f32c(gout, "apply_mask")
        B, T, N = x.shape
        S = m.shape[1]
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gm = torch.empty_like(m) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(x.device):
            _n.check(_n.lib().gss_apply_mask_bwd(x.data_ptr(), m.data_ptr(), g.data_ptr(), B, S, T, N,
                                                 gx.data_ptr() if gx is not None else None,
                                                 gm.data_ptr() if gm is not None else None, _stream()))
        return gx, gm


def apply_mask(mix_feature, mask):
    """``mix_feature [B,T,N]``, ``mask [B,S,T,N/2]`` -> ``[B*S,T,N]``: one real gain per
    complex bin (both packed halves; slot 0 = DC+Nyquist share a gain, the pairing
    of ops.py:234-237), output row ``b*S+s`` (modules.py:396-399).  Differentiable in both
    arguments (``gss_apply_mask_bwd``)."""
    if torch.is_grad_enabled() and (getattr(mix_feature, "requires_grad", False) or getattr(mask, "requires_grad", False)):
        return _ApplyMask.apply(mix_feature, mask)
    return _apply_mask_fwd(mix_feature, mask)


def _apply_mask_fwd(mix_feature, mask):
    x = _f32c(mix_feature, "apply_mask")
    m = _f32c(mask, "apply_mask")
    assert x.dim() == 3 and m.dim() == 4, "apply_mask: mix [B,T,N], mask [B,S,T,N/2]"
    B, T, N = x.shape
    S = m.shape[1]
    assert m.shape == (B, S, T, N // 2), f"apply_mask: mask shape {tuple(m.shape)} != {(B, S, T, N // 2)}"
    out = torch.empty((B * S, T, N), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _n.check(_n.lib().gss_apply_mask(x.data_ptr(), m.data_ptr(), B, S, T, N, out.data_ptr(), _stream()))
    return out


def mask_istft(wave, mask, fft_size=None, hop=None, out=None):
    """Fused synthesis: STFT of the mixture (recomputed from ``wave [B,n]``), per-source
    mask ``[B,S,T,N/2]``, inverse STFT with overlap-add -> ``[B*S, (T-1)*H]``.  Differentiable in
    ``mask`` and ``wave`` (iSTFT adjoint -> ``gss_apply_mask_bwd`` -> STFT adjoint): a mask separator can be
    trained against a waveform-domain loss through the native kernels."""
    _dev(wave, "mask_istft"); _dev(mask, "mask_istft")
    assert wave.dim() == 2 and mask.dim() == 4, "mask_istft: wave [B,n], mask [B,S,T,N/2]"
    N, H = _nh(fft_size if fft_size is not None else 2 * mask.shape[-1], hop)
    if torch.is_grad_enabled() and (wave.requires_grad or mask.requires_grad):
        assert out is None, "mask_istft: out= is not supported when gradients are required"
        return _MaskIstft.apply(wave, mask, N, H)
    return _mask_istft_fwd(wave, mask, N, H, out)


class _MaskIstft(torch.autograd.Function):
    @staticmethod
    def forward(ctx, wave, mask, N, H):
        w, m = _f32c(wave, "mask_istft"), _f32c(mask, "mask_istft")
        ctx.save_for_backward(w, m)
        ctx.cfg = (N, H)
        return _mask_istft_fwd(w, m, N, H, None)

    @staticmethod
    def backward(ctx, gout):
        w, m = ctx.saved_tensors
        N, H = ctx.cfg
        B, S, T = m.shape[0], m.shape[1], m.shape[2]
        gy = _istft_adjoint(gout, T, N, H)                      # [B*S, T, N]
        x = _stft_fwd(w, N, H)                                  # mixture spectrum, recomputed as in the forward
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gm = torch.empty_like(m) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(x.device):
            _n.check(_n.lib().gss_apply_mask_bwd(x.data_ptr(), m.data_ptr(), gy.data_ptr(), B, S, T, N,
                                                 gx.data_ptr() if gx is not None else None,
                                                 gm.data_ptr() if gm is not None else None, _stream()))
        gw = _stft_adjoint(gx, w.shape[-1], N, H) if gx is not None else None
        return gw, gm, None, None


def _mask_istft_fwd(wave, mask, N, H, out=None):
    w = _f32c(wave, "mask_istft")
    m = _f32c(mask, "mask_istft")
    B, n = w.shape
    S = m.shape[1]
    T, _ = _n.frame_count(n, N, H)
    assert m.shape == (B, S, T, N // 2), f"mask_istft: mask shape {tuple(m.shape)} != {(B, S, T, N // 2)}"
    L = (T - 1) * H
    if out is None:
        out = torch.empty((B * S, L), dtype=torch.float32, device=w.device)
    else:
        assert out.is_cuda and out.dtype == torch.float32 and out.shape == (B * S, L) and out.is_contiguous()
    with torch.cuda.device(w.device):
        _n.check(_n.lib().gss_mask_istft(w.data_ptr(), m.data_ptr(), B, S, n, n, N, H, out.data_ptr(), L, _stream()))
    return out


def stft_dual(wave, fft_size=None, hop=None, out_lin=None, out_log=None):
    """One transform, both outputs: ``(feature, to_log_signal(feature))`` = the reference's
    ``s_mixed_signals`` and ``s_mixed_signals_log`` (main.py:328-338) from ``wave [..., n]``.
    The linear features feed ``mask_istft_feature``; the log features feed the separator."""
    w = _f32c(wave, "stft_dual")
    N, H = _nh(fft_size, hop)
    lead, n = w.shape[:-1], w.shape[-1]
    T, _ = _n.frame_count(n, N, H)
    B = 1
    for d in lead:
        B *= d
    shape = lead + (T, N)
    lin = torch.empty(shape, dtype=torch.float32, device=w.device) if out_lin is None else out_lin
    lg = torch.empty(shape, dtype=torch.float32, device=w.device) if out_log is None else out_log
    for o in (lin, lg):
        assert o.is_cuda and o.dtype == torch.float32 and tuple(o.shape) == tuple(shape) and o.is_contiguous()
    with torch.cuda.device(w.device):
        _n.check(_n.lib().gss_stft_packed_dual(w.data_ptr(), B, n, n, N, H, hparams.EPS, lin.data_ptr(), lg.data_ptr(), _stream()))
    return lin, lg


def mask_istft_feature(mix_feature, mask, hop=None, out=None, reverse=False, ae_rows=None):
    """Fused ``apply_mask`` + ``istft`` from the mixture's LINEAR packed features ``[B,T,N]`` (what the
    reference's graph holds as ``s_mixed_signals``, main.py:328-337) and ``mask [B,S,T,N/2]`` ->
    ``[B*S, (T-1)*H]``, row ``b*S+s``.  Same results as ``istft(apply_mask(f, m))`` without materialising the
    ``[B*S,T,N]`` product, and as ``mask_istft(wave, m)`` when ``f = stft(wave)``.  ``reverse=True`` walks the rows
    last-to-first (a cache hint when ``mix_feature`` was written just before; results are identical).
    ``ae_rows`` (float32 ``[B]`` on the device): also receives the per-mixture auto-encoder partial
    ``sum((sum_s separated_s - mixed)^2)`` of main.py:353-361, computed inside the same kernel (FFT_SIZE 256 / 512,
    S <= 4 (S <= 3 at hop N/8); ``ValueError`` otherwise - use ``ae_loss(apply_mask(...))`` there).
    Differentiable in both arguments."""
    _dev(mix_feature, "mask_istft_feature"); _dev(mask, "mask_istft_feature")
    assert mix_feature.dim() == 3 and mask.dim() == 4, "mask_istft_feature: feature [B,T,N], mask [B,S,T,N/2]"
    N, H = _nh(mix_feature.shape[-1], hop)
    if torch.is_grad_enabled() and (mix_feature.requires_grad or mask.requires_grad):
        assert out is None and ae_rows is None, "mask_istft_feature: out= / ae_rows= are not supported when gradients are required"
        return _Istft.apply(apply_mask(mix_feature, mask), H, False)
    f = _f32c(mix_feature, "mask_istft_feature")
    m = _f32c(mask, "mask_istft_feature")
    B, T, _ = f.shape
    S = m.shape[1]
    assert m.shape == (B, S, T, N // 2), f"mask_istft_feature: mask shape {tuple(m.shape)} != {(B, S, T, N // 2)}"
    L = (T - 1) * H
    if out is None:
        out = torch.empty((B * S, L), dtype=torch.float32, device=f.device)
    else:
        assert out.is_cuda and out.dtype == torch.float32 and out.shape == (B * S, L) and out.is_contiguous()
    if ae_rows is not None:
        assert ae_rows.is_cuda and ae_rows.dtype == torch.float32 and ae_rows.numel() == B and ae_rows.is_contiguous()
    with torch.cuda.device(f.device):
        _n.check(_n.lib().gss_mask_istft_feature_ae(f.data_ptr(), m.data_ptr(), B, S, T, N, H,
                                                    _n.FLAG_REVERSE if reverse else 0, out.data_ptr(), L,
                                                    ae_rows.data_ptr() if ae_rows is not None else None, _stream()))
    return out


def metric_vector(ae_rows=None, snr=None, elems_per_row=1, out=None):
    """The per-batch metric vector ``[sum_b mean_i max_k snr[b,i,k], sum_b ae_rows[b] / elems_per_row, 0, B]`` that
    ``parallel.allreduce_metrics`` sums over the ranks (batch means of main.py:353-361, :446-457), built on the
    device by one small kernel so that the collective can follow on the same stream."""
    assert ae_rows is not None or snr is not None, "metric_vector: nothing to reduce"
    ref = ae_rows if ae_rows is not None else snr
    _dev(ref, "metric_vector")
    B = int(ae_rows.numel()) if ae_rows is not None else int(snr.shape[0])
    m, n = (int(snr.shape[1]), int(snr.shape[2])) if snr is not None else (1, 1)
    if snr is not None:
        snr = _f32c(snr, "metric_vector")
        assert snr.dim() == 3 and snr.shape[0] == B, "metric_vector: snr must be [B, m, n]"
    vec = torch.empty(4, dtype=torch.float32, device=ref.device) if out is None else out
    with torch.cuda.device(ref.device):
        _n.check(_n.lib().gss_metric_finalise(ae_rows.data_ptr() if ae_rows is not None else None,
                                              snr.data_ptr() if snr is not None else None, B, m, n,
                                              float(elems_per_row), vec.data_ptr(), _stream()))
    return vec


# ---------------------------------------------------------------------------
# element-wise compression (reference names)
# ---------------------------------------------------------------------------
def _logexp(s_signal, fn_name):
    x = _f32c(s_signal, fn_name)
    N = hparams.FFT_SIZE
    assert x.shape[-1] == N, f"{fn_name}: last axis {x.shape[-1]} != FFT_SIZE {N}"
    out = torch.empty_like(x)
    rows = x.numel() // N
    with torch.cuda.device(x.device):
        _n.check(getattr(_n.lib(), fn_name)(x.data_ptr(), out.data_ptr(), rows, N, hparams.EPS, _stream()))
    return out


def _logexp_bwd(x, gout, fn_name):
    N = hparams.FFT_SIZE
    g = _f32c(gout, fn_name)
    gin = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _n.check(getattr(_n.lib(), fn_name)(x.data_ptr(), g.data_ptr(), gin.data_ptr(), x.numel() // N, N, hparams.EPS, _stream()))
    return gin


class _ToLog(torch.autograd.Function):
    """to_log_signal with the hand-written backward kernel (the reference back-propagates through
    it: main.py:338 sits under the optimisers of main.py:481-484)."""
    @staticmethod
    def forward(ctx, x):
        x = _f32c(x, "to_log_signal")
        ctx.save_for_backward(x)
        return _logexp(x, "gss_to_log")

    @staticmethod
    def backward(ctx, gout):
        (x,) = ctx.saved_tensors
        return _logexp_bwd(x, gout, "gss_to_log_bwd")


class _ToExp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x = _f32c(x, "to_exp_signal")
        ctx.save_for_backward(x)
        return _logexp(x, "gss_to_exp")

    @staticmethod
    def backward(ctx, gout):
        (x,) = ctx.saved_tensors
        return _logexp_bwd(x, gout, "gss_to_exp_bwd")


def to_log_signal(s_signal):
    """ops.py:228-238: each bin pair ``(k, k+N/2)`` scaled by ``0.5*log1p(a2)*rsqrt(a2+EPS)``.
    Differentiable (``gss_to_log_bwd``)."""
    if torch.is_grad_enabled() and isinstance(s_signal, torch.Tensor) and s_signal.requires_grad:
        return _ToLog.apply(s_signal)
    return _logexp(s_signal, "gss_to_log")


def to_exp_signal(s_signal):
    """ops.py:241-251: ``a = sqrt(re^2+im^2+EPS)``, scale ``expm1(a)/a`` (not the inverse of to_log, K7).
    Differentiable (``gss_to_exp_bwd``)."""
    if torch.is_grad_enabled() and isinstance(s_signal, torch.Tensor) and s_signal.requires_grad:
        return _ToExp.apply(s_signal)
    return _logexp(s_signal, "gss_to_exp")


# ---------------------------------------------------------------------------
# metrics (reference names)
# ---------------------------------------------------------------------------
def batch_cross_snr(clear_signal, noisy_signal):
    """ops.py:191-225: ``[B,m,...]`` x ``[B,n,...]`` -> ``[B,m,n]``."""
    c = _f32c(clear_signal, "batch_cross_snr")
    z = _f32c(noisy_signal, "batch_cross_snr")
    assert c.dim() == z.dim()
    assert c.dim() >= 2
    assert c.shape[0] == z.shape[0] and c.shape[2:] == z.shape[2:], "batch_cross_snr: trailing shapes differ"
    B, m, n = c.shape[0], c.shape[1], z.shape[1]
    L = 1
    for d in c.shape[2:]:
        L *= d
    out = torch.empty((B, m, n), dtype=torch.float32, device=c.device)
    with torch.cuda.device(c.device):
        _n.check(_n.lib().gss_cross_snr(c.data_ptr(), z.data_ptr(), B, m, n, L, hparams.EPS, out.data_ptr(), _stream()))
    return out


def batch_snr(clear_signal, noisy_signal):
    """ops.py:162-189: ``[B,...]`` x ``[B,...]`` -> ``[B]`` (the m = n = 1 case of the cross SNR)."""
    c = _dev(clear_signal, "batch_snr")
    z = _dev(noisy_signal, "batch_snr")
    assert c.dim() == z.dim()
    assert c.shape == z.shape, "batch_snr: shapes differ"
    B = c.shape[0]
    return batch_cross_snr(c.reshape(B, 1, -1), z.reshape(B, 1, -1)).reshape(B)


def ae_loss(separated, mixed, n_out):
    """main.py:353-361: ``mean((sum_s sep[b,s] - mix[b])^2)`` with ``separated [B*S,T,N]``."""
    s = _f32c(separated, "ae_loss")
    x = _f32c(mixed, "ae_loss")
    B = x.shape[0]
    assert s.shape[0] == B * n_out and s.shape[1:] == x.shape[1:]
    L = x.numel() // B
    part = torch.empty((B,), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _n.check(_n.lib().gss_ae_partial(s.data_ptr(), x.data_ptr(), B, n_out, L, part.data_ptr(), _stream()))
    return part.sum() / (B * L)


def wav16_normalise(wave):
    """main.py:112-116: per clip shift to min, scale to 32767, truncate -> int16 ``[R, len]``."""
    x = _f32c(wave, "wav16_normalise")
    x2 = x.reshape(-1, x.shape[-1])
    R, n = x2.shape
    mm = torch.empty((R, 2), dtype=torch.float32, device=x.device)
    pcm = torch.empty((R, n), dtype=torch.int16, device=x.device)
    with torch.cuda.device(x.device):
        _n.check(_n.lib().gss_wav16_normalise(x2.data_ptr(), R, n, n, mm.data_ptr(), pcm.data_ptr(), _stream()))
    return pcm.reshape(x.shape)


# ---------------------------------------------------------------------------
# feature-domain mixing (main.py:328-338)
# ---------------------------------------------------------------------------
def mix_signals(s_src_signals, n_signal=None, noise=None, noise_stddev=0.1, generator=None, log=False):
    """main.py:328-337: ``src [B*n_sig, T, N]`` packed features -> mixture ``[B, T, N]`` =
    ``sum_i src[b*n_sig+i] + noise``.  ``noise``: a tensor ``[B,T,N]``, ``None`` to draw
    ``N(0, noise_stddev^2)`` on the device (the reference's ``tf.random_normal(stddev=0.1)``;
    ``generator`` makes the draw reproducible), or ``False`` for no noise.  ``log=True`` returns
    ``(mix, to_log_signal(mix))`` from one fused pass (main.py:338)."""
    x = _f32c(s_src_signals, "mix_signals")
    n_sig = hparams.MAX_N_SIGNAL if n_signal is None else int(n_signal)
    assert x.dim() == 3 and x.shape[0] % n_sig == 0, "mix_signals: src must be [B*n_sig, T, N]"
    B, T, N = x.shape[0] // n_sig, x.shape[1], x.shape[2]
    if noise is None:
        noise = torch.randn((B, T, N), dtype=torch.float32, device=x.device, generator=generator) * noise_stddev
    elif noise is False:
        noise = None
    else:
        noise = _f32c(noise, "mix_signals")
        assert noise.shape == (B, T, N), f"mix_signals: noise shape {tuple(noise.shape)} != {(B, T, N)}"
    mix = torch.empty((B, T, N), dtype=torch.float32, device=x.device)
    mix_log = torch.empty_like(mix) if log else None
    with torch.cuda.device(x.device):
        _n.check(_n.lib().gss_mix_features(x.data_ptr(), noise.data_ptr() if noise is not None else None, B, n_sig, T, N,
                                           _n.FLAG_LOG if log else 0, hparams.EPS, mix.data_ptr(),
                                           mix_log.data_ptr() if log else None, _stream()))
    return (mix, mix_log) if log else mix


def snr_metric(s_src_signals, s_separated_signals, n_signal=None):
    """main.py:446-457: cross SNR of every (source, output) pair, max over outputs, mean over
    batch and sources.  ``src [B*n_sig,T,N]``, ``sep [B*S,T,N]`` -> scalar tensor."""
    n_sig = hparams.MAX_N_SIGNAL if n_signal is None else int(n_signal)
    B = s_src_signals.shape[0] // n_sig
    T, N = s_src_signals.shape[-2:]
    cs = batch_cross_snr(s_src_signals.reshape(B, n_sig, T, N), s_separated_signals.reshape(B, -1, T, N))
    return cs.max(dim=-1).values.mean()


# ---------------------------------------------------------------------------
# demo-mode edges (main.py:83-95)
# ---------------------------------------------------------------------------
def resample_pad_size(length, fft_size=None):
    """main.py:93: zero-padding after the resample branch, ``N - ((L-1) mod N) - 1``."""
    N = hparams.FFT_SIZE if fft_size is None else int(fft_size)
    return N - ((int(length) - 1) % N) - 1


def resample(wave, num):
    """``scipy.signal.resample(x, num)`` (Fourier method, main.py:92) along the last axis of a real CUDA tensor:
    hand-written Bluestein chirp-z transforms over a power-of-two Stockham FFT in float64 (``gss_resample_f64``;
    SciPy itself promotes the int16 samples to float64).  float32 input is computed in float64 and cast back."""
    x = _dev(wave, "resample")
    out_dtype = torch.float64 if wave.dtype == torch.float64 else torch.float32
    x = x.to(torch.float64).contiguous()
    n, num = x.shape[-1], int(num)
    assert n >= 1 and num >= 1
    rows = x.numel() // n
    y = torch.empty(x.shape[:-1] + (num,), dtype=torch.float64, device=x.device)
    nbytes = int(_n.lib().gss_resample_workspace_bytes(n, num))
    if nbytes == 0:
        raise ValueError(f"resample: unsupported lengths n={n}, num={num}")
    ws = torch.empty(nbytes // 8, dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        _n.check(_n.lib().gss_resample_f64(x.data_ptr(), rows, n, n, num, y.data_ptr(), num, ws.data_ptr(), nbytes, _stream()))
    return y.to(out_dtype)
