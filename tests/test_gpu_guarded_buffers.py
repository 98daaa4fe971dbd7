"""Poor man's memcheck (compute-sanitizer is closed on this pool, profiles/r2_compute_sanitizer_closed.txt): every kernel of the
path runs on tensors carved out of a larger allocation whose guard zones (before, between and after the operands) hold NaN.
An out-of-bounds WRITE breaks a guard pattern; an out-of-bounds READ pulls NaN into the result, which must stay bit-identical to
the run on ordinary tensors.  Ragged lengths, odd row counts, several chunks per row, every FFT size and hop family."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GUARD = 4096          # floats


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


class Arena:
    """tensors separated by NaN guard zones inside one allocation; every tensor start is 16-byte aligned"""
    def __init__(self, T, shapes):
        self.T = T
        sizes = [int(np.prod(s)) for s in shapes]
        total = GUARD + sum(((n + 3) // 4) * 4 + GUARD for n in sizes)
        self.buf = T.full((total,), float("nan"), device="cuda")
        self.views, self.spans, off = [], [], GUARD
        for s, n in zip(shapes, sizes):
            self.views.append(self.buf[off:off + n].view(*s))
            self.spans.append((off, off + n))
            off += ((n + 3) // 4) * 4 + GUARD

    def guards_intact(self):
        mask = self.T.ones_like(self.buf, dtype=self.T.bool)
        for lo, hi in self.spans:
            mask[lo:hi] = False
        return bool(self.T.isnan(self.buf[mask]).all())


CASES = [(512, 128, 4797, 3, 3), (512, 256, 6001, 2, 2), (512, 64, 3000, 1, 4), (512, 128, 48000, 5, 3), (256, 128, 3001, 3, 4),
         (256, 64, 16000, 2, 3), (256, 32, 1500, 1, 1), (1024, 256, 9001, 2, 3), (2048, 512, 20000, 1, 2), (4096, 1024, 20000, 2, 3),
         (128, 32, 1000, 2, 2)]


@pytest.mark.parametrize("N,H,n,B,S", CASES)
def test_kernels_stay_inside_their_buffers(T, ops, N, H, n, B, S):
    from gan_sass_tf_b200 import _native
    g = T.Generator(device="cuda").manual_seed(N + n)
    Tn, _ = _native.frame_count(n, N, H)
    L = (Tn - 1) * H
    x0 = T.randn(B, n, device="cuda", generator=g) * 0.1
    m0 = T.rand(B, S, Tn, N // 2, device="cuda", generator=g)
    # reference run on ordinary tensors
    f_ref = ops.stft(x0, N, H)
    fl_ref = ops.stft_log(x0, N, H)
    y_ref = ops.istft(f_ref, H)
    ye_ref = ops.istft(fl_ref, H, exp=True)
    w_ref = ops.mask_istft(x0, m0, N, H)
    feature = N >= 256
    if feature:
        lin_ref, lg_ref = ops.stft_dual(x0, N, H)
        wf_ref = ops.mask_istft_feature(lin_ref, m0, H)
    # guarded run: inputs and outputs inside the arena
    a = Arena(T, [(B, n), (B, S, Tn, N // 2), (B, Tn, N), (B, Tn, N), (B * S, L), (B * S, L), (B,)])
    x, m, lin, lg, w, wf, rows = a.views
    x.copy_(x0); m.copy_(m0)
    f = ops.stft(x, N, H)
    fl = ops.stft_log(x, N, H)
    assert T.equal(f, f_ref) and T.equal(fl, fl_ref)
    assert T.equal(ops.istft(f, H), y_ref) and T.equal(ops.istft(fl, H, exp=True), ye_ref)
    ops.mask_istft(x, m, N, H, out=w)
    assert T.equal(w, w_ref)
    if feature:
        ops.stft_dual(x, N, H, out_lin=lin, out_log=lg)
        assert T.equal(lin, lin_ref) and T.equal(lg, lg_ref)
        ops.mask_istft_feature(lin, m, H, out=wf)
        assert T.equal(wf, wf_ref)
        if N <= 512 and (S <= 3 or (S == 4 and 8 * H != N)):
            rows.fill_(float("nan"))
            ops.mask_istft_feature(lin, m, H, out=wf, ae_rows=rows)
            assert T.equal(wf, wf_ref) and bool(T.isfinite(rows).all())
    T.cuda.synchronize()
    assert a.guards_intact(), "a kernel wrote outside its buffers"
    assert not bool(T.isnan(w).any())


def test_elementwise_and_reduction_kernels_stay_inside(T, ops):
    from gan_sass_tf_b200.app import hparams
    old = hparams.FFT_SIZE
    hparams.FFT_SIZE = 256
    try:
        g = T.Generator(device="cuda").manual_seed(1)
        B, n_sig, S, Tn, N = 3, 2, 3, 37, 256
        a = Arena(T, [(B * n_sig, Tn, N), (B * S, Tn, N), (B, Tn, N), (B, S, Tn, N // 2), (5, 3001)])
        src, sep, mix, mask, wav = a.views
        src.copy_(T.rand(src.shape, device="cuda", generator=g)); sep.copy_(T.rand(sep.shape, device="cuda", generator=g))
        mask.copy_(T.rand(mask.shape, device="cuda", generator=g)); wav.copy_(T.randn(wav.shape, device="cuda", generator=g))
        mx, mxl = ops.mix_signals(src, n_sig, noise=False, log=True)
        mix.copy_(mx)
        outs = [ops.to_log_signal(mix), ops.to_exp_signal(mxl), ops.apply_mask(mix, mask),
                ops.batch_cross_snr(src.view(B, n_sig, Tn, N), sep.view(B, S, Tn, N)), ops.ae_loss(sep, mix, S).reshape(1),
                ops.wav16_normalise(wav).float(), ops.resample(wav.double(), 1777).float()]
        T.cuda.synchronize()
        assert a.guards_intact()
        assert all(bool(T.isfinite(o).all()) for o in outs)
    finally:
        hparams.FFT_SIZE = old


@pytest.mark.parametrize("N,H", [(512, 128), (256, 64), (1024, 256)])
def test_results_do_not_depend_on_timing(T, ops, N, H):
    """Poor man's racecheck for the shared-memory stages (TMA refills behind team barriers, double-buffered mask stage,
    single feature stage): the same launch repeated under different contention - alone, next to a bandwidth hog on another
    stream, next to another instance of itself - must give identical bits every time."""
    from gan_sass_tf_b200 import _native
    n, B, S = 24000, 48, 3
    g = T.Generator(device="cuda").manual_seed(N)
    Tn, _ = _native.frame_count(n, N, H)
    x = T.randn(B, n, device="cuda", generator=g) * 0.1
    m = T.rand(B, S, Tn, N // 2, device="cuda", generator=g)
    lin = ops.stft(x, N, H)
    ref_f = ops.mask_istft_feature(lin, m, H)
    ref_w = ops.mask_istft(x, m, N, H)
    ref_s = ops.stft_log(x, N, H)
    hog_a = T.empty(64 << 20, device="cuda"); hog_b = T.empty(64 << 20, device="cuda")
    side = T.cuda.Stream()
    for rep in range(24):
        mode = rep % 3
        if mode == 1:
            with T.cuda.stream(side):
                for _ in range(4):
                    hog_b.copy_(hog_a)
        elif mode == 2:
            with T.cuda.stream(side):
                other = ops.mask_istft_feature(lin, m, H)
        out_f = ops.mask_istft_feature(lin, m, H, reverse=bool(rep & 1))
        out_w = ops.mask_istft(x, m, N, H)
        out_s = ops.stft_log(x, N, H)
        T.cuda.synchronize()
        assert T.equal(out_f, ref_f) and T.equal(out_w, ref_w) and T.equal(out_s, ref_s), f"repetition {rep} differs"
        if mode == 2:
            assert T.equal(other, ref_f)
