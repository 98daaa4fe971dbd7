"""GPU parity of the feature-fed synthesis (gss_mask_istft_feature) and the dual-output STFT
(gss_stft_packed_dual): against the CPU oracle, against the waveform-fed kernel, and against the
unfused composition apply_mask -> istft (reference data flow main.py:328-342).
"""
import numpy as np
import pytest

from oracle import ref_oracle as R

pytestmark = pytest.mark.gpu
REL_L2 = 1e-5


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


def dev(T, a):
    return T.from_numpy(np.ascontiguousarray(a)).cuda()


CASES = [
    (512, 128, 4797, 2, 3), (512, 128, 6000, 1, 4), (512, 128, 5000, 2, 1), (512, 128, 5000, 1, 2), (512, 128, 3000, 1, 5),
    (512, 256, 5000, 2, 3), (512, 64, 3000, 1, 3), (256, 128, 3000, 2, 4), (256, 64, 2500, 3, 3), (256, 32, 1500, 1, 2),
    (512, 128, 512, 1, 3), (512, 128, 640, 2, 3), (512, 128, 48000, 2, 3),
    (1024, 256, 9000, 1, 3), (1024, 512, 9000, 2, 2), (1024, 128, 5000, 1, 1), (2048, 512, 9000, 1, 3), (4096, 1024, 20000, 1, 3),
    (4096, 2048, 20000, 2, 4),
]


@pytest.mark.parametrize("N,H,n,B,S", CASES)
def test_mask_istft_feature_matches_oracle(T, ops, N, H, n, B, S):
    rng = np.random.default_rng(N + n + S)
    x = (rng.standard_normal((B, n)) * 0.1).astype(np.float32)
    Tn, _ = R.frame_count(n, N, H)
    mask = rng.random((B, S, Tn, N // 2)).astype(np.float32)
    xd, md = dev(T, x), dev(T, mask)
    feat = ops.stft(xd, N, H)
    y = ops.mask_istft_feature(feat, md, H)
    assert y.shape == (B * S, (Tn - 1) * H)
    ref = R.mask_istft_np(x, mask, N, H).reshape(B * S, -1)
    assert R.rel_l2(y.cpu().numpy(), ref) < REL_L2
    # the reversed walk is the same computation item by item: identical bits
    assert T.equal(ops.mask_istft_feature(feat, md, H, reverse=True), y)
    # the waveform-fed kernel and the unfused composition agree to float32 rounding
    y1 = ops.mask_istft(xd, md, N, H)
    y2 = ops.istft(ops.apply_mask(feat, md), H)
    assert R.rel_l2(y.cpu().numpy(), y1.cpu().numpy()) < 3e-6
    assert R.rel_l2(y.cpu().numpy(), y2.cpu().numpy()) < 3e-6


@pytest.mark.parametrize("N,H,n,B", [(512, 128, 4797, 3), (512, 256, 6000, 2), (512, 64, 3000, 1), (256, 128, 3000, 3),
                                     (256, 64, 16000, 2), (1024, 256, 9000, 2), (2048, 512, 9000, 1)])
@pytest.mark.parametrize("scale", [0.1, 3.0])
def test_stft_dual_equals_the_two_single_transforms(T, ops, N, H, n, B, scale):
    rng = np.random.default_rng(N + n)
    x = dev(T, (rng.standard_normal((B, n)) * scale).astype(np.float32))
    lin, lg = ops.stft_dual(x, N, H)
    assert T.equal(lin, ops.stft(x, N, H))
    if N in (256, 512):
        assert T.equal(lg, ops.stft_log(x, N, H))          # same kernel body, same bits
    ref = R.to_log_signal(R.stft_feature_np(x.cpu().numpy(), N, H, np.float64, np.float64))
    assert R.rel_l2(lg.cpu().numpy(), ref) < REL_L2


def test_feature_path_full_size_c2(T, ops):
    """C2 (256 x 3 s, N = 512, H = 128, S = 3) through stft_dual -> mask_istft_feature: masks that sum to one give the
    mixture back (>= 100 dB), a sub-batch agrees with the oracle."""
    N, H, n, B, S = 512, 128, 48000, 256, 3
    g = T.Generator(device="cuda").manual_seed(1234)
    x = T.randn(B, n, device="cuda", generator=g) * 0.1
    lin, lg = ops.stft_dual(x, N, H)
    m = T.rand(B, S, 376, N // 2, device="cuda", generator=g) + 0.05
    m = m / m.sum(dim=1, keepdim=True)
    y = ops.mask_istft_feature(lin, m, H, reverse=True)
    rec = y.reshape(B, S, -1).sum(dim=1)
    err = (rec - x).double().pow(2).sum() / x.double().pow(2).sum()
    assert 10 * np.log10(1.0 / float(err)) >= 100.0
    sub = slice(101, 103)
    ref = R.mask_istft_np(x[sub].cpu().numpy(), m[sub].cpu().numpy(), N, H).reshape(2 * S, -1)
    assert R.rel_l2(y.reshape(B, S, -1)[sub].reshape(2 * S, -1).cpu().numpy(), ref) < REL_L2


def test_mask_istft_feature_gradients(T, ops):
    """differentiable through apply_mask + iSTFT adjoint: compare with autograd of the unfused ops."""
    N, H, n, B, S = 512, 128, 3000, 2, 3
    g = T.Generator(device="cuda").manual_seed(5)
    x = T.randn(B, n, device="cuda", generator=g) * 0.1
    f = ops.stft(x, N, H).requires_grad_(True)
    m = T.rand(B, S, f.shape[1], N // 2, device="cuda", generator=g).requires_grad_(True)
    w = T.randn(B * S, (f.shape[1] - 1) * H, device="cuda", generator=g)
    (ops.mask_istft_feature(f, m, H) * w).sum().backward()
    gf, gm = f.grad.clone(), m.grad.clone()
    f.grad = None; m.grad = None
    (ops.istft(ops.apply_mask(f, m), H) * w).sum().backward()
    assert T.allclose(gf, f.grad, rtol=1e-5, atol=1e-7) and T.allclose(gm, m.grad, rtol=1e-5, atol=1e-7)


def test_feature_entry_rejects_bad_args(T, ops):
    f = T.zeros(1, 9, 128, device="cuda")
    m = T.zeros(1, 1, 9, 64, device="cuda")
    with pytest.raises(ValueError):
        ops.mask_istft_feature(f, m, 32)                   # FFT_SIZE 128: no feature-fed kernel
    f = T.zeros(1, 9, 512, device="cuda")
    with pytest.raises(AssertionError):
        ops.mask_istft_feature(f, T.zeros(1, 1, 8, 256, device="cuda"), 128)


@pytest.mark.parametrize("N,H,n,B,S", [(512, 128, 9000, 3, 3), (512, 128, 48000, 2, 3), (512, 256, 5000, 2, 2), (512, 64, 4000, 1, 1),
                                       (512, 128, 6001, 2, 4), (256, 128, 5000, 4, 3), (256, 64, 3000, 2, 4),
                                       (256, 128, 16256, 8, 4), (512, 256, 9000, 3, 4)])      # the reference's defaults: S = 4 at hop N/2
def test_fused_autoencoder_partial(T, ops, N, H, n, B, S):
    """ae_rows from inside the synthesis kernel == main.py:353-361 on the masked features (oracle: NumPy float64)."""
    rng = np.random.default_rng(N + n + S)
    x = (rng.standard_normal((B, n)) * 0.1).astype(np.float32)
    Tn, _ = R.frame_count(n, N, H)
    mask = rng.random((B, S, Tn, N // 2)).astype(np.float32)
    feat = ops.stft(dev(T, x), N, H)
    rows = T.full((B,), 7.0, device="cuda")                      # the call zeroes it
    y = ops.mask_istft_feature(feat, dev(T, mask), H, ae_rows=rows)
    assert T.equal(y, ops.mask_istft_feature(feat, dev(T, mask), H))       # the extra output does not change the waveforms
    f64 = feat.cpu().numpy().astype(np.float64)
    sep = R.apply_mask(f64, mask.astype(np.float64)).reshape(B, S, Tn, N)
    ref_rows = ((sep.sum(axis=1) - f64) ** 2).sum(axis=(1, 2))
    np.testing.assert_allclose(rows.cpu().numpy(), ref_rows, rtol=2e-5)
    # the reference's scalar: mean over every element (R.ae_loss restates main.py:353-361)
    vec = ops.metric_vector(ae_rows=rows, elems_per_row=Tn * N).cpu().numpy()
    assert vec[3] == B and vec[2] == 0 and vec[0] == 0
    assert abs(vec[1] / B - R.autoencoder_loss(sep.reshape(B * S, Tn, N), f64, B, S)) <= 2e-5 * abs(vec[1] / B)


def test_fused_autoencoder_partial_unsupported_shapes(T, ops):
    f = T.zeros(1, 9, 512, device="cuda")
    rows = T.zeros(1, device="cuda")
    with pytest.raises(ValueError):
        ops.mask_istft_feature(f, T.zeros(1, 5, 9, 256, device="cuda"), 128, ae_rows=rows)     # S = 5 needs two passes
    f = T.zeros(1, 9, 1024, device="cuda")
    with pytest.raises(ValueError):
        ops.mask_istft_feature(f, T.zeros(1, 3, 9, 512, device="cuda"), 256, ae_rows=rows)     # team kernels: no fused partial


@pytest.mark.parametrize("B,m,n,L", [(8, 3, 4, 128 * 256), (2, 2, 3, 48000), (1, 1, 1, 1001), (3, 2, 11, 5000), (40, 3, 4, 376 * 512)])
def test_cluster_reductions_match_oracle(T, ops, B, m, n, L):
    """cross-SNR (one pass over `clear` for all outputs, cluster reduction), AE partial, min / max, metric vector."""
    rng = np.random.default_rng(B * 1000 + L)
    clear = rng.standard_normal((B, m, L)).astype(np.float32)
    noisy = (clear[:, :1].repeat(n, axis=1) + 0.3 * rng.standard_normal((B, n, L))).astype(np.float32)
    snr = ops.batch_cross_snr(dev(T, clear), dev(T, noisy))
    ref = R.batch_cross_snr(clear.astype(np.float64), noisy.astype(np.float64))
    np.testing.assert_allclose(snr.cpu().numpy(), ref, atol=2e-3)
    vec = ops.metric_vector(snr=snr).cpu().numpy()
    assert abs(vec[0] / B - ref.max(axis=2).mean()) < 2e-3 and vec[3] == B
    mix = rng.standard_normal((B, L)).astype(np.float32)
    ae = float(ops.ae_loss(dev(T, noisy.reshape(B * n, 1, L)), dev(T, mix.reshape(B, 1, L)), n))
    ref_ae = float(np.mean((noisy.astype(np.float64).sum(axis=1) - mix.astype(np.float64)) ** 2))
    assert abs(ae - ref_ae) <= 1e-4 * abs(ref_ae)
    pcm = ops.wav16_normalise(dev(T, clear.reshape(B * m, L))).cpu().numpy()
    refp = np.stack([R.wav16_normalise(r) for r in clear.reshape(B * m, L)])
    assert np.abs(pcm.astype(np.int32) - refp.astype(np.int32)).max() <= 1


@pytest.mark.parametrize("N,H", [(512, 128), (1024, 256)])
def test_bench_step_in_a_cuda_graph(T, ops, N, H):
    """the round-2 step (dual STFT, feature-fed synthesis with its memset node and fused partial, metric vector, and the
    cluster-launched cross-SNR) only enqueues on the caller's stream: captured once, replayed on new data."""
    n, B, S = 9000, 4, 3
    Tn, _ = R.frame_count(n, N, H)
    x = T.zeros(B, n, device="cuda")
    m = T.zeros(B, S, Tn, N // 2, device="cuda")
    lin = T.empty(B, Tn, N, device="cuda"); lg = T.empty(B, Tn, N, device="cuda")
    out = T.empty(B * S, (Tn - 1) * H, device="cuda")
    rows = T.zeros(B, device="cuda") if N == 512 else None
    vec = T.zeros(4, device="cuda")
    src = T.zeros(B, 2, Tn, N, device="cuda")

    def step():
        ops.stft_dual(x, N, H, out_lin=lin, out_log=lg)
        ops.mask_istft_feature(lin, m, H, out=out, ae_rows=rows)
        sep = ops.apply_mask(lin, m)
        snr = ops.batch_cross_snr(src, sep.reshape(B, S, Tn, N))
        ops.metric_vector(ae_rows=rows, snr=snr, elems_per_row=Tn * N, out=vec)
        return snr
    step()                                                             # first call: table fill, attribute opt-ins
    T.cuda.synchronize()
    g = T.cuda.CUDAGraph()
    with T.cuda.graph(g):
        snr = step()
    rng = np.random.default_rng(7)
    for seed in (1, 2):
        xs = (np.random.default_rng(seed).standard_normal((B, n)) * 0.1).astype(np.float32)
        ms = rng.random((B, S, Tn, N // 2)).astype(np.float32)
        ss = (rng.standard_normal((B, 2, Tn, N)) * 0.01).astype(np.float32)
        x.copy_(dev(T, xs)); m.copy_(dev(T, ms)); src.copy_(dev(T, ss))
        g.replay()
        T.cuda.synchronize()
        ref_lin = R.stft_feature_np(xs, N, H, np.float64, np.float64)
        assert R.rel_l2(lin.cpu().numpy(), ref_lin) < REL_L2
        assert R.rel_l2(out.cpu().numpy(), R.mask_istft_np(xs, ms, N, H).reshape(B * S, -1)) < REL_L2
        sep64 = R.apply_mask(ref_lin, ms.astype(np.float64)).reshape(B, S, Tn, N)
        ref_snr = R.batch_cross_snr(ss.astype(np.float64), sep64)
        np.testing.assert_allclose(snr.cpu().numpy(), ref_snr, atol=2e-3)
        v = vec.cpu().numpy()
        assert abs(v[0] / B - ref_snr.max(axis=2).mean()) < 2e-3 and v[3] == B
        if rows is not None:
            ref_ae = R.autoencoder_loss(sep64.reshape(B * S, Tn, N), ref_lin, B, S)
            assert abs(v[1] / B - ref_ae) <= 2e-5 * abs(ref_ae)
