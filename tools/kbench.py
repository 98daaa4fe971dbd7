"""Per-kernel timings at the C2 shape (B=256, n=48000, N=512, H=128): device time per launch,
rotating inputs.  Usage: python tools/kbench.py [N H B n]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gan_sass_tf_b200.app import ops
from gan_sass_tf_b200 import _native

N, H, B, n = (int(v) for v in sys.argv[1:5]) if len(sys.argv) >= 5 and sys.argv[1].isdigit() else (512, 128, 256, 48000)
if os.environ.get("GSS_PATH"):                              # 1: no register-exchange kernels, 2: per-frame fallback only
    _native.load_experimental()                             # the switches exist in lib/libgss_experimental.so only
    _native.set_path(int(os.environ["GSS_PATH"]))
T, _ = _native.frame_count(n, N, H)
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
waves = [(torch.randn(B, n, device=dev, generator=g) * 0.1).clamp_(-1, 1) for _ in range(3)]


def timeit(name, fn, bytes_, reps=20):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:28s} {ms * 1e3:8.1f} us   {bytes_ / ms / 1e6:7.0f} GB/s   ({B * T / 2 / ms / 1e3:.1f} M pairs/s)")


feat = [ops.stft(w, N, H) for w in waves]
timeit("stft", lambda i: ops.stft(waves[i % 3], N, H), 4 * B * (n + T * N))
timeit("stft_log", lambda i: ops.stft_log(waves[i % 3], N, H), 4 * B * (n + T * N))
timeit("istft", lambda i: ops.istft(feat[i % 3], H), 4 * B * (T * N + (T - 1) * H))
timeit("istft_exp", lambda i: ops.istft(feat[i % 3], H, exp=True), 4 * B * (T * N + (T - 1) * H))
for S in (1, 2, 3, 4):
    masks = [torch.rand(B, S, T, N // 2, device=dev, generator=g) for _ in range(3)]
    out = torch.empty(B * S, (T - 1) * H, device=dev)
    timeit(f"mask_istft S={S}", lambda i: ops.mask_istft(waves[i % 3], masks[i % 3], N, H, out=out),
           4 * B * (n + S * T * N // 2 + S * (T - 1) * H))
    del masks
if N in (256, 512, 1024, 2048, 4096):      # feature-fed synthesis and the dual-output STFT
    lin = [torch.empty(B, T, N, device=dev) for _ in range(3)]
    lg = torch.empty(B, T, N, device=dev)
    timeit("stft_dual", lambda i: ops.stft_dual(waves[i % 3], N, H, out_lin=lin[i % 3], out_log=lg), 4 * B * (n + T * N))
    for S in (1, 3):
        masks = [torch.rand(B, S, T, N // 2, device=dev, generator=g) for _ in range(3)]
        out = torch.empty(B * S, (T - 1) * H, device=dev)
        for rev in (False, True):
            timeit(f"mask_istft_feature S={S}{' rev' if rev else ''}", lambda i: ops.mask_istft_feature(lin[i % 3], masks[i % 3], H, out=out, reverse=rev),
                   4 * B * (n + S * T * N // 2 + S * (T - 1) * H))
        if S == 3:
            def step_old(i):
                ops.stft(waves[i % 3], N, H, log=True); ops.mask_istft(waves[i % 3], masks[i % 3], N, H, out=out)
            def step_new(i, rev=True):
                ops.stft_dual(waves[i % 3], N, H, out_lin=lin[i % 3], out_log=lg); ops.mask_istft_feature(lin[i % 3], masks[i % 3], H, out=out, reverse=rev)
            stepb = 4 * B * (n + T * N) + 4 * B * (n + S * T * N // 2 + S * (T - 1) * H)
            timeit("STEP stft_log + mask_istft", step_old, stepb)
            timeit("STEP dual + feature rev", step_new, stepb)
            timeit("STEP dual + feature fwd", lambda i: step_new(i, False), stepb)
        del masks
if "--torch" in sys.argv:   # courtesy baseline: cuFFT through torch.stft / torch.istft on the same GPU
    w_ = torch.hann_window(N, periodic=True, device=dev)
    spec = [torch.stft(w, N, H, window=w_, center=True, pad_mode="constant", return_complex=True) for w in waves]
    timeit("torch.stft (cuFFT)", lambda i: torch.stft(waves[i % 3], N, H, window=w_, center=True, pad_mode="constant", return_complex=True), 4 * B * (n + T * N))
    timeit("torch.istft (cuFFT)", lambda i: torch.istft(spec[i % 3], N, H, window=w_, center=True, length=n), 4 * B * (T * N + (T - 1) * H))
    del spec
if "--dftgemm" in sys.argv:
    # BASELINE config C5 asks for "radix FFT vs tensor-core DFT-GEMM variant".  Stand-in for the variant: the STFT
    # as one GEMM  frames [B*T', N] x packed DFT matrix [N, N]  through cuBLAS (library), in fp32 (CUDA cores),
    # single-pass TF32 (tensor cores, ~1e-3 relative error: fails the 1e-5 parity bar) and 3xTF32 (hi/lo split of
    # both operands, three tensor-core GEMMs: the cheapest tensor-core form that keeps fp32-grade products).
    # Zero-extension / tail padding is ignored (interior frames only) - this favours the GEMM.
    k = torch.arange(N // 2, device=dev, dtype=torch.float64)
    nn = torch.arange(N, device=dev, dtype=torch.float64)
    win = (0.5 - 0.5 * torch.cos(2 * torch.pi * nn / N)) * (2.0 / N)
    ang = 2 * torch.pi * nn[:, None] * k[None, :] / N
    Wm = torch.cat([torch.cos(ang), -torch.sin(ang)], dim=1) * win[:, None]          # [N, N]: Re | Im halves
    Wm[:, N // 2] = torch.cos(torch.pi * nn) * win                                     # Nyquist rides in the Im-DC slot
    Wm = Wm.float().contiguous()

    def tf32_split(a):
        hi = (a.view(torch.int32) & -8192).view(torch.float32)                         # keep 10 mantissa bits
        return hi, a - hi
    Whi, Wlo = tf32_split(Wm)

    def frames_of(w):
        return w.unfold(-1, N, H).reshape(-1, N)                                       # materialises [B*T', N] (4x the samples at hop N/4)

    def gemm(i, mode):
        f = frames_of(waves[i % 3])
        if mode == "fp32":
            torch.backends.cuda.matmul.allow_tf32 = False
            return f @ Wm
        torch.backends.cuda.matmul.allow_tf32 = True
        if mode == "tf32":
            return f @ Wm
        fhi, flo = tf32_split(f)
        return fhi @ Whi + (fhi @ Wlo + flo @ Whi)
    ref64 = (frames_of(waves[0]).double() @ Wm.double())
    for mode in ("fp32", "tf32", "3xtf32"):
        err = float((gemm(0, mode).double() - ref64).norm() / ref64.norm())
        timeit(f"DFT-GEMM cuBLAS {mode} (rel {err:.1e})", lambda i: gemm(i, mode), 4 * B * (n + T * N))
    torch.backends.cuda.matmul.allow_tf32 = False
    ours = ops.stft(waves[0], N, H)[:, N // (2 * H): N // (2 * H) + (n - N) // H + 1].reshape(-1, N)
    print(f"   (libgss stft on the same interior frames vs the fp64 GEMM: rel {float((ours.double() - ref64).norm() / ref64.norm()):.1e})")
    del ref64, ours
    # synthesis side: packed features [B*T, N] x inverse packed-DFT matrix [N, N] (window folded in) -> windowed frames
    # [B*T, N]; the overlap-add of those frames (another pass over 4x the output) is NOT included - this favours the GEMM.
    ck = torch.full((N // 2,), 2.0, device=dev, dtype=torch.float64); ck[0] = 1.0
    hann = 0.5 - 0.5 * torch.cos(2 * torch.pi * nn / N)
    Wi = torch.cat([torch.cos(ang).T * ck[:, None], -torch.sin(ang).T * ck[:, None]], dim=0) * (hann * 0.5)[None, :]   # [N, N]: rows = packed slots
    Wi[N // 2] = torch.cos(torch.pi * nn) * hann * 0.5                                    # the Nyquist slot (rides in Im-DC)
    Wi = Wi.float().contiguous()
    Wihi, Wilo = tf32_split(Wi)
    fm = [f.reshape(-1, N) for f in feat]

    def igemm(i, mode):
        f = fm[i % 3]
        if mode == "fp32":
            torch.backends.cuda.matmul.allow_tf32 = False
            return f @ Wi
        torch.backends.cuda.matmul.allow_tf32 = True
        if mode == "tf32":
            return f @ Wi
        fhi, flo = tf32_split(f)
        return fhi @ Wihi + (fhi @ Wilo + flo @ Wihi)
    ref64 = fm[0].double() @ Wi.double()
    for mode in ("fp32", "tf32", "3xtf32"):
        err = float((igemm(0, mode).double() - ref64).norm() / ref64.norm())
        timeit(f"iDFT-GEMM cuBLAS {mode} (rel {err:.1e}), no overlap-add", lambda i: igemm(i, mode), 4 * B * (T * N + (T - 1) * H))
    torch.backends.cuda.matmul.allow_tf32 = False
    del ref64
x = feat[0]
from gan_sass_tf_b200.app import hparams
hparams.FFT_SIZE = N
timeit("to_log", lambda i: ops.to_log_signal(feat[i % 3]), 8 * B * T * N)
timeit("to_exp", lambda i: ops.to_exp_signal(feat[i % 3]), 8 * B * T * N)
m3 = torch.rand(B, 3, T, N // 2, device=dev, generator=g)
timeit("apply_mask S=3", lambda i: ops.apply_mask(feat[i % 3], m3), 4 * B * T * N * (1 + 1.5 + 3))
