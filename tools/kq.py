"""Quick A/B timing of the two C2 kernels (stft_log, mask_istft S=3) for tuning builds selected with GSS_LIB.
Prints device time per launch (CUDA events, rotating inputs) and an output fingerprint to compare variants."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gan_sass_tf_b200.app import ops
from gan_sass_tf_b200 import _native

N, H, B, n, S = 512, 128, int(os.environ.get('KQ_B', 256)), int(os.environ.get('KQ_n', 48000)), 3
T, _ = _native.frame_count(n, N, H)
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
waves = [(torch.randn(B, n, device=dev, generator=g) * 0.1).clamp_(-1, 1) for _ in range(3)]
masks = [torch.rand(B, S, T, N // 2, device=dev, generator=g) for _ in range(3)]
out = torch.empty(B * S, (T - 1) * H, device=dev)


def timeit(fn, reps=30):
    for i in range(5):
        fn(i)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    return best * 1e3


tag = os.path.basename(os.environ.get("GSS_LIB", "libgss.so"))
extra = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("GSS_") and k != "GSS_LIB")
t_st = timeit(lambda i: ops.stft_log(waves[i % 3], N, H)) if "--no-stft" not in sys.argv else float("nan")
t_sy = timeit(lambda i: ops.mask_istft(waves[i % 3], masks[i % 3], N, H, out=out))
if "--istft" in sys.argv:
    feats = [ops.stft(w, N, H) for w in waves]
    t_i = timeit(lambda i: ops.istft(feats[i % 3], H))
    t_ie = timeit(lambda i: ops.istft(feats[i % 3], H, exp=True))
    y = ops.istft(feats[0], H)
    print(f"{tag:24s} istft {t_i:7.1f} us  istft_exp {t_ie:7.1f} us | fp {y.double().abs().sum().item():.6f}")
if "--feat" in sys.argv:
    lins = [ops.stft(w, N, H) for w in waves]
    lg = torch.empty_like(lins[0])
    t_d = timeit(lambda i: ops.stft_dual(waves[i % 3], N, H, out_lin=lins[i % 3], out_log=lg))
    t_f = timeit(lambda i: ops.mask_istft_feature(lins[i % 3], masks[i % 3], H, out=out))
    rows = torch.zeros(B, device=dev)
    t_a = timeit(lambda i: ops.mask_istft_feature(lins[i % 3], masks[i % 3], H, out=out, ae_rows=rows))
    ops.mask_istft_feature(lins[0], masks[0], H, out=out)
    ref = ops.istft(ops.apply_mask(lins[0][:32], masks[0][:32]), H)
    err = float((out[:96] - ref).norm() / ref.norm())
    print(f"{tag:24s} {extra:28s} stft_dual {t_d:7.1f} us  mask_istft_feature {t_f:7.1f} us  +AE {t_a:7.1f} us | fp out {out.double().abs().sum().item():.6f} rel vs unfused {err:.2e}")
ops.mask_istft(waves[0], masks[0], N, H, out=out)
f = ops.stft_log(waves[0], N, H)
print(f"{tag:24s} {extra:28s} stft_log {t_st:7.1f} us  mask_istft {t_sy:7.1f} us  | fp out {out.double().abs().sum().item():.6f} feat {f.double().abs().sum().item():.6f}")
