import sys, os
sys.path.insert(0, os.getcwd())
import torch
from gan_sass_tf_b200.app import ops
from gan_sass_tf_b200 import _native
dev = torch.device("cuda"); g = torch.Generator(device=dev).manual_seed(0)
N, H, S = 512, 128, 3
res = []
for B, n in ((256, 48000), (1024, 64000), (100, 48000), (37, 160000), (8, 960000), (1, 960000), (600, 16000)):
    T, _ = _native.frame_count(n, N, H)
    w = (torch.randn(B, n, device=dev, generator=g) * 0.1).clamp_(-1, 1)
    m = torch.rand(B, S, T, N // 2, device=dev, generator=g)
    out = torch.empty(B * S, (T - 1) * H, device=dev)
    for _ in range(3): ops.mask_istft(w, m, N, H, out=out)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): ops.mask_istft(w, m, N, H, out=out)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / 10)
    res.append(f"B={B} n={n}: {best*1e3:.1f} us  fp {out.double().abs().sum().item():.4f}")
    del w, m, out
print(os.path.basename(os.environ.get("GSS_LIB", "libgss.so")), " | ".join(res))
