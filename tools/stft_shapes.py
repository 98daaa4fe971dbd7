"""STFT (+log) timings of the N = 512 kernel at several batch sizes, for tuning builds selected with GSS_LIB."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gan_sass_tf_b200.app import ops
dev = torch.device("cuda"); g = torch.Generator(device=dev).manual_seed(0)
N, H = 512, 128
res = []
for B, n in ((256, 48000), (1024, 48000), (1024, 64000)):
    w = [(torch.randn(B, n, device=dev, generator=g) * 0.1).clamp_(-1, 1) for _ in range(3)]
    for lg in (False, True):
        for i in range(3): ops.stft(w[i], N, H, log=lg)
        torch.cuda.synchronize(); best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(12): ops.stft(w[i % 3], N, H, log=lg)
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 12)
        res.append(f"B={B} n={n} log={int(lg)}: {best*1e3:.1f}")
    del w
print(os.path.basename(os.environ.get("GSS_LIB", "libgss.so")), " | ".join(res))
