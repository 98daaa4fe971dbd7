import sys, os
sys.path.insert(0, os.getcwd())
import torch
from gan_sass_tf_b200.app import ops
dev = torch.device("cuda"); g = torch.Generator(device=dev).manual_seed(0)
N, H = 512, 128
res = []
for B, n in ((256, 48000), (1024, 48000), (1024, 64000)):
    w = [(torch.randn(B, n, device=dev, generator=g) * 0.1).clamp_(-1, 1) for _ in range(3)]
    f = [ops.stft(x, N, H) for x in w]
    for ex in (False, True):
        for i in range(3): ops.istft(f[i], H, exp=ex)
        torch.cuda.synchronize(); best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(12): ops.istft(f[i % 3], H, exp=ex)
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 12)
        res.append(f"B={B} n={n} exp={int(ex)}: {best*1e3:.1f}")
    del w, f
print(os.path.basename(os.environ.get("GSS_LIB", "libgss.so")), " | ".join(res))
