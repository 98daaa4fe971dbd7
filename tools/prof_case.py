"""Minimal launcher for ncu captures: the C2 step through one path, a few times.
Usage: python tools/prof_case.py [wave|feature] [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gan_sass_tf_b200.app import ops
from gan_sass_tf_b200 import _native

path = sys.argv[1] if len(sys.argv) > 1 else "feature"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
N, H, B, n, S = 512, 128, 256, 48000, 3
T, _ = _native.frame_count(n, N, H)
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
waves = [(torch.randn(B, n, device=dev, generator=g) * 0.1).clamp_(-1, 1) for _ in range(3)]
masks = [torch.rand(B, S, T, N // 2, device=dev, generator=g) for _ in range(3)]
lin = [torch.empty(B, T, N, device=dev) for _ in range(3)]
lg = torch.empty(B, T, N, device=dev)
out = torch.empty(B * S, (T - 1) * H, device=dev)
for i in range(reps):
    k = i % 3
    if path == "feature":
        ops.stft_dual(waves[k], N, H, out_lin=lin[k], out_log=lg)
        ops.mask_istft_feature(lin[k], masks[k], H, out=out, reverse=True)
    else:
        ops.stft(waves[k], N, H, log=True)
        ops.mask_istft(waves[k], masks[k], N, H, out=out)
torch.cuda.synchronize()
print("prof_case done", path, reps, float(out.double().abs().sum()))
