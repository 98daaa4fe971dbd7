"""Minimal launcher for ncu captures of the team kernels: the feature-fed step at one FFT size, B = 1024 x 3 s, hop N/4.
Usage: python tools/prof_team.py N [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gan_sass_tf_b200.app import ops
from gan_sass_tf_b200 import _native

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
H, B, n, S = N // 4, 1024, 48000, 3
T, _ = _native.frame_count(n, N, H)
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
waves = [(torch.randn(B, n, device=dev, generator=g) * 0.1).clamp_(-1, 1) for _ in range(2)]
masks = [torch.rand(B, S, T, N // 2, device=dev, generator=g) for _ in range(2)]
lin = [torch.empty(B, T, N, device=dev) for _ in range(2)]
lg = torch.empty(B, T, N, device=dev)
out = torch.empty(B * S, (T - 1) * H, device=dev)
for i in range(reps):
    k = i % 2
    ops.stft_dual(waves[k], N, H, out_lin=lin[k], out_log=lg)
    ops.mask_istft_feature(lin[k], masks[k], H, out=out)
torch.cuda.synchronize()
print("prof_team done", N, reps, float(out.double().abs().sum()))
