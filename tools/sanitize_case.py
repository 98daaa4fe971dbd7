"""Small launches of every kernel family for ``compute-sanitizer`` (memcheck / racecheck / synccheck):
the three N = 512 register-exchange kernels (+ the feature-fed synthesis), their N = 256 half-warp forms,
one team kernel set (N = 1024) and the element-wise / reduction kernels.  Shapes are tiny (the sanitizer
slows kernels by 10-100x) but cover ragged tails, several chunks per row and S = 1..3.

Usage: compute-sanitizer --tool memcheck python tools/sanitize_case.py [N ...]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gan_sass_tf_b200.app import ops
from gan_sass_tf_b200 import _native

sizes = [int(v) for v in sys.argv[1:]] or [512, 256, 1024]
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
for N in sizes:
    for H in (N // 4, N // 2):
        for B, n in ((3, 6 * N + 37), (2, 40 * N)):
            x = torch.randn(B, n, device=dev, generator=g) * 0.1
            T, _ = _native.frame_count(n, N, H)
            f = ops.stft(x, N, H)
            fl = ops.stft_log(x, N, H)
            y = ops.istft(f, H)
            ye = ops.istft(fl, H, exp=True)
            for S in (1, 3):
                m = torch.rand(B, S, T, N // 2, device=dev, generator=g)
                w = ops.mask_istft(x, m, N, H)
                if hasattr(ops, "mask_istft_feature"):
                    w2 = ops.mask_istft_feature(f, m, H)
            torch.cuda.synchronize()
            err = (y[:, :n] - x).abs().max().item()
            assert err < 1e-4, (N, H, err)
    print(f"N={N}: ok", flush=True)
# element-wise and reduction kernels
from gan_sass_tf_b200.app import hparams
hparams.FFT_SIZE = 256
a = torch.rand(6, 16, 256, device=dev, generator=g)
b = torch.rand(8, 16, 256, device=dev, generator=g)
ops.to_exp_signal(ops.to_log_signal(a))
ops.batch_cross_snr(a.view(2, 3, 16, 256), b.view(2, 4, 16, 256))
ops.ae_loss(b, a[:2], 4)
ops.wav16_normalise(torch.randn(3, 1000, device=dev, generator=g))
ops.mix_signals(a, 3, log=True)
torch.cuda.synchronize()
print("sanitize_case: done")
