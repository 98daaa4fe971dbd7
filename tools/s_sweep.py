import sys, os
sys.path.insert(0, "/root/repo")
import torch
from gan_sass_tf_b200.app import ops
from gan_sass_tf_b200 import _native
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(0)
def timeit(fn, reps=20):
    for i in range(3): fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): fn(i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for N, H, B, n in ((512, 128, 256, 48000), (256, 128, 1024, 48000), (256, 128, 8, 16256), (512, 256, 256, 48000)):
    T, _ = _native.frame_count(n, N, H)
    waves = [(torch.randn(B, n, device=dev, generator=g) * 0.1) for _ in range(3)]
    lins = [ops.stft(w, N, H) for w in waves]
    for S in (2, 3, 4):
        masks = [torch.rand(B, S, T, N // 2, device=dev, generator=g) for _ in range(3)]
        out = torch.empty(B * S, (T - 1) * H, device=dev)
        tf = timeit(lambda i: ops.mask_istft_feature(lins[i % 3], masks[i % 3], H, out=out))
        tw = timeit(lambda i: ops.mask_istft(waves[i % 3], masks[i % 3], N, H, out=out))
        print(f"N={N} H={H} B={B} S={S}: feature {tf:8.1f} us   wave {tw:8.1f} us")
