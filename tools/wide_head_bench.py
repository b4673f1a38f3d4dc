"""Times the 128-wide attention kernels (ViT-H/14 @224 geometry: 257 tokens, 16 heads of 80 padded to 128) next to the
64-wide ones on the same number of heads and tokens; CUDA events, 20 launches after 3 warm-ups."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-image-models_b200"))
from b200timm import ops  # noqa: E402


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for B, N, H in ((64, 257, 16), (64, 577, 16), (256, 197, 12)):
    for hd in (64, 128):
        qkv = torch.randn(B, N, 3 * H * hd, device="cuda").bfloat16()
        dout = torch.randn(B, N, H * hd, device="cuda").bfloat16()
        out, lse = ops.attention_fwd(qkv, B, N, H, hd, hd ** -0.5, need_lse=True)
        tf = timeit(lambda: ops.attention_fwd(qkv, B, N, H, hd, hd ** -0.5, need_lse=True))
        tb = timeit(lambda: ops.attention_bwd(qkv, out, dout, lse, B, N, H, hd, hd ** -0.5))
        fl = 4.0 * B * H * N * N * hd
        print(f"B={B} N={N} H={H} hd={hd}: fwd {tf * 1e3:.1f} us ({fl / tf / 1e9:.0f} TFLOP/s)  bwd {tb * 1e3:.1f} us "
              f"({2.5 * fl / tb / 1e9:.0f} TFLOP/s)")
