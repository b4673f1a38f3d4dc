"""The fused-qkv weight gradient (dW [2304, 768] from M = 50432 rows) as one split-K launch (54 of 74 CTA pairs busy) and as
three launches over row chunks (72 pairs each): CUDA events, 20 repetitions."""
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
    return e0.elapsed_time(e1) / n * 1e3


for M, N, K in ((50432, 2304, 768), (36928, 3072, 1024)):
    dy = torch.randn(M, N, device="cuda").bfloat16()
    x = torch.randn(M, K, device="cuda").bfloat16()
    out = torch.zeros(N, K, device="cuda")
    for rnd in range(2):
        ops._WGRAD_NSPLIT = False
        t0 = timeit(lambda: ops.linear_wgrad(dy, x, out=out))
        ops._WGRAD_NSPLIT = True
        t1 = timeit(lambda: ops.linear_wgrad(dy, x, out=out))
        fl = 2.0 * M * N * K
        print(f"wgrad [{N}, {K}] over {M} rows: one launch {t0:.1f} us ({fl / t0 / 1e6:.0f} TFLOP/s)   row chunks {t1:.1f} us "
              f"({fl / t1 / 1e6:.0f} TFLOP/s)")
