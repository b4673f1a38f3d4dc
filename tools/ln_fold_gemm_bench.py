"""Per-GEMM cost of the folded-LayerNorm epilogues at the bench shapes (M = 50432): the same GEMM with and without the
producer (ln_stats_out) / consumer (ln_in) arguments; CUDA events, 20 launches after 3 warm-ups."""
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


M, D = 50432, 768
x = torch.randn(M, D, device="cuda").bfloat16()
h4 = torch.randn(M, 4 * D, device="cuda").bfloat16()
res = torch.randn(M, D, device="cuda").bfloat16()
stats = torch.randn(ops.ln_slots(M, D) + 1, M, 2, device="cuda").abs() + 1.0
for name, a, N, epi, producer in (("qkv", x, 3 * D, ops.EPI_BIAS, False), ("fc1+GELU", x, 4 * D, ops.EPI_BIAS_GELU, False),
                                 ("proj+res", x, D, ops.EPI_BIAS_RESIDUAL, True), ("fc2+res", h4, D, ops.EPI_BIAS_RESIDUAL, True)):
    K = a.shape[1]
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).bfloat16()
    b = torch.randn(N, device="cuda").bfloat16()
    bf = b.float()
    c1 = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    for rnd in range(2):
        if producer:
            t0 = timeit(lambda: ops.linear_fwd(a, w, b, epilogue=epi, aux=res, out=out))
            t1 = timeit(lambda: ops.linear_fwd(a, w, b, epilogue=epi, aux=res, out=out, ln_stats_out=stats))
        else:
            t0 = timeit(lambda: ops.linear_fwd(a, w, b, epilogue=epi, out=out))
            t1 = timeit(lambda: ops.linear_fwd(a, w, bf, epilogue=epi, out=out, ln_in=(stats, c1, 1e-6)))
        print(f"{name:10s} plain {t0:7.1f} us   folded {t1:7.1f} us   ({(t1 / t0 - 1) * 100:+.1f} %)")
