"""Per-kernel timings on the GPU box: each libb200timm kernel at the ViT-B/16 batch-256 shapes, with the
PyTorch-eager library kernel (cuBLASLt / ATen / SDPA) timed beside it as the on-box bar.
CUDA-event timing, L2 flushed between iterations.  Prints one JSON line per kernel."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-image-models_b200"))
from b200timm import ops  # noqa: E402

DEV = "cuda"
FLUSH = None


def timeit(fn, iters=10, warmup=3):
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(256 * 1024 * 1024, dtype=torch.int8, device=DEV)
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        FLUSH.zero_()
        s = torch.cuda.Event(enable_timing=True)
        e = torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def report(name, ms, flops=None, bytes_=None, ref_ms=None):
    d = {"kernel": name, "ms": round(ms, 4)}
    if flops:
        d["tflops"] = round(flops / ms / 1e9, 1)
    if bytes_:
        d["gbs"] = round(bytes_ / ms / 1e6, 1)
    if ref_ms:
        d["torch_ms"] = round(ref_ms, 4)
        d["speedup_vs_torch"] = round(ref_ms / ms, 3)
    print(json.dumps(d), flush=True)


def main():
    B = int(os.environ.get("MB_BATCH", "256"))
    M = B * 197
    D = 768
    g = torch.Generator(device=DEV).manual_seed(0)
    x = torch.randn(M, D, device=DEV, generator=g).bfloat16()
    for name, N, K, epi in (("qkv", 2304, 768, ops.EPI_BIAS), ("proj+res", 768, 768, ops.EPI_BIAS_RESIDUAL),
                            ("fc1+gelu", 3072, 768, ops.EPI_BIAS_GELU), ("fc2+res", 768, 3072, ops.EPI_BIAS_RESIDUAL)):
        a = torch.randn(M, K, device=DEV, generator=g).bfloat16()
        w = (torch.randn(N, K, device=DEV, generator=g) * 0.02).bfloat16()
        bias = torch.randn(N, device=DEV, generator=g).bfloat16()
        aux = torch.randn(M, N, device=DEV, generator=g).bfloat16() if epi == ops.EPI_BIAS_RESIDUAL else None
        out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
        ms = timeit(lambda: ops.gemm(a, w, M, N, K, epilogue=epi, bias=bias, aux=aux, out=out))
        if epi == ops.EPI_BIAS_GELU:
            ref = timeit(lambda: torch.nn.functional.gelu(torch.nn.functional.linear(a, w, bias)))
        elif epi == ops.EPI_BIAS_RESIDUAL:
            ref = timeit(lambda: aux + torch.nn.functional.linear(a, w, bias))
        else:
            ref = timeit(lambda: torch.nn.functional.linear(a, w, bias))
        report(f"gemm {name} M={M} N={N} K={K}", ms, flops=2.0 * M * N * K, ref_ms=ref)
    # backward-shaped GEMMs: dgrad (B MN-major) and wgrad (A, B MN-major, fp32 out)
    dy = torch.randn(M, 3072, device=DEV, generator=g).bfloat16()
    w1 = (torch.randn(3072, 768, device=DEV, generator=g) * 0.02).bfloat16()
    ms = timeit(lambda: ops.gemm(dy, w1, M, 768, 3072, b_mn=True))
    report(f"gemm dgrad fc1 M={M} N=768 K=3072", ms, flops=2.0 * M * 768 * 3072, ref_ms=timeit(lambda: dy @ w1))
    gw = torch.zeros(3072, 768, device=DEV)
    for sk in (1, 4, 8):
        ms = timeit(lambda: ops.gemm(dy, x, 3072, 768, M, a_mn=True, b_mn=True, out=gw, split_k=sk))
        report(f"gemm wgrad fc1 3072x768 K={M} split_k={sk}", ms, flops=2.0 * M * 768 * 3072,
               ref_ms=timeit(lambda: dy.t() @ x))
    # LayerNorm
    w = torch.randn(D, device=DEV, generator=g).bfloat16()
    b = torch.randn(D, device=DEV, generator=g).bfloat16()
    ms = timeit(lambda: ops.layernorm_fwd(x, w, b, 1e-6))
    report(f"layernorm fwd [{M},{D}] bf16", ms, bytes_=2.0 * M * D * 2,
           ref_ms=timeit(lambda: torch.nn.functional.layer_norm(x, (D,), w, b, 1e-6)))
    y, mean, rstd = ops.layernorm_fwd(x, w, b, 1e-6, save_stats=True)
    ms = timeit(lambda: ops.layernorm_bwd(y, x, w, mean, rstd))
    report(f"layernorm bwd [{M},{D}] bf16", ms, bytes_=3.0 * M * D * 2)
    # attention
    H, hd = 12, 64
    qkv = torch.randn(B, 197, 3 * H * hd, device=DEV, generator=g).bfloat16()
    ms = timeit(lambda: ops.attention_fwd(qkv, B, 197, H, hd, 0.125))
    q, k, v = qkv.view(B, 197, 3, H, hd).permute(2, 0, 3, 1, 4).unbind(0)
    ref = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(B, 197, D))
    report(f"attention fwd B={B} N=197 H=12", ms, flops=4.0 * B * H * 197 * 197 * hd, bytes_=4.0 * M * D * 2, ref_ms=ref)
    # patch embed helpers
    img = torch.randn(B, 3, 224, 224, device=DEV, generator=g).bfloat16()
    ms = timeit(lambda: ops.patchify(img, 16))
    report(f"patchify B={B}", ms, bytes_=2.0 * img.numel() * 2)
    # torch eager copy as a bandwidth sanity check
    y2 = torch.empty_like(x)
    report("torch copy (bandwidth sanity)", timeit(lambda: y2.copy_(x)), bytes_=2.0 * M * D * 2)


if __name__ == "__main__":
    main()
