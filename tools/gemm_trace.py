"""Where does a GEMM launch spend its time?  Per-CTA %globaltimer stamps from the b200_gemm debug hook
(include/b200timm.h: b200_set_trace), summarised as min / median / max over the 148 CTAs, relative to the
earliest CTA entry.     python tools/gemm_trace.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-image-models_b200"))
from b200timm import _lib, ops  # noqa: E402

NAMES = ["entry", "setup done", "first operands", "first acc ready", "last acc ready", "stores drained", "exit"]


def trace(M, N, K, epilogue=0, aux=False, dgrad=False):
    """dgrad=True: out[M,N] = a[M,K] @ w[K,N] with w consumed MN-major (the backward-data layout)."""
    lib = _lib.load()
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(K, N, device="cuda") if dgrad else torch.randn(N, K, device="cuda")).bfloat16()
    bias = None if dgrad else torch.randn(N, device="cuda").bfloat16()
    res = torch.randn(M, N, device="cuda").bfloat16() if aux else None

    def run():
        if dgrad:
            return ops.linear_dgrad(a, w, epilogue=epilogue, aux=res)
        return ops.linear_fwd(a, w, bias, epilogue=epilogue, aux=res)

    buf = torch.zeros(148 * 12, dtype=torch.int64, device="cuda")
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(20):
        run()
    t1.record()
    torch.cuda.synchronize()
    ev_us = t0.elapsed_time(t1) / 20 * 1e3
    # steady state: the traced launch is queued right behind ten untraced ones (the trace pointer is latched at launch)
    buf.zero_()
    torch.cuda.synchronize()
    for _ in range(10):
        run()
    lib.b200_set_trace(0, buf.data_ptr())
    out = run()
    lib.b200_set_trace(0, None)
    torch.cuda.synchronize()
    ref = (a.float() @ w.float()).bfloat16() if dgrad else torch.nn.functional.linear(a, w, bias)
    if epilogue == ops.EPI_DGELU:
        x = res.float()
        ref = (ref.float() * (0.5 * (1 + torch.erf(x * 0.7071067811865476)) + x * torch.exp(-0.5 * x * x) * 0.3989422804014327)).bfloat16()
    elif epilogue == ops.EPI_BIAS_GELU:
        ref = torch.nn.functional.gelu(ref.float()).bfloat16()
    elif aux:
        ref = ref + res
    err = (out.float() - ref.float()).abs().max().item() / ref.float().abs().max().item()
    t = buf.view(148, 12).cpu()
    base = t[:, 0].min().item()
    tiles = (M + 127) // 128 * ((N + 255) // 256)
    print(f"\nM={M} N={N} K={K} epi={epilogue}: events {ev_us:.1f} us/launch, {tiles} tiles = {tiles / 148:.2f} waves, "
          f"span entry..exit {(t[:, 6].max().item() - base) / 1e3:.1f} us")
    for i, n in enumerate(NAMES):
        v = (t[:, i] - base).float() / 1e3
        print(f"  {n:18s} min {v.min():7.2f}  med {v.median():7.2f}  max {v.max():7.2f} us")
    for i, n in ((7, "MMA wait operands"), (8, "MMA wait acc free"), (9, "TMA wait slot")):
        v = t[:, i].float() / 1.9e3
        print(f"  {n:18s} min {v.min():7.2f}  med {v.median():7.2f}  max {v.max():7.2f} us (cycles / 1.9 GHz)")
    ex = (t[:, 6] - base).float() / 1e3
    slow = torch.argsort(ex, descending=True)[:8]
    print(f"  tiles per CTA: min {t[:, 10].min().item()} max {t[:, 10].max().item()};  rel err vs torch {err:.2e};  "
          f"slowest CTAs (cta:smid:exit us) " + " ".join(f"{i}:{t[i, 11].item()}:{ex[i]:.1f}" for i in slow.tolist()))
    if os.environ.get("GEMM_TRACE_DUMP"):
        tag = f"{M}x{N}x{K}_e{epilogue}_" + ("2cta" if os.environ.get("B200_GEMM_2CTA", "0") == "1" else "1cta")
        torch.save(t, os.path.join(ROOT, "gpurun_out", f"trace_{tag}.pt"))
    q = torch.quantile(ex, torch.tensor([0.1, 0.25, 0.5, 0.75, 0.9, 1.0]))
    print("  exit quantiles 10/25/50/75/90/100%: " + " ".join(f"{v:.1f}" for v in q.tolist()))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "heavy":
        trace(50432, 3072, 768, dgrad=True)
        trace(50432, 3072, 768, epilogue=ops.EPI_DGELU, aux=True, dgrad=True)
        trace(50432, 3072, 768, epilogue=ops.EPI_BIAS_GELU)
        sys.exit(0)
    trace(50432, 768, 768)
    trace(50432, 768, 768, epilogue=ops.EPI_BIAS_RESIDUAL, aux=True)
    trace(50432, 2304, 768)
    trace(50432, 3072, 768, epilogue=ops.EPI_BIAS_GELU)
    trace(50432, 768, 3072)
    trace(148 * 128, 256, 768)
