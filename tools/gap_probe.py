"""Idle time between consecutive kernels on one stream, from the GEMM per-CTA %globaltimer trace: a chain
GEMM_a -> GEMM_b (back to back) and GEMM_b -> LayerNorm -> GEMM_c (a kernel with a different shared-memory carve-out in
between), queued deep enough that launch latency is hidden.     python tools/gap_probe.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-image-models_b200"))
from b200timm import _lib, ops  # noqa: E402


def main():
    lib = _lib.load()
    M, D = 50432, 768
    x = torch.randn(M, D, device="cuda").bfloat16()
    w1 = (torch.randn(3072, D, device="cuda") * 0.05).bfloat16()
    w2 = (torch.randn(D, 3072, device="cuda") * 0.05).bfloat16()
    b1, b2 = torch.zeros(3072, device="cuda").bfloat16(), torch.zeros(D, device="cuda").bfloat16()
    lw, lb = torch.ones(D, device="cuda").bfloat16(), torch.zeros(D, device="cuda").bfloat16()
    bufs = [torch.zeros(148 * 12, dtype=torch.int64, device="cuda") for _ in range(3)]

    def chain(trace):
        if trace:
            lib.b200_set_trace(0, bufs[0].data_ptr())
        h = ops.linear_fwd(x, w1, b1, epilogue=ops.EPI_BIAS_GELU)                       # GEMM_a
        if trace:
            lib.b200_set_trace(0, bufs[1].data_ptr())
        y = ops.linear_fwd(h, w2, b2, epilogue=ops.EPI_BIAS_RESIDUAL, aux=x)            # GEMM_b (back to back)
        n = ops.layernorm_fwd(y, lw, lb, 1e-6)[0]                                       # LayerNorm
        if trace:
            lib.b200_set_trace(0, bufs[2].data_ptr())
        ops.linear_fwd(n, w1, b1, epilogue=ops.EPI_BIAS_GELU)                           # GEMM_c
        if trace:
            lib.b200_set_trace(0, None)

    for _ in range(20):
        chain(False)
    chain(True)
    for _ in range(3):
        chain(False)
    torch.cuda.synchronize()
    t = [b.view(148, 12).cpu() for b in bufs]
    ent = [(v[:, 0][v[:, 0] > 0]).min().item() for v in t]
    ext = [v[:, 6].max().item() for v in t]
    first_exit = [(v[:, 6][v[:, 6] > 0]).min().item() for v in t]
    print(f"GEMM_a: {1e-3 * (ext[0] - ent[0]):.1f} us   GEMM_b: {1e-3 * (ext[1] - ent[1]):.1f} us   GEMM_c: {1e-3 * (ext[2] - ent[2]):.1f} us")
    print(f"last exit of GEMM_a -> first entry of GEMM_b: {1e-3 * (ent[1] - ext[0]):.2f} us "
          f"(first CTA exit of GEMM_a -> last: {1e-3 * (ext[0] - first_exit[0]):.2f} us)")
    print(f"last exit of GEMM_b -> first entry of GEMM_c: {1e-3 * (ent[2] - ext[1]):.2f} us  (a LayerNorm of ~33 us runs in between)")


if __name__ == "__main__":
    main()
