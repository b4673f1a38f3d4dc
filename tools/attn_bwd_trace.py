"""Cycle budget of the attention backward kernel (N <= 208): per compute-warp phase and per issuer wait, averaged per item,
from the b200_set_trace(2, ...) debug hook.     python tools/attn_bwd_trace.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-image-models_b200"))
from b200timm import _lib, ops  # noqa: E402


def main(B=256, N=197, H=12):
    lib = _lib.load()
    qkv = torch.randn(B * N, 3 * H * 64, device="cuda").bfloat16()
    dout = torch.randn(B, N, H * 64, device="cuda").bfloat16()
    out, lse = ops.attention_fwd(qkv, B, N, H, 64, 0.125, need_lse=True)
    db = torch.zeros(3 * H * 64, device="cuda")
    for fused in (False, True):
        for _ in range(3):
            ops.attention_bwd(qkv, out, dout, lse, B, N, H, 64, 0.125, dbias=db if fused else None)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(20):
            ops.attention_bwd(qkv, out, dout, lse, B, N, H, 64, 0.125, dbias=db if fused else None)
        t1.record()
        torch.cuda.synchronize()
        print(f"B={B} N={N} H={H} dbias={fused}: {t0.elapsed_time(t1) / 20 * 1e3:.1f} us/launch (incl. the delta kernel)")
    buf = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
    lib.b200_set_trace(2, buf.data_ptr())
    ops.attention_bwd(qkv, out, dout, lse, B, N, H, 64, 0.125)
    lib.b200_set_trace(2, None)
    torch.cuda.synchronize()
    t = buf.view(148, 16).float()
    items = t[:, 15].clamp(min=1)
    names = ["compute: wait S^T/dP^T", "compute: prev-tile dV/dK store", "compute: softmax pass (TMEM ld + math + st)",
             "compute: wait last MMAs", "compute: tail stores (dV, dK, dQ)", "compute: next item's stats",
             None, None, "issuer: wait operands (TMA)", "issuer: wait P^T/dS^T", "issuer: wait accumulators read",
             "issuer: wait last MMAs (prefetch)", "issuer: total"]
    for i, n in enumerate(names):
        if n:
            v = t[:, i] / items
            print(f"  {n:44s} {v.mean():8.0f} cycles / item   (min {v.min():.0f} max {v.max():.0f})")
    print(f"  items per CTA {items.min():.0f}..{items.max():.0f}")


if __name__ == "__main__":
    main()
