"""Cycle budget of the attention forward kernel (N <= 208): per softmax phase and per issuer wait, averaged per item,
from the b200_set_trace(1, ...) debug hook.     python tools/attn_trace.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-image-models_b200"))
from b200timm import _lib, ops  # noqa: E402


def main(B=256, N=197, H=12):
    lib = _lib.load()
    qkv = torch.randn(B * N, 3 * H * 64, device="cuda").bfloat16()
    for _ in range(3):
        ops.attention_fwd(qkv, B, N, H, 64, 0.125)
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(20):
        ops.attention_fwd(qkv, B, N, H, 64, 0.125)
    t1.record()
    torch.cuda.synchronize()
    print(f"B={B} N={N} H={H}: {t0.elapsed_time(t1) / 20 * 1e3:.1f} us/launch")
    buf = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
    for _ in range(5):
        ops.attention_fwd(qkv, B, N, H, 64, 0.125)
    lib.b200_set_trace(1, buf.data_ptr())
    ops.attention_fwd(qkv, B, N, H, 64, 0.125)
    lib.b200_set_trace(1, None)
    torch.cuda.synchronize()
    t = buf.view(148, 16).float()
    items = t[:, 15].clamp(min=1)
    names = ["softmax: wait S", "softmax: max pass", "softmax: exp pass + P store", "softmax: wait O", "softmax: epilogue",
             "issuer 0: wait operands", "issuer 0: wait TMEM free", "issuer 0: wait P",
             "issuer 1: wait operands", "issuer 1: wait TMEM free", "issuer 1: wait P", None, None, None, "issuer 0: total"]
    for i, n in enumerate(names):
        if n:
            v = t[:, i] / items
            print(f"  {n:30s} {v.mean():8.0f} cycles / item   (min {v.min():.0f} max {v.max():.0f})")
    print(f"  items per CTA {items.min():.0f}..{items.max():.0f}")


if __name__ == "__main__":
    main()
