"""Optimizer-step microbenchmark on ViT-B/16's parameter list (152 tensors, 86.6 M elements): b200timm.optim vs
torch.optim (foreach, the default on CUDA) vs torch.optim(fused=True).  CUDA events, 20 steps after 5 warm-up steps;
reports ms per step and the HBM rate of the libb200timm kernel against its algorithmic bytes
(SGD: p r/w + g r + buf r/w; AdamW: p r/w + g r + m, v r/w; state fp32).
    python tools/optim_bench.py [--dtype bf16|fp32]"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "pytorch-image-models_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--steps", type=int, default=20)
    args = ap.parse_args()
    from b200timm import optim as BO
    from b200timm.vision_transformer import create_model
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    model = create_model("vit_base_patch16_224")
    shapes = [p.shape for p in model.parameters()]
    numel = sum(p.numel() for p in model.parameters())
    esz = 2 if dtype == torch.bfloat16 else 4
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except OSError:
        pass

    def params():
        ps = [torch.nn.Parameter(torch.randn(s, device="cuda").to(dtype)) for s in shapes]
        for p in ps:
            p.grad = torch.randn_like(p)
        return ps

    def timed(opt, clip=None):
        for _ in range(5):
            if clip:
                clip()
            opt.step()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(args.steps):
            if clip:
                clip()
            opt.step()
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) / args.steps

    out = {"dtype": args.dtype, "tensors": len(shapes), "elements": numel}
    for name, mk_mine, mk_torch, bytes_per in (
            ("sgd_nesterov", lambda p: BO.SGD(p, lr=1e-4, momentum=0.9, nesterov=True, weight_decay=1e-4),
             lambda p, **k: torch.optim.SGD(p, lr=1e-4, momentum=0.9, nesterov=True, weight_decay=1e-4, **k), 3 * esz + 8),
            ("adamw", lambda p: BO.AdamW(p, lr=1e-4, weight_decay=0.05),
             lambda p, **k: torch.optim.AdamW(p, lr=1e-4, weight_decay=0.05, **k), 3 * esz + 16)):
        ps = params()
        ms = timed(mk_mine(ps))
        out[name] = {"b200_ms": round(ms, 4), "b200_GBps": round(numel * bytes_per / ms / 1e6, 1),
                     "algorithmic_MB": round(numel * bytes_per / 1e6, 1)}
        ps = params()
        out[name]["torch_foreach_ms"] = round(timed(mk_torch(ps)), 4)
        ps = params()
        try:
            out[name]["torch_fused_ms"] = round(timed(mk_torch(ps, fused=True)), 4)
        except Exception as e:      # noqa: BLE001
            out[name]["torch_fused_ms"] = f"unavailable: {type(e).__name__}"
        ps = params()
        o = mk_mine(ps)
        out[name]["b200_with_fused_clip_ms"] = round(timed(o, lambda: BO.clip_grad_norm_(ps, 1.0, optimizer=o)), 4)
        ps = params()
        o = mk_torch(ps)
        out[name]["torch_with_clip_ms"] = round(timed(o, lambda: torch.nn.utils.clip_grad_norm_(ps, 1.0)), 4)
    if peaks:
        out["hbm_peak_GBps"] = peaks.get("hbm_gbps") or peaks.get("hbm", {})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
