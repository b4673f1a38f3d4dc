"""Device-resident inference throughput of one mirrored model (ViT or ConvNeXt) -- the non-headline BASELINE configs.
    python tools/bench_model.py convnext_base --batch 256 --steps 20
Prints one JSON line: img/s, ms/step and the per-kernel-type time split from CUDA events."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-image-models_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("model")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--img-size", type=int, default=224)
    ap.add_argument("--detail", action="store_true", help="per-op CUDA-event time split of one step")
    ap.add_argument("--train", action="store_true", help="forward + cross-entropy + backward + SGD step")
    args = ap.parse_args()
    from b200timm import ops
    if args.model.startswith("naflex:"):
        return naflex(args, ops)
    if args.model.startswith("convnext"):
        from b200timm.convnext import create_model
        model = create_model(args.model)
    else:
        from b200timm.vision_transformer import create_model
        kw = {} if args.img_size == 224 or "336" in args.model else {"img_size": args.img_size}
        model = create_model(args.model, **kw)
    torch.manual_seed(42)
    model = model.cuda().to(torch.bfloat16).eval()
    size = 336 if "336" in args.model else args.img_size
    xs = [torch.randn(args.batch, 3, size, size, device="cuda").bfloat16() for _ in range(2)]
    if args.train:
        return train(args, model, xs, ops)
    with torch.no_grad():
        for i in range(args.warmup):
            model(xs[i & 1])
        torch.cuda.synchronize()
        ops.reset_counters()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for i in range(args.steps):
            model(xs[i & 1])
        t1.record()
        torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / args.steps
    if args.detail:
        with torch.no_grad():
            ops.enable_op_timing(True)
            model(xs[0])
            det = ops.op_timing_summary()
            ops.enable_op_timing(False)
        tot = sum(t for _, t in det.values())
        for k, (n, t) in sorted(det.items(), key=lambda kv: -kv[1][1]):
            print(f"  {t:8.3f} ms {100 * t / tot:5.1f}%  n={n:3d}  {k}")
    print(json.dumps({"model": args.model, "batch": args.batch, "img_size": size, "ms_per_step": round(ms, 3),
                      "img_per_s": round(args.batch / ms * 1e3, 1), "launches_per_step": ops.launch_count() // args.steps}))


def naflex(args, ops):
    """BASELINE configs[4]: vit_large_patch14_clip_336 as a NaFlexVit on a variable-sequence batch (patch dicts with
    random aspect ratios, <= 576 patches, zero padded + patch_valid), bf16 inference.
        python tools/bench_model.py naflex:vit_large_patch14_clip_336 --batch 64"""
    from b200timm.naflexvit import create_model
    name = args.model.split(":", 1)[1]
    torch.manual_seed(42)
    model = create_model(name).cuda().to(torch.bfloat16).eval()
    g = torch.Generator().manual_seed(0)
    B, P, N = args.batch, 14, 576

    def batch():
        patches = torch.zeros(B, N, P * P * 3)
        coord = torch.zeros(B, N, 2, dtype=torch.int64)
        valid = torch.zeros(B, N, dtype=torch.bool)
        for b in range(B):
            h = int(torch.randint(12, 33, (1,), generator=g))
            w = min(48, max(1, N // h))
            n = h * w
            patches[b, :n] = torch.randn(n, P * P * 3, generator=g)
            yy, xx = torch.meshgrid(torch.arange(h), torch.arange(w), indexing="ij")
            coord[b, :n] = torch.stack([yy.flatten(), xx.flatten()], -1)
            valid[b, :n] = True
        return patches.cuda().bfloat16(), coord.cuda(), valid.cuda()

    xs = [batch() for _ in range(2)]
    tokens = sum(int(v.sum()) for _, _, v in xs) / 2
    if args.train:      # forward + cross-entropy + backward + SGD on the ragged batch
        model.train()
        from b200timm.optim import SGD
        opt = SGD(model.parameters(), lr=1e-3, momentum=0.9, nesterov=True)
        target = torch.randint(0, 1000, (B,), device="cuda")

        def step(x):
            opt.zero_grad(set_to_none=True)
            loss = torch.nn.functional.cross_entropy(model(*x).float(), target)
            loss.backward()
            opt.step()
            return loss

        for i in range(args.warmup):
            step(xs[i & 1])
        torch.cuda.synchronize()
        ops.reset_counters()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for i in range(args.steps):
            loss = step(xs[i & 1])
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / args.steps
        print(json.dumps({"model": args.model, "mode": "train", "batch": B, "max_patches": N,
                          "valid_patches_per_batch": tokens, "ms_per_step": round(ms, 3),
                          "img_per_s": round(B / ms * 1e3, 1), "loss": round(loss.item(), 4),
                          "launches_per_step": ops.launch_count() // args.steps}))
        return
    with torch.no_grad():
        for i in range(args.warmup):
            model(*xs[i & 1])
        torch.cuda.synchronize()
        ops.reset_counters()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for i in range(args.steps):
            model(*xs[i & 1])
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / args.steps
        if args.detail:
            ops.enable_op_timing(True)
            model(*xs[0])
            det = ops.op_timing_summary()
            ops.enable_op_timing(False)
            tot = sum(t for _, t in det.values())
            for k, (n, t) in sorted(det.items(), key=lambda kv: -kv[1][1]):
                print(f"  {t:8.3f} ms {100 * t / tot:5.1f}%  n={n:3d}  {k}")
    print(json.dumps({"model": args.model, "batch": B, "max_patches": N, "valid_patches_per_batch": tokens,
                      "ms_per_step": round(ms, 3), "img_per_s": round(B / ms * 1e3, 1),
                      "launches_per_step": ops.launch_count() // args.steps}))


def train(args, model, xs, ops):
    model.train()
    from b200timm.optim import SGD
    opt = SGD(model.parameters(), lr=1e-3, momentum=0.9, nesterov=True)
    target = torch.randint(0, 1000, (args.batch,), device="cuda")

    def step(x):
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.cross_entropy(model(x).float(), target)
        loss.backward()
        opt.step()
        return loss

    for i in range(args.warmup):
        step(xs[i & 1])
    torch.cuda.synchronize()
    ops.reset_counters()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(args.steps):
        loss = step(xs[i & 1])
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / args.steps
    if args.detail:
        ops.enable_op_timing(True)
        step(xs[0])
        det = ops.op_timing_summary()
        ops.enable_op_timing(False)
        tot = sum(t for _, t in det.values())
        for k, (n, t) in sorted(det.items(), key=lambda kv: -kv[1][1]):
            print(f"  {t:8.3f} ms {100 * t / tot:5.1f}%  n={n:3d}  {k}")
    print(json.dumps({"model": args.model, "mode": "train", "batch": args.batch, "img_size": xs[0].shape[-1],
                      "ms_per_step": round(ms, 3), "img_per_s": round(args.batch / ms * 1e3, 1), "loss": round(loss.item(), 4),
                      "launches_per_step": ops.launch_count() // args.steps}))


if __name__ == "__main__":
    main()
