"""Training-step benchmark (BASELINE.json configs[3]): vit_base_patch16_224, bf16, batch 256 per GPU, DDP over
NCCL when world_size > 1.  One step = zero_grad -> forward -> cross entropy -> backward (+ gradient all-reduce) ->
SGD step, i.e. benchmark.py:396-420 wrapped the way timm/task/classification.py:64-65 wraps the model.  The optimizer
is the one benchmark.py builds -- create_optimizer_v2(opt='sgd', lr=1e-4): Nesterov SGD, momentum 0.9, no weight decay
(benchmark.py:383-386, timm/optim/_optim_factory.py:355-361) -- as b200timm.optim.SGD (one multi-tensor launch;
B200_TORCH_OPTIM=1 swaps torch.optim.SGD in for A/B runs).  Forward and backward of the model run on libb200timm
kernels, and so does the cross-entropy loss (b200timm.loss; B200_TORCH_LOSS=1 for torch's)."""
import json
import os

import torch
import torch.distributed as dist

from . import ops
from .vision_transformer import create_model

MODEL = "vit_base_patch16_224"
METRIC = "images/sec (infer+train, bf16) vit_base_patch16_224 at 1/2/4/8 B200"
TRAIN_FLOPS_PER_IMG = 3 * 35.13e9


def run(args, rank, world, dev):
    torch.manual_seed(42 + rank)
    model = create_model(MODEL).to(dev).to(torch.bfloat16).train()
    net = model
    if world > 1:
        # DDP as timm wraps it (timm/task/classification.py:64-65); B200_DDP_OPTS="bucket_view,cap=<MB>,static" switches on
        # gradient_as_bucket_view / a larger bucket_cap_mb / static_graph for experiments
        kw = {}
        for tok in filter(None, os.environ.get("B200_DDP_OPTS", "").split(",")):
            if tok == "bucket_view":
                kw["gradient_as_bucket_view"] = True
            elif tok == "static":
                kw["static_graph"] = True
            elif tok.startswith("cap="):
                kw["bucket_cap_mb"] = int(tok[4:])
        net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[dev.index], broadcast_buffers=True, **kw)
    if os.environ.get("B200_TORCH_OPTIM") == "1":
        opt = torch.optim.SGD(model.parameters(), lr=1e-4, momentum=0.9, nesterov=True)
    else:
        from .optim import SGD
        opt = SGD(model.parameters(), lr=1e-4, momentum=0.9, nesterov=True)
    B = args.batch
    xs = [torch.randn(B, 3, 224, 224, device=dev).to(torch.bfloat16) for _ in range(2)]
    ys = [torch.empty(B, dtype=torch.long, device=dev).random_(1000) for _ in range(2)]
    host_x = [t.cpu().pin_memory() for t in xs]
    host_y = [t.cpu().pin_memory() for t in ys]
    if os.environ.get("B200_TORCH_LOSS") == "1":
        ce = torch.nn.CrossEntropyLoss()
        loss_fn = lambda out, y: ce(out.float(), y)      # noqa: E731
    else:
        from .loss import CrossEntropyLoss               # nn.CrossEntropyLoss of benchmark.py:380 on b200_cross_entropy_*
        loss_fn = CrossEntropyLoss()

    def step(x, y):
        opt.zero_grad(set_to_none=True)
        loss = loss_fn(net(x), y)
        loss.backward()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(xs[i & 1], ys[i & 1])
    barrier()
    ops.reset_counters()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    import time
    t0.record()
    h0 = time.perf_counter()
    for i in range(args.steps):
        step(xs[i & 1], ys[i & 1])
    host_ms = (time.perf_counter() - h0) * 1e3 / args.steps    # time the host needs to ENQUEUE one step
    t1.record()
    barrier()
    ms = t0.elapsed_time(t1)
    launches = ops.launch_count()

    # end to end: inputs from pinned host memory every step, loss read back every step
    def e2e_step(i):
        x = host_x[i & 1].to(dev, non_blocking=True)
        y = host_y[i & 1].to(dev, non_blocking=True)
        return step(x, y).item()

    e2e_step(0)
    barrier()
    t0.record()
    for i in range(args.steps):
        e2e_step(i)
    t1.record()
    barrier()
    e2e_ms = t0.elapsed_time(t1)
    times = torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, e2e_ms = times.tolist()
    total = B * args.steps * world
    value = total / (ms / 1e3)
    peaks = {}
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            peaks = json.load(f)
    peak = peaks.get("bf16_tflops_sustained", 1400.0)
    tf = TRAIN_FLOPS_PER_IMG * value / world / 1e12
    return {
        "metric": METRIC, "value": round(value, 1), "unit": "img/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{MODEL} bf16 TRAIN step batch={B}/GPU (BASELINE.json configs[3]): fwd + CE + bwd + "
                               "Nesterov SGD step (b200timm.optim), random-init weights, random targets", "global_batch": B * world,
                   "parallelism": f"dp{world} (torch DDP, NCCL gradient all-reduce)" if world > 1 else "single GPU",
                   "flush": "two alternating 77 MB input batches; ~25 GB of activations / gradients streamed per step"},
        "e2e": {"value": round(total / (e2e_ms / 1e3), 1), "unit": "img/s",
                "h2d_bytes_per_step": host_x[0].numel() * 2 + host_y[0].numel() * 8, "d2h_bytes_per_step": 4},
        "gpu_launches": launches, "host_enqueue_ms_per_step": round(host_ms, 2),
        "roofline": {"bound": "tensor", "kernel": "whole training step (3 x forward FLOPs)", "achieved": round(tf, 1),
                     "peak": peak, "unit": "TFLOP/s", "frac": round(tf / peak, 4), "traffic": None},
    }
