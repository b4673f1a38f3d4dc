"""Training-step measurement (BASELINE.json configs[3]): vit_base_patch16_224, bf16, batch 256 per GPU, torch DDP over NCCL
when world_size > 1.  One step = zero_grad -> forward -> cross entropy -> backward (+ gradient all-reduce) -> SGD step,
i.e. benchmark.py:396-420 wrapped the way timm/task/classification.py:64-65 wraps the model.  The optimizer is the one
benchmark.py builds -- create_optimizer_v2(opt='sgd', lr=1e-4): Nesterov SGD, momentum 0.9 (benchmark.py:383-386,
timm/optim/_optim_factory.py:355-361) -- as b200timm.optim.SGD (one multi-tensor launch).  Forward, backward, loss and
optimizer run on libb200timm kernels; the whole step is one CUDA graph (b200timm.train_step.TrainStep) unless
--no-graph.  Environment switches for A/B runs: B200_TORCH_OPTIM=1 / B200_TORCH_LOSS=1 swap torch's optimizer / loss in,
B200_DDP_OPTS="cap=<MB>,static,no_bucket_view" tune the DDP wrapper."""
import os
import time

import torch
import torch.distributed as dist

from . import ops
from .train_step import TrainStep, wrap_ddp

TRAIN_FLOPS_PER_IMG = 3 * 35.13e9


def build(model_name, dev, world, dp="ddp"):
    """model, net (the model or its DDP wrapper), optimizer, loss, grad_sync (PeerGradSync or None)."""
    from .vision_transformer import create_model
    model = create_model(model_name).to(dev).to(torch.bfloat16).train()
    net = model
    sync = None
    if world > 1 and dp == "peer":
        # same start on every rank, as DDP's constructor guarantees by broadcasting rank 0's parameters
        for p in model.parameters():
            dist.broadcast(p.data, 0)
        from .grad_sync import PeerGradSync
        flag = torch.ones(1, device=dev)
        try:
            sync = PeerGradSync(model)
        except Exception as e:      # no NVLink multicast / symmetric memory on this node: every rank falls back together
            flag.zero_()
            if dist.get_rank() == 0:
                print(f"[b200timm] PeerGradSync unavailable ({type(e).__name__}: {e}); using torch DDP", flush=True)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if flag.item() == 0:
            sync, dp = None, "ddp"
    if world > 1 and sync is None:
        kw = {"bucket_cap_mb": 1024}      # one bucket: see DESIGN.md section 6 (the overlapped buckets cost more than they hide)
        for tok in filter(None, os.environ.get("B200_DDP_OPTS", "").split(",")):
            if tok == "no_bucket_view":
                kw["gradient_as_bucket_view"] = False
            elif tok == "static":
                kw["static_graph"] = True
            elif tok.startswith("cap="):
                kw["bucket_cap_mb"] = int(tok[4:])
        net = wrap_ddp(model, dev, **kw)
    if os.environ.get("B200_TORCH_OPTIM") == "1":
        opt = torch.optim.SGD(model.parameters(), lr=1e-4, momentum=0.9, nesterov=True)
    else:
        from .optim import SGD
        opt = SGD(model.parameters(), lr=1e-4, momentum=0.9, nesterov=True)
    if os.environ.get("B200_TORCH_LOSS") == "1":
        ce = torch.nn.CrossEntropyLoss()
        loss_fn = lambda out, y: ce(out.float(), y)      # noqa: E731
    else:
        from .loss import CrossEntropyLoss               # nn.CrossEntropyLoss of benchmark.py:380 on b200_cross_entropy_*
        loss_fn = CrossEntropyLoss()
    return model, net, opt, loss_fn, sync


def run(args, rank, world, dev, model_name="vit_base_patch16_224", clocks=None):
    """Returns the training record: value (img/s, all ranks), ms_per_step, e2e, launches, roofline of the GEMM launches."""
    torch.manual_seed(42 + rank)
    dp = getattr(args, "dp", "ddp")
    model, net, opt, loss_fn, sync = build(model_name, dev, world, dp)
    B = args.batch
    # two device-resident input batches (2 x 77 MB > 126 MB L2) alternate between steps; a step streams ~25 GB
    xs = [torch.randn(B, 3, 224, 224, device=dev).to(torch.bfloat16) for _ in range(2)]
    ys = [torch.empty(B, dtype=torch.long, device=dev).random_(1000) for _ in range(2)]
    host_x = [t.cpu().pin_memory() for t in xs]
    host_y = [t.cpu().pin_memory() for t in ys]
    step = TrainStep(net, loss_fn, opt, xs[0], ys[0], graph=not args.no_graph, grad_sync=sync)
    runners = [step.capture_on(xs[i], ys[i]) for i in range(2)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        runners[i & 1]()
    barrier()
    if clocks is not None:
        clocks.start()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record()
    h0 = time.perf_counter()
    for i in range(args.steps):
        runners[i & 1]()
    host_ms = (time.perf_counter() - h0) * 1e3 / args.steps    # time the host needs to ENQUEUE one step
    t1.record()
    barrier()
    ms = t0.elapsed_time(t1)
    clock_info = clocks.stop() if clocks is not None else None
    clocks_per_rank = None
    if world > 1:
        # every rank's SM clock under load: the slowest (most power-capped) GPU sets the pace of a synchronous step
        infos = [None] * world
        dist.all_gather_object(infos, None if clock_info is None else
                               {"sm_mhz": clock_info.get("sm_mhz"), "reasons": clock_info.get("reasons")})
        clocks_per_rank = infos

    # ---- end to end: inputs from pinned host memory every step (H2D on a copy stream into the buffer the other graph is
    # not reading), the loss read back on the host every step (the previous step's, so that one step stays in flight)
    copy_stream = torch.cuda.Stream(device=dev)
    compute = torch.cuda.current_stream()
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    host_loss = torch.zeros(2, dtype=torch.float32).pin_memory()
    loss_ready = [torch.cuda.Event() for _ in range(2)]

    def upload(slot, src):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])
            xs[slot].copy_(host_x[src], non_blocking=True)
            ys[slot].copy_(host_y[src], non_blocking=True)
            ready[slot].record(copy_stream)

    def e2e_loop(n):
        losses = []
        for s in range(2):
            consumed[s].record(compute)
        upload(0, 0)
        for i in range(n):
            cur, nxt = i & 1, (i + 1) & 1
            if i + 1 < n:
                upload(nxt, nxt)
            compute.wait_event(ready[cur])
            loss = runners[cur]()
            consumed[cur].record(compute)
            host_loss[cur:cur + 1].copy_(loss.detach().reshape(1), non_blocking=True)
            loss_ready[cur].record(compute)
            if i >= 1:
                loss_ready[nxt].synchronize()
                losses.append(float(host_loss[nxt]))
        compute.synchronize()
        losses.append(float(host_loss[(n - 1) & 1]))
        return losses

    e2e_loop(2)
    barrier()
    t0.record()
    losses = e2e_loop(args.steps)
    t1.record()
    barrier()
    e2e_ms = t0.elapsed_time(t1)

    # ---- roofline.achieved of the dominant kernel: the same K steps eagerly with CUDA events around every GEMM launch
    ops.reset_counters()
    step._eager(xs[0], ys[0])
    eager_launches = ops.launch_count()
    ops.enable_gemm_timing(True)
    for i in range(min(args.steps, 10)):
        step._eager(xs[i & 1], ys[i & 1])
    gemm_flops, gemm_ms, gemm_n = ops.gemm_timing_summary()
    ops.enable_gemm_timing(False)
    barrier()

    times = torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, e2e_ms = times.tolist()
    total = B * args.steps * world
    value = total / (ms / 1e3)
    launches_per_step = step.launches if step.graph_enabled else eager_launches
    ddp = ""
    if world > 1 and sync is not None:
        ddp = (f"dp{world} (b200timm.grad_sync.PeerGradSync: every block's gradients are pushed to all ranks by multimem.red.add "
               f"over the NVLink multicast address from the kernel that finalises them -- no all-reduce kernel; 173 MB bf16 "
               f"gradients / step / rank; NCCL process group for setup / barriers only)")
    elif world > 1:
        ddp = (f"dp{world} (torch DistributedDataParallel, gradient_as_bucket_view, NCCL all-reduce of the gradient bucket(s); "
               f"173 MB bf16 gradients / step / rank)")
    return {
        "value": round(value, 1), "unit": "img/s", "ms_per_step": round(ms / args.steps, 3),
        "workload": f"{model_name} bf16 TRAIN step batch={B}/GPU (BASELINE.json configs[3]): zero_grad + fwd + CE + bwd + "
                    "Nesterov SGD step (benchmark.py:396-420), random-init weights, random targets",
        "parallelism": ddp if world > 1 else "single GPU (no collective)",
        "launch": "eager" if not step.graph_enabled else "one CUDA graph per resident input buffer (whole step incl. NCCL + optimizer)",
        "e2e": {"value": round(total / (e2e_ms / 1e3), 1), "unit": "img/s",
                "h2d_bytes_per_step": host_x[0].numel() * 2 + host_y[0].numel() * 8, "d2h_bytes_per_step": 4,
                "note": "pinned host batch + targets -> H2D on a copy stream (double buffered) -> graph replay of the step -> "
                        "loss D2H and read on the host every step (one step in flight)", "last_loss": round(losses[-1], 4)},
        "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,
        "host_enqueue_ms_per_step": round(host_ms, 3),
        "gemm": {"flops": gemm_flops, "ms": gemm_ms, "n": gemm_n},
        "whole_step_tflops": round(TRAIN_FLOPS_PER_IMG * value / world / 1e12, 1),
        "clocks": clock_info, "clocks_per_rank": clocks_per_rank,
    }
