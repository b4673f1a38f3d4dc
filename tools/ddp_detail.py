#!/usr/bin/env python
"""Where does the DDP step lose time?  Same process, same box: the graphed ViT-B training step (a) without DDP, (b) DDP
default buckets, (c) DDP one bucket (all-reduce after the backward), plus the eager per-op CUDA-event split of each.
    torchrun --nproc-per-node 2 tools/ddp_detail.py"""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-image-models_b200"))
from b200timm import loss as BL, ops, optim as BO  # noqa: E402
from b200timm.train_step import TrainStep, prepare_ddp_env, wrap_ddp  # noqa: E402
from b200timm.vision_transformer import create_model  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    prepare_ddp_env()
    dist.init_process_group("nccl", device_id=dev)
    B = 256
    xs = [torch.randn(B, 3, 224, 224, device=dev).bfloat16() for _ in range(2)]
    ys = [torch.randint(0, 1000, (B,), device=dev) for _ in range(2)]
    out = {}
    variants = [("no_ddp", None), ("ddp_default", {}), ("ddp_one_bucket", {"bucket_cap_mb": 1024}), ("peer_push", "peer")]
    if os.environ.get("VARIANTS"):
        variants = [v for v in variants if v[0] in os.environ["VARIANTS"].split(",")]
    for name, kw in variants:
        torch.manual_seed(1)
        model = create_model("vit_base_patch16_224").to(dev).to(torch.bfloat16).train()
        sync = None
        if kw == "peer":
            from b200timm.grad_sync import PeerGradSync
            net, sync = model, PeerGradSync(model)
        else:
            net = model if kw is None else wrap_ddp(model, dev, **kw)
        opt = BO.SGD(model.parameters(), lr=1e-4, momentum=0.9, nesterov=True)
        step = TrainStep(net, BL.CrossEntropyLoss(), opt, xs[0], ys[0], grad_sync=sync)
        runners = [step.capture_on(xs[i], ys[i]) for i in range(2)]
        for i in range(4):
            runners[i & 1]()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(20):
            runners[i & 1]()
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 20
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)       # the slower rank sets the pace of every synchronised variant
        ms_max = t.item()
        # eager per-op split on a side stream (DDP's AccumulateGrad nodes live there)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step._eager(xs[0], ys[0])
            ops.enable_op_timing(True)
            step._eager(xs[1], ys[1])
            det = ops.op_timing_summary()
            ops.enable_op_timing(False)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        out[name] = {"graph_ms": round(ms, 3), "graph_ms_max_over_ranks": round(ms_max, 3), "ops": {k: round(t, 3) for k, (n, t) in det.items()},
                     "ops_total": round(sum(t for _, t in det.values()), 3)}
        del runners, step, net, model, opt, sync
        torch.cuda.empty_cache()
        dist.barrier()
    if rank == 0:
        names = list(out)
        print(json.dumps({n: out[n]["graph_ms"] for n in names}), "max over ranks:",
              json.dumps({n: out[n]["graph_ms_max_over_ranks"] for n in names}), {n: out[n]["ops_total"] for n in names})
        keys = sorted(out[names[0]]["ops"], key=lambda k: -out[names[0]]["ops"][k])
        for k in keys[:30]:
            print("  ".join(f"{out[n]['ops'].get(k, 0):8.3f}" for n in names), " ", k)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
