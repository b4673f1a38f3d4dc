#!/usr/bin/env python
"""N-rank GPU check of the data-parallel path (run with torchrun, N = 2 is enough):

  1. the DDP-averaged gradients of the KERNEL model (b200timm autograd functions under torch DistributedDataParallel +
     NCCL) equal the gradients of the same model on the whole global batch computed by one rank without DDP;
  2. the same for b200timm.grad_sync.PeerGradSync (gradients pushed to all ranks by multimem.red.add from the kernels that
     finalise them -- no all-reduce);
  3. every rank ends with bit-identical gradients;
  4. three graphed steps (TrainStep, whole step incl. the exchange in one CUDA graph) == three eager steps, for both.

    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tools/ddp_parity.py [--out profiles/r02_ddp_parity.json]
All DDP work runs on side streams (a DDP backward on the legacy default stream cannot be captured later)."""
import argparse
import copy
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-image-models_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--model", default="vit_base_patch16_224")
    ap.add_argument("--per-rank", type=int, default=8)
    ap.add_argument("--modes", default="ddp,peer")
    args = ap.parse_args()
    from b200timm import loss as BL, optim as BO
    from b200timm.grad_sync import PeerGradSync
    from b200timm.train_step import TrainStep, prepare_ddp_env, wrap_ddp
    from b200timm.vision_transformer import create_model
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    prepare_ddp_env()
    dist.init_process_group("nccl", device_id=dev)
    side = torch.cuda.Stream()
    torch.manual_seed(7)                                       # same weights and same global batch on every rank
    base = create_model(args.model).to(dev).to(torch.bfloat16).train()
    G = args.per_rank * world
    x = torch.randn(G, 3, 224, 224, device=dev).bfloat16()
    y = torch.randint(0, 1000, (G,), device=dev)
    sl = slice(rank * args.per_rank, (rank + 1) * args.per_rank)
    xs, ys = x[sl].contiguous(), y[sl].contiguous()
    loss_fn = BL.CrossEntropyLoss()
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        loss_fn(base(x), y).backward()                         # single-rank reference on the global batch
    torch.cuda.synchronize()
    ref = {n: p.grad.float().clone() for n, p in base.named_parameters()}
    base.zero_grad(set_to_none=True)
    scale = max(g.norm().item() for g in ref.values())
    res = {"world_size": world, "model": args.model, "per_rank_batch": args.per_rank,
           "tolerance": "relative L2 <= 3e-2 per tensor (mean of N bf16 shard gradients vs one bf16 global-batch gradient); "
                        "spread across ranks == 0; graphed vs eager: parameters within 2e-2, losses within 2e-2"}
    ok = True
    for mode in args.modes.split(","):
        model = copy.deepcopy(base)
        sync, net = None, model
        with torch.cuda.stream(side):
            if mode == "ddp":
                net = wrap_ddp(model, dev, bucket_cap_mb=1024)
            else:
                sync = PeerGradSync(model)
            opt = BO.SGD(model.parameters(), lr=0.0, momentum=0.0)      # lr 0: the step leaves the weights alone
            step = TrainStep(net, loss_fn, opt, xs, ys, graph=False, grad_sync=sync)
            step(xs, ys)
        torch.cuda.synchronize()
        worst, worst_name = 0.0, ""
        for n, p in model.named_parameters():
            r = ref[n]
            if r.norm().item() > 1e-3 * scale:
                e = (p.grad.float() - r).norm().item() / r.norm().item()
                if e > worst:
                    worst, worst_name = e, n
        flat = torch.cat([p.grad.float().flatten() for p in model.parameters()])
        mx, mn = flat.clone(), flat.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        spread = (mx - mn).abs().max().item()
        # graphed == eager over three real steps
        state = copy.deepcopy(model.state_dict())
        outs = []
        for graph in (False, True):
            model.load_state_dict(state)
            with torch.cuda.stream(side):
                opt = BO.SGD(model.parameters(), lr=1e-3, momentum=0.9, nesterov=True)
                st = TrainStep(net, loss_fn, opt, xs, ys, graph=graph, grad_sync=sync)
                if not graph:
                    for _ in range(st.warmup):                 # the graphed run warms up with the same eager steps
                        st(xs, ys)
                losses = [float(st(xs, ys).detach()) for _ in range(3)]
            torch.cuda.synchronize()
            outs.append((losses, torch.cat([p.detach().float().flatten() for p in model.parameters()]), st.launches))
        pdiff = (outs[0][1] - outs[1][1]).abs().max().item()
        ldiff = max(abs(a - b) for a, b in zip(outs[0][0], outs[1][0]))
        good = worst <= 3e-2 and spread == 0.0 and pdiff <= 2e-2 and ldiff <= 2e-2
        ok = ok and good
        res[mode] = {"vs_global_batch_worst_rel_l2": worst, "worst_param": worst_name, "max_grad_spread_across_ranks": spread,
                     "graphed_vs_eager_max_param_diff": pdiff, "graphed_vs_eager_max_loss_diff": ldiff,
                     "eager_losses": outs[0][0], "graphed_losses": outs[1][0], "graph_launches_per_step": outs[1][2], "ok": good}
        del st, opt, net, sync, model
        torch.cuda.empty_cache()
        dist.barrier()
    res["ok"] = bool(ok)
    if rank == 0:
        print(json.dumps(res))
        if args.out:
            with open(args.out, "w") as f:
                json.dump(res, f, indent=1)
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        raise SystemExit(1)


if __name__ == "__main__":
    main()
