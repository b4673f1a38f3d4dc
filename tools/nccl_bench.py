#!/usr/bin/env python
"""All-reduce microbenchmark on the box's GPUs (what the DDP gradient exchange of ViT-B costs in isolation):
    torchrun --nproc-per-node N tools/nccl_bench.py
Prints one JSON line: time of one 173 MB bf16 all-reduce, of the same bytes in 28 MB buckets, alg / bus bandwidth."""
import json
import os

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = 86567656
    flat = torch.randn(n, device="cuda").bfloat16()
    chunks = list(flat.split(14175744))
    res = {"world": world}
    for name, bufs in (("one_173MB", [flat]), ("buckets_28MB", chunks)):
        for _ in range(3):
            for b in bufs:
                dist.all_reduce(b)
        torch.cuda.synchronize()
        dist.barrier()
        a, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            for b in bufs:
                dist.all_reduce(b)
        b_.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b_) / 10
        t = torch.tensor([ms], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
        algbw = n * 2 / ms / 1e6
        res[name] = {"ms": round(ms, 4), "algbw_GBs": round(algbw, 1), "busbw_GBs": round(algbw * 2 * (world - 1) / world, 1)}
    if rank == 0:
        print(json.dumps(res))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
