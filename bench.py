#!/usr/bin/env python
"""Benchmark of the B200 hot path:  images/sec, vit_base_patch16_224, bf16, batch 256 per GPU
(BASELINE.json configs[1]; --mode train times the training step of configs[3]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode infer|train] [--impl ours|reference]

N > 1 is launched by the driver through torch.distributed.run (one rank per GPU).  Inference has no data-path
collective (replicas only): every rank runs its own batch; training is DDP with the NCCL
all-reduce of the gradient buckets.  One JSON line is printed by rank 0.

`--impl reference` times the reference's CPU implementation of the same path (the oracle port, since timm
itself is pure Python + PyTorch and cannot travel to the GPU box) on the host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "pytorch-image-models_b200"))
sys.path.insert(0, ROOT)

MODEL = "vit_base_patch16_224"
BATCH = 256
METRIC = "images/sec (infer+train, bf16) vit_base_patch16_224 at 1/2/4/8 B200"
FWD_FLOPS_PER_IMG = 35.13e9  # SURVEY.md 8d / BASELINE.md (2 * 17.564 GMAC)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return {"bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                "hbm_gbs": d["hbm_gbs"], "source": "measured"}
    # fallback stated in /opt/skills/guides/B200_PROFILING.md
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def recorded_traffic():
    """roofline.traffic: DRAM bytes per GEMM launch (mean over the four block GEMMs) from the committed ncu --set full
    capture -- bench.py cannot run ncu itself, so the number is the recorded one and says where it came from."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        g = t["gemm_bf16_kernel"]
        return {"traffic": g["mean_traffic_mb"] * 1e6, "traffic_unit": "bytes/launch (mean of qkv, proj, fc1, fc2)",
                "traffic_algorithmic": g["mean_algorithmic_mb"] * 1e6, "traffic_source": t["source"]}
    except (OSError, KeyError, ValueError):
        return {"traffic": None}


_THREADS = {}


def pick_cpu_threads(dtype):
    """Thread count for the CPU arm: the fastest of a few candidates between the cgroup CPU quota and the affinity mask,
    each timed on one small forward.  (A box whose mask shows 128 CPUs but whose quota is far smaller ran the port 40x
    slower with 128 spinning OpenMP threads than with the right count -- the reference arm should not lose that way.)"""
    if dtype in _THREADS:
        return _THREADS[dtype]
    from oracle import vit_oracle as O
    aff = max(1, len(os.sched_getaffinity(0)))
    cand = {aff, max(1, aff // 2), max(1, aff // 4)}
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            cand.add(max(1, min(aff, -(-int(quota) // int(period)))))
    except (OSError, ValueError):
        pass
    cand |= {t for t in (64, 32, 16, 8) if t <= aff}
    cfg = dict(O.VIT_CONFIGS[MODEL])
    sd = {k: v.to(dtype) for k, v in O.random_state_dict(cfg, seed=42).items()}
    img = torch.randn(4, 3, 224, 224).to(dtype)
    best, best_t = aff, float("inf")
    with torch.no_grad():
        for t in sorted(cand, reverse=True):
            torch.set_num_threads(t)
            O.vit_forward_eager(img[:1], sd, cfg)
            t0 = time.perf_counter()
            O.vit_forward_eager(img, sd, cfg)
            dt = time.perf_counter() - t0
            if dt < best_t:
                best, best_t = t, dt
    _THREADS[dtype] = best
    return best


def cpu_port_throughput(batch, steps, warmup, dtype=torch.bfloat16, threads=None):
    """The reference path on the host CPU: oracle.vit_oracle.vit_forward_eager issues the same torch CPU
    library calls timm's eager forward does (F.conv2d / layer_norm / linear / sdpa / gelu).  img/s."""
    from oracle import vit_oracle as O
    # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers; the CPU arm uses every core it may run on
    torch.set_num_threads(threads or pick_cpu_threads(dtype))
    cfg = dict(O.VIT_CONFIGS[MODEL])
    sd = {k: v.to(dtype) for k, v in O.random_state_dict(cfg, seed=42).items()}
    g = torch.Generator().manual_seed(0)
    img = torch.randn(batch, 3, 224, 224, generator=g).to(dtype)
    with torch.no_grad():
        for _ in range(warmup):
            O.vit_forward_eager(img, sd, cfg)
        t0 = time.perf_counter()
        for _ in range(steps):
            O.vit_forward_eager(img, sd, cfg)
        dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps, torch.get_num_threads()


def run_reference(args, rank, world):
    if rank != 0:
        return
    batch = 16  # CPU throughput is batch-insensitive; one step = one bounded sample of the workload
    if os.environ.get("B200_REF_TINY"):   # unit tests of the harness only
        batch = 1
    ips32, _, _ = cpu_port_throughput(batch, 2, 1, dtype=torch.float32)
    ips, spp, cores = cpu_port_throughput(batch, args.steps, min(args.warmup, 2), dtype=torch.bfloat16)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(ips, 2), "unit": "img/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 2), "ms_per_step": round(spp * 1e3, 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{MODEL} bf16 infer (configs[1]), CPU sample batch={batch}/step, random-init weights",
                   "flush": "n/a (CPU)"},
        "cpu_baseline": {"value": round(ips, 2), "unit": "img/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} steps x batch {batch}, bf16, oracle.vit_forward_eager (timm's eager call "
                                   f"sequence on torch CPU kernels); same port in fp32: {ips32:.2f} img/s"},
        "e2e": {"value": round(ips, 2), "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--mode", choices=["infer", "train"], default="infer")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    from b200timm import ops
    from b200timm.vision_transformer import create_model
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU path for the product")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    if args.mode == "train":
        from b200timm import train_bench
        line = train_bench.run(args, rank, world, dev)
        if rank == 0:
            print(json.dumps(line), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    torch.manual_seed(42 + rank)
    model = create_model(MODEL).to(dev).to(torch.bfloat16).eval()
    B = args.batch
    # two device-resident input batches (2 x 77 MB > 126 MB L2) alternate between steps; the ~8 GB of
    # activations each step writes evict everything else
    dev_in = [torch.randn(B, 3, 224, 224, device=dev).to(torch.bfloat16) for _ in range(2)]
    host_in = [t.cpu().pin_memory() for t in dev_in]
    host_out = torch.empty(B, 1000, dtype=torch.bfloat16).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # The public inference entry: the model's forward captured into CUDA graphs (b200timm.graph), one per resident
    # input buffer for the device-resident measurement, one on a staging buffer for the host-buffer (e2e) path.
    if args.no_graph:
        runners = [lambda i=i: model(dev_in[i]) for i in range(2)]
        staged = model
    else:
        from b200timm.graph import GraphedForward
        gf = GraphedForward(model, dev_in[0])
        runners = [gf.capture_on(dev_in[0]), gf.capture_on(dev_in[1])]
        staged = None
    launches_per_step = 0
    with torch.no_grad():
        ops.reset_counters()
        model(dev_in[0])                      # one eager step: counts our kernel launches per step
        launches_per_step = ops.launch_count()

    # ---------------------------------------------------------------- device-resident throughput (`value`)
    with torch.no_grad():
        for i in range(args.warmup):
            runners[i & 1]()
        barrier()
        clocks = ClockSampler(local_rank)
        if rank == 0:
            clocks.start()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        start.record()
        for i in range(args.steps):
            runners[i & 1]()
        end.record()
        barrier()
        ms = start.elapsed_time(end)
        launches = launches_per_step * args.steps
        clock_info = clocks.stop() if rank == 0 else None
        # roofline.achieved of the dominant kernel: the same K steps again, eagerly, with CUDA events bracketing
        # every b200_gemm launch on the launching stream (events cannot be recorded inside a replayed graph)
        ops.enable_gemm_timing(True)
        for i in range(args.steps):
            model(dev_in[i & 1])
        gemm_flops, gemm_ms, gemm_n = ops.gemm_timing_summary()
        ops.enable_gemm_timing(False)

        # ------------------------------------------------------------ end to end with host buffers (`e2e`)
        copy_stream = torch.cuda.Stream(device=dev)
        compute = torch.cuda.current_stream()
        stage = [torch.empty_like(dev_in[0]) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        consumed = [torch.cuda.Event() for _ in range(2)]

        if args.no_graph:
            e2e_fwd = [lambda i=i: model(stage[i]) for i in range(2)]
        else:
            e2e_fwd = [gf.capture_on(stage[0]), gf.capture_on(stage[1])]

        def e2e_loop(n):
            with torch.cuda.stream(copy_stream):
                stage[0].copy_(host_in[0], non_blocking=True)
                ready[0].record(copy_stream)
            for i in range(n):
                cur, nxt = i & 1, (i + 1) & 1
                if i + 1 < n:  # upload batch i+1 on the copy stream while batch i computes
                    with torch.cuda.stream(copy_stream):
                        if i >= 1:
                            copy_stream.wait_event(consumed[nxt])
                        stage[nxt].copy_(host_in[nxt], non_blocking=True)
                        ready[nxt].record(copy_stream)
                compute.wait_event(ready[cur])
                logits = e2e_fwd[cur]()
                consumed[cur].record(compute)
                host_out.copy_(logits, non_blocking=True)
            compute.synchronize()

        e2e_loop(2)
        barrier()
        t0 = torch.cuda.Event(enable_timing=True)
        t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        e2e_loop(args.steps)
        t1.record()
        barrier()
        e2e_ms = t0.elapsed_time(t1)

    times = torch.tensor([ms, e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, e2e_ms = times.tolist()

    if rank == 0:
        peaks = measured_peaks()
        total_imgs = B * args.steps * world
        value = total_imgs / (ms / 1e3)
        e2e = total_imgs / (e2e_ms / 1e3)
        gemm_tflops = gemm_flops / (gemm_ms * 1e9) if gemm_ms > 0 else 0.0
        peak = peaks["bf16_tflops_sustained"]
        line = {
            "metric": METRIC, "value": round(value, 1), "unit": "img/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{MODEL} bf16 infer batch={B}/GPU (BASELINE.json configs[1]), random-init weights, "
                                   "torch.randn inputs as benchmark.py:286-290",
                       "global_batch": B * world, "parallelism": f"replicas x{world} (no data-path collective)",
                       "launch": "eager" if args.no_graph else "CUDA graph replay (one graph per resident input buffer)",
                       "flush": "inputs alternate between two 77 MB batches (> 126 MB L2 together); each step "
                                "streams ~8 GB of activations"},
            "e2e": {"value": round(e2e, 1), "unit": "img/s", "h2d_bytes_per_step": host_in[0].numel() * 2,
                    "d2h_bytes_per_step": host_out.numel() * 2,
                    "note": "pinned bf16 host batch -> H2D on a copy stream (double buffered) -> forward (CUDA-graph replay "
                            "of the model's kernels) -> logits D2H, every step inside the timed region"},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "gemm_bf16_kernel (CTA pairs, 256x256 tiles; qkv/proj/fc1/fc2/head GEMMs)",
                         "achieved": round(gemm_tflops, 1), "peak": peak, "unit": "TFLOP/s",
                         "frac": round(gemm_tflops / peak, 4), **recorded_traffic(),
                         "launches_timed": gemm_n, "timing": "CUDA events around every GEMM launch over K eagerly launched steps",
                         "peak_source": peaks["source"] + " (sustained cuBLAS bf16)",
                         "whole_step_tflops": round(FWD_FLOPS_PER_IMG * value / world / 1e12, 1)},
            "clocks": clock_info,
        }
        if world == 1 and not args.no_cpu_baseline:
            ips, spp, cores = cpu_port_throughput(16, 10, 2)
            line["cpu_baseline"] = {"value": round(ips, 2), "unit": "img/s", "cores": cores, "kind": "port",
                                    "sample": "10 steps x batch 16, bf16, oracle.vit_forward_eager (timm's eager call "
                                              "sequence on torch CPU kernels) on the host cores"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
