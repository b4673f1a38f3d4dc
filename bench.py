#!/usr/bin/env python
"""Benchmark of the B200 hot path:  images/sec (infer + train), vit_base_patch16_224, bf16, batch 256 per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--mode both|train|infer] [--impl ours|reference]

The headline `value` is the TRAINING step (BASELINE.json configs[3]): zero_grad -> forward -> cross entropy -> backward
-> Nesterov SGD step, data parallel over N GPUs.  The gradient exchange (the only multi-GPU step on this path)
is b200timm.grad_sync by default -- every block's gradients are pushed to all ranks by multimem.red.add over the NVLink
multicast address from the kernel that finalises them -- or torch DDP's NCCL all-reduce with --dp ddp (also the automatic
fall-back when the node has no multicast).  The inference half of the metric (configs[1]) rides in the same JSON
line as the sibling record `infer` (replicas, no data-path collective).  N > 1 is launched by the driver through
torch.distributed.run (one rank per GPU); rank 0 prints ONE JSON line.  At N = 1 the line also carries

  cpu_baseline          the unmodified reference's benchmark.py --device cpu on the host cores (oracle/_ref),
  plugin_benchmark_py   the unmodified reference's benchmark.py run through the plugin on this GPU (the drop-in boundary),
  gpu_eager_baseline    the same script without the plugin: timm's own eager PyTorch path on this GPU (informational).

`--impl reference` times the reference's own CPU implementation (oracle/_ref/scripts/benchmark.py --device cpu, the
unmodified script and package installed by oracle/install_ref.py) on a bounded sample of the same workload.
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


def recorded_traffic(train=False):
    """roofline.traffic: DRAM bytes per GEMM launch (mean over the block's GEMMs: 4 forward ones for inference, all 12 of
    forward + backward for training) from the committed ncu --set full captures -- bench.py cannot run ncu itself, so the
    number is the recorded one and says where it came from."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        g = t["gemm_bf16_kernel_train" if train else "gemm_bf16_kernel"]
        return {"traffic": g["mean_traffic_mb"] * 1e6,
                "traffic_unit": "bytes/launch (mean of the block's " + ("12 forward + backward GEMMs)" if train else "qkv, proj, fc1, fc2)"),
                "traffic_algorithmic": g["mean_algorithmic_mb"] * 1e6, "traffic_source": g.get("source", t["source"])}
    except (OSError, KeyError, ValueError):
        return {"traffic": None}


# ------------------------------------------------------------------------------------------------ the reference on the CPU
def cpu_threads():
    """Threads of the CPU arm = the cores this container may actually use: min(affinity mask, cgroup CPU quota).  The GPU
    boxes show 128 CPUs in the mask under a quota of 16 (cpu.max = 1600000 100000; round 1's run-time calibration found
    the same 16 the slow way, and 128 spinning OpenMP threads ran 40x slower) -- deterministic, so per-N ratios compare."""
    n = len(os.sched_getaffinity(0))
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = min(n, max(1, int(quota) // int(period)))
    except (OSError, ValueError):
        pass
    return max(1, n)


def ref_dir():
    from oracle import install_ref
    return install_ref.TARGET if install_ref.installed() else None


def reference_benchmark_cpu(iters, warm, batch=16, bench="both", timeout=1500):
    """The UNMODIFIED reference script (oracle/_ref/scripts/benchmark.py, timm from oracle/_ref) on the host CPU in a fresh
    interpreter: `benchmark.py --model vit_base_patch16_224 --bench both --device cpu --precision bfloat16 -b 16`
    (SURVEY.md 8d).  Returns its result dict (infer_samples_per_sec, train_samples_per_sec, ...) or None."""
    ref = ref_dir()
    if ref is None:
        return None
    out = os.path.join("/tmp", f"b200_ref_bench_{os.getpid()}.json")
    env = dict(os.environ, PYTHONPATH=ref, OMP_NUM_THREADS=str(cpu_threads()), MKL_NUM_THREADS=str(cpu_threads()),
               CUDA_VISIBLE_DEVICES="")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    cmd = [sys.executable, os.path.join(ref, "scripts", "benchmark.py"), "--model", MODEL, "--bench", bench, "--device", "cpu",
           "--precision", "bfloat16", "-b", str(batch), "--num-warm-iter", str(max(1, warm)),
           "--num-bench-iter", str(max(5, iters)), "--results-file", out, "--results-format", "json"]
    try:
        subprocess.run(cmd, env=env, cwd="/tmp", check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, timeout=timeout)
        with open(out) as f:
            res = json.load(f)
        os.remove(out)
        return res[0] if isinstance(res, list) else res
    except (subprocess.SubprocessError, OSError, ValueError):
        return None


def cpu_port_throughput(batch, steps, warmup, dtype=torch.bfloat16):
    """Fallback when oracle/_ref is absent: oracle.vit_oracle.vit_forward_eager issues the same torch CPU library calls
    timm's eager forward does.  img/s (inference only)."""
    from oracle import vit_oracle as O
    torch.set_num_threads(cpu_threads())
    cfg = dict(O.VIT_CONFIGS[MODEL])
    sd = {k: v.to(dtype) for k, v in O.random_state_dict(cfg, seed=42).items()}
    img = torch.randn(batch, 3, 224, 224, generator=torch.Generator().manual_seed(0)).to(dtype)
    with torch.no_grad():
        for _ in range(warmup):
            O.vit_forward_eager(img, sd, cfg)
        t0 = time.perf_counter()
        for _ in range(steps):
            O.vit_forward_eager(img, sd, cfg)
        dt = time.perf_counter() - t0
    return batch * steps / dt, dt / steps


def cpu_baseline_record(iters, warm):
    batch = 1 if os.environ.get("B200_REF_TINY") else 16      # B200_REF_TINY: unit tests of the harness only
    res = reference_benchmark_cpu(iters, warm, batch=batch)
    if res is not None:
        return {"value": res["train_samples_per_sec"], "unit": "img/s", "cores": cpu_threads(), "kind": "reference",
                "sample": f"unmodified benchmark.py --bench both --device cpu --precision bfloat16 -b {batch}, "
                          f"{max(5, iters)} timed iterations each (oracle/_ref, {cpu_threads()} threads): value = TRAIN img/s",
                "train_ms_per_step": res["train_step_time"], "infer_value": res["infer_samples_per_sec"],
                "infer_ms_per_step": res["infer_step_time"]}
    ips, spp = cpu_port_throughput(batch, max(2, min(iters, 10)), 1)
    return {"value": round(ips, 2), "unit": "img/s", "cores": cpu_threads(), "kind": "port",
            "sample": f"oracle/_ref not installed: INFERENCE only, oracle.vit_forward_eager, batch {batch}, bf16"}


def run_reference(args, rank, world):
    if rank != 0:
        return
    rec = cpu_baseline_record(args.steps, min(args.warmup, 2))
    ms = rec.get("train_ms_per_step")
    line = {
        "impl": "reference", "metric": METRIC, "value": rec["value"], "unit": "img/s", "n_gpus": args.gpus,
        "steps": max(5, args.steps), "warmup": max(1, min(args.warmup, 2)), "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{MODEL} bf16 TRAIN step (configs[3]) on the host CPU, bounded sample: batch 16 / step; the "
                               "inference half (configs[1]) in `infer`", "flush": "n/a (CPU)",
                   "parallelism": "host CPU only (rank 0), no GPU"},
        "cpu_baseline": rec,
        "e2e": {"value": rec["value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if "infer_value" in rec:
        line["infer"] = {"value": rec["infer_value"], "unit": "img/s", "ms_per_step": rec["infer_ms_per_step"],
                         "e2e": {"value": rec["infer_value"], "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- the reference's script on the GPU
def run_benchmark_py(extra, with_plugin, timeout=600):
    """`benchmark.py --model vit_base_patch16_224 -b 256 --precision bfloat16 ...` -- the reference's UNMODIFIED script
    (oracle/_ref) in a fresh interpreter on this GPU, either through `python -m b200timm.run` (plugin installed, strict:
    any hot-path call the kernels do not cover raises instead of deferring to timm) or plainly (timm's eager path)."""
    ref = ref_dir()
    if ref is None:
        return None
    out = os.path.join("/tmp", f"b200_gpu_bench_{os.getpid()}.json")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ref, os.path.join(ROOT, "pytorch-image-models_b200")]),
               B200TIMM_STRICT="1", B200TIMM_REPORT=out + ".plugin")
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT", "OMP_NUM_THREADS"):
        env.pop(k, None)
    script = os.path.join(ref, "scripts", "benchmark.py")
    cmd = [sys.executable] + (["-m", "b200timm.run"] if with_plugin else []) + [
        script, "--model", MODEL, "-b", str(BATCH), "--precision", "bfloat16", "--num-warm-iter", "3", "--num-bench-iter", "10",
        "--no-retry", "--results-file", out, "--results-format", "json"] + extra
    try:
        subprocess.run(cmd, env=env, cwd="/tmp", check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, timeout=timeout)
        with open(out) as f:
            res = json.load(f)
        os.remove(out)
        res = res[0] if isinstance(res, list) else res
        if with_plugin and os.path.exists(out + ".plugin"):
            with open(out + ".plugin") as f:
                res["plugin"] = json.load(f)
            os.remove(out + ".plugin")
        return res
    except subprocess.CalledProcessError as e:
        return {"error": (e.stderr or b"").decode(errors="replace")[-400:]}
    except (subprocess.SubprocessError, OSError, ValueError) as e:
        return {"error": repr(e)[:400]}


def gpu_script_records():
    """plugin_benchmark_py / gpu_eager_baseline: benchmark.py --bench both, with and without the plugin."""
    recs = {}
    for key, plug in (("plugin_benchmark_py", True), ("gpu_eager_baseline", False)):
        r = run_benchmark_py(["--bench", "both"], plug)
        if r is None:
            recs[key] = {"unavailable": "oracle/_ref not installed"}
        elif "error" in r:
            recs[key] = r
        else:
            recs[key] = {"infer_img_s": r.get("infer_samples_per_sec"), "infer_ms_per_step": r.get("infer_step_time"),
                         "train_img_s": r.get("train_samples_per_sec"), "train_ms_per_step": r.get("train_step_time"),
                         "cmd": ("python -m b200timm.run " if plug else "python ") + "oracle/_ref/scripts/benchmark.py --model "
                                f"{MODEL} --bench both -b {BATCH} --precision bfloat16 --num-warm-iter 3 --num-bench-iter 10",
                         "note": ("unmodified script, plugin installed with B200TIMM_STRICT=1 (torch's SGD / CE as the script "
                                  "builds them; the script synchronises every step)") if plug else
                                 "unmodified script and timm on the same GPU: PyTorch eager (cuBLAS / SDPA / ATen)"}
            if "plugin" in r:
                recs[key]["plugin"] = r["plugin"]
    return recs


def other_config_records():
    """BASELINE.json configs[2] (convnext_base infer, batch 256) and configs[4] (NaFlex ViT-L/14 variable-sequence infer,
    batch 64) measured by tools/bench_model.py (device-resident, CUDA events, two alternating input batches) in fresh
    interpreters: sub-records of the N = 1 line, so that the driver's run carries them."""
    recs = {}
    env = dict(os.environ)
    for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    for key, argv, flops in (("convnext_base_infer", ["convnext_base", "--batch", "256"], 30.71e9),
                             ("naflex_vit_large_patch14_infer", ["naflex:vit_large_patch14_clip_336", "--batch", "64"], None)):
        try:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "bench_model.py")] + argv + ["--steps", "10", "--warmup", "3"],
                               env=env, cwd=ROOT, check=True, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300, text=True)
            d = json.loads([l for l in r.stdout.splitlines() if l.startswith("{")][-1])
            rec = {"value": d["img_per_s"], "unit": "img/s", "ms_per_step": d["ms_per_step"], "batch": d["batch"],
                   "launches_per_step": d.get("launches_per_step"), "cmd": "python tools/bench_model.py " + " ".join(argv)}
            if flops:
                tf = flops * d["img_per_s"] / 1e12
                rec["whole_step_tflops"] = round(tf, 1)
                rec["whole_step_frac"] = round(tf / measured_peaks()["bf16_tflops_sustained"], 4)
            for k in ("valid_tokens_per_img", "tokens_per_s"):
                if k in d:
                    rec[k] = d[k]
            recs[key] = rec
        except (subprocess.SubprocessError, OSError, ValueError, IndexError, KeyError) as e:
            recs[key] = {"error": repr(e)[:300]}
    return recs


# --------------------------------------------------------------------------------------------------------- inference record
def bench_infer(args, rank, world, dev, clocks=None):
    import torch.distributed as dist
    from b200timm import ops
    from b200timm.vision_transformer import create_model
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

    if args.no_graph:
        runners = [lambda i=i: model(dev_in[i]) for i in range(2)]
    else:
        from b200timm.graph import GraphedForward
        gf = GraphedForward(model, dev_in[0])
        runners = [gf.capture_on(dev_in[0]), gf.capture_on(dev_in[1])]
    with torch.no_grad():
        ops.reset_counters()
        model(dev_in[0])                      # one eager step: counts our kernel launches per step
        launches_per_step = ops.launch_count()
        for i in range(args.warmup):
            runners[i & 1]()
        barrier()
        if clocks is not None:
            clocks.start()
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        start.record()
        for i in range(args.steps):
            runners[i & 1]()
        end.record()
        barrier()
        ms = start.elapsed_time(end)
        clock_info = clocks.stop() if clocks is not None else None
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
    total_imgs = B * args.steps * world
    value = total_imgs / (ms / 1e3)
    del runners, model
    return {
        "value": round(value, 1), "unit": "img/s", "ms_per_step": round(ms / args.steps, 4),
        "workload": f"{MODEL} bf16 infer batch={B}/GPU (BASELINE.json configs[1]), random-init weights, torch.randn inputs as "
                    "benchmark.py:286-290",
        "parallelism": f"replicas x{world} (no data-path collective)",
        "launch": "eager" if args.no_graph else "CUDA graph replay (one graph per resident input buffer)",
        "e2e": {"value": round(total_imgs / (e2e_ms / 1e3), 1), "unit": "img/s", "h2d_bytes_per_step": host_in[0].numel() * 2,
                "d2h_bytes_per_step": host_out.numel() * 2,
                "note": "pinned bf16 host batch -> H2D on a copy stream (double buffered) -> forward (CUDA-graph replay "
                        "of the model's kernels) -> logits D2H, every step inside the timed region"},
        "gpu_launches": launches_per_step * args.steps, "launches_per_step": launches_per_step,
        "gemm": {"flops": gemm_flops, "ms": gemm_ms, "n": gemm_n},
        "whole_step_tflops": round(FWD_FLOPS_PER_IMG * value / world / 1e12, 1),
        "clocks": clock_info,
    }


def roofline_record(rec, peaks, train=False):
    g = rec.pop("gemm")
    tf = g["flops"] / (g["ms"] * 1e9) if g["ms"] > 0 else 0.0
    sustained, burst = peaks["bf16_tflops_sustained"], peaks["bf16_tflops"]
    return {"bound": "tensor", "kernel": "gemm_bf16_kernel (CTA pairs, 256x256 tiles; every Linear's forward / dgrad / wgrad GEMM)",
            "achieved": round(tf, 1), "peak": sustained, "unit": "TFLOP/s", "frac": round(tf / sustained, 4),
            "peak_burst": burst, "frac_burst": round(tf / burst, 4), **recorded_traffic(train),
            "launches_timed": g["n"], "timing": "CUDA events around every GEMM launch over eagerly launched steps",
            "peak_source": peaks["source"] + " (cuBLAS bf16: sustained for launches timed inside a step, burst beside it)",
            "whole_step_tflops": rec.pop("whole_step_tflops"),
            "whole_step_frac": round(rec.get("_wsf", 0.0), 4)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--mode", choices=["both", "infer", "train"], default="both")
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-script-records", action="store_true", help="skip plugin_benchmark_py / gpu_eager_baseline")
    ap.add_argument("--no-graph", action="store_true", help="launch kernels eagerly instead of replaying CUDA graphs")
    ap.add_argument("--dp", choices=["ddp", "peer"], default=os.environ.get("B200_DP", "peer"),
                    help="N > 1 gradient exchange: torch DDP + NCCL all-reduce, or b200timm.grad_sync (multimem pushes fused "
                         "into the gradient-finalising kernels)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    from b200timm import train_bench
    from b200timm.train_step import prepare_ddp_env
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU path for the product")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        prepare_ddp_env()
        dist.init_process_group("nccl", device_id=dev)

    peaks = measured_peaks()
    train = infer = None
    if args.mode in ("both", "train"):
        train = train_bench.run(args, rank, world, dev, MODEL, ClockSampler(local_rank))
        torch.cuda.empty_cache()
    if args.mode in ("both", "infer"):
        infer = bench_infer(args, rank, world, dev, ClockSampler(local_rank) if rank == 0 and train is None else None)
        torch.cuda.empty_cache()

    if rank == 0:
        head = train if train is not None else infer
        for rec, flops in ((train, 3 * FWD_FLOPS_PER_IMG), (infer, FWD_FLOPS_PER_IMG)):
            if rec is not None:
                rec["_wsf"] = rec["whole_step_tflops"] / peaks["bf16_tflops_sustained"]
                rec["roofline"] = roofline_record(rec, peaks, train=rec is train)
                rec.pop("_wsf")
        clocks = head.pop("clocks")
        if infer is not None:
            infer.pop("clocks", None)
        workload = head["workload"]
        if train is not None and infer is not None:
            workload += "  ||  sibling record `infer`: " + infer["workload"]
        line = {
            "metric": METRIC, "value": head["value"], "unit": "img/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload, "global_batch": args.batch * world, "parallelism": head["parallelism"],
                       "launch": head["launch"],
                       "flush": "inputs alternate between two 77 MB batches (> 126 MB L2 together); a training step streams "
                                "~25 GB of activations / gradients, an inference step ~8 GB"},
            "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "launches_per_step": head["launches_per_step"],
            "roofline": head["roofline"], "clocks": clocks,
        }
        if "host_enqueue_ms_per_step" in head:
            line["host_enqueue_ms_per_step"] = head["host_enqueue_ms_per_step"]
        if head.get("clocks_per_rank"):
            line["clocks_per_rank"] = head["clocks_per_rank"]
        if train is not None and infer is not None:
            line["infer"] = {k: infer[k] for k in ("value", "unit", "ms_per_step", "parallelism", "launch", "e2e", "gpu_launches",
                                                   "launches_per_step", "roofline")}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_record(5, 1)
        if world == 1 and not args.no_script_records:
            line.update(gpu_script_records())
            line["other_configs"] = other_config_records()
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
