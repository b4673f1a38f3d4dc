"""gpurun_out/{launches_train,launches_infer}.csv + prof_train_block_raw.csv (tools/gpu_ncu_step.sh) -> profiles/r02_*.md and
profiles/traffic.json (the per-launch DRAM traffic bench.py copies into roofline.traffic)."""
import collections
import csv
import json
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, SRC = os.path.join(ROOT, "profiles"), os.path.join(ROOT, "gpurun_out")


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\((?:int|bool|unsigned int)\)", "", name)      # template-argument casts: (int)-1, (bool)1
    name = re.sub(r"\(.*", "", name)
    name = name.replace("b200::", "").replace("(anonymous namespace)::", "")
    return name[:96]


def launch_list(src, dst, title, cmd):
    with open(os.path.join(SRC, src)) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg, tot = collections.OrderedDict(), 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u.startswith("n") else (v * 1e3 if u.startswith("m") else v)
        key = (short(row["Kernel Name"]), row["Grid Size"], row["Block Size"])
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    ours = sum(t for (k, _, _), (n, t) in agg.items() if not k.startswith("at::") and "vectorized" not in k and "elementwise" not in k)
    with open(os.path.join(OUT, dst), "w") as f:
        f.write(f"# {title}\n\n`{cmd}`\n\nTimes are cold-cache and serialised (ncu replays every kernel): compare SHARES, not absolutes; a number "
                f"printed under ncu is never a bench value.\n\ntotal {tot:.1f} us over {sum(a[0] for a in agg.values())} launches; "
                f"libb200timm kernels: {100 * ours / tot:.1f} % of the time (the rest: torch memsets / copies / the autograd glue)\n\n"
                "| us | share | launches | avg us | kernel (grid, block) |\n|---|---|---|---|---|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {t:.1f} | {100 * t / tot:.1f}% | {n} | {t / n:.1f} | `{k[0]}` {k[1]} {k[2]} |\n")
    return agg, tot


def unit_scale(u):
    return {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)


def block_table(src="prof_train_block_raw.csv", dst="r02_train_block_kernels.md", header=None):
    rows = list(csv.reader(open(os.path.join(SRC, src))))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, k, default=0.0):
        try:
            return float(r[col[k]].replace(",", ""))
        except (KeyError, ValueError, IndexError):
            return default
    out, traffic = [], {}
    for r in rows[2:]:
        name = short(r[col["Kernel Name"]])
        dur = val(r, "gpu__time_duration.sum")
        du = units[col["gpu__time_duration.sum"]]
        dur_us = dur / 1e3 if du.startswith("n") else (dur * 1e3 if du.startswith("m") else dur)
        rd = val(r, "dram__bytes_read.sum") * unit_scale(units[col["dram__bytes_read.sum"]])
        wr = val(r, "dram__bytes_write.sum") * unit_scale(units[col["dram__bytes_write.sum"]])
        out.append((name, r[col["Grid Size"]] if "Grid Size" in col else "", dur_us, (rd + wr) / 1e6,
                    (rd + wr) / max(dur_us, 1e-9) / 1e3,
                    val(r, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                    val(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                    val(r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                    val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active", val(r, "sm__inst_issued.avg.pct_of_peak_sustained_active")),
                    int(val(r, "launch__registers_per_thread"))))
    peak = 6569.6
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except (OSError, KeyError, ValueError):
        pass
    with open(os.path.join(OUT, dst), "w") as f:
        f.write(header if header is not None else "# r02: every kernel of ONE ViT-B/16 training step at depth 1 (batch 256, bf16) under `ncu --set full --clock-control none`\n\n"
                "`ncu --set full --clock-control none --profile-from-start off python tools/ncu_train_step.py --depth 1` (tools/gpu_ncu_step.sh; "
                "the report is converted to CSV on the GPU box, `gpurun_out/prof_train_block_raw.csv`).  Launch order = forward "
                "(patch embed, block, head), loss, backward, optimizer.  DRAM GB/s = (dram read + write bytes) / duration against the "
                f"measured HBM peak {peak:.0f} GB/s.  Captures are serialised cold-cache replays: use them for pipe utilisation, traffic "
                "and stall shares; steady-state timings come from CUDA events (bench.py, tools/*).\n\n"
"")
        f.write("| # | kernel | grid | us | DRAM MB | DRAM GB/s | % HBM peak | tensor % | FMA % | XU % | regs |\n|---|---|---|---|---|---|---|---|---|---|---|\n")
        for i, (name, grid, us, mb, gbs, tc, fma, xu, issue, regs) in enumerate(out):
            f.write(f"| {i} | `{name}` | {grid} | {us:.1f} | {mb:.1f} | {gbs:.0f} | {100 * gbs / peak:.0f} | {tc:.0f} | {fma:.0f} | {xu:.0f} | {regs} |\n")
    return out


def main_b():
    """The final-tree captures (tools/gpu_ncu_r02b.sh) -> profiles/r02b_*.md and the inference entry of profiles/traffic.json."""
    launch_list("launches_infer_b.csv", "r02b_launches_infer.md",
                "r02b: ncu launch list of ONE inference forward of the bench workload (ViT-B/16 bf16, batch 256), eager -- final tree: "
                "LayerNorm folded into the GEMMs",
                "ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off python tools/ncu_train_step.py --infer")
    note = ("DRAM GB/s = (dram read + write bytes) / duration against the measured HBM peak.  Serialised cold-cache replays: use them for "
            "pipe utilisation, traffic and stall shares; steady-state timings come from CUDA events (bench.py, tools/*).\n\n")
    rows = block_table("prof_infer2_raw.csv", "r02b_infer_block_kernels.md",
                       "# r02b: every kernel of a depth-2 ViT-B/16 inference forward (batch 256, bf16) under `ncu --set full --clock-control none`\n\n"
                       "`ncu --set full --clock-control none --profile-from-start off python tools/ncu_train_step.py --infer --depth 2` "
                       "(tools/gpu_ncu_r02b.sh).  Block 0: LayerNorm kernel + plain qkv GEMM, proj GEMM producing statistics (LNF = 2), "
                       "ln_finalize_kernel, fc1 + GELU consuming them (LNF = 1), fc2 producing; block 1: qkv consuming.  " + note)
    block_table("prof_new_raw.csv", "r02b_new_kernels.md",
                "# r02b: the kernels added late in round 2 under `ncu --set full --clock-control none` (tools/ncu_new_kernels.py)\n\n"
                "128-wide attention forward / backward at ViT-H/14 geometry (B = 64, N = 257, H = 16: attn_fwd_long_kernel<.., 2>, "
                "attn_bwd_delta_kernel<16>, attn_bwd_long_kernel<.., 2>, attn_dq_convert_kernel), the depthwise convolution producing "
                "LayerNorm statistics (dwconv_tma_kernel<7, true>, ConvNeXt-B stage 3) with ln_finalize_kernel and the consuming fc1 GEMM.  " + note)
    gem = [r for r in rows if r[0].startswith("gemm_bf16_kernel") and r[2] > 40]
    for r in gem:
        print(f"{r[2]:7.1f} us {r[3]:7.1f} MB tensor {r[5]:.0f}%  {r[0][:90]}")
    return rows


def main():
    launch_list("launches_train.csv", "r02_launches_train.md",
                "r02: ncu launch list of ONE training step of the bench workload (ViT-B/16 bf16, batch 256: zero_grad + fwd + CE + bwd + SGD), eager",
                "ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off python tools/ncu_train_step.py")
    launch_list("launches_infer.csv", "r02_launches_infer.md",
                "r02: ncu launch list of ONE inference forward of the bench workload (ViT-B/16 bf16, batch 256), eager",
                "ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off python tools/ncu_train_step.py --infer")
    rows = block_table()
    gem = [r for r in rows if r[0].startswith("gemm_bf16_kernel") and r[2] > 100]
    print(len(rows), "kernels;", len(gem), "large GEMMs")
    for r in gem:
        print(f"{r[2]:7.1f} us {r[3]:7.1f} MB tensor {r[5]:.0f}%  {r[0][:70]}")
    return rows


if __name__ == "__main__":
    import sys
    main_b() if len(sys.argv) > 1 and sys.argv[1] == "b" else main()
