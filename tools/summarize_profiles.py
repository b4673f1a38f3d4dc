"""Turn the ncu artefacts of a gpurun call (gpurun_out/launches.csv, *.ncu-rep) into the tracked summaries
under profiles/ (per-kernel share of a step from the launch list; key metrics of the full captures)."""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
SRC = os.path.join(ROOT, "gpurun_out")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def launches():
    path = os.path.join(SRC, "launches.csv")
    if not os.path.exists(path):
        return
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg, tot = collections.OrderedDict(), 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"].startswith("n") else (v * 1e3 if row["Metric Unit"].startswith("m") else v)
        key = (re.sub(r"\(.*", "", row["Kernel Name"]), row["Grid Size"], row["Block Size"])
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
        tot += v
    with open(os.path.join(OUT, f"{tag}_launches.md"), "w") as f:
        f.write(f"# {tag}: ncu launch list of 2 timed steps of `bench.py` (ViT-B/16 bf16 infer, batch 256)\n\n"
                "`ncu --metrics gpu__time_duration.sum --clock-control none -k regex:<our kernels> -s 267 -c 178 python bench.py "
                "--steps 2 --warmup 3`.  Times are cold-cache and serialised: compare SHARES, not absolutes.\n\n"
                f"total {tot:.1f} us over {sum(a[0] for a in agg.values())} launches\n\n| us | share | launches | avg us | kernel (grid, block) |\n|---|---|---|---|---|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {t:.1f} | {100 * t / tot:.1f}% | {n} | {t / n:.1f} | `{k[0]}` {k[1]} {k[2]} |\n")


def full(name):
    rep = os.path.join(SRC, name + ".ncu-rep")
    if not os.path.exists(rep):
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(os.path.join(OUT, f"{tag}_{name}.md"), "w") as f:
        f.write(f"# {tag}: `ncu --set full --clock-control none --import-source on` -- {name}\n\n"
                "(values per launch; a number printed under ncu is never a bench value)\n")
        for row in rows[2:]:
            d = dict(zip(hdr, row))
            f.write(f"\n## {d.get('Kernel Name', '?')[:110]}\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in d:
                    f.write(f"| {k} | {d[k]} | {units[hdr.index(k)]} |\n")
            try:
                rd = float(d["dram__bytes_read.sum"]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[units[hdr.index("dram__bytes_read.sum")]]
                wr = float(d["dram__bytes_write.sum"]) * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[units[hdr.index("dram__bytes_write.sum")]]
                f.write(f"| traffic = dram read + write | {(rd + wr) / 1e6:.1f} | MB |\n")
            except (KeyError, ValueError):
                pass


os.makedirs(OUT, exist_ok=True)
launches()
for n in ("prof_gemm", "prof_attn", "prof_bwd", "prof_dwconv", "prof_optim"):
    full(n)
print(os.listdir(OUT))
