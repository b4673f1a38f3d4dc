"""profiles/<tag>_kernels.md: one row per hot-path kernel from the `ncu --set full` captures in gpurun_out/
(tools/gpu_ncu_all.sh + tools/gpu_final.sh): duration, DRAM bytes and throughput vs the measured HBM peak, tensor /
FMA / XU pipe utilisation, registers, the top stall reason."""
import csv
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
peaks = {"hbm": 6569.6}
try:
    mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    peaks["hbm"] = float(mp.get("hbm_gbs", mp.get("hbm_GBps", peaks["hbm"])))
except (OSError, ValueError, TypeError):
    pass

M = {"t": "gpu__time_duration.sum", "rd": "dram__bytes_read.sum", "wr": "dram__bytes_write.sum",
     "tc": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
     "fma": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
     "xu": "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
     "issue": "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "regs": "launch__registers_per_thread", "grid": "launch__grid_size", "block": "launch__block_size",
     "warps": "sm__warps_active.avg.pct_of_peak_sustained_active"}
STALLS = ["long_scoreboard", "short_scoreboard", "math_pipe_throttle", "mio_throttle", "barrier", "wait", "not_selected",
          "tex_throttle", "lg_throttle", "dispatch_stall", "membar", "sleeping", "branch_resolving", "no_instruction"]


def unit_scale(u):
    return {"nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6, "byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0,
            "Gbyte": 1e3}.get(u, 1.0)


rows = []
for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", "prof_*.ncu-rep"))):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(raw.splitlines()))
    if len(r) < 3:
        continue
    hdr, units = r[0], r[1]
    for vals in r[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        g = lambda k: float(d[M[k]].replace(",", "")) * (unit_scale(u[M[k]]) if k in ("t", "rd", "wr") else 1.0) if d.get(M[k]) else 0.0  # noqa: E731
        st = {s: float(d.get(f"smsp__average_warps_issue_stalled_{s}_per_issue_active.ratio", "0") or 0) for s in STALLS}
        top = max(st, key=st.get)
        name = d.get("Kernel Name", "?").split("(")[0].replace("void ", "").replace("b200::", "")
        t = g("t")
        mb = g("rd") + g("wr")
        rows.append((os.path.basename(rep)[5:-8], name[:58], t, mb, mb / t * 1e3 / 1e3 if t else 0, g("tc"), g("fma"), g("xu"),
                     g("issue"), int(g("regs")), f"{int(g('grid'))}x{int(g('block'))}", f"{top} {st[top]:.1f}"))
out = os.path.join(ROOT, "profiles", f"{tag}_kernels.md")
with open(out, "w") as f:
    f.write(f"# {tag}: every hot-path kernel under `ncu --set full --clock-control none` (one launch each, production shapes)\n\n"
            "DRAM GB/s = (dram read + write bytes) / duration, against the measured HBM peak "
            f"{peaks['hbm']:.0f} GB/s.  Captures are serialised, cold-cache replays: use them for pipe utilisation, traffic and "
            "stall reasons; steady-state timings are in README.md.\n\n"
            "| capture | kernel | us | DRAM MB | DRAM GB/s | % HBM peak | tensor % | FMA % | XU % | issue % | regs | grid x block | top stall (cycles/issue) |\n"
            "|---|---|---|---|---|---|---|---|---|---|---|---|---|\n")
    for c, n, t, mb, gbs, tc, fma, xu, iss, regs, gb, top in rows:
        f.write(f"| {c} | `{n}` | {t:.1f} | {mb:.1f} | {mb / t * 1e3:.0f} | {100 * mb / t * 1e3 / peaks['hbm']:.0f} | {tc:.0f} | {fma:.0f} | "
                f"{xu:.0f} | {iss:.0f} | {regs} | {gb} | {top} |\n")
print(open(out).read())
