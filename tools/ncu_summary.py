"""One line per kernel launch of an ncu --set full report: time, DRAM traffic, issue / pipe utilisation, occupancy, top stalls.
    python tools/ncu_summary.py report.ncu-rep"""
import csv
import io
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = {h: i for i, h in enumerate(rows[0])}


def g(r, k, d=0.0):
    try:
        return float(r[hdr[k]].replace(",", ""))
    except Exception:
        return d


stalls = [k for k in rows[0] if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    name = r[hdr["Kernel Name"]].split("(")[0]
    t = g(r, "gpu__time_duration.sum")
    rd, wr = g(r, "dram__bytes_read.sum"), g(r, "dram__bytes_write.sum")
    unit_t = rows[1][hdr["gpu__time_duration.sum"]]
    unit_b = rows[1][hdr["dram__bytes_read.sum"]]
    top = sorted(((g(r, k), k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]) for k in stalls), reverse=True)[:4]
    print(f"{name:28s} {t:9.1f} {unit_t}  dram r {rd:7.1f} w {wr:7.1f} {unit_b}  dram% {g(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):5.1f}  "
          f"issue% {g(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active'):5.1f}  warps% {g(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):5.1f}  "
          f"regs {g(r, 'launch__registers_per_thread'):3.0f}  inst {g(r, 'smsp__inst_executed.sum'):.3g}  lanes/inst {g(r, 'smsp__thread_inst_executed_per_inst_executed.ratio'):4.1f}  "
          f"fma% {g(r, 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'):4.1f} xu% {g(r, 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'):4.1f} "
          f"lsu% {g(r, 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active'):4.1f}  stalls " + ", ".join(f"{n} {v:.1f}" for v, n in top))
