"""Summarise an .ncu-rep (read here, no GPU): python bench_tools/ncu_read.py gpurun_out/x.ncu-rep [warp_planes]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
wp = float(sys.argv[2]) if len(sys.argv) > 2 else 45279.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
d = dict(zip(rows[0], rows[2]))
keys = ["gpu__time_duration.sum", "launch__registers_per_thread", "sm__inst_executed.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__occupancy_limit_registers"]
for k in keys:
    if k in d:
        print(f"{k} = {d[k]}")
st = {k: float(v) for k, v in d.items() if re.match(r"smsp__average_warps_issue_stalled_.*_per_issue_active\.ratio$", k) and v not in ("", "n/a")}
print("stalls/issue:", ", ".join(f"{k.split('stalled_')[1].split('_per_')[0]}={v:.2f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:9]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
heads = [i for i, r in enumerate(rows) if "Instructions Executed" in r]   # one table per captured launch: the first
hdr = rows[heads[0]]
iS, iE = hdr.index("Source"), hdr.index("Instructions Executed")
ops, tot = collections.Counter(), 0
for r in rows[heads[0] + 1:(heads[1] - 1 if len(heads) > 1 else len(rows))]:
    if len(r) <= iE or not r[iE].isdigit() or not r[iS].split():
        continue
    p = r[iS].split()
    b = p[1] if p[0].startswith("@") else p[0]
    n = int(r[iE])
    ops[b.split(".")[0]] += n
    tot += n
print(f"warp instructions {tot}  per warp-plane {tot / wp:.1f}")
print("  ".join(f"{k}={v / wp:.1f}" for k, v in ops.most_common(24)))
