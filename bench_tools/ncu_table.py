"""One line per captured launch of an .ncu-rep (read here, no GPU):
python bench_tools/ncu_table.py gpurun_out/x.ncu-rep [kernel-name filter]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
cols = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "us"), ("launch__grid_size", "grid"),
        ("launch__block_size", "blk"), ("launch__registers_per_thread", "regs"),
        ("dram__bytes_read.sum", "rd"), ("dram__bytes_write.sum", "wr"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("smsp__inst_executed.sum", "inst"),
        ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "ld_sect"),
        ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "ld_req"),
        ("lts__t_sector_hit_rate.pct", "l2hit%"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar"),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
        ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg")]
idx = [(hdr.index(c), n) for c, n in cols if c in hdr]
print(" | ".join(f"{n}[{units[i]}]" if units[i] else n for i, n in idx))
for r in rows[2:]:
    if flt and flt not in r[hdr.index("Kernel Name")]:
        continue
    print(" | ".join(r[i][:28] for i, _ in idx))
