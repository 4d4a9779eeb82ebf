"""Summarise an .ncu-rep (ncu --set full) into a small CSV of the metrics the roofline discussion uses.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/out.csv"""
import csv
import io
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__sass_inst_executed_op_integer_pred_on.sum",
]
PREFIX = ("smsp__average_warps_issue_stalled_", "smsp__average_warp_latency_issue_stalled_")

raw = subprocess.check_output(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], text=True)
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
out = [["metric", "unit"] + [f"launch{i}" for i in range(len(data))]]
name_col = hdr.index("Kernel Name")
out.append(["Kernel Name", ""] + [r[name_col] for r in data])
for j, h in enumerate(hdr):
    if h in KEEP or (h.startswith(PREFIX) and h.endswith("_per_issue_active.ratio")):
        out.append([h, units[j]] + [r[j] for r in data])
csv.writer(open(sys.argv[2], "w")).writerows(out)
print(f"{len(out) - 2} metrics x {len(data)} launches -> {sys.argv[2]}")
