"""Text summary of one kernel of an .ncu-rep (the metrics DESIGN.md quotes), for profiles/.  Usage: ncu_summary.py rep [kernel-substring]"""
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "sm__cycles_active.avg",
        "sm__cycles_elapsed.max"]
STALL = "smsp__average_warps_issue_stalled_"

rep = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
ik = hdr.index("Kernel Name")
for r in rows[2:]:
    if want not in r[ik]:
        continue
    print("# ncu --set full --clock-control none --import-source on;", rep.split("/")[-1])
    print("#", r[ik])
    d = dict(zip(hdr, zip(units, r)))
    for k in KEYS:
        if k in d:
            print(f"{k} = {d[k][1]} {d[k][0]}")
    for k in hdr:
        if k.startswith(STALL) and k.endswith("_per_issue_active.ratio") and float(d[k][1] or 0) >= 0.05:
            print(f"{k} = {d[k][1]} {d[k][0]}")
    break
