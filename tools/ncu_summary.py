#!/usr/bin/env python
"""Key counters per kernel launch of an .ncu-rep (raw page): duration, DRAM bytes/throughput, L2, occupancy, issue, stalls.
usage: python tools/ncu_summary.py report.ncu-rep"""
import csv
import subprocess
import sys

out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
ci = {h: i for i, h in enumerate(hdr)}
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__cycles_elapsed.avg.per_second"]
STALL = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    print("=" * 100)
    print(r[ci["Kernel Name"]][:110] if "Kernel Name" in ci else r[4][:110])
    for k in KEYS:
        if k in ci:
            print(f"  {k:75s} {r[ci[k]]:>18s} {units[ci[k]]}")
    st = sorted(((float(r[ci[h]] or 0), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for h in STALL), reverse=True)[:6]
    print("  top stalls (warps per issue):", ", ".join(f"{n}={v:.2f}" for v, n in st))
