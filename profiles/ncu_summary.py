#!/usr/bin/env python
"""Key metrics of an ncu report: ncu -i X.ncu-rep --page raw --csv | python profiles/ncu_summary.py"""
import csv
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__cycles_elapsed.max"]
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print("kernel:", name[:110])
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print(f"  {w:70s} {vals[i]:>16s} {units[i]}")
