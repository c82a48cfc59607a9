#!/usr/bin/env python
"""Summarise an .ncu-rep (one kernel) into the metrics B200_PROFILING.md names: usage: ncu_summary.py rep [rep...]"""
import csv, io, subprocess, sys
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.sum", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.avg", "launch__grid_size", "launch__block_size",
        "smsp__average_warp_latency_issue_stalled_barrier.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "lts__t_bytes.sum", "smsp__cycles_active.avg"]
for rep in sys.argv[1:]:
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, vals = rows[0], rows[2] if len(rows) > 2 else None
    units = rows[1]
    print("==", rep, "kernel:", vals[hdr.index("Kernel Name")][:60] if vals else None)
    for i, h in enumerate(hdr):
        if h in KEYS or "warp_issue_stalled" in h and h.endswith("_per_warp_active.pct"):
            try:
                v = float(vals[i].replace(",", ""))
            except ValueError:
                continue
            if "stalled" in h and v < 3: continue
            print(f"  {h:95s} {vals[i]:>18s} {units[i]}")
