#!/usr/bin/env python
"""Key metrics per kernel of an ncu report: ncu -i X.ncu-rep --page raw --csv | ncu_summary.py"""
import csv, sys
rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.avg.per_second"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("=====", d.get("Kernel Name"))
    for k in keys:
        if k in d: print(f"  {k:78s} {d[k]:>18s} {units[hdr.index(k)]}")
    for k in hdr:
        if "average_warps_issue_stalled" in k and "per_issue_active" in k and "not_issued" not in k:
            try: v = float(d[k].replace(',', ''))
            except ValueError: continue
            if v > 0.05: print(f"     stall {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):24s} {v:.3f}")
