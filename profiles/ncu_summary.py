#!/usr/bin/env python
"""Key metrics of one kernel from `ncu -i X.ncu-rep --page raw --csv` (stdin)."""
import csv
import sys

rows = list(csv.reader(sys.stdin))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__shared_mem_per_block_dynamic",
        "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled", "launch__block_size", "launch__grid_size", "sm__inst_executed_pipe_alu", "sm__inst_executed_pipe_fma",
        "sm__pipe_alu_cycles_active", "sm__pipe_fma_cycles_active", "sm__inst_executed_pipe_lsu", "lts__t_sector_hit_rate"]
for vals in rows[2:]:
    print("==", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
    for h, u, v in zip(hdr, units, vals):
        if any(h.startswith(w) for w in want) and "pct_of_peak_sustained_elapsed" not in h.replace("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "") and ".per_second" not in h:
            try:
                if float(v) == 0.0 and "stalled" in h:
                    continue
            except ValueError:
                pass
            print(f"  {h} [{u}] {v}")
