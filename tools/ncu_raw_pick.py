#!/usr/bin/env python
"""Picks the metrics that matter for the roofline out of `ncu -i X.ncu-rep --page raw --csv` (file argument)."""
import csv
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct",
        "sm__pipe_tensor_cycles_active", "sm__inst_executed_pipe_tensor", "sm__warps_active.avg.pct",
        "launch__registers_per_thread", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__throughput.avg.pct", "sm__throughput.avg.pct",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size",
        "smsp__average_warp", "smsp__issue_active.avg.pct", "sm__inst_executed.sum ", "lts__t_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__sass_data_bytes_mem_shared", "smsp__inst_executed_op_shared",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "dram__cycles_active"]
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("kernel:", d.get("Kernel Name", "")[:90], "| grid", d.get("Grid Size"), "block", d.get("Block Size"))
    for i, h in enumerate(hdr):
        if any(k in h for k in KEYS) and r[i] not in ("", "n/a"):
            print(f"   {h:95s} {r[i]:>18s} {units[i]}")
