#!/usr/bin/env python
"""Markdown summary of ncu captures summarised on the GPU box as <name>.raw.csv (+ <name>.source.csv):
key metrics of the first kernel instance, stall-reason totals and the hottest SASS lines.
usage: python scripts/summarise_ncu.py gpurun_out/prof4_k1_ref_ef64 [more basenames...] > profiles/xxx.md"""
import csv, io, os, subprocess, sys

KEYS = ["Kernel Name", "Grid Size", "Block Size", "launch__cluster_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__cycles_elapsed.avg.per_second", "smsp__inst_executed.sum"]
here = os.path.dirname(os.path.abspath(__file__))
for base in sys.argv[1:]:
    print(f"## {os.path.basename(base)}\n```")
    rows = list(csv.reader(open(base + ".raw.csv")))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    for k in KEYS:
        if k in d:
            print(f"{k:62s} {d[k][0]} {d[k][1]}")
    print("```")
    if os.path.exists(base + ".source.csv"):
        out = subprocess.run([sys.executable, os.path.join(here, "ncu_stalls.py"), "8"], stdin=open(base + ".source.csv"), capture_output=True, text=True).stdout
        print("Stall samples and hottest SASS lines:\n```\n" + out.rstrip() + "\n```")
