#!/usr/bin/env python
"""Selected rows of an ncu report's raw page as metric,unit,value CSV (what profiles/*_summary.csv hold).
usage: ncu_summary.py report.ncu-rep > summary.csv"""
import csv, subprocess, sys
KEEP = ("gpu__time_duration", "dram__bytes", "gpu__dram_throughput", "launch__", "sm__inst_executed", "sm__issue_active",
        "smsp__issue_active", "smsp__inst_executed.sum", "sm__warps_active", "smsp__average_warp", "smsp__pcsamp_warps_issue_stalled",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared", "l1tex__data_pipe_lsu_wavefronts_mem_shared", "sm__pipe_",
        "lts__t_bytes", "lts__t_sector_hit_rate", "l1tex__t_sector_hit_rate", "sm__throughput", "smsp__cycles_active.avg",
        "sm__cycles_elapsed.max", "smsp__thread_inst_executed.sum")
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units, val = rows[0], rows[1], rows[-1]
w = csv.writer(sys.stdout)
w.writerow(["metric", "unit", "value"])
for h, u, v in sorted(zip(hdr, units, val)):
    if h.startswith(KEEP) or h == "Kernel Name":
        w.writerow([h, u, v])
