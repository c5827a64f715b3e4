#!/usr/bin/env python
"""One entry of profiles/traffic.json from an `ncu --set full` report: DRAM bytes, issue-slot and pipe utilisation,
warp instructions and duration of the (single) captured launch.   usage: ncu_facts.py report.ncu-rep"""
import csv, json, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u, v = rows[0], rows[1], rows[-1]
d = dict(zip(h, v)); un = dict(zip(h, u))
def num(k):
    return float(d[k].replace(",", ""))
def bytes_of(k):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[un[k]]
    return num(k) * scale
ms = num("gpu__time_duration.sum") * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[un["gpu__time_duration.sum"]]
print(json.dumps({
    "dram_bytes": int(bytes_of("dram__bytes_read.sum") + bytes_of("dram__bytes_write.sum")),
    "dram_read_bytes": int(bytes_of("dram__bytes_read.sum")), "dram_write_bytes": int(bytes_of("dram__bytes_write.sum")),
    "ms": round(ms, 4), "warp_instructions": int(num("smsp__inst_executed.sum")),
    "issue_active_pct": round(num("sm__issue_active.avg.pct_of_peak_sustained_elapsed"), 2),
    "fma_pipe_pct": round(num("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"), 2),
    "alu_pipe_pct": round(num("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"), 2),
    "smem_wavefronts_pct": round(num("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"), 2),
    "registers": int(num("launch__registers_per_thread")), "kernel": d.get("Kernel Name", "")}))
