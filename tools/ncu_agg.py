#!/usr/bin/env python
"""Aggregate an ncu report's source page per function / per line (instructions executed, stall samples).
usage: ncu_agg.py report.ncu-rep [func|lines FILE.cuh [minpct]]"""
import csv, glob, os, re, subprocess, sys
rep = sys.argv[1]
mode = sys.argv[2] if len(sys.argv) > 2 else "func"
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "2026-simple-c-tts_b200", "csrc", "gpu")
srcs = {os.path.basename(f): open(f).read().splitlines() for f in glob.glob(root + "/*.cu*")}
def fn(path, ln):
    src = srcs.get(os.path.basename(path))
    if not src: return os.path.basename(path)
    name = "?"
    for l in src[:ln]:
        m = re.match(r'^(?:__device__|__global__).*?(\w+)\(', l)
        if m: name = m.group(1)
    return name
tot = tots = 0; by = {}; bys = {}; lines = {}
cur = None
for r in rows:
    if r and r[0] == "File Path": cur = r[1]; continue
    if r and r[0] == "Line No": hdr = r; iI = hdr.index("Instructions Executed"); iS = hdr.index("# Samples"); continue
    if len(r) > 10 and r[0].isdigit():
        try: inst = int(r[iI]); smp = int(r[iS])
        except ValueError: continue
        k = fn(cur, int(r[0]))
        by[k] = by.get(k, 0) + inst; bys[k] = bys.get(k, 0) + smp; tot += inst; tots += smp
        lines.setdefault(os.path.basename(cur), []).append((int(r[0]), inst, smp, r[1]))
print("total warp instructions", tot, "stall samples", tots)
if mode == "func":
    for k, v in sorted(by.items(), key=lambda x: -x[1])[:30]:
        print(f"{k:28s} {100*v/tot:6.2f}% inst  {100*bys[k]/tots:6.2f}% samples")
else:
    f = sys.argv[3]; minpct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.25
    for ln, inst, smp, src in sorted(lines.get(f, [])):
        if 100*inst/tot >= minpct or 100*smp/tots >= minpct:
            print(f"{ln:4d} {100*inst/tot:5.2f}%i {100*smp/tots:5.2f}%s {src.strip()[:110]}")
