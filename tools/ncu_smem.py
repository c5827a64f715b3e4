#!/usr/bin/env python
"""Shared-memory wavefronts (total / excessive = bank conflicts) per source line of an ncu report."""
import csv, os, subprocess, sys
rep = sys.argv[1]; minpct = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur = None; tot = totx = 0; L = []
for r in rows:
    if r and r[0] == "File Path": cur = os.path.basename(r[1]); continue
    if r and r[0] == "Line No": hdr = r; iW = hdr.index("L1 Wavefronts Shared"); iX = hdr.index("L1 Wavefronts Shared Excessive"); continue
    if len(r) > 10 and r[0].isdigit():
        try: w = int(r[iW]); x = int(r[iX])
        except ValueError: continue
        if w: L.append((cur, int(r[0]), w, x, r[1])); tot += w; totx += x
print("shared wavefronts", tot, "excessive", totx, f"({100*totx/max(tot,1):.1f}%)")
for f, ln, w, x, src in sorted(L, key=lambda t: -t[2]):
    if 100 * w / tot >= minpct: print(f"{f}:{ln:4d} {100*w/tot:5.1f}% of wavefronts, {100*x/max(w,1):5.1f}% excessive  {src.strip()[:90]}")
