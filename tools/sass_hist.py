#!/usr/bin/env python
"""Instruction histogram per kernel of libctts_gpu.so (cuobjdump -sass): what profiles/*_sass_histogram.txt holds.
usage: sass_hist.py [library.so] > profiles/rNN_sass_histogram.txt"""
import collections, os, re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                        "2026-simple-c-tts_b200", "libctts_gpu.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
kern, hist, arch = None, {}, set()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        hist[kern] = collections.Counter()
        continue
    m = re.search(r"arch = (sm_\w+)", line)
    if m:
        arch.add(m.group(1))
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)(\.[A-Z0-9_.]*)?", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
print("library:", os.path.basename(so), " architectures:", ", ".join(sorted(arch)))
BLACKWELL = ("UTMALDG", "UTMASTG", "UBLKCP", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "FFMA2", "FMUL2", "FADD2", "LDGSTS", "HMMA")
for k, h in hist.items():
    n = sum(h.values())
    print(f"\n{k}: {n} SASS instructions")
    print("  " + ", ".join(f"{op} {c}" for op, c in h.most_common(28)))
    print("  tensor / TMA / packed-FP32 / cp.async mnemonics: " + (", ".join(f"{op} {h[op]}" for op in BLACKWELL if h[op]) or "none"))
