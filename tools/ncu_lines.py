#!/usr/bin/env python
"""Executed warp-instructions and stall samples per CUDA source line of the first kernel of an .ncu-rep (needs -lineinfo
and --import-source on).  usage: python tools/ncu_lines.py REP [top_n] [warps]   (warps: divide counts by this number)"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 60; warps = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
fname, L, tot = "", [], 0
for r in csv.reader(io.StringIO(txt)):
    if len(r) >= 2 and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0].isdigit():
        e, s = int(r[7] or 0), int(r[6] or 0)
        L.append((e, s, fname, int(r[0]), r[1].strip()[:120])); tot += e
print("total warp-instructions", tot, " per warp", tot / warps)
for e, s, f, ln, src in sorted(L, reverse=True)[:top]:
    print(f"{e/warps:10.1f} {s:6d} {f}:{ln:<5d} {src}")
