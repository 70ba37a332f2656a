#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --print-source cuda,sass --csv` per CUDA source line:
share of executed warp instructions and of stall samples.  Usage: ncu_lines.py file.csv [min_pct]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
minp = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
fname = ""
out = []
hdr = None
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if len(r) > 5 and r[0] == "Line No":
        hdr = {n: k for k, n in enumerate(r)}
        continue
    if hdr is None or len(r) < 10 or r[2] != "-":
        continue  # only per-CUDA-line aggregate rows (Address column "-")
    try:
        inst = int(float(r[hdr["Instructions Executed"]] or 0))
        samp = int(float(r[hdr["# Samples"]] or 0))
    except ValueError:
        continue
    out.append((fname, int(r[0]), r[1].strip(), inst, samp))
ti = sum(o[3] for o in out) or 1
ts = sum(o[4] for o in out) or 1
print(f"total warp-instructions {ti}, samples {ts}")
for f, ln, src, inst, samp in out:
    if 100 * inst / ti >= minp or 100 * samp / ts >= minp:
        print(f"{f}:{ln:<4} inst {100 * inst / ti:5.1f}%  samples {100 * samp / ts:5.1f}%  | {src[:100]}")
