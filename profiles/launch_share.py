#!/usr/bin/env python
"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv`)
per (kernel, grid size): launches, total / mean duration and share of all GPU time in the capture.
Usage: launch_share.py launches.csv"""
import collections
import csv
import sys

tot = collections.defaultdict(lambda: [0, 0.0])
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]
ki, vi, mi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name"), hdr.index("Grid Size")
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum":
        continue
    t = tot[r[ki].split('(')[0].replace('void ', '') + ' grid=' + r[gi].replace(' ', '')]
    t[0] += 1
    t[1] += float(r[vi].replace(",", ""))
allns = sum(v[1] for v in tot.values()) or 1.0
print(f"{'launches':>8} {'total us':>10} {'mean us':>9} {'share':>7}  kernel")
for k, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:8d} {ns / 1e3:10.1f} {ns / n / 1e3:9.2f} {100 * ns / allns:6.1f}%  {k[:110]}")
