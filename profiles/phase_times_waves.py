#!/usr/bin/env python
"""Phase times split by wave (first-wave CTAs vs CTAs scheduled later) for a multi-wave launch:
tells cold-start effects (instruction cache, first-touch) from steady-state costs."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from mobilenet_yolo_pytorch_b200 import _lib, ops

NAMES = ["decode", "scan+scatter", "rank", "pairs", "sweep", "out-prefix", "store"]
dev = torch.device("cuda", 0)
name, N = sys.argv[1], int(sys.argv[2])
wl = dict(bench.WORKLOADS[name], N=N)
tables = bench.anchor_tables(wl)
h0, h1 = (h.to(dev) for h in bench.make_heads(wl, N, seed=0))
big = torch.empty(64 << 20, dtype=torch.float32, device=dev)
dbg = torch.zeros((N, 16), dtype=torch.int64, device=dev)
for i in range(3):
    big.fill_(float(i))
    if i == 2:
        _lib.load().b200yolo_debug_phase_stamps(dbg.data_ptr())
    ops.decode_nms_padded(h0, h1, tables, wl["C"], wl["conf"])
torch.cuda.synchronize()
_lib.load().b200yolo_debug_phase_stamps(None)
t = dbg.cpu().numpy().astype(np.float64)[:, :8]
t0 = t[:, 0].min()
d = np.diff(t, axis=1) / 1e3
first = (t[:, 0] - t0) < 1000.0
print(f"== {name} N={N}: span {(t[:, 7].max() - t0) / 1e3:.1f} us; first-wave CTAs {first.sum()}, later {(~first).sum()}")
for k in range(7):
    a, b = d[first, k], d[~first, k]
    print(f"   {NAMES[k]:<13} first wave median {np.median(a):6.2f}   later waves median {np.median(b) if len(b) else float('nan'):6.2f} us")
print(f"   CTA total     first wave median {np.median((t[first, 7] - t[first, 0])) / 1e3:6.2f}   later {np.median((t[~first, 7] - t[~first, 0])) / 1e3 if (~first).any() else float('nan'):6.2f}")
