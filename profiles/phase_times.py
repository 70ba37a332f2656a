#!/usr/bin/env python
"""Per-phase wall time of the fused decode+NMS kernel from in-kernel %globaltimer stamps
(b200yolo_debug_phase_stamps).  Prints, per workload, the median over images of each phase
and the span of the whole launch (first CTA start -> last CTA end)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from mobilenet_yolo_pytorch_b200 import _lib, ops

NAMES = ["init", "decode", "scan+scatter", "rank", "pairs", "sweep", "out-prefix", "store"]
dev = torch.device("cuda", 0)
for name in sys.argv[1:] or ["cfg2", "cfg2_sparse"]:
    wl = bench.WORKLOADS[name]
    N = wl["N"]
    tables = bench.anchor_tables(wl)
    sets = [tuple(h.to(dev) for h in bench.make_heads(wl, N, seed=s)) for s in range(3)]
    big = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    dbg = torch.zeros((N, 16), dtype=torch.int64, device=dev)
    for i in range(4):
        big.fill_(float(i))
        if i == 3:
            _lib.load().b200yolo_debug_phase_stamps(dbg.data_ptr())
        ops.decode_nms_padded(sets[i % 3][0], sets[i % 3][1], tables, wl["C"], wl["conf"])
    torch.cuda.synchronize()
    _lib.load().b200yolo_debug_phase_stamps(None)
    raw = dbg.cpu().numpy().astype(np.float64)
    rr = raw[:, 8:16]
    print("   decode rounds (warp 0 of each CTA, us after CTA start, median):",
          " ".join(f"{np.median(rr[:, k] - raw[:, 0]) / 1e3:.2f}" for k in range(7) if rr[:, k].max() > 0),
          "| init done at", f"{np.median(rr[:, 7] - raw[:, 0]) / 1e3:.2f}")
    t = raw[:, :8]
    t0 = t[:, 0].min()
    d = np.diff(t, axis=1) / 1e3
    print(f"== {name}: launch span {(t[:, 7].max() - t0) / 1e3:.1f} us; CTA start spread {(t[:, 0].max() - t0) / 1e3:.1f} us; "
          f"CTA duration median {np.median(t[:, 7] - t[:, 0]) / 1e3:.1f} max {(t[:, 7] - t[:, 0]).max() / 1e3:.1f} us")
    for k in range(7):
        print(f"   {NAMES[k + 1]:<13} median {np.median(d[:, k]):6.2f}  p90 {np.percentile(d[:, k], 90):6.2f}  max {d[:, k].max():6.2f} us"
              f"   (ends at median {np.median(t[:, k + 1] - t0) / 1e3:6.2f} us)")
