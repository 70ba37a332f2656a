#!/usr/bin/env python
"""Quick device-time check of the fused kernel: us per step for a workload, launched through
b200yolo_decode_nms_batches (overlapping launches) and in plain stream order.
    python profiles/quick_perf.py [cfg2 cfg2_sparse cfg3 cfg5 ...] [--steps 200]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from mobilenet_yolo_pytorch_b200 import _lib, ops

ap = argparse.ArgumentParser()
ap.add_argument("workloads", nargs="*", default=["cfg2", "cfg2_sparse"])
ap.add_argument("--steps", type=int, default=200)
ap.add_argument("--n", type=int, default=0)
a = ap.parse_args()
dev = torch.device("cuda", 0)
peak = 6452.8
for name in a.workloads:
    wl = bench.WORKLOADS[name]
    N = a.n or wl["N"]
    tables = bench.anchor_tables(wl)
    K = bench.cells_per_image(wl)
    in_bytes = N * bench.bytes_in_per_image(wl)
    R = max(4, int(np.ceil(400e6 / in_bytes)))
    sets = [tuple(h.to(dev) for h in bench.make_heads(wl, N, seed=r)) for r in range(R)]
    ring = int(os.environ.get("OUT_RING", "2"))   # result buffers in rotation (>= 2: chained launches, no wait before the stores)
    ring = a.steps if ring == 0 else ring          # 0: every step has its own result buffer
    outs = [(torch.empty((N, K, 7), dtype=torch.float32, device=dev), torch.empty((N,), dtype=torch.int32, device=dev)) for _ in range(ring)]
    out, cnt = outs[0]
    if K > _lib.load().b200yolo_max_cells(0):
        print(name, "large-image path: skipped here")
        continue
    plan = ops.BatchPlan([(sets[i % R][0], sets[i % R][1]) + outs[i % ring] for i in range(a.steps)], tables, wl["C"], wl["conf"])
    plan.run(0, min(10, a.steps))
    torch.cuda.synchronize()
    kept = float(cnt.sum().item())
    res = {}
    base_flags = int(os.environ.get("B200YOLO_FLAGS", "0"))
    for label, flags in (("overlapped", 0), ("stream_order", 2)):
        _lib.load().b200yolo_debug_set_flags(flags | base_flags)
        plan.run(0, min(10, a.steps))
        torch.cuda.synchronize()
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan.run()
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1) / a.steps * 1e3)
        res[label] = best
    _lib.load().b200yolo_debug_set_flags(base_flags)
    algo = in_bytes + 28 * kept + 4 * N
    print(f"{name:12s} N={N} kept/img={kept / N:7.1f}  overlapped {res['overlapped']:7.2f} us ({algo / res['overlapped'] / 1e3:6.0f} GB/s, "
          f"frac {algo / res['overlapped'] / 1e3 / peak:.3f})  stream order {res['stream_order']:7.2f} us  "
          f"images/s {N / res['overlapped'] * 1e6:.3e}", flush=True)
