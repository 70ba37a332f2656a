#!/usr/bin/env python
"""Per-phase time of the fused decode+NMS kernel from in-kernel stamps (b200yolo_debug_phase_stamps):
%globaltimer (256 ns resolution) for the launch span across CTAs, the SM cycle counter for the phases
inside a CTA.  Prints, per workload, the median over images of each phase.
    PHASE_N=32 python profiles/phase_times.py cfg2 cfg2_sparse
    PHASE_STEADY=1 ...: 16 overlapping launches from one C call (flag 512), the stamps of launch 9 are reported"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from mobilenet_yolo_pytorch_b200 import _lib, ops

# stamp slots in kernel order (decode_nms.cuh)
ORDER = [(0, "start"), (15, "init"), (8, "decode round 1"), (9, "decode round 2"), (10, "decode round 3"),
         (11, "decode round 4+"), (1, "decode done (barrier)"), (12, "bucket scan per class + barrier"),
         (13, "class starts (warp 0) + barrier"), (3, "tables (warp 0) | scatter, rank (others) + barrier"),
         (4, "pairs + sweep + barrier"), (5, "wait for the previous launch"), (7, "output (warp 0's tiles)")]
dev = torch.device("cuda", 0)
MHZ = 1965.0
for name in sys.argv[1:] or ["cfg2", "cfg2_sparse"]:
    wl = bench.WORKLOADS[name]
    N = int(os.environ.get("PHASE_N", wl["N"]))
    tables = bench.anchor_tables(wl)
    sets = [tuple(h.to(dev) for h in bench.make_heads(wl, N, seed=s)) for s in range(3)]
    big = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    dbg = torch.zeros((N, 32), dtype=torch.int64, device=dev)
    if os.environ.get("PHASE_STEADY"):
        K = bench.cells_per_image(wl)
        sets = [tuple(h.to(dev) for h in bench.make_heads(wl, N, seed=s)) for s in range(9)]
        out = torch.empty((N, K, 7), dtype=torch.float32, device=dev)
        cnt = torch.empty((N,), dtype=torch.int32, device=dev)
        plan = ops.BatchPlan([(sets[i % 9][0], sets[i % 9][1], out, cnt) for i in range(16)], tables, wl["C"], wl["conf"])
        ring = torch.zeros((16, N, 32), dtype=torch.int64, device=dev)
        base = int(os.environ.get("B200YOLO_FLAGS", "0"))
        _lib.load().b200yolo_debug_set_flags(base | 512)
        _lib.load().b200yolo_debug_phase_stamps(ring.data_ptr())
        plan.run()
        torch.cuda.synchronize()
        _lib.load().b200yolo_debug_phase_stamps(ring.data_ptr())
        plan.run()
        torch.cuda.synchronize()
        _lib.load().b200yolo_debug_phase_stamps(None)
        _lib.load().b200yolo_debug_set_flags(base)
        allg = ring.cpu().numpy().astype(np.float64)
        span = allg[:, :, 7].max() - allg[:, :, 0].min()
        print(f"== {name}: 16 overlapped stamped launches span {span / 1e3:.1f} us = {span / 16e3:.2f} us per launch")
        dbg = ring[9]
    else:
        for i in range(4):
            big.fill_(float(i))
            if i == 3:
                _lib.load().b200yolo_debug_phase_stamps(dbg.data_ptr())
            ops.decode_nms_padded(sets[i % 3][0], sets[i % 3][1], tables, wl["C"], wl["conf"])
        torch.cuda.synchronize()
        _lib.load().b200yolo_debug_phase_stamps(None)
    raw = dbg.cpu().numpy().astype(np.float64)
    g, c = raw[:, :16], raw[:, 16:]
    t0 = g[:, 0].min()
    print(f"== {name} N={N}: launch span {(g[:, 7].max() - t0) / 1e3:.1f} us; CTA start spread {(g[:, 0].max() - t0) / 1e3:.1f} us; "
          f"CTA duration (SM clock @ {MHZ:.0f} MHz) median {np.median(c[:, 7] - c[:, 0]) / MHZ:.2f} max {(c[:, 7] - c[:, 0]).max() / MHZ:.2f} us")
    prev = 0
    for slot, label in ORDER[1:]:
        if c[:, slot].max() == 0:
            continue
        d = (c[:, slot] - c[:, prev]) / MHZ
        print(f"   {label:<36} median {np.median(d):6.2f}  p90 {np.percentile(d, 90):6.2f}  max {d.max():6.2f} us"
              f"   (ends at median {np.median(c[:, slot] - c[:, 0]) / MHZ:6.2f} us)")
        prev = slot
