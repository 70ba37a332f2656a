#!/usr/bin/env python
"""Time the fused kernel on cfg2 dense / sparse heads under different B200YOLO debug flags.
Usage: flag_sweep.py FLAGS [FLAGS ...]   (integers; see include/b200yolo.h)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from mobilenet_yolo_pytorch_b200 import _lib, ops

dev = torch.device("cuda", 0)
res = {}
for name in ("cfg2", "cfg2_sparse"):
    wl = bench.WORKLOADS[name]
    N = wl["N"]
    tables = bench.anchor_tables(wl)
    R = 9
    sets = [tuple(h.to(dev) for h in bench.make_heads(wl, N, seed=s)) for s in range(R)]
    out = torch.empty((N, bench.cells_per_image(wl), 7), dtype=torch.float32, device=dev)
    cnt = torch.empty((N,), dtype=torch.int32, device=dev)
    for fl in [int(a) for a in sys.argv[1:]] or [0]:
        _lib.load().b200yolo_debug_set_flags(fl)
        for i in range(20):
            ops.decode_nms_padded(sets[i % R][0], sets[i % R][1], tables, wl["C"], wl["conf"], out=out, out_count=cnt)
        torch.cuda.synchronize()
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(200):
                ops.decode_nms_padded(sets[i % R][0], sets[i % R][1], tables, wl["C"], wl["conf"], out=out, out_count=cnt)
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1) / 200 * 1e3)
        print(f"{name:12s} flags {fl:6d} (delay {(fl >> 8) / 10:.1f} us): {best:7.2f} us/launch", flush=True)
    _lib.load().b200yolo_debug_set_flags(0)
