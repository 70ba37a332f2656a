#!/usr/bin/env python
"""The three separate entry points (YOLOLoss.forward x2 + utils.box.nms = 3 launches, padded forms, no host sync)
against the fused call on cfg2 dense / sparse heads."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from mobilenet_yolo_pytorch_b200 import ops

dev = torch.device("cuda", 0)
for name in ("cfg2", "cfg2_sparse"):
    wl = bench.WORKLOADS[name]
    N, C = wl["N"], wl["C"]
    tables = bench.anchor_tables(wl)
    R = 9
    sets = [tuple(h.to(dev) for h in bench.make_heads(wl, N, seed=s)) for s in range(R)]

    def sep(i):
        h0, h1 = sets[i % R]
        r0, c0 = ops.decode_head_padded(h0, tables[0], C, wl["conf"])
        r1, c1 = ops.decode_head_padded(h1, tables[1], C, wl["conf"])
        return ops.nms_padded(r0, c0, r1, c1, C)

    def fused(i):
        h0, h1 = sets[i % R]
        return ops.decode_nms_padded(h0, h1, tables, C, wl["conf"])

    for label, fn in (("separate (3 launches)", sep), ("fused (1 launch)", fused)):
        for i in range(10):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(100):
            fn(i)
        e1.record()
        e1.synchronize()
        print(f"{name:12s} {label:22s} {e0.elapsed_time(e1) * 10:8.2f} us/step")
