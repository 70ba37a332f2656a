#!/usr/bin/env python
"""Tiny driver for ncu captures: launches the fused decode+NMS kernel a few times on
dense (randn) and sparse (objectness logits shifted by -2.6) cfg2 heads.

    ncu --set full --clock-control none --import-source on -k regex:decode_nms -s 4 -c 2 \
        -o gpurun_out/prof_fused python profiles/run_profile.py --workload cfg2
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from mobilenet_yolo_pytorch_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cfg2")
ap.add_argument("--iters", type=int, default=6)
ap.add_argument("--loss", action="store_true")
a = ap.parse_args()
wl = bench.WORKLOADS[a.workload]
dev = torch.device("cuda", 0)
tables = bench.anchor_tables(wl)
sets = [tuple(h.to(dev) for h in bench.make_heads(wl, wl["N"], seed=s)) for s in range(3)]
big = torch.empty(64 << 20, dtype=torch.float32, device=dev)
for i in range(a.iters):
    big.fill_(float(i))  # 256 MB write: evicts the heads from L2 between launches
    h0, h1 = sets[i % 3]
    ops.decode_nms_padded(h0, h1, tables, wl["C"], wl["conf"])
torch.cuda.synchronize()
print("done")
