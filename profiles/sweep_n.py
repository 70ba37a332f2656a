#!/usr/bin/env python
"""Kernel time of the fused decode+NMS launch as a function of the batch size (how CTAs share an SM):
    python profiles/sweep_n.py [workload ...]
CUDA-event timing, rotating input sets larger than L2."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from mobilenet_yolo_pytorch_b200 import ops

dev = torch.device("cuda", 0)
names = [a for a in sys.argv[1:] if not a.startswith("-")] or ["cfg2", "cfg2_sparse"]
for name in names:
    wl = bench.WORKLOADS[name]
    tables = bench.anchor_tables(wl)
    K = bench.cells_per_image(wl)
    for N in (74, 148, 222, 256, 296, 444, 592, 1184, 2368):
        per = N * bench.bytes_in_per_image(wl)
        R = max(3, int(np.ceil(300e6 / per)))
        sets = [tuple(h.to(dev) for h in bench.make_heads(wl, N, seed=100 + r)) for r in range(R)]
        out = torch.empty((N, K, 7), dtype=torch.float32, device=dev)
        cnt = torch.empty((N,), dtype=torch.int32, device=dev)
        for i in range(5):
            ops.decode_nms_padded(sets[i % R][0], sets[i % R][1], tables, wl["C"], wl["conf"], out=out, out_count=cnt)
        torch.cuda.synchronize()
        steps = 60
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            ops.decode_nms_padded(sets[i % R][0], sets[i % R][1], tables, wl["C"], wl["conf"], out=out, out_count=cnt)
        e1.record()
        e1.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / steps
        kept = float(cnt.sum().item())
        gbs = (per + 28 * kept + 4 * N) / (us * 1e-6) / 1e9
        print(f"{name:12s} N={N:5d}  {us:8.2f} us/launch  {N / us:7.2f} img/us  {us / N * 148:6.2f} us*SM/img  {gbs:7.1f} GB/s algorithmic")
        del sets
