#!/usr/bin/env python
"""e2e (host buffers -> host detections) rate of b200yolo_decode_nms_host on cfg2.  The chunk size comes from the
B200YOLO_HOST_CHUNK environment variable (read once per process)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from mobilenet_yolo_pytorch_b200 import ops

wl = bench.WORKLOADS["cfg2"]
N, C = wl["N"], wl["C"]
tables = bench.anchor_tables(wl)
K = bench.cells_per_image(wl)
hh0, hh1 = bench.make_heads(wl, N, seed=7, pin=True)
ho = torch.empty((N, K, 7), dtype=torch.float32).pin_memory()
hc = torch.empty((N,), dtype=torch.int32).pin_memory()
for _ in range(5):
    ops.decode_nms_host(hh0, hh1, tables, C, wl["conf"], device=0, out=ho, out_count=hc)
rates = []
for rep in range(5):
    t0 = time.perf_counter()
    for _ in range(20):
        ops.decode_nms_host(hh0, hh1, tables, C, wl["conf"], device=0, out=ho, out_count=hc)
    rates.append(N * 20 / (time.perf_counter() - t0))
print("chunk", os.environ.get("B200YOLO_HOST_CHUNK", "default (N/4)"), " ".join(f"{r / 1e3:.0f}k" for r in rates), "img/s")
# pure copy reference: the same bytes, one cudaMemcpy each way
d0, d1 = torch.empty_like(hh0, device="cuda"), torch.empty_like(hh1, device="cuda")
do = torch.empty_like(ho, device="cuda")
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    d0.copy_(hh0, non_blocking=True)
    d1.copy_(hh1, non_blocking=True)
    ho.copy_(do, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 20
print(f"copies only (46.5 MB H2D + 13 MB D2H, one stream): {dt * 1e3:.3f} ms -> {N / dt / 1e3:.0f}k img/s bound (serial), "
      f"H2D alone would be {46.464e6 / dt / 1e9:.1f} GB/s if overlapped")
