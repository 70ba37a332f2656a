#!/usr/bin/env python
"""HBM rate of the seg-head loss kernel (b200yolo_seg_loss): input (N,C,H,W) + truth (N,H,W,C) read once."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from mobilenet_yolo_pytorch_b200 import seg_loss

dev = torch.device("cuda", 0)
for (N, C, H, W) in ((32, 3, 24, 40), (256, 3, 96, 160), (256, 1, 384, 640)):
    R = 6
    xs = [torch.randn(N, C, H, W, device=dev) for _ in range(R)]
    ts = [(torch.rand(N, H, W, C, device=dev) < 0.3).float() for _ in range(R)]
    for i in range(5):
        seg_loss._seg_sums(xs[i % R], ts[i % R])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(60):
        seg_loss._seg_sums(xs[i % R], ts[i % R])
    e1.record()
    e1.synchronize()
    us = e0.elapsed_time(e1) / 60 * 1e3
    by = 8.0 * N * C * H * W
    print(f"seg_loss N={N} C={C} {H}x{W}: {us:8.1f} us/step, {by / 1e6:7.1f} MB -> {by / us / 1e3:7.1f} GB/s")
