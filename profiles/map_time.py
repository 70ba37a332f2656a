#!/usr/bin/env python
"""Time b200yolo_map_eval on a VOC07-test-sized evaluation (4952 images, 20 classes, ~2.5 objects and DETS
detections per image) against the CPU oracle on a bounded sample.  Usage: map_time.py [N] [DETS]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import oracle
from mobilenet_yolo_pytorch_b200 import eval_mAP

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4952
DETS = int(sys.argv[2]) if len(sys.argv) > 2 else 40
C = 21
r = np.random.RandomState(0)
L = [[] for _ in range(6)]
for b in range(N):
    ng, nd = r.randint(1, 5), r.randint(DETS // 2, DETS + 1)
    g = np.sort(r.rand(ng, 2, 2), axis=1).reshape(ng, 4).astype(np.float32)
    gl = r.randint(1, C, ng)
    db = np.sort(r.rand(nd, 2, 2), axis=1).reshape(nd, 4).astype(np.float32)
    dl = r.randint(1, C, nd)
    pick, near = r.randint(0, ng, nd), r.rand(nd) < 0.5
    db[near] = g[pick[near]] + r.randn(int(near.sum()), 4).astype(np.float32) * 0.02
    dl[near] = gl[pick[near]]
    for lst, v in zip(L, (db, dl.astype(np.int64), r.rand(nd).astype(np.float32), g, gl.astype(np.int64), np.zeros(ng, np.uint8))):
        lst.append(v)
dev = torch.device("cuda", 0)
t = [[torch.from_numpy(x).to(dev) for x in lst] for lst in L]
db, doff, D = eval_mAP._pack(t[0], dev, torch.float32, 4)
dl, _, _ = eval_mAP._pack(t[1], dev, torch.int32)
ds, _, _ = eval_mAP._pack(t[2], dev, torch.float32)
tb, toff, T = eval_mAP._pack(t[3], dev, torch.float32, 4)
tl, _, _ = eval_mAP._pack(t[4], dev, torch.int32)
td, _, _ = eval_mAP._pack(t[5], dev, torch.uint8)
for _ in range(3):
    ap, tp, fp = eval_mAP.map_eval(db, dl, ds, doff, tb, tl, td, toff, C)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ap, tp, fp = eval_mAP.map_eval(db, dl, ds, doff, tb, tl, td, toff, C)
e1.record()
e1.synchronize()
ms = e0.elapsed_time(e1) / 20
n_cpu = min(N, 200)
t0 = time.perf_counter()
o = oracle.calculate_map(*[lst[:n_cpu] for lst in L], C)
cpu_s = time.perf_counter() - t0
print(f"map_eval: N={N} images, D={D} detections, T={T} objects, {C - 1} classes: {ms * 1e3:.1f} us per evaluation (2 kernels); "
      f"mAP {ap.mean().item():.4f}; numpy oracle on {n_cpu} images: {cpu_s:.2f} s ({cpu_s / n_cpu * N:.1f} s scaled to N)")
