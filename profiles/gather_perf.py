#!/usr/bin/env python
"""Cost of the gather code path by itself: PeerGather.run_steps with a world of ONE rank (local stores only, plus the
signal / wait launches) against the plain batched launches, cfg2 dense and sparse.
    python profiles/gather_perf.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import bench
from mobilenet_yolo_pytorch_b200 import ops
from mobilenet_yolo_pytorch_b200 import dist as b2dist

os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29533")
dist.init_process_group("gloo", rank=0, world_size=1)
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
steps = 200
for name in ("cfg2", "cfg2_sparse"):
    wl = bench.WORKLOADS[name]
    N, K = wl["N"], bench.cells_per_image(wl)
    tables = bench.anchor_tables(wl)
    R = 9
    sets = [tuple(h.to(dev) for h in bench.make_heads(wl, N, seed=r)) for r in range(R)]
    out = torch.empty((N, K, 7), dtype=torch.float32, device=dev)
    cnt = torch.empty((N,), dtype=torch.int32, device=dev)
    plan = ops.BatchPlan([(sets[i % R][0], sets[i % R][1], out, cnt) for i in range(steps)], tables, wl["C"], wl["conf"])
    pg = b2dist.PeerGather(N, K, device=dev)
    gplan = pg.run_steps([sets[i % R] for i in range(steps)], tables, wl["C"], wl["conf"])
    torch.cuda.synchronize()

    def t(fn):
        fn()
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1) / steps * 1e3)
        return best
    a = t(lambda: plan.run())
    b = t(lambda: pg.run_steps(None, None, wl["C"], wl["conf"], plan=gplan))
    print(f"{name}: plain {a:.2f} us/step, gather path (1 rank: local stores + signal/wait launches) {b:.2f} us/step", flush=True)
    pg.close()
dist.destroy_process_group()
