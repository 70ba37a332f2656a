#!/usr/bin/env python
"""Where the fixed cost of a K-step timed region goes (cfg2 dense / sparse): K launches issued by one C call (a) as bench.py
times them, (b) behind a ~100 us spin kernel so that all launches are queued before the GPU reaches the start event (host
launch latency excluded: diagnosis only, never a bench value), (c) the same K launches replayed from one CUDA graph.
    python profiles/fill_drain.py [--steps 20]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from mobilenet_yolo_pytorch_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, nargs="*", default=[10, 20, 40])
a = ap.parse_args()
dev = torch.device("cuda", 0)
for name in ("cfg2", "cfg2_sparse"):
    wl = bench.WORKLOADS[name]
    N = wl["N"]
    tables = bench.anchor_tables(wl)
    K = bench.cells_per_image(wl)
    R = max(4, int(np.ceil(400e6 / (N * bench.bytes_in_per_image(wl)))))
    sets = [tuple(h.to(dev) for h in bench.make_heads(wl, N, seed=r)) for r in range(R)]
    out = torch.empty((N, K, 7), dtype=torch.float32, device=dev)
    cnt = torch.empty((N,), dtype=torch.int32, device=dev)
    for steps in a.steps:
        plan = ops.BatchPlan([(sets[i % R][0], sets[i % R][1], out, cnt) for i in range(steps)], tables, wl["C"], wl["conf"])
        plan.run()
        torch.cuda.synchronize()

        def timed(fn, gate):
            best = []
            for rep in range(5):
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                if gate:
                    torch.cuda._sleep(400000)
                e0.record()
                fn()
                e1.record()
                e1.synchronize()
                best.append(e0.elapsed_time(e1) * 1e3)
            return float(np.median(best)), float(np.min(best))

        plain = timed(plan.run, False)
        gated = timed(plan.run, True)
        s = torch.cuda.Stream()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(s):
            plan.run()
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                plan.run()
        torch.cuda.synchronize()
        g.replay()
        graph = timed(g.replay, False)
        graph_gated = timed(g.replay, True)
        print(f"{name:12s} K={steps:3d}  total us (median/min of 5):  one C call {plain[0]:7.1f}/{plain[1]:7.1f}   behind a spin kernel "
              f"{gated[0]:7.1f}/{gated[1]:7.1f}   graph replay {graph[0]:7.1f}/{graph[1]:7.1f}   graph behind spin {graph_gated[0]:7.1f}/{graph_gated[1]:7.1f}"
              f"   per step: {plain[0] / steps:.2f} / {gated[0] / steps:.2f} / {graph[0] / steps:.2f} / {graph_gated[0] / steps:.2f}", flush=True)
