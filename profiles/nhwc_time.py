#!/usr/bin/env python
"""Channels-last heads: the fused kernel consuming them directly (b200yolo_decode_nms_nhwc) against an NCHW copy
(.contiguous()) followed by the planar kernel, and against the planar kernel on NCHW heads."""
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
    nchw = [tuple(h.to(dev) for h in bench.make_heads(wl, N, seed=s)) for s in range(R)]
    nhwc = [tuple(h.contiguous(memory_format=torch.channels_last) for h in hs) for hs in nchw]
    out = torch.empty((N, bench.cells_per_image(wl), 7), dtype=torch.float32, device=dev)
    cnt = torch.empty((N,), dtype=torch.int32, device=dev)

    def planar(i):
        ops.decode_nms_padded(nchw[i % R][0], nchw[i % R][1], tables, C, wl["conf"], out=out, out_count=cnt)

    def direct(i):
        ops.decode_nms_padded(nhwc[i % R][0], nhwc[i % R][1], tables, C, wl["conf"], out=out, out_count=cnt)

    def copy_then_planar(i):
        ops.decode_nms_padded(nhwc[i % R][0].contiguous(), nhwc[i % R][1].contiguous(), tables, C, wl["conf"], out=out, out_count=cnt)

    for label, fn in (("NCHW heads, planar kernel", planar), ("channels-last heads, direct", direct),
                      ("channels-last heads, NCHW copy + planar", copy_then_planar)):
        for i in range(10):
            fn(i)
        torch.cuda.synchronize()
        best = 1e9
        for rep in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(200):
                fn(i)
            e1.record()
            e1.synchronize()
            best = min(best, e0.elapsed_time(e1) * 5)
        print(f"{name:12s} {label:42s} {best:7.2f} us/step")
