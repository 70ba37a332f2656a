#!/usr/bin/env python
"""SURVEY section 7's first-choice parity test, as a measurement: the unmodified reference on device='cuda' (oracle/_ref:
ATen elementwise kernels + torchvision's CUDA nms) against this library on the same head tensors, compared BIT FOR BIT,
with the default decode (SFU sigmoid / exp, multiplication by 1/W) and with flag 128 ("exact": IEEE sigmoid / expf, true
division), and what the exact decode costs.
    python profiles/exact_decode.py"""
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from mobilenet_yolo_pytorch_b200 import _lib, ops
from oracle import ref_loader

ref = ref_loader.load()
dev = torch.device("cuda", 0)
lib = _lib.load()
for name, N in (("cfg2", 64), ("cfg2_sparse", 64), ("cfg3", 16)):
    wl = bench.WORKLOADS[name]
    C = wl["C"]
    tables = bench.anchor_tables(wl)
    h0, h1 = [h.to(dev) for h in bench.make_heads(wl, N, seed=0)]
    losses = []
    for i in range(2):
        l = ref.YOLOLoss(wl["anchors"], bench.MASK[i], C, wl["img"], 0.5, 0.5, val_conf=wl["conf"])
        if any(H != W for (H, W) in wl["grids"]):
            l.pre_maps = types.MethodType(ref_loader.fixed_pre_maps, l)
        losses.append(l)
    with torch.no_grad():
        preds = [losses[0](h0), losses[1](h1)]
        want = ref.nms(preds, C)
    for flags, label in ((0, "default"), (128, "exact (flag 128)")):
        lib.b200yolo_debug_set_flags(flags)
        r0, c0, i0 = ops.decode_head_padded(h0, tables[0], C, wl["conf"], want_ids=True)
        r1, c1, i1 = ops.decode_head_padded(h1, tables[1], C, wl["conf"], want_ids=True)
        out, cnt = ops.decode_nms_padded(h0, h1, tables, C, wl["conf"])
        torch.cuda.synchronize()
        same_cand = same_bits = tot = 0
        for b in range(N):
            for rows, cc, pr in ((r0, c0, preds[0]), (r1, c1, preds[1])):
                k = int(cc[b])
                if k == pr[b].shape[0]:
                    same_cand += 1
                    same_bits += int((rows[b, :k] == pr[b]).sum())
                tot += pr[b].numel()
        same_img = sum(int(int(cnt[b]) == want[b].shape[0] and torch.equal(out[b, :int(cnt[b])], want[b])) for b in range(N))
        close_img = sum(int(int(cnt[b]) == want[b].shape[0] and torch.allclose(out[b, :int(cnt[b])], want[b], rtol=1e-5, atol=1e-6)) for b in range(N))
        # cost: the fused kernel, overlapped launches
        K = bench.cells_per_image(wl)
        o = torch.empty((N, K, 7), dtype=torch.float32, device=dev)
        c = torch.empty((N,), dtype=torch.int32, device=dev)
        plan = ops.BatchPlan([(h0, h1, o, c)] * 100, tables, C, wl["conf"])
        plan.run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); plan.run(); e1.record(); e1.synchronize()
        us = e0.elapsed_time(e1) * 10
        print(f"{name:12s} {label:18s}: candidate lists with the reference's length {same_cand}/{2 * N}, decoded floats bit-equal "
              f"{same_bits}/{tot} ({100.0 * same_bits / max(tot, 1):.4f} %), images whose final detections are bit-equal "
              f"{same_img}/{N} (within 1e-5: {close_img}/{N}), fused kernel {us:.1f} us per {N}-image launch (L2-resident heads)", flush=True)
    lib.b200yolo_debug_set_flags(0)
