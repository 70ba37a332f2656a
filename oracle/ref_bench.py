#!/usr/bin/env python
"""Time the UNMODIFIED reference (oracle/_ref snapshot or /root/reference) on this box: decode + NMS
(`YOLOLoss.forward(input)` x2 + `utils.box.nms`, the call sequence of models/mbv2_yolo.py:158-160) or the loss
(`YOLOLoss.forward(input, targets)` x2).  Prints ONE JSON line.

TEST / MEASUREMENT INFRASTRUCTURE -- executed only by bench.py's reference legs, as a subprocess:

    CUDA_VISIBLE_DEVICES="" python oracle/ref_bench.py --device cpu  --workload cfg2 --n 64 --reps 5
    python oracle/ref_bench.py --device cuda --workload cfg2 --n 64 --reps 5     # torchvision's sm_100 NMS kernel

The reference picks its device at import (quirk Q4: models/yolo_loss.py:10, utils/box.py:4), hence the separate
process per device.  `--what nms_only` times utils.box.nms alone on pre-decoded candidates (the torchvision
kernel driven per (image, class) exactly as utils/box.py:16-29 drives it) next to ONE torchvision
`batched_nms` call over the whole batch -- the strongest way to use that kernel, not the reference's.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--device", default="cpu", choices=["cpu", "cuda"])
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--what", default="decode_nms", choices=["decode_nms", "nms_only", "loss"])
    ap.add_argument("--n", type=int, default=64)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--budget", type=float, default=60.0, help="stop repeating after this many seconds")
    a = ap.parse_args()

    import torch
    if a.device == "cpu" and torch.cuda.is_available():
        raise SystemExit("ref_bench --device cpu must run with CUDA_VISIBLE_DEVICES='' (reference quirk Q4)")
    if a.device == "cuda" and not torch.cuda.is_available():
        raise SystemExit("ref_bench --device cuda: no CUDA device")
    if a.threads > 0:
        torch.set_num_threads(a.threads)
    import bench  # workload tables + the seeded head generator (no GPU work at import)
    from oracle import ref_loader
    ref = ref_loader.load()
    dev = torch.device(a.device)

    if a.what == "loss":
        wl = bench.WORKLOADS["cfg2"]
    else:
        wl = bench.WORKLOADS[a.workload]
    N, C = a.n, wl["C"]
    heads = [h.to(dev) for h in bench.make_heads(wl, N, seed=0)]
    square = all(H == W for (H, W) in wl["grids"])
    losses = []
    for i in range(2):
        if a.what == "loss":
            l = ref.YOLOLoss(wl["anchors"], bench.MASK[i], C, wl["img"], bench.VOC_IGNORE[i], bench.VOC_IOU_THRESH,
                             iou_weighting=bench.VOC_IOU_WEIGHTING)
        else:
            l = ref.YOLOLoss(wl["anchors"], bench.MASK[i], C, wl["img"], 0.5, 0.5, val_conf=wl["conf"])
        if not square:
            l.pre_maps = types.MethodType(ref_loader.fixed_pre_maps, l)   # quirk Q1
        losses.append(l)

    def sync():
        if a.device == "cuda":
            torch.cuda.synchronize()

    extra = {}
    if a.what == "decode_nms":
        def run():
            with torch.no_grad():
                preds = [losses[i](heads[i]) for i in range(2)]      # mbv2_yolo.py:158
                return ref.nms(preds, C)                              # :160
    elif a.what == "nms_only":
        with torch.no_grad():
            preds = [losses[i](heads[i]) for i in range(2)]

        def run():
            with torch.no_grad():
                return ref.nms(preds, C)
    else:
        targets = [torch.from_numpy(t) for t in bench.make_targets(N, 100, C, 1)]

        def run():
            with torch.no_grad():
                return [losses[i](heads[i], targets) for i in range(2)]

    out = run()  # warm-up (allocator, cuDNN/torchvision kernels, thread pool)
    sync()
    times = []
    t_begin = time.perf_counter()
    for _ in range(max(1, a.reps)):
        t0 = time.perf_counter()
        out = run()
        sync()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_begin > a.budget:
            break
    med = statistics.median(times)
    if a.what != "loss":
        extra["kept_rows_per_image"] = sum(int(o.shape[0]) for o in out) / N
    if a.what == "nms_only":
        # one torchvision batched_nms over the whole batch (image*C + class as the category): the best case for
        # torchvision's kernel, NOT what the reference does
        import torchvision
        per = [torch.cat((preds[0][b], preds[1][b]), 0) for b in range(N)]
        allr = torch.cat(per, 0)
        img = torch.cat([torch.full((p.shape[0],), b, dtype=torch.int64, device=dev) for b, p in enumerate(per)])
        cat = img * C + allr[:, 6].to(torch.int64)
        sc = allr[:, 5] * allr[:, 4]
        torchvision.ops.batched_nms(allr[:, :4], sc, cat, 0.45)
        sync()
        bt = []
        for _ in range(max(3, a.reps)):
            t0 = time.perf_counter()
            keep = torchvision.ops.batched_nms(allr[:, :4], sc, cat, 0.45)
            sync()
            bt.append(time.perf_counter() - t0)
        extra["torchvision_batched_nms"] = {"images_per_s": N / statistics.median(bt), "ms": 1e3 * statistics.median(bt),
                                            "kept_rows_per_image": int(keep.numel()) / N, "candidates": int(allr.shape[0])}
    import torchvision
    print(json.dumps({
        "what": a.what, "device": a.device, "workload": a.workload if a.what != "loss" else "cfg4_loss", "n": N,
        "reps": len(times), "images_per_s": N / med, "median_s": med, "min_s": min(times), "max_s": max(times),
        "threads": torch.get_num_threads(), "cores": len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count(),
        "torch": torch.__version__, "torchvision": torchvision.__version__, "reference_root": ref.root, **extra,
    }), flush=True)


if __name__ == "__main__":
    main()
