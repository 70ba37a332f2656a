#!/usr/bin/env python
"""BASELINE.json configs[0]: the REAL models/mbv2_yolo.py (MobileNetV2-YOLO, VOC 20 classes, 352x352, batch 1,
random-init weights, synthetic image) driven the way inference.py:109-126 drives it, from the unmodified
reference files (oracle/_ref snapshot or /root/reference).

TEST / MEASUREMENT INFRASTRUCTURE -- used by tests/test_reference_integration.py and bench.py's reference legs.

    CUDA_VISIBLE_DEVICES="" python oracle/ref_config1.py --out /tmp/cfg1_cpu.npz      # the reference's CPU path
    python oracle/ref_config1.py --out /tmp/cfg1_cuda.npz                             # its CUDA path (quirk Q4)

`build_model(patch=None)`: `patch` is called with the reference's modules BEFORE the model is constructed, so
that a drop-in (`mobilenet_yolo_pytorch_b200.patch_reference`) replaces `YOLOLoss` / `nms` inside
models/mbv2_yolo.py exactly as INTEGRATION.md describes; this file itself never imports the product.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def make_image(n=1, size=352):
    import torch
    g = torch.Generator().manual_seed(0)
    return torch.rand(n, 3, size, size, generator=g) - 0.5


def build_model(patch=None, seed=0, val_conf=0.3):
    """-> (model.eval() on CPU, config, reference namespace).  Same seed => same random-init weights."""
    import torch
    from oracle import ref_loader
    ns = ref_loader.load()
    m, config = ref_loader.build_voc_model()
    if patch is not None:
        patch(ns, m)
    torch.manual_seed(seed)
    model = m.yolo(config=config)          # inference.py:38
    model.eval()                           # :44
    model.yolo_losses[0].val_conf = val_conf   # :46-47
    model.yolo_losses[1].val_conf = val_conf
    return model, config, ns


def capture_heads(model):
    """Forward hooks on the two head convolutions (mbv2_yolo.py:144,153): the tensors the hot path consumes."""
    got = {}
    h0 = model.yolo_headS32.register_forward_hook(lambda mod, inp, out: got.__setitem__("out0", out.detach()))
    h1 = model.yolo_headS16.register_forward_hook(lambda mod, inp, out: got.__setitem__("out1", out.detach()))
    return got, (h0, h1)


def run(device: str, reps: int = 5, patch=None):
    import torch
    model, config, ns = build_model(patch)
    dev = torch.device(device)
    model = model.to(dev)
    x = make_image().to(dev)
    got, hooks = capture_heads(model)
    with torch.no_grad():
        dets = model(x)                    # inference.py:121
    for h in hooks:
        h.remove()
    times = []
    with torch.no_grad():
        for _ in range(reps):
            if device == "cuda":
                torch.cuda.synchronize()
            t0 = time.perf_counter()
            d = model(x)
            n = int(d[0].shape[0])  # (forces the result like inference.py's drawing loop does)
            if device == "cuda":
                torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
    return {"dets": dets[0].detach().cpu().numpy(), "out0": got["out0"].cpu().numpy(), "out1": got["out1"].cpu().numpy(),
            "latency_ms": 1e3 * statistics.median(times), "n": n}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    import numpy as np
    import torch
    device = "cuda" if torch.cuda.is_available() else "cpu"
    r = run(device, a.reps)
    if a.out:
        np.savez(a.out, dets=r["dets"], out0=r["out0"], out1=r["out1"])
    print(json.dumps({"config": "MobileNetV2-YOLO 352x352 VOC, batch 1, random-init (seed 0), val_conf 0.3",
                      "device": device, "latency_ms": r["latency_ms"], "detections": int(r["dets"].shape[0]),
                      "threads": torch.get_num_threads()}), flush=True)


if __name__ == "__main__":
    main()
