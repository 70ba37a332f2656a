"""Import the UNMODIFIED reference modules from oracle/_ref (see snapshot_reference.py) or, in the build
container, straight from /root/reference.

TEST / MEASUREMENT INFRASTRUCTURE -- only tests/, bench.py's reference legs and oracle/ref_bench.py use it.

The reference's utils/__init__.py:3-9 pulls in matplotlib and `progress` through utils/logger.py:4-8 and
utils/misc.py; neither is part of the detection path and neither is installed here, so they are stubbed in
memory.  models/mobilenetv2.py:164 downloads ImageNet weights in the constructor;
`stub_weight_download()` makes that return an empty state dict, which IS BASELINE config 1's "random-init
weights".

Note quirk Q4: models/yolo_loss.py:10 and utils/box.py:4 fix a module-global `device` at import time
('cuda' when a GPU is visible).  To run the reference's CPU path on a GPU box, import it in a process
started with CUDA_VISIBLE_DEVICES="" (oracle/ref_bench.py does).
"""
from __future__ import annotations

import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
SNAP = os.path.join(HERE, "_ref")


def reference_root() -> str | None:
    """oracle/_ref when the snapshot exists, else the checkout (build container), else None."""
    if os.path.exists(os.path.join(SNAP, "models", "yolo_loss.py")):
        return SNAP
    ref = os.environ.get("REFERENCE_ROOT", "/root/reference")
    if os.path.exists(os.path.join(ref, "models", "yolo_loss.py")):
        return ref
    return None


def _stubs():
    for n in ["matplotlib", "matplotlib.pyplot", "progress", "progress.bar"]:
        sys.modules.setdefault(n, types.ModuleType(n))
    mpl = sys.modules["matplotlib"]
    if not hasattr(mpl, "use"):
        mpl.use = lambda *a, **k: None
    if not hasattr(mpl, "pyplot"):
        mpl.pyplot = sys.modules["matplotlib.pyplot"]
    bar = sys.modules["progress.bar"]
    for name in ("Bar", "IncrementalBar"):
        if not hasattr(bar, name):
            setattr(bar, name, object)


def load(root: str | None = None):
    """Returns a namespace with the reference's YOLOLoss, nms, iou module and (lazily) the model module."""
    root = root or reference_root()
    if root is None:
        raise RuntimeError("the reference is not available: no oracle/_ref snapshot and no /root/reference "
                           "(run oracle/snapshot_reference.py in the build container)")
    if root not in sys.path:
        sys.path.insert(0, root)
    _stubs()
    import models.yolo_loss as yl  # noqa: E402  (the reference's own module)
    import utils.box as box  # noqa: E402
    import utils.iou as iou  # noqa: E402
    ns = types.SimpleNamespace(root=root, yolo_loss=yl, YOLOLoss=yl.YOLOLoss, box=box, nms=box.nms, iou=iou)
    return ns


def stub_weight_download():
    """models/mobilenetv2.py:164 -> {} (no network; random-init weights, BASELINE config 1)."""
    import models.mobilenetv2 as mb
    mb.load_state_dict_from_url = lambda *a, **k: {}
    return mb


def build_voc_model(root: str | None = None):
    """models.mbv2_yolo.yolo(config) exactly as inference.py:28-47 builds it (yaml.safe_load instead of the
    Loader-less yaml.load that PyYAML 6 rejects), random-init."""
    import yaml
    ns = load(root)
    stub_weight_download()
    import models.mbv2_yolo as m
    with open(os.path.join(ns.root, "models", "voc", "config.yaml")) as f:
        config = yaml.safe_load(f)
    return m, config


def fixed_pre_maps(self, bs, is_cuda, anchors, in_w, in_h):
    """Quirk Q1: the intended meshgrid for non-square grids (yolo_loss.py:71-72 only works when H == W); equal to
    the original on square grids (asserted by tests/golden/make_golden.py)."""
    import numpy as np
    import torch
    this = torch.FloatTensor(np.array(anchors)[self.mask])
    A = self.num_mask
    anchor_wh = this.view(1, A, 1, 1, 2).expand(bs, A, in_h, in_w, 2).contiguous()
    gx = torch.linspace(0, in_w - 1, in_w).view(1, 1, 1, in_w, 1).expand(bs, A, in_h, in_w, 1)
    gy = torch.linspace(0, in_h - 1, in_h).view(1, 1, in_h, 1, 1).expand(bs, A, in_h, in_w, 1)
    g = torch.cat((gx, gy), 4).contiguous()
    if is_cuda:
        g, anchor_wh = g.cuda(), anchor_wh.cuda()
    return g, anchor_wh
