#!/usr/bin/env python
"""Recipe: place the UNMODIFIED reference files of the detection hot path under oracle/_ref/.

TEST / MEASUREMENT INFRASTRUCTURE -- nothing under mobilenet_yolo_pytorch_b200/ imports this.

The reference is pure Python, so there is nothing to compile: "building" the real
reference for this path means making its few source files importable where /root/reference does not
exist (the GPU box).  oracle/_ref/ is git-ignored (the files never enter this repository's history) but
not gpurun-ignored, so the snapshot travels with the working tree like the built .so files do.  It is
used for exactly two things:

  * bench.py --impl reference / extra.reference_cuda: timing the reference's own CPU and CUDA paths on
    the GPU box;
  * tests/test_reference_integration.py: running the real models/mbv2_yolo.py (BASELINE config 1) patched
    and unpatched.

Run by __graft_entry__.build() whenever /root/reference is present; a no-op elsewhere.  Files are copied
byte for byte; their sha256 is written to oracle/_ref/MANIFEST.json so a stale or edited snapshot is seen.
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
REF = os.environ.get("REFERENCE_ROOT", "/root/reference")

# the path's files + the caller of BASELINE config 1 and what it imports
FILES = [
    "models/__init__.py",
    "models/yolo_loss.py",       # A1-A4, A9-A16
    "models/seg_loss.py",        # f4 (imported by mbv2_yolo.py:7)
    "models/mbv2_yolo.py",       # A17: yolo.forward, the call site :158-160
    "models/mobilenetv2.py",     # backbone of config 1 (cuDNN work, unchanged)
    "models/voc/config.yaml",
    "models/bdd100k/config.yaml",
    "utils/__init__.py",
    "utils/box.py",              # A5-A6
    "utils/iou.py",              # A8
    "utils/eval_mAP.py",         # f2
    "utils/misc.py",             # imported by utils/__init__.py:3
    "utils/logger.py",           # imported by utils/__init__.py:4
    "LICENSE",
]


def sha256(path: str) -> str:
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def snapshot(force: bool = False) -> bool:
    """Copy FILES from the reference checkout.  Returns True when oracle/_ref is usable afterwards."""
    if not os.path.isdir(REF):
        return os.path.exists(os.path.join(DEST, "MANIFEST.json"))
    manifest = {}
    for rel in FILES:
        src = os.path.join(REF, rel)
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if force or not os.path.exists(dst) or sha256(dst) != sha256(src):
            shutil.copyfile(src, dst)
        manifest[rel] = sha256(dst)
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": "eric612/Mobilenet-YOLO-Pytorch (unmodified files, see LICENSE)", "sha256": manifest}, f, indent=1)
    return True


if __name__ == "__main__":
    print("oracle/_ref ready" if snapshot(force=True) else "no reference checkout and no snapshot")
