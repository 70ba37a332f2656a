"""Drop-in for the reference's ``utils/box.py`` (``nms``, ``wh_to_x2y2``)."""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def wh_to_x2y2(bbox):
    """utils/box.py:6-10 (in place, same operation order)."""
    bbox[..., 0] = bbox[..., 0] - bbox[..., 2] / 2
    bbox[..., 1] = bbox[..., 1] - bbox[..., 3] / 2
    bbox[..., 2] = bbox[..., 2] + bbox[..., 0]
    bbox[..., 3] = bbox[..., 3] + bbox[..., 1]


def _pad(lst, device):
    """Generic python list of (n_b,7) tensors -> (N, stride, 7) + counts."""
    n = [int(t.shape[0]) for t in lst]
    stride = max(n + [1])
    padded = torch.zeros((len(lst), stride, 7), dtype=torch.float32, device=device)
    for b, t in enumerate(lst):
        if n[b]:
            padded[b, :n[b]] = t.to(device=device, dtype=torch.float32)
    return padded, torch.tensor(n, dtype=torch.int32, device=device)


def nms(preds, num_classes, return_indices: bool = False):
    """utils.box.nms(preds, num_classes) (utils/box.py:11-31): ``preds`` is a pair of
    per-head candidate lists; returns list[N] of (k_b, 7) detections, class-ascending
    blocks in descending ``conf*class_score`` order, IoU threshold 0.45.

    Lists produced by this package's ``YOLOLoss.forward`` carry their padded device
    buffer and are consumed without repacking; any other list of CUDA tensors is
    packed first."""
    assert len(preds) == 2  # only do two layers yolo (box.py:13)
    assert len(preds[0]) == len(preds[1])
    p0, p1 = preds
    if len(p0) == 0:
        return ([], []) if return_indices else []
    # both heads still undecoded (YOLOLoss.lazy_eval): the fused kernel does decode + NMS in one launch
    if (isinstance(p0, ops.LazyCandidates) and isinstance(p1, ops.LazyCandidates) and p0.pending and p1.pending
            and p0.conf_thr == p1.conf_thr and p0.num_classes == p1.num_classes == num_classes
            and p0.head.shape[:2] == p1.head.shape[:2] and p0.head.device == p1.head.device):
        tables = np.stack((p0.anchor_wh, p1.anchor_wh)).astype(np.float32)
        res = ops.decode_nms_padded(p0.head, p1.head, tables, num_classes, p0.conf_thr, want_idx=return_indices)
        counts = res[1].cpu().tolist()
        dets = [res[0][b, :k] for b, k in enumerate(counts)]
        if return_indices:
            return dets, [res[2][b, :k] for b, k in enumerate(counts)]
        return dets
    p0 = p0.materialise() if isinstance(p0, ops.LazyCandidates) else p0
    p1 = p1.materialise() if isinstance(p1, ops.LazyCandidates) else p1
    device = next((t.device for t in list(p0) + list(p1) if isinstance(t, torch.Tensor) and t.is_cuda), None)
    if device is None:
        raise RuntimeError("nms needs CUDA tensors: the b200yolo kernels have no CPU fallback")
    packed = []
    for p in (p0, p1):
        if isinstance(p, ops.CandidateList):
            packed.append((p.padded, p.counts))
        else:
            packed.append(_pad(p, device))
    out, oc, oi = ops.nms_padded(packed[0][0], packed[0][1], packed[1][0], packed[1][1], num_classes, want_idx=True)
    counts = oc.cpu().tolist()
    dets = [out[b, :k] for b, k in enumerate(counts)]
    if return_indices:
        return dets, [oi[b, :k] for b, k in enumerate(counts)]
    return dets
