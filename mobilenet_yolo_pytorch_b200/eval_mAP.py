"""Drop-in for the reference's ``utils/eval_mAP.py::calculate_mAP``: the per-class
greedy detection <-> ground-truth matching and the 11-point average precision run in two kernels of
libb200yolo.so instead of a Python loop over every (class, image, detection)."""
from __future__ import annotations

from typing import List, Sequence

import numpy as np
import torch

from . import _lib, ops

IOU_THRESHOLD = 0.5  # utils/eval_mAP.py:51


def _pack(lst: Sequence[torch.Tensor], device, dtype, width=None):
    """list over images of (n_b[, width]) tensors -> one (sum n_b[, width]) device tensor + (N+1,) int32 offsets."""
    counts = [int(t.shape[0]) if t.dim() else 0 for t in lst]
    offs = np.zeros(len(lst) + 1, np.int32)
    np.cumsum(counts, out=offs[1:])
    parts = [t.to(device=device, dtype=dtype).reshape((-1, width) if width else (-1,)) for t, n in zip(lst, counts) if n]
    shape = (0, width) if width else (0,)
    data = torch.cat(parts, 0).contiguous() if parts else torch.empty(shape, dtype=dtype, device=device)
    return data, torch.from_numpy(offs).to(device), int(offs[-1])


def map_eval(det_boxes, det_labels, det_scores, det_off, true_boxes, true_labels, true_difficulties, true_off,
             n_classes: int, iou_thr: float = IOU_THRESHOLD):
    """b200yolo_map_eval on packed device tensors: returns (ap, tp_sum, fp_sum), each (n_classes-1,) float32 on the
    device."""
    dev = det_off.device
    if dev.type != "cuda":
        raise RuntimeError("map_eval needs CUDA tensors: the b200yolo kernels have no CPU fallback")
    D, T, N = int(det_boxes.shape[0]), int(true_boxes.shape[0]), int(det_off.shape[0]) - 1
    thr = np.ascontiguousarray(torch.arange(start=0, end=1.1, step=.1).numpy().astype(np.float32))  # eval_mAP.py:120
    lib = _lib.load()
    with torch.cuda.device(dev):
        out = torch.empty((3, n_classes - 1), dtype=torch.float32, device=dev)
        ws_bytes = int(lib.b200yolo_map_eval_workspace_bytes(D, T, N, n_classes))
        ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
        _lib.check(lib.b200yolo_map_eval(
            det_boxes.data_ptr(), det_labels.data_ptr(), det_scores.data_ptr(), det_off.data_ptr(), D,
            true_boxes.data_ptr(), true_labels.data_ptr(), true_difficulties.data_ptr(), true_off.data_ptr(), T,
            N, n_classes, float(np.float32(iou_thr)), thr.ctypes.data, len(thr), out[0].data_ptr(), out[1].data_ptr(),
            out[2].data_ptr(), ws.data_ptr(), ws_bytes, ops._stream(det_off)))
    return out[0], out[1], out[2]


def calculate_mAP(det_boxes: List[torch.Tensor], det_labels, det_scores, true_boxes, true_labels, true_difficulties,
                  classes_name):
    """Same arguments and return value as utils/eval_mAP.py:134-188: lists over images; returns
    (average_precisions dict, mean_average_precision, class_true_positive dict, class_false_positive dict)."""
    assert len(det_boxes) == len(det_labels) == len(det_scores) == len(true_boxes) == len(true_labels) == len(
        true_difficulties)
    n_classes = len(classes_name)
    device = next((t.device for t in list(det_boxes) + list(true_boxes) if isinstance(t, torch.Tensor) and t.is_cuda), None)
    if device is None:
        raise RuntimeError("calculate_mAP needs CUDA tensors: the b200yolo kernels have no CPU fallback")
    db, doff, _ = _pack(det_boxes, device, torch.float32, 4)
    dl, _, _ = _pack(det_labels, device, torch.int32)
    ds, _, _ = _pack(det_scores, device, torch.float32)
    tb, toff, _ = _pack(true_boxes, device, torch.float32, 4)
    tl, _, _ = _pack(true_labels, device, torch.int32)
    td, _, _ = _pack(true_difficulties, device, torch.uint8)
    ap, tp, fp = map_eval(db, dl, ds, doff, tb, tl, td, toff, n_classes)
    host = torch.stack((ap, tp, fp)).cpu()                      # the one D2H copy
    names = list(classes_name)
    average_precisions = {names[c + 1]: v for c, v in enumerate(host[0].tolist())}
    class_true_positive = {names[c + 1]: v for c, v in enumerate(host[1].tolist())}
    class_false_positive = {names[c + 1]: v for c, v in enumerate(host[2].tolist())}
    return average_precisions, host[0].mean().item(), class_true_positive, class_false_positive
