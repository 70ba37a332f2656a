"""mobilenet_yolo_pytorch_b200 -- the B200 (sm_100a) detection hot path of
eric612/Mobilenet-YOLO-Pytorch behind the reference's own entry points:

    YOLOLoss                      <- models/yolo_loss.py::YOLOLoss
    nms, wh_to_x2y2               <- utils/box.py
    find_intersection/union/jaccard_overlap  <- utils/iou.py
    decode_nms                    <- the inference branch of models/mbv2_yolo.py:158-160, fused
    calculate_mAP                 <- utils/eval_mAP.py
    SegLoss                       <- models/seg_loss.py

All computation happens in libb200yolo.so (hand-written CUDA, C ABI in
include/b200yolo.h); PyTorch only provides device memory, streams and
torch.distributed.  There is no CPU fallback.
"""
from . import _lib, dist, ops
from .box import nms, wh_to_x2y2
from .eval_mAP import calculate_mAP
from .fused import adjust_confidence, decode_nms, decode_nms_padded, head_anchor_table, patch_reference
from .iou import find_intersection, find_jaccard_overlap, find_union
from .seg_loss import SegLoss
from .yolo_loss import YOLOLoss

__all__ = ["YOLOLoss", "nms", "wh_to_x2y2", "find_intersection", "find_union", "find_jaccard_overlap", "decode_nms",
           "decode_nms_padded", "head_anchor_table", "patch_reference", "ops", "dist", "calculate_mAP", "SegLoss", "adjust_confidence"]
