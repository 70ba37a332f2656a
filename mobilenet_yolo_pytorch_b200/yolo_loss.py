"""Drop-in for the reference's ``models/yolo_loss.py::YOLOLoss`` (constructor,
mutable attributes and both ``forward`` return conventions, yolo_loss.py:33-50,
206-241) whose work is done by the sm_100a kernels in libb200yolo.so."""
from __future__ import annotations

from typing import Optional

import torch
import torch.nn as nn

from . import _lib, ops


class YOLOLoss(nn.Module):
    lazy_eval = False

    """Same signature as the reference (yolo_loss.py:33).

    ``forward(input)``            -> list[N] of (n_b, 7) rows
        ``[x1, y1, x2, y2, conf, class_score, class_index]`` with ``conf > val_conf``
        in (a, j, i) row-major order (yolo_loss.py:180-204).
    ``forward(input, targets)``   -> ``(loss, recall, avg_iou, obj, no_obj, cls_score, count)``
        in the order of yolo_loss.py:236; ``targets`` is the reference's list[N] of
        CPU ``(n_b, 5)`` tensors ``[cls(1-based), cx, cy, w, h]``.

    ``val_conf``, ``img_size``, ``ignore_threshold``, ``iou_thresh``, ``iou_weighting``,
    ``anchors`` and ``mask`` stay plain mutable attributes because the reference's
    callers overwrite them (mbv2_yolo.py:139-140, inference.py:46-47, train.py:149-150,
    417-418).

    ``lazy_eval`` (class or instance attribute, default False): ``forward(input)`` then returns an
    ``ops.LazyCandidates`` -- nothing is launched until it is used, and ``utils.box.nms`` on the two heads' results
    (models/mbv2_yolo.py:158-160) runs the fused decode + NMS kernel, so the reference's own call sites get the
    one-launch path.  ``patch_reference(..., fuse_inference=True)`` switches it on.

    ``lazy_stats`` (attribute, default False): ``forward(input, targets)`` then performs NO host synchronisation: the
    loss and the six statistics come back as 0-dim float32 device tensors (computed by ``b200yolo_loss_finalize_dev``),
    and an out-of-range ground-truth box -- the reference's IndexError -- is reported by ``check()`` or by the next
    ``forward`` call instead of this one.  ``targets`` may also be an ``ops.PackedTargets`` (ground truth already on the
    device): together they take a 512-image step from 0.8 ms of host work to the kernels' own time.

    ``process_group``: optional torch.distributed group.  When set, the batch is a
    shard of a data-parallel batch and the 16 partial sums are all-reduced (NCCL)
    before the batch-global normalisation of yolo_loss.py:55,224,170-178.
    """

    def __init__(self, anchors, mask, num_classes, img_size, ignore_threshold, iou_thresh, val_conf=0.1,
                 iou_weighting=0.01, process_group=None):
        super().__init__()
        self.anchors = anchors
        self.mask = mask
        self.num_mask = len(mask)
        self.num_anchors = len(anchors)
        self.num_classes = num_classes
        self.bbox_attrs = 5 + num_classes
        self.img_size = img_size
        self.ignore_threshold = ignore_threshold
        self.val_conf = val_conf
        self.label_smooth_eps = 0.1  # yolo_loss.py:48; the kernel hard-codes the resulting 0.95 / 0.05
        self.iou_thresh = iou_thresh
        self.iou_weighting = iou_weighting
        self.process_group = process_group
        self.last_sums: Optional[torch.Tensor] = None
        self.lazy_stats = False
        self._pending_status = None           # lazy_stats: (pinned host word, event) of the previous forward
        self._status_slots = None
        self._status_turn = 0

    # yolo_loss.py:214
    def scaled_anchors(self):
        return ops.scaled_anchors(self.anchors, self.img_size)

    def head_anchor_wh(self):
        return self.scaled_anchors()[list(self.mask)]

    def box_ciou(self, box1: torch.Tensor, box2: torch.Tensor):
        """yolo_loss.py:257-293: (iou - ciou_term, iou), each (n1, n2) (the reference calls it with one box each)."""
        return ops.pairwise(box1, box2, 4), ops.pairwise(box1, box2, 2)

    def box_giou(self, box1: torch.Tensor, box2: torch.Tensor):
        """yolo_loss.py:295-317 (dead code upstream, kept for completeness): (iou - giou_term, iou)."""
        return ops.pairwise(box1, box2, 3), ops.pairwise(box1, box2, 2)

    def get_pred_boxes(self, input: torch.Tensor):
        if self.lazy_eval:
            ops._require_cuda(input, "head")
            return ops.LazyCandidates(input, self.head_anchor_wh(), self.num_classes, self.val_conf)
        rows, count, ids = ops.decode_head_padded(input, self.head_anchor_wh(), self.num_classes, self.val_conf,
                                                  want_ids=True)
        return ops._as_list(rows, count, ids)

    @staticmethod
    def _raise_status(st: int) -> None:
        if st == 1:
            raise IndexError("a GT box maps outside the grid or has a class outside [1, num_classes] "
                             "(the reference raises IndexError at yolo_loss.py:149)")
        if st == 2:
            raise RuntimeError("more than 1024 GT boxes in one image are not supported by the target-assignment kernel")

    def check(self) -> None:
        """lazy_stats: raise what the last ``forward(input, targets)`` would have raised (one host read)"""
        if self._pending_status is not None:
            word, ev = self._pending_status
            self.__dict__["_pending_status"] = None
            ev.synchronize()                   # the copy was queued right behind that forward's kernel
            self._raise_status(int(word))

    def _prepared(self):
        """scaled anchors, mask and fp32-rounded thresholds as the C call takes them, rebuilt when an attribute changed"""
        key = (tuple(map(tuple, self.anchors)), tuple(self.img_size), tuple(self.mask), self.ignore_threshold, self.iou_thresh,
               self.iou_weighting)
        c = self.__dict__.get("_prep")
        if c is None or c[0] != key:
            import numpy as np
            c = (key, (ops._host_f32(self.scaled_anchors()).reshape(-1, 2),
                       np.ascontiguousarray(np.asarray(self.mask, dtype=np.int32)),
                       tuple(float(np.float32(v)) for v in key[3:])))
            self.__dict__["_prep"] = c
        return c[1]

    def _post_status(self, status: torch.Tensor) -> None:
        """lazy_stats: queue the status word's copy to pinned host memory behind the kernel and remember its event, so
        that the next forward (or check()) reads it without waiting for anything launched after it"""
        if self._status_slots is None or self._status_slots[0].device != status.device:
            host = torch.zeros((2,), dtype=torch.int32).pin_memory()
            self.__dict__["_status_slots"] = (status, host, (torch.cuda.Event(), torch.cuda.Event()), (host[0:1], host[1:2]))
        _, host, evs, views = self._status_slots
        k = self.__dict__["_status_turn"] = self._status_turn ^ 1
        views[k].copy_(status, non_blocking=True)
        if torch.cuda.current_device() == status.device.index:
            evs[k].record()
        else:
            evs[k].record(torch.cuda.current_stream(status.device))
        self.__dict__["_pending_status"] = (views[k], evs[k])

    def forward(self, input: torch.Tensor, targets=None):
        if targets is None:
            return self.get_pred_boxes(input)
        if isinstance(targets, ops.PackedTargets):
            gt, gt_off, G, max_gt = targets.gt, targets.gt_off, targets.G, targets.max_gt
        else:
            gt, gt_off, G, counts = ops.pack_targets(targets, input.device)
            max_gt = max(counts + [1])
        need_grad = torch.is_grad_enabled() and input.requires_grad
        x = input.detach()
        N, _, H, W = x.shape
        state = torch.empty((N, self.num_mask * H * W), dtype=torch.uint8, device=x.device) if need_grad else None
        if self.lazy_stats:
            # no host synchronisation and the host work of one call (ops.target_loss_lazy); shards of a data-parallel
            # batch all-reduce the 16 partial sums (and the status word) between the two launches
            sa, m, thr = self._prepared()
            between = None
            if self.process_group is not None:
                import torch.distributed as dist
                pg = self.process_group

                def between(sums_, status_):
                    dist.all_reduce(sums_, op=dist.ReduceOp.SUM, group=pg)
                    dist.all_reduce(status_, op=dist.ReduceOp.MAX, group=pg)
            sums, status, res = ops.target_loss_lazy(x, gt, gt_off, G, sa, m, self.num_classes, thr[0], thr[1], thr[2], max_gt, state,
                                                     between=between)
            d = self.__dict__                  # (nn.Module.__setattr__ costs 5 us per assignment)
            d["last_sums"] = sums
            self.check()                       # the previous call's status: its copy completed long ago, no stall
            self._post_status(status)
            loss, recall, avg_iou, obj, no_obj, cls_score, count = res.unbind(0)
            if need_grad:
                loss = _LossGrad.apply(input, loss, self, gt, gt_off, G, max_gt, state, sums)
            return loss, recall, avg_iou, obj, no_obj, cls_score, count
        sums, status = ops.target_loss_sums(x, gt, gt_off, G, self.scaled_anchors(), self.mask, self.num_classes,
                                            self.ignore_threshold, self.iou_thresh, max_gt=max_gt, cell_state=state)
        if self.process_group is not None:
            import torch.distributed as dist
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=self.process_group)
            dist.all_reduce(status, op=dist.ReduceOp.MAX, group=self.process_group)
        self.last_sums = sums
        host = torch.cat((sums, status.to(torch.float64))).cpu().numpy()  # one D2H sync (the reference has ~5 per GT)
        self._raise_status(int(host[-1]))
        r = ops.loss_finalize(host[:_lib.S_COUNT], self.iou_weighting)
        loss = torch.tensor(r[0], dtype=torch.float32, device=input.device)
        if need_grad:
            # loss.backward() (train.py:282) runs b200yolo_target_loss_backward on the saved cell states and the
            # batch-global sums
            loss = _LossGrad.apply(input, loss, self, gt, gt_off, G, max_gt, state, sums)
        no_obj = torch.tensor(r[4], dtype=torch.float32, device=input.device) if r[6] > 0 else 0
        # yolo_loss.py:236 -- loss, recall, avg_iou, obj, no_obj (tensor), cls_score, count/bs
        return loss, float(r[1]), float(r[2]), float(r[3]), no_obj, float(r[5]), float(r[6])


class _LossGrad(torch.autograd.Function):
    """Attaches the analytic gradient of the fused loss to the autograd graph: forward returns the loss value
    computed by the kernels, backward launches ``b200yolo_target_loss_backward``."""

    @staticmethod
    def forward(ctx, input, loss_value, module, gt, gt_off, G, max_gt, state, sums):
        ctx.save_for_backward(input.detach(), gt, gt_off, state, sums)
        ctx.cfg = (G, max_gt, module.scaled_anchors(), list(module.mask), module.num_classes, module.iou_thresh,
                   module.iou_weighting)
        return loss_value.clone()

    @staticmethod
    def backward(ctx, grad_output):
        x, gt, gt_off, state, sums = ctx.saved_tensors
        G, max_gt, sa, mask, C, iou_thr, iou_w = ctx.cfg
        grad = ops.target_loss_backward(x, gt, gt_off, G, sa, mask, C, iou_thr, state, sums, iou_w, grad_out=grad_output,
                                        max_gt=max_gt)
        return (grad,) + (None,) * 8
