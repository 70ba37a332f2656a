"""Drop-in for the reference's ``models/seg_loss.py::SegLoss``: the drivable-area head
of the BDD100k multi-task model (models/mbv2_yolo.py:163,170)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from . import _lib, ops

_WS = {}


def _seg_sums(x: torch.Tensor, truth: torch.Tensor) -> torch.Tensor:
    N, C, H, W = x.shape
    lib = _lib.load()
    with ops._on_device(x.device):
        sums = torch.empty((8,), dtype=torch.float64, device=x.device)
        key = (x.device.index, torch.cuda.current_stream(x.device).cuda_stream)
        ws = _WS.get(key)
        if ws is None:
            ws = _WS[key] = torch.empty((int(lib.b200yolo_seg_loss_workspace_bytes()),), dtype=torch.uint8, device=x.device)
        _lib.check(lib.b200yolo_seg_loss(x.data_ptr(), truth.data_ptr(), N, C, H, W, sums.data_ptr(), ws.data_ptr(), ws.numel(),
                                         ops._stream(x)))
    return sums


class _SegGrad(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input, loss_value, truth):
        ctx.save_for_backward(input.detach(), truth)
        return loss_value.clone()

    @staticmethod
    def backward(ctx, grad_output):
        x, truth = ctx.saved_tensors
        N, C, H, W = x.shape
        with ops._on_device(x.device):
            grad = torch.empty_like(x)
            go = grad_output.detach().to(device=x.device, dtype=torch.float32).reshape(1).contiguous()
            _lib.check(_lib.load().b200yolo_seg_loss_backward(x.data_ptr(), truth.data_ptr(), N, C, H, W, go.data_ptr(),
                                                              grad.data_ptr(), ops._stream(x)))
        return grad, None, None


class SegLoss(nn.Module):
    """Same constructor and return conventions as models/seg_loss.py:33-81.

    ``forward(input, targets)`` -> ``(0.05 * mse(sigmoid(input), targets), mean sigmoid over targets >= 0.5,
    mean sigmoid over targets < 0.5)`` with ``targets`` of shape (N, H, W, C); ``forward(input)`` -> numpy array
    ``sigmoid(input)[0]`` of shape (C, H, W)."""

    def __init__(self, num_classes):
        super().__init__()
        self.num_classes = num_classes

    def forward(self, input: torch.Tensor, targets=None):
        ops._require_cuda(input, "input")
        x = input.detach().contiguous()
        if targets is None:
            C, H, W = x.shape[1:]
            out = torch.empty((C, H, W), dtype=torch.float32, device=x.device)
            with ops._on_device(x.device):
                _lib.check(_lib.load().b200yolo_seg_sigmoid(x.data_ptr(), C * H * W, out.data_ptr(), ops._stream(x)))
            return out.cpu().numpy()                                   # seg_loss.py:78-80
        truth = targets.to(device=x.device, dtype=torch.float32).contiguous()
        if tuple(truth.shape) != (x.shape[0], x.shape[2], x.shape[3], x.shape[1]):
            raise RuntimeError(f"targets must be (N, H, W, C) = {(x.shape[0], x.shape[2], x.shape[3], x.shape[1])}, got {tuple(truth.shape)}")
        s = _seg_sums(x, truth).cpu().numpy()                           # one D2H sync (the reference has two .item())
        with np.errstate(divide="ignore", invalid="ignore"):
            loss_v = np.float64(0.05) * s[0] / s[1]
            obj, no_obj = s[2] / s[3], s[4] / s[5]                      # mean of an empty selection is NaN, like torch
        loss = torch.tensor(loss_v, dtype=torch.float32, device=x.device)
        if torch.is_grad_enabled() and input.requires_grad:
            loss = _SegGrad.apply(input, loss, truth)
        return loss, float(obj), float(no_obj)
