"""Drop-in mirror of the reference's `code/loss.py` (YOLOLoss) -- FORWARD ONLY on the sm_100a kernel.

`YOLOLoss().forward(predictions, targets, anchors)` returns the same list `[5*box, 1*object, 0.5*no_obj,
1*class]` (loss.py:78-81) computed by one fused pass (csrc/loss.cu) instead of mask-index gathers plus
BCEWithLogits / MSE / CrossEntropy launches; it also reproduces the reference's in-place updates of
`predictions[..., 1:3]` and `targets[..., 2:4]` (loss.py:71-72).  This serves the no-grad validation loop
(train.py:94-156).  The backward pass (SURVEY config #4) is not built: a tensor that requires grad raises.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from ._lib import YoloB200Error, lib, ptr, stream_ptr


class YOLOLoss(nn.Module):
    def __init__(self):
        super().__init__()
        self.lambda_box = 5
        self.lambda_obj = 1
        self.lambda_noobj = 0.5
        self.lambda_class = 1

    def forward(self, predictions, targets, anchors):
        if not predictions.is_cuda:
            raise YoloB200Error("YOLOLoss needs CUDA tensors (no CPU fallback)")
        if predictions.requires_grad and torch.is_grad_enabled():
            raise YoloB200Error("YOLOLoss backward is not built on this path yet: call it under torch.no_grad()")
        if predictions.dtype != torch.float32 or targets.dtype != torch.float32 or targets.device != predictions.device:
            raise YoloB200Error("predictions and targets must be fp32 tensors on the same CUDA device")
        B, A, S, S2, Cc = predictions.shape
        if A != 3 or S != S2 or tuple(targets.shape) != (B, 3, S, S, 6):
            raise YoloB200Error(f"shapes {tuple(predictions.shape)} / {tuple(targets.shape)} are not one YOLO scale")
        dev = predictions.device
        anc = torch.as_tensor(anchors, dtype=torch.float32).reshape(-1).cpu()
        sums = torch.zeros(6, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            lib.yolo_loss_fwd(ptr(predictions), (C.c_int64 * 5)(*predictions.stride()), ptr(targets),
                              (C.c_int64 * 5)(*targets.stride()), B, S, Cc - 5, (C.c_float * 6)(*anc.tolist()), 1,
                              ptr(sums), stream_ptr(dev))
        s_noobj, n_noobj, s_obj, s_box, s_cls, n_obj = sums.unbind()
        zero = torch.zeros((), dtype=torch.float64, device=dev)
        has_obj = n_obj > 0
        no_obj_loss = s_noobj / n_noobj                       # mean over no-object cells (NaN when there are none)
        object_loss = torch.where(has_obj, s_obj / n_obj, zero)
        box_loss = torch.where(has_obj, s_box / (4 * n_obj), zero)
        class_loss = torch.where(has_obj, s_cls / n_obj, zero)
        f32 = lambda v: v.to(torch.float32)  # noqa: E731
        return [self.lambda_box * f32(box_loss), self.lambda_obj * f32(object_loss),
                self.lambda_noobj * f32(no_obj_loss), self.lambda_class * f32(class_loss)]
