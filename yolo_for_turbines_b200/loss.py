"""Drop-in mirror of the reference's `code/loss.py` (YOLOLoss) on the sm_100a kernels.

`YOLOLoss().forward(predictions, targets, anchors)` returns the same list `[5*box, 1*object, 0.5*no_obj,
1*class]` (loss.py:78-81) computed by one fused pass (csrc/loss.cu) instead of mask-index gathers plus
BCEWithLogits / MSE / CrossEntropy launches.

  * under torch.no_grad() (the validation loop, train.py:94-156) it also reproduces the reference's in-place
    updates of `predictions[..., 1:3]` and `targets[..., 2:4]` (loss.py:71-72);
  * when `predictions` requires grad (the training loop, train.py:56-67) the four terms are outputs of one
    autograd node whose backward is the fused kernel `yolo_loss_bwd`; the in-place updates are then NOT applied
    (they would only overwrite tensors the training loop never reads again).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from ._lib import YoloB200Error, lib, ptr, stream_ptr


def _terms_from_sums(sums, lambdas):
    s_noobj, n_noobj, s_obj, s_box, s_cls, n_obj = sums.unbind()
    zero = torch.zeros((), dtype=torch.float64, device=sums.device)
    has_obj = n_obj > 0
    no_obj_loss = s_noobj / n_noobj                       # mean over no-object cells (NaN when there are none)
    object_loss = torch.where(has_obj, s_obj / n_obj, zero)
    box_loss = torch.where(has_obj, s_box / (4 * n_obj), zero)
    class_loss = torch.where(has_obj, s_cls / n_obj, zero)
    lb, lo, ln, lc = lambdas
    return [(lb * box_loss).float(), (lo * object_loss).float(), (ln * no_obj_loss).float(), (lc * class_loss).float()]


def _launch_fwd(predictions, targets, anc, mutate):
    B, _, S, _, Cc = predictions.shape
    dev = predictions.device
    sums = torch.zeros(6, dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        lib.yolo_loss_fwd(ptr(predictions), (C.c_int64 * 5)(*predictions.stride()), ptr(targets),
                          (C.c_int64 * 5)(*targets.stride()), B, S, Cc - 5, anc, mutate, ptr(sums), stream_ptr(dev))
    return sums


class _YoloLossFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, predictions, targets, anc, lambdas):
        pred = predictions.detach()
        sums = _launch_fwd(pred, targets, anc, 0)
        ctx.save_for_backward(pred, targets, sums)
        ctx.anc, ctx.lambdas = anc, lambdas
        return tuple(_terms_from_sums(sums, lambdas))

    @staticmethod
    def backward(ctx, g_box, g_obj, g_noobj, g_cls):
        pred, targets, sums = ctx.saved_tensors
        B, _, S, _, Cc = pred.shape
        dev = pred.device
        # the kernel carries the reference's lambdas (5, 1, 0.5, 1); fold upstream gradients and any change of them in
        ups = [0.0 if g is None else float(g) for g in (g_box, g_obj, g_noobj, g_cls)]
        tw = (C.c_float * 4)(*[u * lam / ref for u, lam, ref in zip(ups, ctx.lambdas, (5.0, 1.0, 0.5, 1.0))])
        d = torch.empty(pred.shape, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            lib.yolo_loss_bwd(ptr(pred), (C.c_int64 * 5)(*pred.stride()), ptr(targets), (C.c_int64 * 5)(*targets.stride()),
                              B, S, Cc - 5, ctx.anc, ptr(sums), 1.0, tw, ptr(d), (C.c_int64 * 5)(*d.stride()), 0, stream_ptr(dev))
        return d, None, None, None


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
        if predictions.dtype != torch.float32 or targets.dtype != torch.float32 or targets.device != predictions.device:
            raise YoloB200Error("predictions and targets must be fp32 tensors on the same CUDA device")
        B, A, S, S2, Cc = predictions.shape
        if A != 3 or S != S2 or tuple(targets.shape) != (B, 3, S, S, 6):
            raise YoloB200Error(f"shapes {tuple(predictions.shape)} / {tuple(targets.shape)} are not one YOLO scale")
        anc = (C.c_float * 6)(*torch.as_tensor(anchors, dtype=torch.float32).reshape(-1).cpu().tolist())
        lambdas = (float(self.lambda_box), float(self.lambda_obj), float(self.lambda_noobj), float(self.lambda_class))
        if predictions.requires_grad and torch.is_grad_enabled():
            return list(_YoloLossFn.apply(predictions, targets, anc, lambdas))
        sums = _launch_fwd(predictions, targets, anc, 1)
        return _terms_from_sums(sums, lambdas)
