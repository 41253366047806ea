"""Device mirror of the reference's inference pre-processing (code/config.py:101-113 `set_only_image_transforms`,
called by code/demo.py:37-39) and of `plot_original`'s inverse box mapping (code/utils.py:475-501).

    x = letterbox_batch([img0, img1, ...], 416)        # uint8 HWC images of any size -> (B, 3, 416, 416) fp32 CUDA
    boxes = unletterbox_boxes(kept_rows, h, w, 416)     # rows normalised to the square -> to the original image
    x = LetterboxPlan(64, 480, 640, 416).run(frames)    # serving loop: fixed frame size, one copy + one launch per batch

`letterbox_batch(...)[i]` equals `set_only_image_transforms(S)(image=img_i)["image"]`: albumentations' LongestMaxSize
(round-half-even output size, cv2.INTER_LINEAR uint8 fixed point) + centred zero PadIfNeeded + /255 + ToTensorV2, in
ONE launch for the whole batch (csrc/layout.cu).  No CPU fallback: images are uploaded if they are not on the device.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np
import torch

from ._lib import YoloB200Error, lib, ptr, stream_ptr


def letterbox_geometry(h: int, w: int, size: int) -> Tuple[int, int, int, int]:
    """(new_h, new_w, top, left): LongestMaxSize's output size (Python round = half to even, albumentations'
    py3round) and PadIfNeeded's centred offsets (top/left get the floor)."""
    scale = size / float(max(h, w))
    nh, nw = (h, w) if scale == 1.0 else (int(round(h * scale)), int(round(w * scale)))
    top = int((size - nh) / 2.0) if nh < size else 0
    left = int((size - nw) / 2.0) if nw < size else 0
    return nh, nw, top, left


class _Desc(C.Structure):
    _fields_ = [("data", C.c_void_p), ("h", C.c_int32), ("w", C.c_int32), ("nh", C.c_int32), ("nw", C.c_int32),
                ("top", C.c_int32), ("left", C.c_int32)]


def letterbox_batch(images: Sequence, size: int, device=None, out: torch.Tensor = None) -> torch.Tensor:
    if device is None:
        device = next((im.device for im in images if isinstance(im, torch.Tensor) and im.is_cuda), torch.device("cuda"))
    device = torch.device(device)
    if device.type != "cuda":
        raise YoloB200Error("letterbox_batch runs on a CUDA device only (no CPU fallback)")
    if C.sizeof(_Desc) != lib.yolo_letterbox_desc_bytes():
        raise YoloB200Error("letterbox descriptor layout mismatch between preprocess.py and libyolo_b200.so")
    keep: List[torch.Tensor] = []
    descs = (_Desc * max(len(images), 1))()
    channels = None
    for i, im in enumerate(images):
        t = torch.from_numpy(np.ascontiguousarray(im)) if isinstance(im, np.ndarray) else im
        if t.dtype != torch.uint8 or t.dim() != 3:
            raise YoloB200Error(f"image {i}: expected a uint8 HWC image, got {t.dtype} {tuple(t.shape)}")
        t = t.to(device, non_blocking=True).contiguous()
        keep.append(t)
        h, w, c = t.shape
        channels = c if channels is None else channels
        if c != channels:
            raise YoloB200Error("all images of a batch must have the same channel count")
        nh, nw, top, left = letterbox_geometry(h, w, size)
        descs[i] = _Desc(t.data_ptr(), h, w, nh, nw, top, left)
    B = len(images)
    if out is None:
        out = torch.empty(B, channels or 3, size, size, dtype=torch.float32, device=device)
    if B == 0:
        return out
    raw = torch.frombuffer(bytearray(bytes(descs)), dtype=torch.uint8)[: B * C.sizeof(_Desc)]
    table = raw.to(device)
    with torch.cuda.device(device):
        lib.yolo_letterbox_u8(ptr(table), B, size, channels, ptr(out), stream_ptr(device))
    out._yb_keepalive = (keep, table)   # sources and table must outlive the asynchronous launch
    return out


class LetterboxPlan:
    """Static serving-loop variant of `letterbox_batch` for frames of ONE size: a device uint8 staging buffer
    [B, h, w, c], the descriptor table and the fp32 [B, c, size, size] output are allocated once, so a step is one
    H2D copy of the uint8 frames (a quarter of the fp32 tensor's bytes even before any down-scaling) and one launch:

        lp = LetterboxPlan(64, 480, 640, 416)
        x = lp.run(frames_u8_pinned)          # [64, 480, 640, 3] uint8, host (pinned) or device -> [64, 3, 416, 416] fp32

    Same arithmetic as `letterbox_batch` (it is the same kernel, yolo_letterbox_u8)."""

    def __init__(self, batch: int, h: int, w: int, size: int, channels: int = 3, device=None):
        device = torch.device("cuda" if device is None else device)
        if device.type != "cuda":
            raise YoloB200Error("LetterboxPlan runs on a CUDA device only (no CPU fallback)")
        if C.sizeof(_Desc) != lib.yolo_letterbox_desc_bytes():
            raise YoloB200Error("letterbox descriptor layout mismatch between preprocess.py and libyolo_b200.so")
        self.batch, self.h, self.w, self.size, self.channels, self.device = batch, h, w, size, channels, device
        self.staging = torch.empty(batch, h, w, channels, dtype=torch.uint8, device=device)
        self.out = torch.empty(batch, channels, size, size, dtype=torch.float32, device=device)
        nh, nw, top, left = letterbox_geometry(h, w, size)
        descs = (_Desc * batch)()
        per = h * w * channels
        for i in range(batch):
            descs[i] = _Desc(self.staging.data_ptr() + i * per, h, w, nh, nw, top, left)
        self.table = torch.frombuffer(bytearray(bytes(descs)), dtype=torch.uint8).to(device)

    def run(self, frames: torch.Tensor) -> torch.Tensor:
        """Asynchronous on the current stream; `frames` is uint8 [B, h, w, c] (pinned host memory for a non-blocking copy)."""
        if frames.dtype != torch.uint8 or tuple(frames.shape) != tuple(self.staging.shape):
            raise YoloB200Error(f"expected uint8 frames of shape {tuple(self.staging.shape)}, got {frames.dtype} {tuple(frames.shape)}")
        self.staging.copy_(frames, non_blocking=True)
        with torch.cuda.device(self.device):
            lib.yolo_letterbox_u8(ptr(self.table), self.batch, self.size, self.channels, ptr(self.out), stream_ptr(self.device))
        return self.out


def unletterbox_boxes(boxes, orig_h: int, orig_w: int, size: int):
    """plot_original's mapping (utils.py:475-501) of rows [cx, cy, w, h, score, cls] from the letterboxed square back
    to the original image; `boxes` may be a list of rows (returns a list, like the reference) or a tensor."""
    scale = min(size / orig_w, size / orig_h)
    new_w, new_h = int(orig_w * scale), int(orig_h * scale)
    pad_w, pad_h = (size - new_w) // 2, (size - new_h) // 2
    if isinstance(boxes, torch.Tensor):
        out = boxes.clone()
        out[:, 0] = (boxes[:, 0] * size - pad_w) / new_w
        out[:, 1] = (boxes[:, 1] * size - pad_h) / new_h
        out[:, 2] = boxes[:, 2] * size / new_w
        out[:, 3] = boxes[:, 3] * size / new_h
        return out
    return [[(b[0] * size - pad_w) / new_w, (b[1] * size - pad_h) / new_h, (b[2] * size) / new_w, (b[3] * size) / new_h,
             b[4], b[5]] for b in boxes]
