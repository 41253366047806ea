"""Device mirror of the reference's training-target encoder: the anchor-assignment part of `YOLODataset.__getitem__`
(code/dataset.py:119-167) plus `collate_fn`'s per-scale stacking (code/utils.py:694-700), for a whole batch in one
launch (csrc/targets.cu).  Image loading, CSV splits and Albumentations stay out of scope.

    targets = encode_targets(boxes_per_image, config.TURBINE_ANCHORS, image_size=416)   # 3 x (B, 3, S, S, 6) fp32 CUDA
    losses = trainer.step(x, targets)
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence

import torch

from ._lib import YoloB200Error, lib, ptr, stream_ptr

IGNORE_IOU_THRESHOLD = 0.5  # dataset.py:53


def encode_targets(boxes_per_image: Sequence, anchors, image_size: int = None, grid_sizes: Sequence[int] = None,
                   device="cuda", ignore_iou_threshold: float = IGNORE_IOU_THRESHOLD) -> List[torch.Tensor]:
    """boxes_per_image[b]: rows (x, y, w, h, class) in YOLO format (fractions of the image), in the order the reference
    would visit them (later boxes see cells claimed by earlier ones).  Returns the three target tensors the training
    step consumes.  No CPU fallback."""
    device = torch.device(device)
    if device.type != "cuda":
        raise YoloB200Error("encode_targets runs on a CUDA device only (no CPU fallback)")
    if grid_sizes is None:
        if image_size is None or image_size % 32:
            raise YoloB200Error("give image_size (a multiple of 32) or grid_sizes")
        grid_sizes = [image_size // 32, image_size // 16, image_size // 8]      # dataset.py:115
    flat = [float(v) for scale in anchors for a in scale for v in a]
    if len(flat) != 18:
        raise YoloB200Error("expected 3 scales x 3 anchors x (w, h)")
    rows, offsets = [], [0]
    for b in boxes_per_image:
        t = torch.as_tensor(b, dtype=torch.float64).reshape(-1, 5)
        rows.append(t)
        offsets.append(offsets[-1] + t.shape[0])
    B = len(rows)
    boxes = (torch.cat(rows) if rows else torch.zeros(0, 5, dtype=torch.float64)).to(device).contiguous()
    if boxes.numel() == 0:
        boxes = torch.zeros(1, 5, dtype=torch.float64, device=device)
    off = torch.tensor(offsets, dtype=torch.int32, device=device)
    outs = [torch.empty(B, 3, s, s, 6, dtype=torch.float32, device=device) for s in grid_sizes]
    with torch.cuda.device(device):
        lib.yolo_encode_targets(ptr(boxes), ptr(off), B, (C.c_float * 18)(*flat), int(grid_sizes[0]), int(grid_sizes[1]),
                                int(grid_sizes[2]), float(ignore_iou_threshold), ptr(outs[0]), ptr(outs[1]), ptr(outs[2]),
                                stream_ptr(device))
    outs[0]._yb_keepalive = (boxes, off)
    return outs
