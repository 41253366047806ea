"""CPU oracle of the reference's training-target encoder -- TEST INFRASTRUCTURE ONLY.

Restates `YOLODataset.__getitem__`'s target construction (code/dataset.py:119-167; `iou_aligned` code/utils.py:22-36)
and the target part of `collate_fn` (code/utils.py:694-700: stack per scale -> (B, 3, S, S, 6)):

  per image: targets[s] = zeros(3, S_s, S_s, 6); for every box (x, y, w, h, class) IN ORDER:
    iou = iou_aligned([w, h], 9 anchors) in fp32; anchors visited by descending IoU (dataset.py:131);
    scale = a // 3, anchor = a % 3, i = int(S*y), j = int(S*x) in Python doubles (dataset.py:142);
    "taken" = targets[scale][anchor, i, j, 0] != 0  -- element 0 (x_cell), NOT the objectness flag (dataset.py:143);
    not taken and scale has no anchor yet -> obj = 1, [S*x - j, S*y - i, w*S, h*S] (doubles -> fp32), class (dataset.py:146-156);
    elif not taken and iou > 0.5        -> obj = -1 (ignore)                                        (dataset.py:160-161)
Pinned by tests/test_oracle_targets.py against tests/golden/targets.npz, produced by driving the UNMODIFIED
`YOLODataset.__getitem__` (oracle/gen_golden_targets.py).  Ties between anchor IoUs are broken towards the lower anchor
index here (torch.argsort's order for ties is unspecified); the fixtures contain none.
"""
from __future__ import annotations

from typing import List, Sequence

import torch

IGNORE_IOU_THRESHOLD = 0.5  # dataset.py:53


def iou_aligned(box1: torch.Tensor, box2: torch.Tensor) -> torch.Tensor:
    inter = torch.min(box1[..., 0], box2[..., 0]) * torch.min(box1[..., 1], box2[..., 1])   # utils.py:34
    union = box1[..., 0] * box1[..., 1] + box2[..., 0] * box2[..., 1] - inter                  # utils.py:35
    return inter / union


def encode_image(boxes: Sequence[Sequence[float]], anchors, grid_sizes: Sequence[int]) -> List[torch.Tensor]:
    anc = torch.tensor(anchors[0] + anchors[1] + anchors[2])                                   # dataset.py:39
    targets = [torch.zeros((3, s, s, 6)) for s in grid_sizes]                                  # dataset.py:123
    for box in boxes:
        iou = iou_aligned(torch.tensor(box[2:4]), anc)                                         # dataset.py:130
        order = iou.argsort(descending=True, dim=0, stable=True)                               # dataset.py:131
        x, y, w, h, cls = box
        has_anchor = [False] * 3
        for a in order.tolist():
            s_idx, a_idx = a // 3, a % 3
            S = grid_sizes[s_idx]
            i, j = int(S * y), int(S * x)
            taken = targets[s_idx][a_idx, i, j, 0]
            if not taken and not has_anchor[s_idx]:
                targets[s_idx][a_idx, i, j, 4] = 1
                targets[s_idx][a_idx, i, j, 5] = int(cls)
                targets[s_idx][a_idx, i, j, :4] = torch.tensor([S * x - j, S * y - i, w * S, h * S])
                has_anchor[s_idx] = True
            elif not taken and iou[a] > IGNORE_IOU_THRESHOLD:
                targets[s_idx][a_idx, i, j, 4] = -1
    return targets


def encode_batch(boxes_per_image, anchors, grid_sizes) -> List[torch.Tensor]:
    per = [encode_image(b, anchors, grid_sizes) for b in boxes_per_image]
    return [torch.stack(t) for t in zip(*per)]                                                  # utils.py:699
