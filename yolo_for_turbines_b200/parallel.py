"""Data-parallel plumbing: one process per GPU, images sharded across ranks.

The reference has no distributed runtime at all; forward, decode and NMS are
per-image, so the inference path shards with NO collective.  The one exchange step is the mAP
evaluation (utils.py:193 consumes detections of the whole data set): ranks all-gather their
[img, cx, cy, w, h, score, cls] rows -- a few KB to MB over NCCL/NVLink -- in rank order, which is the
reference's global order (ascending image index, then NMS output order) when ranks own contiguous
image ranges.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of `n_items` owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_rows(rows: torch.Tensor, group=None) -> torch.Tensor:
    """All-gathers variable-length [k_r, C] row blocks and concatenates them in rank order."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return rows
    world = dist.get_world_size(group)
    rows = rows.contiguous()
    n = torch.tensor([rows.shape[0]], dtype=torch.int64, device=rows.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n, group=group)
    counts = [int(c.item()) for c in counts]
    cap = max(max(counts), 1)
    padded = torch.zeros(cap, rows.shape[1], dtype=rows.dtype, device=rows.device)
    padded[: rows.shape[0]] = rows
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded, group=group)
    return torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)


def distributed_mAP(local_dets: torch.Tensor, local_gts: torch.Tensor, iou_threshold=0.5, box_format="center",
                    num_classes=20, group=None):
    """calc_mAP over the union of every rank's detections / ground truths (image ids must be global)."""
    from .utils import calc_mAP

    dets = gather_rows(local_dets.reshape(-1, 7), group)
    gts = gather_rows(local_gts.reshape(-1, 7), group)
    return calc_mAP(dets, gts, iou_threshold, box_format, num_classes)
