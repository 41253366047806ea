"""Deterministic synthetic inputs shared by the golden generator, the tests, smoke() and bench.py.

Test infrastructure only.  Nothing here is copied from the reference; it only fixes HOW seeded
weights, Darknet files and box sets are produced so that the reference (in this container), the
oracle and the CUDA path all see identical bytes without committing 248 MB of weights.
"""
from __future__ import annotations

import hashlib
from typing import Dict

import numpy as np
import torch


def _gen(seed_text: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(hashlib.sha256(seed_text.encode()).digest()[:7], "little"))
    return g


def synth_state_dict(template: Dict[str, torch.Tensor], seed: int = 0, head_gain: float = 1.0) -> Dict[str, torch.Tensor]:
    """Random but well-conditioned values for every tensor of a reference-keyed state_dict
    (keys `layers.N....{conv,batch_norm}.*`).  Per-key generators => independent of dict order."""
    out = {}
    for k, t in template.items():
        g = _gen(f"{seed}:{k}")
        if k.endswith("conv.weight"):
            fan_in = t.shape[1] * t.shape[2] * t.shape[3]
            # gains keep activations O(1) through 75 convs + 23 residual adds (measured: head std 0.6-0.9)
            gain = 0.3 if (".layers." in k and k.endswith(".1.conv.weight")) else 1.4
            if ".pred_block.1." in k:
                gain = head_gain
            v = torch.randn(t.shape, generator=g) * (gain * (1.0 / fan_in) ** 0.5)
        elif k.endswith("conv.bias"):
            v = torch.randn(t.shape, generator=g) * 0.5
        elif k.endswith("batch_norm.weight"):
            v = 0.75 + 0.5 * torch.rand(t.shape, generator=g)
        elif k.endswith("batch_norm.bias") or k.endswith("running_mean"):
            v = 0.1 * torch.randn(t.shape, generator=g)
        elif k.endswith("running_var"):
            v = 0.75 + 0.5 * torch.rand(t.shape, generator=g)
        elif k.endswith("num_batches_tracked"):
            v = torch.zeros(t.shape, dtype=t.dtype)
        else:
            raise KeyError(f"unexpected state_dict key {k}")
        out[k] = v.to(t.dtype)
    return out


def synth_darknet_file(path: str, n_floats: int) -> None:
    """A Darknet-format file (5 x int32 header + flat fp32) with a cheap deterministic pattern.
    running_var slots may come out negative -- irrelevant for loader parity (values are only copied)."""
    i = np.arange(n_floats, dtype=np.uint64)
    vals = (((i * np.uint64(2654435761)) % np.uint64(1000003)).astype(np.float64) / 1000003.0 - 0.5).astype(np.float32)
    with open(path, "wb") as f:
        np.array([0, 2, 0, 32013312, 0], dtype=np.int32).tofile(f)
        vals.tofile(f)


def synth_boxes(n: int, num_classes: int, seed: int, tie_frac: float = 0.01, wh=(0.02, 0.3)) -> torch.Tensor:
    """SURVEY.md 8d config 5 box sets: cx,cy~U(0,1), w,h~U(wh), score~U(0,1), cls~U{0..nc-1},
    with a `tie_frac` subset of exactly repeated scores (stable-sort coverage)."""
    g = _gen(f"boxes:{seed}:{n}:{num_classes}")
    b = torch.empty(n, 6)
    b[:, 0:2] = torch.rand(n, 2, generator=g)
    b[:, 2:4] = wh[0] + (wh[1] - wh[0]) * torch.rand(n, 2, generator=g)
    b[:, 4] = torch.rand(n, generator=g)
    b[:, 5] = torch.randint(0, num_classes, (n,), generator=g).float()
    nt = int(n * tie_frac)
    if nt > 1:
        idx = torch.randperm(n, generator=g)[:nt]
        b[idx, 4] = b[idx[0], 4]
    return b
