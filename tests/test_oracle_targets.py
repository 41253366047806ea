"""Pins oracle/target_oracle.py against targets produced by the UNMODIFIED reference `YOLODataset.__getitem__`
(tests/golden/targets.npz, oracle/gen_golden_targets.py).  CPU only; bit-exact."""
import os

import numpy as np
import torch

from oracle import target_oracle as to
from oracle import yolo_oracle as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "targets.npz")


def golden_cases():
    z = np.load(GOLD)
    for c in range(int(z["n"])):
        size, which = (int(v) for v in z[f"c{c}/meta"])
        anchors = orc.TURBINE_ANCHORS if which else orc.ANCHORS
        grid = [size // 32, size // 16, size // 8]
        dense = []
        for s in range(3):
            t = torch.zeros(3, grid[s], grid[s], 6)
            idx, val = z[f"c{c}/t{s}_idx"], z[f"c{c}/t{s}_val"]
            if len(idx):
                t[idx[:, 0], idx[:, 1], idx[:, 2]] = torch.from_numpy(val)
            dense.append(t)
        yield c, z[f"c{c}/boxes"].tolist(), anchors, grid, dense


def test_target_oracle_matches_reference():
    n_ignore = 0
    for c, boxes, anchors, grid, dense in golden_cases():
        got = to.encode_image(boxes, anchors, grid)
        for s in range(3):
            assert torch.equal(got[s], dense[s]), (c, s)
            n_ignore += int((dense[s][..., 4] == -1).sum())
    assert n_ignore > 0   # the fixtures exercise the ignore rule


def test_iou_aligned_kat():
    kat = to.iou_aligned(torch.tensor([0.2, 0.3]), torch.tensor([[0.28, 0.22], [0.38, 0.48]]))
    assert torch.allclose(kat, torch.tensor([0.5670, 0.3289]), atol=1e-4)   # SURVEY 8c KAT
