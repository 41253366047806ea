"""Golden vectors for get_eval_boxes (code/utils.py:276-332), produced by the UNMODIFIED reference.

    python -m oracle.gen_golden_eval          # needs the reference (oracle/ref_loader.py search order)

Test infrastructure only.  A fake model returns seeded head tensors (the reference only needs eval()/train()/__call__),
a fake loader yields two batches of (x, targets); the fixture stores the inputs, the two returned lists and the
sequence of eval()/train() calls the reference made on the model (utils.py:295 and :331 -- train() is unconditional).
Heads are generated so that no objectness sits within 1e-3 of the confidence threshold and no same-class pair IoU
within 1e-3 of the NMS threshold: the CUDA path decodes within 1e-5 of the reference, which then cannot flip a
decision.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from oracle import yolo_oracle as orc  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
NC, IOU_THR, OBJ_THR = 3, 0.45, 0.6
GRIDS = (2, 4, 8)
ANCHORS = [[(0.28, 0.22), (0.38, 0.48), (0.9, 0.78)], [(0.07, 0.15), (0.15, 0.11), (0.14, 0.29)],
           [(0.02, 0.03), (0.04, 0.07), (0.08, 0.06)]]   # code/config.py:47-51 (the COCO set)


class FakeModel:
    def __init__(self, batches):
        self.batches, self.i, self.calls = batches, 0, []

    def eval(self):
        self.calls.append("eval")

    def train(self):
        self.calls.append("train")

    def __call__(self, x):
        out = [o.clone() for o in self.batches[self.i]]
        self.i += 1
        return out


def make_batch(g, bsz):
    heads, tgts = [], []
    lim = float(torch.logit(torch.tensor(OBJ_THR)))
    for s in GRIDS:
        o = torch.randn(bsz, 3, s, s, 5 + NC, generator=g)
        o[..., 2:4] *= 0.5
        o[..., 4] = 1.5 * torch.randn(bsz, 3, s, s, generator=g)
        near = (o[..., 4] - lim).abs() < 0.02
        o[..., 4][near] += 0.05
        heads.append(o)
        t = torch.zeros(bsz, 3, s, s, 6)
        u = torch.rand(bsz, 3, s, s, generator=g)
        t[..., 4] = torch.where(u < 0.15, torch.tensor(1.0), torch.where(u < 0.2, torch.tensor(-1.0), torch.tensor(0.0)))
        t[..., 0:2] = torch.rand(bsz, 3, s, s, 2, generator=g)
        t[..., 2:4] = 0.3 + 2.0 * torch.rand(bsz, 3, s, s, 2, generator=g)
        t[..., 5] = torch.randint(0, NC, (bsz, 3, s, s), generator=g).float()
        tgts.append(t)
    return heads, tgts


def margins_ok(heads):
    """No same-class pair of threshold-passing candidates with an IoU within 1e-3 of the NMS threshold."""
    bsz = heads[0].shape[0]
    per_img = [[] for _ in range(bsz)]
    for i, o in enumerate(heads):
        s = o.shape[2]
        rows = orc.cells_to_boxes(o.clone(), torch.tensor(ANCHORS[i]) * s, s, True)
        for b in range(bsz):
            per_img[b] += rows[b]
    for rows in per_img:
        t = torch.tensor([r for r in rows if r[4] > OBJ_THR], dtype=torch.float32)
        if t.numel() == 0:
            continue
        if bool(((t[:, 4] - OBJ_THR).abs() < 1e-3).any()):
            return False
        for k in range(t.shape[0]):
            iou = orc.calc_iou(t[k:k + 1, :4], t[:, :4], "center")
            same = t[:, 5] == t[k, 5]
            if bool((((iou - IOU_THR).abs() < 1e-3) & same).any()):
                return False
    return True


def main():
    _, rutils, _, _ = ref_loader.load()
    seed = 4242
    while True:
        g = torch.Generator().manual_seed(seed)
        batches = [make_batch(g, 2), make_batch(g, 3)]
        if all(margins_ok(h) for h, _ in batches):
            break
        seed += 1
    model = FakeModel([h for h, _ in batches])
    loader = [(torch.zeros(len(h[0]), 3, 64, 64), [t.clone() for t in tg]) for h, tg in batches]
    preds, trues = rutils.get_eval_boxes(loader, model, IOU_THR, ANCHORS, OBJ_THR, box_format="center", device="cpu")
    out = {"seed": np.int64(seed), "iou_thr": np.float64(IOU_THR), "obj_thr": np.float64(OBJ_THR),
           "anchors": np.asarray(ANCHORS, dtype=np.float64), "calls": np.asarray(model.calls),
           "preds": np.asarray(preds, dtype=np.float64).reshape(-1, 7), "trues": np.asarray(trues, dtype=np.float64).reshape(-1, 7)}
    for bi, (h, tg) in enumerate(batches):
        for si in range(3):
            out[f"b{bi}_head{si}"] = h[si].numpy()
            out[f"b{bi}_tgt{si}"] = tg[si].numpy()
    np.savez_compressed(os.path.join(GOLD, "eval_boxes.npz"), **out)
    print(f"seed {seed}: {len(preds)} predictions, {len(trues)} true boxes, model calls {model.calls}")


if __name__ == "__main__":
    main()
