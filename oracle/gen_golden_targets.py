"""Generates tests/golden/targets.npz by driving the UNMODIFIED reference `YOLODataset.__getitem__`
(code/dataset.py:119-167) on synthetic label lists: the object is created without __init__ (no CSV / image folder
needed), `load_image`, `load_boxes` and `apply_augmentations` are replaced by stubs that hand back the synthetic
boxes unchanged, everything after them is the reference's own code.

    python -m oracle.gen_golden_targets      # needs /root/reference -- not the GPU box
"""
import importlib
import os
import sys
import tempfile

import numpy as np
import pandas as pd
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_loader  # noqa: E402
from oracle import yolo_oracle as orc  # noqa: E402


def synth_boxes(n, nc, rng, anchors=None):
    """YOLO-format rows (x, y, w, h, class) as Python floats, the way albumentations hands them to the encoder.
    A third of the boxes are shaped like one of the anchors, so that several anchors pass the ignore threshold."""
    rows = []
    flat = [a for scale in anchors for a in scale] if anchors else []
    for k in range(n):
        w, h = float(rng.uniform(0.01, 0.9)), float(rng.uniform(0.01, 0.9))
        if flat and k % 3 == 2:
            aw, ah = flat[int(rng.integers(0, len(flat)))]
            w, h = min(0.95, float(aw * rng.uniform(0.8, 1.25))), min(0.95, float(ah * rng.uniform(0.8, 1.25)))
        x, y = float(rng.uniform(w / 2, 1 - w / 2)), float(rng.uniform(h / 2, 1 - h / 2))
        rows.append([min(x, 0.999999), min(y, 0.999999), w, h, float(rng.integers(0, nc))])
    return rows


def main():
    rmodel, rutils, rloss, rcfg = ref_loader.load()
    sys.modules.update(config=rcfg, utils=rutils)      # dataset.py imports them by bare name
    sys.path.insert(0, ref_loader.REF_CODE_DIRS[0])
    try:
        ds_mod = importlib.import_module("dataset")
    finally:
        sys.path.remove(ref_loader.REF_CODE_DIRS[0])
        for k in ("config", "utils", "dataset"):
            sys.modules.pop(k, None)
    rng = np.random.default_rng(11)
    out, n_case = {}, 0
    with tempfile.TemporaryDirectory() as td:
        for anchors_name, anchors in (("coco", orc.ANCHORS), ("turbine", orc.TURBINE_ANCHORS)):
            for size, counts in ((416, [0, 1, 3, 12, 40]), (608, [2, 25]), (320, [60])):
                grid = [size // 32, size // 16, size // 8]
                for n in counts:
                    boxes = synth_boxes(n, 80 if anchors_name == "coco" else 2, rng, anchors)
                    if n >= 12:   # crowd a few boxes into one cell so that "anchor taken" and the ignore rule fire
                        for k in range(1, 6):
                            boxes[k][0], boxes[k][1] = boxes[0][0] + 1e-3 * k, boxes[0][1] - 1e-3 * k
                            boxes[k][2], boxes[k][3] = boxes[0][2] * (1 + 0.03 * k), boxes[0][3] * (1 - 0.02 * k)
                    ds = object.__new__(ds_mod.YOLODataset)
                    ds.annotations = pd.DataFrame([["img.jpg", "label.txt"]])
                    ds.annotation_folder = td
                    open(os.path.join(td, "label.txt"), "w").close()
                    ds.anchors = torch.tensor(anchors[0] + anchors[1] + anchors[2])
                    ds.num_anchors, ds.num_anchors_per_scale = 9, 3
                    ds.grid_sizes, ds.ignore_iou_threshold = grid, 0.5
                    ds.load_image = lambda idx: None
                    ds.load_boxes = lambda path, idx, b=boxes: b
                    ds.apply_augmentations = lambda img, bx, idx: (torch.zeros(3, 8, 8), bx)
                    _, targets = ds[0]
                    out[f"c{n_case}/boxes"] = np.asarray(boxes, dtype=np.float64).reshape(-1, 5)
                    out[f"c{n_case}/meta"] = np.asarray([size, 0 if anchors_name == "coco" else 1], dtype=np.int32)
                    for s in range(3):
                        t = targets[s].numpy()
                        idx = np.argwhere(t[..., 4] != 0)                     # sparse: the non-empty cells only
                        out[f"c{n_case}/t{s}_idx"] = idx.astype(np.int32)
                        out[f"c{n_case}/t{s}_val"] = t[idx[:, 0], idx[:, 1], idx[:, 2]]
                    print(anchors_name, size, n, [int((targets[s][..., 4] == 1).sum()) for s in range(3)],
                          [int((targets[s][..., 4] == -1).sum()) for s in range(3)])
                    n_case += 1
    out["n"] = np.int32(n_case)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "targets.npz"), **out)


if __name__ == "__main__":
    main()
