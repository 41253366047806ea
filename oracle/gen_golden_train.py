"""Generates tests/golden/train_step.npz by running one training step of the UNMODIFIED reference (CPU, fp32).

    python -m oracle.gen_golden_train      # needs /root/reference (read-only) -- not the GPU box

Test infrastructure only.  Reference code exercised: code/model.py:150 YOLOv3 in train() mode, code/loss.py:29
YOLOLoss on the three scales, and the loss assembly + backward of code/train.py:53-67 (without autocast: CPU fp32).
Stored per case: the seeded input, the synthetic targets, the four loss terms, and for EVERY parameter gradient its
L2 norm and 16 evenly spaced samples (full tensors for the BatchNorm / bias gradients of a few layers); plus the
running statistics of the first and last BatchNorm after the step.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader, synth  # noqa: E402
from oracle import yolo_oracle as orc  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
CASES = {"nc2_leaky_64": (2, "leaky_relu", 64, 2, 5), "nc2_mish_96": (2, "mish", 96, 2, 6), "nc80_leaky_64": (80, "leaky_relu", 64, 3, 7)}
FULL = ("layers.0.batch_norm.weight", "layers.0.batch_norm.bias", "layers.10.layers.3.1.batch_norm.weight",
        "layers.15.pred_block.1.conv.bias", "layers.29.pred_block.1.conv.bias", "layers.16.batch_norm.bias")


def sample_idx(n):
    return np.linspace(0, n - 1, num=min(16, n)).astype(np.int64)


def main():
    rmodel, _, rloss, _ = ref_loader.load()
    out = {}
    for name, (nc, act, size, bsz, seed) in CASES.items():
        m = rmodel.YOLOv3(num_classes=nc, activation=act)
        m.load_state_dict(synth.synth_state_dict(m.state_dict(), seed=seed))
        m.train()
        x = torch.rand(bsz, 3, size, size, generator=torch.Generator().manual_seed(40 + seed))
        tg = orc.synth_targets(bsz, size, nc, 50 + seed)
        outs = m(x)
        lf = rloss.YOLOLoss()
        terms = [0, 0, 0, 0]
        for i, (o, t) in enumerate(zip(outs, tg)):     # train.py:56-65
            per = lf(o, t.clone(), torch.tensor(orc.TURBINE_ANCHORS[i]) * o.shape[2])
            terms = [a + b for a, b in zip(terms, per)]
        sum(terms).backward()
        out[name + "/x"] = x.numpy()
        for i, t in enumerate(tg):
            out[f"{name}/t{i}"] = t.numpy()
        out[name + "/loss"] = np.asarray([float(v.detach()) for v in terms], dtype=np.float64)
        keys, norms, samples = [], [], []
        for k, p in m.named_parameters():
            g = p.grad.detach().flatten()
            keys.append(k)
            norms.append(float(g.double().norm()))
            s = np.zeros(16, dtype=np.float32)
            idx = sample_idx(g.numel())
            s[: len(idx)] = g[idx].numpy()
            samples.append(s)
            if k in FULL:
                out[f"{name}/grad/{k}"] = p.grad.detach().numpy()
        out[name + "/keys"] = np.asarray(keys)
        out[name + "/norms"] = np.asarray(norms, dtype=np.float64)
        out[name + "/samples"] = np.stack(samples)
        sd = m.state_dict()
        for k in ("layers.0.batch_norm.running_mean", "layers.0.batch_norm.running_var",
                  "layers.29.pred_block.0.batch_norm.running_mean", "layers.29.pred_block.0.batch_norm.running_var"):
            out[f"{name}/after/{k}"] = sd[k].numpy()
        print(name, out[name + "/loss"], "grad norm", float(np.sqrt((np.asarray(norms) ** 2).sum())))
    np.savez_compressed(os.path.join(GOLD, "train_step.npz"), **out)


if __name__ == "__main__":
    main()
