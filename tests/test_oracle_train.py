"""Pins the oracle's training-step restatement (oracle.train_step_grads: train-mode forward + YOLOLoss + backward,
code/train.py:53-67) against gradients produced by the UNMODIFIED reference (tests/golden/train_step.npz, written
by oracle/gen_golden_train.py).  CPU only."""
import os

import numpy as np
import torch

from oracle import synth
from oracle import yolo_oracle as orc
from oracle.gen_golden_train import CASES, FULL, sample_idx
from conftest import template_state_dict

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "train_step.npz")


def _template(nc):
    import json
    keys = json.load(open(os.path.join(os.path.dirname(GOLD), "state_dict_keys_nc80.json")))
    t = template_state_dict(keys)
    if nc != 80:
        for k in list(t):
            if ".pred_block.1.conv." in k:
                shape = list(t[k].shape)
                shape[0] = 3 * (nc + 5)
                t[k] = torch.zeros(shape, dtype=t[k].dtype)
    return t


def test_train_step_oracle_matches_reference():
    z = np.load(GOLD)
    for name, (nc, act, size, bsz, seed) in CASES.items():
        if size > 64:
            continue  # one case per activation is enough on CPU; the 96x96 mish case runs below
        sd = synth.synth_state_dict(_template(nc), seed=seed)
        x = torch.from_numpy(z[name + "/x"])
        tg = [torch.from_numpy(z[f"{name}/t{i}"]) for i in range(3)]
        terms, grads = orc.train_step_grads(sd, x, tg, orc.TURBINE_ANCHORS, nc, act)
        assert np.allclose(terms, z[name + "/loss"], rtol=1e-6), name
        keys = [str(k) for k in z[name + "/keys"]]
        for i, k in enumerate(keys):
            g = grads[k].flatten()
            assert abs(float(g.double().norm()) - z[name + "/norms"][i]) <= 1e-5 * max(1.0, z[name + "/norms"][i]), (name, k)
            idx = sample_idx(g.numel())
            assert np.allclose(g[idx].numpy(), z[name + "/samples"][i][: len(idx)], rtol=1e-4, atol=1e-6), (name, k)
        for k in FULL:
            assert np.allclose(grads[k].numpy(), z[f"{name}/grad/{k}"], rtol=1e-4, atol=1e-6), (name, k)
        for k in ("layers.0.batch_norm.running_mean", "layers.29.pred_block.0.batch_norm.running_var"):
            assert np.allclose(sd[k].numpy(), z[f"{name}/after/{k}"], rtol=1e-5, atol=1e-7), (name, k)


def test_train_step_oracle_mish():
    z = np.load(GOLD)
    name = "nc2_mish_96"
    nc, act, size, bsz, seed = CASES[name]
    sd = synth.synth_state_dict(_template(nc), seed=seed)
    x = torch.from_numpy(z[name + "/x"])
    tg = [torch.from_numpy(z[f"{name}/t{i}"]) for i in range(3)]
    terms, grads = orc.train_step_grads(sd, x, tg, orc.TURBINE_ANCHORS, nc, act)
    assert np.allclose(terms, z[name + "/loss"], rtol=1e-6)
    assert np.allclose([float(grads[str(k)].double().norm()) for k in z[name + "/keys"]], z[name + "/norms"], rtol=1e-5)


def test_synth_targets_are_deterministic():
    a, b = orc.synth_targets(2, 64, 2, 3), orc.synth_targets(2, 64, 2, 3)
    assert all(torch.equal(x, y) for x, y in zip(a, b))
    assert [tuple(t.shape) for t in a] == [(2, 3, 2, 2, 6), (2, 3, 4, 4, 6), (2, 3, 8, 8, 6)]
    assert int((a[2][..., 4] == 1).sum()) == 8 and int((a[2][..., 4] == -1).sum()) == 2
