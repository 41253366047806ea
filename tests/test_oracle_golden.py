"""Pins oracle/ (the CPU restatement) against outputs of the UNMODIFIED reference, stored in
tests/golden/ by oracle/gen_golden.py.  Runs without a GPU and without /root/reference."""
import math
import os
import tempfile

import numpy as np
import pytest
import torch

from conftest import template_state_dict
from oracle import synth
from oracle import yolo_oracle as orc


def _rows_equal(a, b):
    """bit-level equality that treats NaN == NaN (kept rows may legitimately hold NaNs)"""
    a = np.asarray(a, dtype=np.float32).reshape(-1, 6)
    b = np.asarray(b, dtype=np.float32).reshape(-1, 6)
    return a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32)) or \
        (a.shape == b.shape and np.array_equal(np.nan_to_num(a, nan=-777.0), np.nan_to_num(b, nan=-777.0)))


def test_nms_python_oracle_matches_reference(gold):
    for m in gold.nms_meta:
        boxes = torch.from_numpy(gold.nms[m["name"] + "/boxes"])
        kept = orc.non_max_suppression(boxes.tolist(), m["iou_thr"], m["obj_thr"], m["fmt"])
        assert len(kept) == m["n_kept"], m["name"]
        assert _rows_equal(kept, gold.nms[m["name"] + "/kept"]), m["name"]


def test_nms_index_oracles_match_reference(gold, oracle_c):
    for m in gold.nms_meta:
        boxes = torch.from_numpy(gold.nms[m["name"] + "/boxes"])
        ref_rows = gold.nms[m["name"] + "/kept"]
        for impl in (orc.nms_keep_indices, oracle_c):
            idx = impl(boxes, m["iou_thr"], m["obj_thr"], m["fmt"])
            assert _rows_equal(boxes[idx].numpy() if idx else np.zeros((0, 6)), ref_rows), (m["name"], impl)


def test_nms_c_oracle_equals_python_oracle_on_larger_sets(oracle_c):
    for seed, (n, nc, conf, fmt) in enumerate([(3000, 80, 0.5, "center"), (2500, 2, 0.01, "center"), (2000, 5, 0.3, "corners")]):
        b = synth.synth_boxes(n, nc, 100 + seed, tie_frac=0.02, wh=(0.05, 0.4))
        assert oracle_c(b, 0.45, conf, fmt) == orc.nms_keep_indices(b, 0.45, conf, fmt)


def test_iou_matches_reference(gold):
    a, b = torch.from_numpy(gold.iou["a"]), torch.from_numpy(gold.iou["b"])
    assert torch.equal(orc.calc_iou(a, b, "center"), torch.from_numpy(gold.iou["center"]))
    assert torch.equal(orc.calc_iou(a, b, "corners"), torch.from_numpy(gold.iou["corners"]))
    assert torch.equal(orc.calc_iou(a[0], b, "center"), torch.from_numpy(gold.iou["bcast"]))
    assert torch.equal(orc.iou_aligned(a[:, 2:], b[:, 2:]), torch.from_numpy(gold.iou["aligned"]))
    kat = orc.iou_aligned(torch.tensor([0.2, 0.3]), torch.tensor([[0.28, 0.22], [0.38, 0.48]]))
    assert torch.equal(kat, torch.from_numpy(gold.iou["aligned_kat"]))
    assert torch.allclose(kat, torch.tensor([0.5670, 0.3289]), atol=1e-4)  # SURVEY 8c KAT
    same = orc.calc_iou(torch.tensor([0.5, 0.5, 0.25, 0.25]), torch.tensor([0.5, 0.5, 0.25, 0.25]))
    assert 0.9999 < float(same) < 1.0  # +1e-6 in the denominator: not 1.0 (utils.py:83)


def test_cells_to_boxes_matches_reference(gold):
    d = gold.decode
    for name in ("zeros", "s13_nc80", "s26_nc2", "s8_nc5", "target"):
        x = torch.from_numpy(d[name + "/in"]).clone()
        anchors = torch.from_numpy(d[name + "/anchors"])
        s = x.shape[2]
        out = orc.cells_to_boxes(x, anchors, s, is_pred=(name != "target"))
        assert np.array_equal(np.asarray(out, dtype=np.float32), d[name + "/out"]), name
        if name + "/mutated" in d.files:
            assert np.array_equal(x.numpy(), d[name + "/mutated"]), name  # in-place mutation preserved
    z = np.asarray(d["zeros/out"])
    assert z.shape == (5, 27, 6)  # utils_test.py:34-40
    assert np.allclose(z[0, 4], [(0.5 + 1) / 3, (0.5 + 1) / 3, 0.28 / 3, 0.22 / 3, 0.5, 0.0], atol=1e-6)


def test_map_matches_reference(gold):
    for c in gold.map:
        res = orc.calc_mAP(c["preds"], c["trues"], c["iou_thr"], c["fmt"], c["num_classes"])
        assert float(res).hex() == c["mAP_hex"], c["name"]
    with pytest.raises(ZeroDivisionError):
        orc.calc_mAP([[0, .5, .5, .1, .1, .9, 0]], [], 0.5, "center", 3)


def test_forward_matches_reference(gold):
    for name, nc, act, seed in (("nc80_leaky_64", 80, "leaky_relu", 0), ("nc2_mish_96", 2, "mish", 1)):
        keys = dict(gold.keys)
        if nc != 80:  # head conv shapes depend on num_classes
            for k, v in keys.items():
                if ".pred_block.1.conv." in k:
                    keys[k] = [3 * (nc + 5)] + v[1:]
        sd = synth.synth_state_dict(template_state_dict(keys), seed=seed)
        x = torch.from_numpy(gold.forward[name + "/x"])
        with torch.no_grad():
            outs = orc.forward(sd, x, nc, act)
        for i, o in enumerate(outs):
            ref = torch.from_numpy(gold.forward[f"{name}/out{i}"])
            assert o.shape == ref.shape
            assert torch.allclose(o, ref, atol=1e-5, rtol=1e-5), (name, i, float((o - ref).abs().max()))
    with pytest.raises(AssertionError):
        orc.forward(sd, torch.full((1, 3, 32, 32), float("nan")), nc, act)


def test_darknet_loader_matches_reference(gold):
    with tempfile.TemporaryDirectory() as td:
        full = os.path.join(td, "yolov3.weights")
        synth.synth_darknet_file(full, 62001757)
        cut = os.path.join(td, "darknet53.conv.74")
        os.symlink(full, cut)
        for tag, path in (("full", full), ("cutoff74", cut)):
            sd = template_state_dict(gold.keys)
            info = orc.read_darknet_weights(path, sd)
            g = gold.loader[tag]
            assert info["param_idx"] == g["param_idx"] == 62001757
            assert info["layer_id"] == g["layer_id"]
            assert info["n_floats"] == g["n_floats"]
            for k, st in g["stats"].items():
                if st["changed"]:
                    assert float(sd[k].double().sum()) == st["sum"], (tag, k)
                    assert float(sd[k].flatten()[0]) == st["first"] and float(sd[k].flatten()[-1]) == st["last"]
                else:
                    assert float(sd[k].abs().sum()) == 0.0, (tag, k)  # untouched by the cutoff
            if tag == "cutoff74":
                assert info["loaded_bn_convs"] == 37  # not 52/74


def test_accuracy_reductions_match_reference(gold):
    a = gold.accuracy
    outs = [torch.from_numpy(a[f"out{i}"]) for i in range(3)]
    tgts = [torch.from_numpy(a[f"tgt{i}"]) for i in range(3)]
    (ca, na, oa), counts = orc.check_model_accuracy(outs, tgts, float(a["thr"]))
    assert [float(ca), float(na), float(oa)] == list(a["result"])
    assert counts[1] == counts[3] > 0 and counts[5] > 0


def test_yolo_loss_matches_reference(gold):
    d = gold.loss
    for name in ("s13_nc2", "s16_nc80", "s8_noobj"):
        p, t = torch.from_numpy(d[name + "/pred"]).clone(), torch.from_numpy(d[name + "/tgt"]).clone()
        out = orc.yolo_loss(p, t, torch.from_numpy(d[name + "/anchors"]))
        assert [float(v) for v in out] == list(d[name + "/loss"]), name
        assert np.array_equal(p.numpy(), d[name + "/pred_after"]) and np.array_equal(t.numpy(), d[name + "/tgt_after"]), name


def test_get_eval_boxes_oracle_matches_reference(gold):
    """oracle.get_eval_boxes vs the lists the UNMODIFIED reference returned (utils.py:276-332, oracle/gen_golden_eval.py)."""
    e = gold.eval_boxes
    batches = [([torch.from_numpy(e[f"b{bi}_head{si}"]) for si in range(3)], [torch.from_numpy(e[f"b{bi}_tgt{si}"]) for si in range(3)])
               for bi in range(2)]
    preds, trues = orc.get_eval_boxes(batches, e["anchors"].tolist(), float(e["iou_thr"]), float(e["obj_thr"]), "center")
    assert np.array_equal(np.asarray(preds, dtype=np.float64).reshape(-1, 7), e["preds"])
    assert np.array_equal(np.asarray(trues, dtype=np.float64).reshape(-1, 7), e["trues"])
    assert e["calls"].tolist() == ["eval", "train"]   # utils.py:295 and the unconditional :331
