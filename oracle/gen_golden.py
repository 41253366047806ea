"""Generates tests/golden/*.npz|json by running the UNMODIFIED reference in this container.

    python -m oracle.gen_golden            # needs /root/reference (read-only) -- not the GPU box

Test infrastructure only.  Every fixture stores the seeded INPUTS and the reference's OUTPUTS, so
the tests can check oracle/ (CPU, everywhere) and the CUDA path (GPU box) against the reference's
own results without the reference being present.  Reference functions exercised (paths relative
to the reference checkout): code/utils.py:38 calc_iou, :22 iou_aligned, :86 cells_to_boxes,
:150 non_max_suppression, :193 calc_mAP; code/model.py:150 YOLOv3 (forward, load_weights).
"""
from __future__ import annotations

import json
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader, synth  # noqa: E402
from oracle import yolo_oracle as orc  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def nms_cases():
    cases = []

    def add(name, boxes, iou_thr, obj_thr, fmt):
        cases.append(dict(name=name, boxes=boxes.clone(), iou_thr=iou_thr, obj_thr=obj_thr, fmt=fmt))

    add("center_nc3_ties", synth.synth_boxes(400, 3, 1, tie_frac=0.05), 0.45, 0.5, "center")
    add("corners_nc80_lowconf", synth.synth_boxes(600, 80, 2), 0.45, 0.01, "corners")
    add("dense_nc1_manyties", synth.synth_boxes(300, 1, 3, tie_frac=0.2, wh=(0.2, 0.5)), 0.3, 0.2, "center")
    b = synth.synth_boxes(256, 2, 4)
    b[::7, 4] = 0.5  # scores exactly at the (strict) threshold are dropped
    add("threshold_equal_scores", b, 0.45, 0.5, "center")
    add("all_filtered", synth.synth_boxes(64, 4, 5), 0.45, 2.0, "center")
    b = synth.synth_boxes(200, 2, 6, wh=(0.1, 0.4))
    b[3, 0] = float("nan")
    b[17, 2] = float("inf")
    b[40, 3] = float("nan")
    b[41, 5] = float("nan")  # NaN class label
    add("nan_inf_boxes", b, 0.45, 0.1, "center")
    b = synth.synth_boxes(150, 2, 7, wh=(0.3, 0.3))
    b[:, 2:4] = 0.3
    b[:, 0] = (torch.arange(150) % 10) * 0.03  # regular lattice => many IoUs land exactly on a few values
    b[:, 1] = (torch.arange(150) // 10) * 0.03
    add("lattice_equal_ious", b, float(torch.tensor(0.45)), 0.0, "corners")
    add("midpoint_alias_is_corners", synth.synth_boxes(120, 2, 8, wh=(0.1, 0.4)), 0.45, 0.3, "midpoint")
    return cases


def main():
    os.makedirs(GOLD, exist_ok=True)
    rmodel, rutils, rloss, rcfg = ref_loader.load()
    torch.manual_seed(0)

    # ---- NMS ------------------------------------------------------------------------------
    out = {}
    meta = []
    for c in nms_cases():
        kept = rutils.non_max_suppression(c["boxes"].tolist(), c["iou_thr"], c["obj_thr"], c["fmt"])
        out[c["name"] + "/boxes"] = c["boxes"].numpy()
        out[c["name"] + "/kept"] = np.asarray(kept, dtype=np.float32).reshape(-1, 6)
        meta.append(dict(name=c["name"], iou_thr=c["iou_thr"], obj_thr=c["obj_thr"], fmt=c["fmt"], n_kept=len(kept)))
        print("nms", c["name"], len(c["boxes"]), "->", len(kept))
    np.savez_compressed(os.path.join(GOLD, "nms.npz"), **out)
    json.dump(meta, open(os.path.join(GOLD, "nms.json"), "w"), indent=1)

    # ---- IoU ------------------------------------------------------------------------------
    g = torch.Generator().manual_seed(11)
    a = torch.rand(257, 4, generator=g)
    b = torch.rand(257, 4, generator=g)
    a[:, 2:] = 0.05 + 0.4 * a[:, 2:]
    b[:, 2:] = 0.05 + 0.4 * b[:, 2:]
    b[:8] = a[:8]  # identical boxes: 0.99998.., not 1.0 (utils.py:83)
    iou = dict(a=a.numpy(), b=b.numpy(),
               center=rutils.calc_iou(a, b, "center").numpy(), corners=rutils.calc_iou(a, b, "corners").numpy(),
               bcast=rutils.calc_iou(a[0], b, "center").numpy(),
               aligned=rutils.iou_aligned(a[:, 2:], b[:, 2:]).numpy(),
               aligned_kat=rutils.iou_aligned(torch.tensor([0.2, 0.3]), torch.tensor([[0.28, 0.22], [0.38, 0.48]])).numpy())
    np.savez_compressed(os.path.join(GOLD, "iou.npz"), **iou)

    # ---- decode ---------------------------------------------------------------------------
    dec = {}
    z = torch.zeros((5, 3, 3, 3, 8))  # utils_test.py:34-40
    anc = torch.tensor([[0.28, 0.22], [0.38, 0.48], [0.9, 0.78]])
    dec["zeros/in"] = z.numpy().copy()
    dec["zeros/anchors"] = anc.numpy()
    dec["zeros/out"] = np.asarray(rutils.cells_to_boxes(z.clone(), anc, 3), dtype=np.float32)
    for name, (bsz, s, nc, scale_i) in {"s13_nc80": (2, 13, 80, 0), "s26_nc2": (3, 26, 2, 1), "s8_nc5": (1, 8, 5, 2)}.items():
        g = torch.Generator().manual_seed(100 + s)
        # NHWC storage viewed as (B,3,S,S,C): the non-contiguous layout ScalePredictionBlock returns
        raw = (2.0 * torch.randn(bsz, s, s, 3, 5 + nc, generator=g)).permute(0, 3, 1, 2, 4)
        raw[0, 0, 0, 0, 5:] = 0.25  # all-equal logits: argmax must return the first index
        raw[0, 1, 1, 1, 7 if nc > 2 else 6] = raw[0, 1, 1, 1, 5]
        anchors = torch.tensor(orc.ANCHORS[scale_i]) * s
        dec[name + "/in"] = raw.contiguous().numpy().copy()
        dec[name + "/anchors"] = anchors.numpy()
        work = raw.clone()
        dec[name + "/out"] = np.asarray(rutils.cells_to_boxes(work, anchors, s, is_pred=True), dtype=np.float32)
        dec[name + "/mutated"] = work.contiguous().numpy().copy()  # the reference writes into its input
    g = torch.Generator().manual_seed(7)
    tgt = torch.rand(2, 3, 8, 8, 6, generator=g)
    tgt[..., 4] = (tgt[..., 4] > 0.8).float()
    tgt[..., 5] = torch.randint(0, 2, (2, 3, 8, 8), generator=g).float()
    anchors = torch.tensor(orc.TURBINE_ANCHORS[2]) * 8
    dec["target/in"] = tgt.numpy().copy()
    dec["target/anchors"] = anchors.numpy()
    dec["target/out"] = np.asarray(rutils.cells_to_boxes(tgt.clone(), anchors, 8, is_pred=False), dtype=np.float32)
    np.savez_compressed(os.path.join(GOLD, "decode.npz"), **dec)

    # ---- mAP ------------------------------------------------------------------------------
    mp = []
    kat_p = [[0, 0.5, 0.5, 0.25, 0.25, 0.9, 0], [0, 0.5, 0.5, 0.1, 0.1, 0.6, 0]]  # utils_test.py:22-32
    mp.append(dict(name="kat_all_detected", preds=kat_p, trues=kat_p, iou_thr=0.5, fmt="center", num_classes=20))
    mp.append(dict(name="kat_one_of_two", preds=kat_p[:1], trues=kat_p, iou_thr=0.5, fmt="center", num_classes=20))
    mp.append(dict(name="kat_no_dets", preds=[], trues=kat_p, iou_thr=0.5, fmt="center", num_classes=20))
    for seed, (n_img, nc, n_gt, n_det, fmt) in enumerate([(3, 3, 12, 40, "center"), (5, 2, 30, 120, "center"),
                                                          (2, 4, 10, 60, "corners"), (4, 20, 40, 150, "center")]):
        g = torch.Generator().manual_seed(500 + seed)
        gts = torch.empty(n_gt, 7)
        gts[:, 0] = torch.randint(0, n_img, (n_gt,), generator=g).float()
        gts[:, 1:3] = 0.2 + 0.6 * torch.rand(n_gt, 2, generator=g)
        gts[:, 3:5] = 0.1 + 0.3 * torch.rand(n_gt, 2, generator=g)
        gts[:, 5] = 1.0
        gts[:, 6] = torch.randint(0, nc, (n_gt,), generator=g).float()
        src = torch.randint(0, n_gt, (n_det,), generator=g)
        dets = gts[src].clone()
        dets[:, 1:5] += 0.05 * torch.randn(n_det, 4, generator=g)  # jitter: some match, some do not
        dets[:, 3:5] = dets[:, 3:5].abs() + 0.01
        dets[:, 5] = torch.rand(n_det, generator=g)
        dets[::9, 5] = dets[0, 5]  # score ties
        flip = torch.rand(n_det, generator=g) < 0.15
        dets[flip, 6] = torch.randint(0, nc, (int(flip.sum()),), generator=g).float()
        order = torch.argsort(dets[:, 0], stable=True)  # callers append image by image (utils.py:317-330)
        dets = dets[order]
        mp.append(dict(name=f"random_{seed}", preds=dets.tolist(), trues=gts.tolist(), iou_thr=0.5, fmt=fmt,
                       num_classes=nc))
    for c in mp:
        res = rutils.calc_mAP(c["preds"], c["trues"], c["iou_thr"], c["fmt"], c["num_classes"])
        c["mAP"] = float(res)
        c["mAP_hex"] = float(res).hex()
        print("mAP", c["name"], c["mAP"])
    json.dump(mp, open(os.path.join(GOLD, "map.json"), "w"))

    # ---- check_model_accuracy (utils.py:334-381) -------------------------------------------------
    class _FakeModel:  # the reference only needs eval()/train() and __call__
        def __init__(self, outs):
            self.outs = outs

        def eval(self):
            pass

        def train(self):
            pass

        def __call__(self, x):
            return [o.clone() for o in self.outs]

    acc = {}
    g = torch.Generator().manual_seed(321)
    nc, thr = 4, 0.6
    outs, tgts = [], []
    for s in (2, 4, 8):
        o = 2.0 * torch.randn(3, 3, s, s, 5 + nc, generator=g)
        # keep objectness logits away from the threshold so that 1-ulp sigmoid differences cannot flip a count
        lim = float(torch.logit(torch.tensor(thr)))
        near = (o[..., 4] - lim).abs() < 0.05
        o[..., 4][near] += 0.2
        t = torch.zeros(3, 3, s, s, 6)
        u = torch.rand(3, 3, s, s, generator=g)
        t[..., 4] = torch.where(u < 0.2, torch.tensor(1.0), torch.where(u < 0.3, torch.tensor(-1.0), torch.tensor(0.0)))
        t[..., 5] = torch.randint(0, nc, (3, 3, s, s), generator=g).float()
        agree = torch.rand(3, 3, s, s, generator=g) < 0.5  # make half of the labels agree with the argmax
        t[..., 5][agree] = torch.argmax(o[..., 5:], dim=-1).float()[agree]
        outs.append(o)
        tgts.append(t)
    res = rutils.check_model_accuracy(_FakeModel(outs), [(torch.zeros(3, 3, 64, 64), [t.clone() for t in tgts])], thr)
    for i in range(3):
        acc[f"out{i}"] = outs[i].numpy()
        acc[f"tgt{i}"] = tgts[i].numpy()
    acc["thr"] = np.float64(thr)
    acc["result"] = np.asarray([float(r) for r in res], dtype=np.float64)  # class, noobj, obj
    np.savez_compressed(os.path.join(GOLD, "accuracy.npz"), **acc)
    print("accuracy", acc["result"])

    # ---- YOLOLoss.forward (loss.py:29-81) ------------------------------------------------------
    lossd = {}
    for name, (bsz, s, nc, scale_i, with_obj) in {"s13_nc2": (4, 13, 2, 0, True), "s16_nc80": (1, 16, 80, 1, True),
                                                  "s8_noobj": (2, 8, 3, 2, False)}.items():
        g = torch.Generator().manual_seed(700 + s)
        pred = 1.5 * torch.randn(bsz, 3, s, s, 5 + nc, generator=g)
        tgt = torch.zeros(bsz, 3, s, s, 6)
        u = torch.rand(bsz, 3, s, s, generator=g)
        if with_obj:
            tgt[..., 4] = torch.where(u < 0.08, torch.tensor(1.0), torch.where(u < 0.12, torch.tensor(-1.0), torch.tensor(0.0)))
        tgt[..., 0:2] = torch.rand(bsz, 3, s, s, 2, generator=g)
        tgt[..., 2:4] = 0.5 + 3.5 * torch.rand(bsz, 3, s, s, 2, generator=g)     # grid units
        tgt[..., 5] = torch.randint(0, nc, (bsz, 3, s, s), generator=g).float()
        anchors = torch.tensor(orc.TURBINE_ANCHORS[scale_i]) * s
        p_work, t_work = pred.clone(), tgt.clone()
        out = rloss.YOLOLoss()(p_work, t_work, anchors)
        lossd[name + "/pred"] = pred.numpy()
        lossd[name + "/tgt"] = tgt.numpy()
        lossd[name + "/anchors"] = anchors.numpy()
        lossd[name + "/loss"] = np.asarray([float(v) for v in out], dtype=np.float64)
        lossd[name + "/pred_after"] = p_work.numpy()   # in-place side effects (loss.py:71-72)
        lossd[name + "/tgt_after"] = t_work.numpy()
        print("loss", name, lossd[name + "/loss"])
    np.savez_compressed(os.path.join(GOLD, "loss.npz"), **lossd)

    # ---- forward --------------------------------------------------------------------------
    fwd = {}
    for name, (nc, act, size, seed) in {"nc80_leaky_64": (80, "leaky_relu", 64, 0), "nc2_mish_96": (2, "mish", 96, 1)}.items():
        m = rmodel.YOLOv3(num_classes=nc, activation=act).eval()
        sd = synth.synth_state_dict(m.state_dict(), seed=seed)
        m.load_state_dict(sd)
        x = torch.rand(2, 3, size, size, generator=torch.Generator().manual_seed(900 + seed))
        with torch.no_grad():
            outs = m(x)
        fwd[name + "/x"] = x.numpy()
        for i, o in enumerate(outs):
            fwd[f"{name}/out{i}"] = o.contiguous().numpy()
        print("forward", name, [tuple(o.shape) for o in outs], float(outs[0].abs().mean()))
    np.savez_compressed(os.path.join(GOLD, "forward.npz"), **fwd)
    keys = {k: list(v.shape) for k, v in rmodel.YOLOv3(num_classes=80).state_dict().items()}
    json.dump(keys, open(os.path.join(GOLD, "state_dict_keys_nc80.json"), "w"))

    # ---- Darknet loader ---------------------------------------------------------------------
    load = {}
    with tempfile.TemporaryDirectory() as td:
        full = os.path.join(td, "yolov3.weights")
        synth.synth_darknet_file(full, 62001757)
        cut = os.path.join(td, "darknet53.conv.74")
        os.symlink(full, cut)
        for tag, path in (("full", full), ("cutoff74", cut)):
            torch.manual_seed(1234)
            m = rmodel.YOLOv3(num_classes=80, weights_path=path)
            before = {k: v.clone() for k, v in m.state_dict().items()}
            m.load_weights()
            after = m.state_dict()
            stats = {}
            for k, v in after.items():
                if v.dtype != torch.float32:
                    continue
                changed = not torch.equal(v, before[k])
                stats[k] = dict(changed=changed, sum=float(v.double().sum()) if changed else 0.0,
                                first=float(v.flatten()[0]) if changed else 0.0,
                                last=float(v.flatten()[-1]) if changed else 0.0)
            load[tag] = dict(param_idx=int(m.param_idx), layer_id=int(m.layer_id), n_floats=int(m.weights.size),
                             cutoff=m.cutoff, stats=stats)
            print("loader", tag, load[tag]["param_idx"], load[tag]["layer_id"],
                  sum(1 for s in stats.values() if s["changed"]), "tensors changed")
    json.dump(load, open(os.path.join(GOLD, "loader.json"), "w"))


if __name__ == "__main__":
    main()
