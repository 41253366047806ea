"""GPU parity: K3 decode, calc_iou / iou_aligned, K7 mAP matching, through the reference-shaped API."""
import numpy as np
import pytest
import torch

from oracle import synth
from oracle import yolo_oracle as orc

pytestmark = pytest.mark.gpu

DECODE_ATOL = 1e-5  # BASELINE.json north_star: decoded boxes within 1e-5 (fp32)


def _close(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.all(np.abs(a - b) <= DECODE_ATOL * np.maximum(1.0, np.abs(b)))


def test_decode_golden(gold):
    from yolo_for_turbines_b200.utils import cells_to_boxes

    d = gold.decode
    for name in ("zeros", "s13_nc80", "s26_nc2", "s8_nc5", "target"):
        x = torch.from_numpy(d[name + "/in"]).clone()
        out = np.asarray(cells_to_boxes(x, torch.from_numpy(d[name + "/anchors"]), x.shape[2], is_pred=(name != "target")),
                         dtype=np.float32)
        ref = d[name + "/out"]
        assert out.shape == ref.shape, name
        assert np.array_equal(out[..., 5], ref[..., 5]), name + ": class index (argmax, first on ties) must be exact"
        assert _close(out[..., :5], ref[..., :5]), (name, float(np.abs(out - ref).max()))
        if name + "/mutated" in d.files:  # the reference rewrites predictions[..., :4] in place
            assert _close(x.numpy()[..., :4], d[name + "/mutated"][..., :4]), name
            assert np.array_equal(x.numpy()[..., 4:], d[name + "/mutated"][..., 4:]), name


def test_decode_strided_head_layout_vs_oracle():
    """The model hands decode a non-contiguous (B,3,S,S,85) view of an NHWC buffer with pitch 256."""
    from yolo_for_turbines_b200.utils import decode_boxes

    B, S, nc = 3, 52, 80
    g = torch.Generator().manual_seed(3)
    buf = 1.5 * torch.randn(B, S, S, 256, generator=g)
    view = torch.as_strided(buf, (B, 3, S, S, 85), (S * S * 256, 85, S * 256, 256, 1))
    anchors = torch.tensor(orc.ANCHORS[2]) * S
    ref = np.asarray(orc.cells_to_boxes(view.clone(), anchors, S), dtype=np.float32)
    out = decode_boxes(torch.as_strided(buf.cuda(), view.shape, view.stride()), anchors, S).cpu().numpy()
    assert np.array_equal(out[..., 5], ref[..., 5])
    assert _close(out[..., :5], ref[..., :5])


def test_iou_golden(gold):
    from yolo_for_turbines_b200.utils import calc_iou, iou_aligned

    a, b = torch.from_numpy(gold.iou["a"]), torch.from_numpy(gold.iou["b"])
    assert np.array_equal(calc_iou(a, b, "center").cpu().numpy(), gold.iou["center"])  # bit-exact op order
    assert np.array_equal(calc_iou(a, b, "corners").cpu().numpy(), gold.iou["corners"])
    assert np.array_equal(calc_iou(a[0], b, "center").cpu().numpy(), gold.iou["bcast"])
    assert np.array_equal(iou_aligned(a[:, 2:], b[:, 2:]).cpu().numpy(), gold.iou["aligned"])
    kat = iou_aligned(torch.tensor([0.2, 0.3]), torch.tensor([[0.28, 0.22], [0.38, 0.48]]))
    assert np.array_equal(kat.cpu().numpy(), gold.iou["aligned_kat"])


def test_map_golden(gold):
    from yolo_for_turbines_b200.utils import calc_mAP

    for c in gold.map:
        res = calc_mAP(c["preds"], c["trues"], c["iou_thr"], c["fmt"], c["num_classes"])
        assert res.dim() == 0
        assert abs(float(res) - c["mAP"]) <= 1e-6, (c["name"], float(res), c["mAP"])
    with pytest.raises(ZeroDivisionError):
        calc_mAP([[0, .5, .5, .1, .1, .9, 0]], [], 0.5, "center", 3)


def test_map_tp_flags_match_oracle_exactly():
    """The matching step (TP/FP per detection) is integer work: bit-exact against the oracle."""
    from yolo_for_turbines_b200.utils import map_match

    g = torch.Generator().manual_seed(77)
    n_img, nc, n_gt, n_det = 6, 4, 80, 600
    gts = torch.empty(n_gt, 7)
    gts[:, 0] = torch.randint(0, n_img, (n_gt,), generator=g).float()
    gts[:, 1:3] = 0.2 + 0.6 * torch.rand(n_gt, 2, generator=g)
    gts[:, 3:5] = 0.1 + 0.3 * torch.rand(n_gt, 2, generator=g)
    gts[:, 5] = 1.0
    gts[:, 6] = torch.randint(0, nc, (n_gt,), generator=g).float()
    dets = gts[torch.randint(0, n_gt, (n_det,), generator=g)].clone()
    dets[:, 1:5] += 0.04 * torch.randn(n_det, 4, generator=g)
    dets[:, 3:5] = dets[:, 3:5].abs() + 0.01
    dets[:, 5] = torch.rand(n_det, generator=g)
    dets[::11, 5] = dets[0, 5]
    dets = dets[torch.argsort(dets[:, 0], stable=True)]
    ref_map, tp_rows = orc.calc_mAP(dets.tolist(), gts.tolist(), 0.5, "center", nc, return_tp=True)
    _, _, _, _, tp = map_match(dets.cuda(), gts.cuda(), 0.5, "center")
    tp = tp.cpu().tolist()
    assert all(tp[r] == v for r, v in tp_rows.items())
    from yolo_for_turbines_b200.utils import calc_mAP
    assert abs(float(calc_mAP(dets.tolist(), gts.tolist(), 0.5, "center", nc)) - float(ref_map)) <= 1e-6


def test_accuracy_counts_match_reference(gold):
    """check_model_accuracy reductions (utils.py:334-381): integer counts exact, ratios equal to the reference's."""
    from yolo_for_turbines_b200.utils import accuracy_counts

    a = gold.accuracy
    outs = [torch.from_numpy(a[f"out{i}"]) for i in range(3)]
    tgts = [torch.from_numpy(a[f"tgt{i}"]) for i in range(3)]
    _, ref_counts = orc.check_model_accuracy(outs, tgts, float(a["thr"]))
    # heads as the model hands them over: non-contiguous (B,3,S,S,C) views of NHWC storage
    dev_outs = [o.permute(0, 2, 3, 1, 4).contiguous().cuda().permute(0, 3, 1, 2, 4) for o in outs]
    c = accuracy_counts(dev_outs, tgts, float(a["thr"])).cpu().tolist()
    assert c == ref_counts
    got = [c[0] / (c[1] + 1e-16), c[4] / (c[5] + 1e-16), c[2] / (c[3] + 1e-16)]
    assert np.allclose(got, a["result"], rtol=1e-6)


def test_yolo_loss_forward_matches_reference(gold):
    """Fused YOLOLoss forward (loss.py:29-81).  Stated tolerance: 1e-5 relative per term (fp32 reference sums vs
    fp64 device accumulation); in-place side effects within 1e-6."""
    from yolo_for_turbines_b200.loss import YOLOLoss

    d = gold.loss
    for name in ("s13_nc2", "s16_nc80", "s8_noobj"):
        p = torch.from_numpy(d[name + "/pred"]).cuda()
        t = torch.from_numpy(d[name + "/tgt"]).cuda()
        with torch.no_grad():
            out = YOLOLoss()(p, t, torch.from_numpy(d[name + "/anchors"]))
        got = np.asarray([float(v) for v in out])
        ref = d[name + "/loss"]
        assert len(out) == 4 and all(v.dim() == 0 and v.dtype == torch.float32 for v in out)
        assert np.allclose(got, ref, rtol=1e-5, atol=1e-7), (name, got, ref)
        assert np.allclose(p.cpu().numpy(), d[name + "/pred_after"], rtol=1e-6, atol=1e-6), name
        assert np.allclose(t.cpu().numpy(), d[name + "/tgt_after"], rtol=1e-6, atol=1e-6), name
    # with a gradient-tracking input the four terms are one autograd node (tests/test_gpu_train_kernels.py pins the values)
    p = torch.zeros(1, 3, 4, 4, 7, device="cuda", requires_grad=True)
    out = YOLOLoss()(p, torch.zeros(1, 3, 4, 4, 6, device="cuda"), [[1, 1]] * 3)
    sum(out).backward()
    assert p.grad is not None and float(p.grad[..., 4].min()) > 0 and float(p.grad[..., :4].abs().max()) == 0


@pytest.mark.gpu
def test_map_ap_kernel_equals_the_reference_tail_per_class():
    """yolo_map_ap (one CTA per class) against utils.py:262-272 written out with torch on the CPU: cumsum, precision /
    recall with (1, 0) prepended, torch.trapz -- per class, including a class without detections, a class without
    ground truth, a one-detection class and a class longer than several scan chunks."""
    from yolo_for_turbines_b200._lib import lib, ptr, stream_ptr

    g = torch.Generator().manual_seed(5)
    sizes = [0, 1, 7, 256, 257, 3000, 0, 12]
    n_gt = [3, 1, 0, 100, 9, 2500, 0, 40]
    tps = [(torch.rand(n, generator=g) < 0.4).float() for n in sizes]
    tp = torch.cat(tps) if sum(sizes) else torch.zeros(0)
    off = [0]
    for n in sizes:
        off.append(off[-1] + n)
    dev = torch.device("cuda")
    ap = torch.full((len(sizes),), -1.0, device=dev)
    tp_d = tp.to(dev)      # named: the launch is asynchronous, temporaries would be recycled under it
    lo_d = torch.tensor(off[:-1], dtype=torch.int32, device=dev)
    hi_d = torch.tensor(off[1:], dtype=torch.int32, device=dev)
    ng_d = torch.tensor(n_gt, dtype=torch.int32, device=dev)
    lib.yolo_map_ap(ptr(tp_d), ptr(lo_d), ptr(hi_d), ptr(ng_d), len(sizes), ptr(ap), stream_ptr(dev))
    got = ap.cpu()
    for c, (t, ng) in enumerate(zip(tps, n_gt)):
        if ng == 0 or t.numel() == 0:
            assert float(got[c]) == 0.0
            continue
        ctp = torch.cumsum(t, 0)
        cfp = torch.cumsum(1 - t, 0)
        prec = torch.cat([torch.ones(1), ctp / (ctp + cfp)])
        rec = torch.cat([torch.zeros(1), ctp / ng])
        ref = float(torch.trapz(prec, rec))
        assert abs(float(got[c]) - ref) <= 1e-6 * max(1.0, abs(ref)), (c, float(got[c]), ref)


@pytest.mark.gpu
@pytest.mark.parametrize("nc,pitch", [(80, 256), (2, 32)])
def test_decode_multi_equals_per_scale_decode(nc, pitch):
    """yolo_decode_multi (all scales in one launch) writes exactly what three yolo_decode calls write; heads that are
    not the model's dense layout take the per-scale route through the same wrapper."""
    from yolo_for_turbines_b200.utils import decode_boxes, decode_boxes_multi

    g = torch.Generator().manual_seed(9)
    B, C = 3, 5 + nc
    heads, anchors = [], []
    for S in (4, 8, 16):
        buf = (torch.randn(B * S * S, pitch, generator=g) * 2).cuda()       # [pixels][pitch], channels (a, c) = a * C + c
        heads.append(buf[:, : 3 * C].view(B, S, S, 3, C).permute(0, 3, 1, 2, 4))
        anchors.append(torch.rand(3, 2, generator=g) * S)
    n = sum(3 * h.shape[2] ** 2 for h in heads)
    ref = torch.zeros(B, n, 6, device="cuda")
    off = 0
    for h, a in zip(heads, anchors):
        decode_boxes(h, a, h.shape[2], True, out=ref, out_offset=off)
        off += 3 * h.shape[2] ** 2
    got = torch.zeros(B, n, 6, device="cuda")
    decode_boxes_multi(heads, anchors, got)
    assert torch.equal(got, ref)
    got2 = torch.zeros(B, n, 6, device="cuda")
    decode_boxes_multi([h.contiguous() for h in heads], anchors, got2)       # (B,3,S,S,C)-contiguous: generic kernel
    assert torch.equal(got2, ref)
