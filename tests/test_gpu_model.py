"""GPU parity of the whole drop-in module: YOLOv3.forward on the tcgen05 path against the oracle's
fp32 forward on identical (synthetic, seeded) weights, plus the fused detect pipeline.

Stated bf16 tolerance for head outputs (75 bf16 layers deep, fp32 accumulate): cosine >= 0.999 and
max-abs error <= 0.05 * (1 + max|ref|) per scale; both are printed."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import template_state_dict
from oracle import synth
from oracle import yolo_oracle as orc

pytestmark = pytest.mark.gpu


def _model(nc=80, act="leaky_relu", seed=0):
    from yolo_for_turbines_b200.model import YOLOv3

    m = YOLOv3(num_classes=nc, activation=act).eval()
    sd = synth.synth_state_dict(m.state_dict(), seed=seed)
    m.load_state_dict(sd)
    return m.cuda(), sd


def _compare(outs, refs, tag):
    for i, (o, r) in enumerate(zip(outs, refs)):
        o = o.float().cpu()
        assert o.shape == r.shape, (tag, i, o.shape, r.shape)
        cos = float(F.cosine_similarity(o.flatten(), r.flatten(), dim=0))
        mx = float((o - r).abs().max())
        print(f"{tag} scale {i}: cosine {cos:.6f} max-abs {mx:.4f} (max|ref| {float(r.abs().max()):.3f})")
        assert cos >= 0.999, (tag, i, cos)
        assert mx <= 0.05 * (1 + float(r.abs().max())), (tag, i, mx)


def test_forward_matches_reference_golden(gold):
    for name, nc, act, seed in (("nc80_leaky_64", 80, "leaky_relu", 0), ("nc2_mish_96", 2, "mish", 1)):
        m, _ = _model(nc, act, seed)
        x = torch.from_numpy(gold.forward[name + "/x"]).cuda()
        with torch.no_grad():
            outs = m(x)
        _compare(outs, [torch.from_numpy(gold.forward[f"{name}/out{i}"]) for i in range(3)], name)


def test_forward_416_matches_oracle_and_graph_replay_is_stable():
    m, sd = _model(80, "leaky_relu", 3)
    x = torch.rand(2, 3, 416, 416, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        refs = orc.forward(sd, x, 80)
        outs = m(x.cuda())           # eager warm-up + graph capture
        outs2 = m(x.cuda())          # graph replay
    assert [tuple(o.shape) for o in outs] == [(2, 3, 13, 13, 85), (2, 3, 26, 26, 85), (2, 3, 52, 52, 85)]
    _compare(outs, refs, "416")
    for a, b in zip(outs, outs2):
        assert torch.equal(a, b)
    # multi-scale: another input size on the same module (config 4's 320..608 sweep)
    x2 = torch.rand(1, 3, 320, 320, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        _compare(m(x2.cuda()), orc.forward(sd, x2, 80), "320")


def test_nan_guards_keep_reference_error_behaviour():
    m, _ = _model(2, "leaky_relu", 0)
    x = torch.rand(1, 3, 64, 64)
    x[0, 1, 5, 5] = float("nan")
    with pytest.raises(AssertionError):   # model.py:175
        m(x.cuda())
    with torch.no_grad():
        m.layers[0].batch_norm.bias[3] = float("nan")
    with pytest.raises(ValueError, match="Nan in layer"):   # model.py:183-184
        m(torch.rand(1, 3, 64, 64).cuda())


def test_state_dict_update_is_picked_up():
    m, sd = _model(2, "leaky_relu", 0)
    x = torch.rand(1, 3, 64, 64).cuda()
    a = m(x)
    sd2 = synth.synth_state_dict(m.state_dict(), seed=5)
    m.load_state_dict(sd2)
    b = m(x)
    assert not torch.equal(a[0], b[0])
    with torch.no_grad():
        _compare(b, orc.forward(sd2, x.cpu(), 2), "reload")


def test_detect_pipeline_matches_oracle_on_device_boxes():
    """End to end: forward -> decode -> NMS.  The conv path is bf16, so kept sets are compared given the
    DEVICE's decoded boxes: oracle NMS over the GPU candidates must return exactly the GPU survivors."""
    from yolo_for_turbines_b200.utils import Detector

    m, sd = _model(80, "leaky_relu", 7)
    x = torch.rand(3, 3, 416, 416, generator=torch.Generator().manual_seed(9)).cuda()
    det = Detector(m, orc.ANCHORS, 0.45, 0.5, "center")
    res, plan = det(x)
    plan.check_status()
    cand = res.boxes.view(3, -1, 6).cpu()
    assert cand.shape[1] == 10647
    got = res.to_lists()
    for b in range(3):
        exp = orc.nms_keep_indices(cand[b], 0.45, 0.5, "center")
        assert np.array_equal(np.asarray(got[b], dtype=np.float32), cand[b][exp].numpy())
    # decoded boxes against the oracle's decode of the SAME head tensors: 1e-5
    _, heads = m.forward_async(x)
    off = 0
    for i, h in enumerate(heads):
        s = h.shape[2]
        ref = np.asarray(orc.cells_to_boxes(h.cpu().clone(), torch.tensor(orc.ANCHORS[i]) * s, s), dtype=np.float32)
        got_d = cand[:, off:off + 3 * s * s].numpy()
        assert np.array_equal(got_d[..., 5], ref[..., 5])
        assert np.all(np.abs(got_d[..., :5] - ref[..., :5]) <= 1e-5 * np.maximum(1, np.abs(ref[..., :5])))
        off += 3 * s * s


def test_standalone_blocks_match_reference_shapes():
    """model_tests.py:16-45 shape checks, with values checked against the oracle arithmetic."""
    from yolo_for_turbines_b200.model import CNNBlock, ResidualBlock, ScalePredictionBlock

    torch.manual_seed(0)
    blk = CNNBlock(3, 32, kernel_size=3, stride=1, padding=1).eval().cuda()
    x = torch.rand(2, 3, 64, 64)
    y = blk(x.cuda())
    assert y.shape == (2, 32, 64, 64)
    ref = F.leaky_relu(F.batch_norm(F.conv2d(x, blk.conv.weight.cpu(), None, 1, 1), blk.batch_norm.running_mean.cpu(),
                                    blk.batch_norm.running_var.cpu(), blk.batch_norm.weight.cpu(), blk.batch_norm.bias.cpu(),
                                    False, 0.1, 1e-5), 0.1)
    assert float((y.cpu() - ref).abs().max()) < 0.05
    rb = ResidualBlock(64, num_blocks=2).eval().cuda()
    assert rb(torch.rand(2, 64, 26, 26).cuda()).shape == (2, 64, 26, 26)
    sp = ScalePredictionBlock(512, num_classes=2).eval().cuda()
    assert sp(torch.rand(2, 512, 13, 13).cuda()).shape == (2, 3, 13, 13, 7)


def test_fused_stem_path_matches_patch_matrix_path():
    """The fused stem (default: TMA image windows, taps built in shared memory) must agree with the patch-matrix path
    (yolo_input_patchify + K=64 GEMM) bit for bit: same bf16 operands, same fp32 accumulation order inside one K=64
    block.  Sizes cover one / two / four row segments per image row (96, 416, 608) and the NaN-input flag."""
    for size, bsz in ((96, 2), (416, 1), (608, 1)):
        m, sd = _model(2, "leaky_relu", 4)
        x = torch.rand(bsz, 3, size, size, generator=torch.Generator().manual_seed(3)).cuda()
        a = m(x)
        eng = m._engine(x.device)
        assert eng.plans[(bsz, size, size)].stem_direct
        eng.stem_direct = False
        eng.plans.clear()
        b = m(x)
        assert not m._engine(x.device).plans[(bsz, size, size)].stem_direct
        for u, v in zip(a, b):
            assert torch.equal(u, v), size
    m, _ = _model(2, "leaky_relu", 4)
    x = torch.rand(1, 3, 96, 96)
    x[0, 2, 95, 95] = float("nan")
    with pytest.raises(AssertionError):   # model.py:175 through the fused stem's own check
        m(x.cuda())


def test_forward_608_matches_oracle():
    """BASELINE configs[2] geometry (19 / 38 / 76 grids: different tile tails and tail splits than 416)."""
    m, sd = _model(80, "leaky_relu", 5)
    x = torch.rand(1, 3, 608, 608, generator=torch.Generator().manual_seed(4))
    with torch.no_grad():
        refs = orc.forward(sd, x, 80)
        outs = m(x.cuda())
    assert [tuple(o.shape) for o in outs] == [(1, 3, 19, 19, 85), (1, 3, 38, 38, 85), (1, 3, 76, 76, 85)]
    _compare(outs, refs, "608")


def test_config2_608_low_conf_pipeline_and_map(oracle_c):
    """BASELINE configs[2] end to end: 608^2, conf 0.01 (22 743 candidates per image, nearly all pass), IoU 0.45 ->
    kept rows bit-exact against the C oracle on the device's decoded boxes; then calc_mAP(num_classes=80) against the
    oracle on synthetic ground truth (20 boxes per image, seed 7)."""
    from yolo_for_turbines_b200.utils import Detector, calc_mAP

    m, sd = _model(80, "leaky_relu", 6)
    B = 2
    x = torch.rand(B, 3, 608, 608, generator=torch.Generator().manual_seed(13)).cuda()
    det = Detector(m, orc.ANCHORS, 0.45, 0.01, "center")
    res, plan = det(x)
    plan.check_status()
    cand = res.boxes.view(B, -1, 6).cpu()
    assert cand.shape[1] == 22743
    kept = res.kept_rows()
    preds = []
    for b in range(B):
        exp = oracle_c(cand[b], 0.45, 0.01, "center")
        assert np.array_equal(kept[b].cpu().numpy(), cand[b][exp].numpy()), b
        preds += [[float(b)] + r for r in kept[b].tolist()]
    g = torch.Generator().manual_seed(7)
    gts = []
    for b in range(B):
        cxy = torch.rand(20, 2, generator=g)
        wh = 0.05 + 0.35 * torch.rand(20, 2, generator=g)
        cls = torch.randint(0, 80, (20,), generator=g).float()
        gts += [[float(b), float(cxy[i, 0]), float(cxy[i, 1]), float(wh[i, 0]), float(wh[i, 1]), 1.0, float(cls[i])] for i in range(20)]
    # the oracle is pure Python, O(D * G): evaluate the 4 000 best-scored detections of each image
    sub = []
    for b in range(B):
        rows = [p for p in preds if p[0] == b]
        sub += sorted(rows, key=lambda r: -r[5])[:4000]
    ref = float(orc.calc_mAP(sub, gts, 0.5, "center", 80))
    got = float(calc_mAP(sub, gts, 0.5, "center", 80))
    print(f"config 2: kept {len(preds)} of {B * 22743}, mAP {got:.6f} (oracle {ref:.6f})")
    assert abs(got - ref) <= 1e-6


def test_standalone_blocks_in_train_mode_use_batch_statistics():
    """The reference's own unit tests call the blocks in default (train) mode (model_tests.py:16-45): BatchNorm then
    normalises with batch statistics and updates the running ones (nn.BatchNorm2d, model.py:61)."""
    from yolo_for_turbines_b200.model import CNNBlock, ResidualBlock, ScalePredictionBlock

    torch.manual_seed(1)
    blk = CNNBlock(3, 32, kernel_size=1).cuda()
    assert blk.training
    x = torch.randn(5, 3, 64, 64)
    y = blk(x.cuda())
    assert y.shape == (5, 32, 64, 64)
    z = F.conv2d(x.bfloat16().float(), blk.conv.weight.detach().cpu().bfloat16().float())
    ref = F.leaky_relu(F.batch_norm(z, None, None, blk.batch_norm.weight.detach().cpu(), blk.batch_norm.bias.detach().cpu(),
                                    True, 0.1, 1e-5), 0.1)
    assert float((y.detach().cpu() - ref).abs().max()) <= 2.0 ** -6 * max(1.0, float(ref.abs().max()))
    assert int(blk.batch_norm.num_batches_tracked) == 1
    exp_mean = 0.1 * z.mean((0, 2, 3))
    exp_var = 0.9 + 0.1 * z.var((0, 2, 3), unbiased=True)
    assert torch.allclose(blk.batch_norm.running_mean.cpu(), exp_mean, atol=2e-3)
    assert torch.allclose(blk.batch_norm.running_var.cpu(), exp_var, rtol=1e-2, atol=1e-3)
    rb = ResidualBlock(128, num_blocks=2).cuda()
    assert rb(torch.randn(5, 128, 64, 64).cuda()).shape == (5, 128, 64, 64)
    sp = ScalePredictionBlock(512, num_classes=2).cuda()
    out = sp(torch.randn(5, 512, 13, 13).cuda())
    assert out.shape == (5, 3, 13, 13, 7) and bool(torch.isfinite(out).all())


@pytest.mark.parametrize("nc,size,bsz", [(80, 416, 2), (2, 96, 3), (80, 608, 1)])
def test_detector_fused_decode_equals_decode_of_stored_heads(nc, size, bsz):
    """Head convs with the anchor decode in their epilogue (the Detector's default) against yolo_decode on the stored
    fp32 heads: same arithmetic on the same accumulators, so candidates and survivors must be bit-identical."""
    from yolo_for_turbines_b200.utils import Detector

    m, sd = _model(nc, "leaky_relu", 8)
    anchors = orc.ANCHORS if nc == 80 else orc.TURBINE_ANCHORS
    x = torch.rand(bsz, 3, size, size, generator=torch.Generator().manual_seed(21)).cuda()
    fused = Detector(m, anchors, 0.45, 0.3, "center")
    plain = Detector(m, anchors, 0.45, 0.3, "center")
    plain.fuse_decode = False
    for _ in range(3):   # eager call, graph capture, graph replay
        ra, plan = fused(x)
        plan.check_status()
        assert plan.cand is not None and fused._state[next(iter(fused._state))]["fused"] is True
        ca, ka, oa = ra.boxes.clone(), ra.keep_idx.clone(), ra.keep_off.clone()
        rb, plan = plain(x)
        plan.check_status()
        assert torch.equal(ca.view(-1, 6), rb.boxes.view(-1, 6))
        assert torch.equal(oa, rb.keep_off)
        n = int(oa[-1])
        assert n > 0 and torch.equal(ka[:n], rb.keep_idx[:n])
