"""CPU-side checks of the drop-in surface: module structure / state_dict keys, the Darknet loader
(against stats of the reference's own loader), the C-ABI library (loads, exports every symbol the
header declares) and the loud failure without a GPU.  No kernel is launched here."""
import ctypes
import inspect
import os
import re
import tempfile

import pytest
import torch

from conftest import ROOT, template_state_dict
from oracle import synth


def test_state_dict_keys_and_shapes_equal_the_reference(gold):
    from yolo_for_turbines_b200.model import YOLOv3

    sd = YOLOv3(num_classes=80).state_dict()
    assert list(sd.keys()) == list(gold.keys.keys())  # 438 keys, same order
    assert all(list(sd[k].shape) == v for k, v in gold.keys.items())
    assert len(sd) == 438
    m2 = YOLOv3(num_classes=2, activation="mish")
    assert sum(p.numel() for p in m2.parameters()) == 61529119
    assert sum(p.numel() for p in YOLOv3().parameters()) == 61949149
    with pytest.raises(ValueError, match="Unsupported activation"):
        YOLOv3(activation="relu")


def test_module_surface_matches_reference_signatures():
    from yolo_for_turbines_b200 import model, utils

    def params(f):
        return [(p.name, p.default) for p in inspect.signature(f).parameters.values() if p.name != "self"]

    E = inspect.Parameter.empty
    assert params(model.CNNBlock.__init__)[:4] == [("in_channels", E), ("out_channels", E), ("batch_norm_act", True),
                                                   ("activation", "leaky_relu")]
    assert params(model.ResidualBlock.__init__) == [("in_channels", E), ("activation", "leaky_relu"),
                                                    ("use_residual", True), ("num_blocks", 1)]
    assert params(model.ScalePredictionBlock.__init__) == [("in_channels", E), ("num_classes", E),
                                                           ("activation", "leaky_relu"), ("anchors_per_scale", 3)]
    assert params(model.YOLOv3.__init__) == [("in_channels", 3), ("num_classes", 80), ("activation", "leaky_relu"),
                                             ("weights_path", None), ("freeze", False)]
    assert params(utils.calc_iou) == [("boxes1", E), ("boxes2", E), ("box_format", "center")]
    assert params(utils.cells_to_boxes) == [("predictions", E), ("anchors", E), ("grid_size", E), ("is_pred", True)]
    assert params(utils.non_max_suppression) == [("boxes", E), ("iou_threshold", E), ("obj_threshold", E),
                                                 ("box_format", "corners")]
    assert params(utils.calc_mAP) == [("pred_boxes", E), ("true_boxes", E), ("iou_threshold", 0.5),
                                      ("box_format", "center"), ("num_classes", 20)]
    assert utils.cells_to_bboxes is utils.cells_to_boxes and utils.intersection_over_union is utils.calc_iou
    assert utils.mean_average_precision is utils.calc_mAP
    m = model.YOLOv3(num_classes=3)
    for attr in ("layers", "param_idx", "layer_id", "weights_path", "freeze", "in_channels", "num_classes", "activation"):
        assert hasattr(m, attr)
    assert len(m.layers) == 30
    with pytest.raises(AttributeError):
        m.load_weights()  # no weights_path => no self.weights, like the reference


def test_darknet_loader_matches_reference_stats(gold, capsys):
    from yolo_for_turbines_b200.model import YOLOv3

    with tempfile.TemporaryDirectory() as td:
        full = os.path.join(td, "yolov3.weights")
        synth.synth_darknet_file(full, 62001757)
        cut = os.path.join(td, "darknet53.conv.74")
        os.symlink(full, cut)
        for tag, path in (("full", full), ("cutoff74", cut)):
            torch.manual_seed(1234)
            m = YOLOv3(num_classes=80, weights_path=path, freeze=(tag == "full"))
            before = {k: v.clone() for k, v in m.state_dict().items()}
            m.load_weights()
            g = gold.loader[tag]
            assert (m.param_idx, m.layer_id, int(m.weights.size), m.cutoff) == (g["param_idx"], g["layer_id"],
                                                                                g["n_floats"], g["cutoff"])
            sd = m.state_dict()
            for k, st in g["stats"].items():
                if st["changed"]:
                    assert float(sd[k].double().sum()) == st["sum"], (tag, k)
                    assert float(sd[k].flatten()[0]) == st["first"] and float(sd[k].flatten()[-1]) == st["last"]
                else:
                    assert torch.equal(sd[k], before[k]), (tag, k)
            if tag == "full":
                assert not any(p.requires_grad for p in m.parameters())  # freeze=True
    assert "loaded successfully" in capsys.readouterr().out


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "yolo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(yolo_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from yolo_for_turbines_b200 import _lib

    assert os.path.isfile(_lib.LIB_PATH), "build with `python -m yolo_for_turbines_b200.build`"
    dll = ctypes.CDLL(_lib.LIB_PATH)
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(dll, s), f"{s} declared in include/yolo_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == syms  # the ctypes table binds exactly the header's surface
    assert _lib.lib.yolo_version() == 100
    assert ctypes.sizeof(_lib.ConvDesc) == 44 * 4
    assert _lib.lib.yolo_nms_workspace_bytes(10647 * 64, 64) > 10647 * 64 * 40


def test_no_cpu_fallback():
    from yolo_for_turbines_b200._lib import YoloB200Error
    from yolo_for_turbines_b200.model import CNNBlock, YOLOv3
    from yolo_for_turbines_b200.utils import decode_boxes

    with pytest.raises(YoloB200Error, match="no CPU fallback"):
        YOLOv3(num_classes=2).eval()(torch.zeros(1, 3, 64, 64))
    with pytest.raises(YoloB200Error, match="no CPU fallback"):
        CNNBlock(3, 8, kernel_size=3, padding=1).eval()(torch.zeros(1, 3, 8, 8))
    with pytest.raises(YoloB200Error):
        decode_boxes(torch.zeros(1, 3, 2, 2, 7), [[1, 1]] * 3, 2)
    if not torch.cuda.is_available():
        from yolo_for_turbines_b200.utils import non_max_suppression
        with pytest.raises(YoloB200Error, match="no CPU fallback"):
            non_max_suppression([[0.5, 0.5, 0.1, 0.1, 0.9, 0]], 0.45, 0.5)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "yolo_for_turbines_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, fn)).read().replace("CPU oracle", ""), fn


def test_letterbox_geometry_host_rule_equals_oracle():
    """preprocess.letterbox_geometry (host side of yolo_letterbox_u8) against the oracle's restatement of
    albumentations' LongestMaxSize / PadIfNeeded integer rules, over many shapes."""
    import numpy as np

    from oracle import preprocess_oracle as po
    from yolo_for_turbines_b200.preprocess import letterbox_geometry

    rng = np.random.default_rng(0)
    for _ in range(2000):
        h, w = int(rng.integers(1, 3000)), int(rng.integers(1, 3000))
        for size in (320, 416, 608):
            assert letterbox_geometry(h, w, size) == po.letterbox_geometry(h, w, size), (h, w, size)


def test_gradient_buckets_cover_the_flat_buffer():
    from yolo_for_turbines_b200.train import make_buckets

    import random
    rnd = random.Random(1)
    for _ in range(50):
        sizes = [rnd.choice([4, 64, 512, 4096, 100000]) for _ in range(rnd.randint(1, 80))]
        offs, n = [], 0
        for s_ in sizes:
            offs.append(n)
            n += s_
        bk = make_buckets([(o, i) for i, o in enumerate(offs)], n, rnd.choice([1, 1000, 50000, 10 ** 9]))
        spans = sorted((lo, hi) for _, lo, hi in bk)
        assert spans[0][0] == 0 and spans[-1][1] == n and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        firsts = [f for f, _, _ in bk]
        assert firsts == sorted(firsts, reverse=True)            # fired in backward (descending op) order
        for first_op, lo, hi in bk:                               # every op whose parameters lie in the slice is >= first_op
            assert all(i >= first_op for i, o in enumerate(offs) if lo <= o < hi)


def test_new_entry_points_refuse_cpu_tensors():
    import torch

    from yolo_for_turbines_b200._lib import YoloB200Error
    from yolo_for_turbines_b200.dataset import encode_targets
    from yolo_for_turbines_b200.loss import YOLOLoss
    from yolo_for_turbines_b200.preprocess import LetterboxPlan, letterbox_batch

    with pytest.raises(YoloB200Error, match="no CPU fallback"):
        letterbox_batch([torch.zeros(4, 4, 3, dtype=torch.uint8)], 32, device="cpu")
    with pytest.raises(YoloB200Error, match="no CPU fallback"):
        LetterboxPlan(2, 8, 8, 32, device="cpu")
    with pytest.raises(YoloB200Error, match="no CPU fallback"):
        encode_targets([[[0.5, 0.5, 0.1, 0.1, 0]]], [[(0.1, 0.1)] * 3] * 3, image_size=64, device="cpu")
    with pytest.raises(YoloB200Error, match="no CPU fallback"):
        YOLOLoss()(torch.zeros(1, 3, 2, 2, 7), torch.zeros(1, 3, 2, 2, 6), [[1, 1]] * 3)


def test_bench_launch_count_matches_the_nms_pipeline():
    """`gpu_launches` in the bench line is a claim about csrc/nms.cu's launch list: one memset, the threshold
    compaction (3), 8-bit radix passes of hist + scan + scatter over (32 score + image) bits and over the (image, class)
    key, class keys, gather, the two NMS kernels, and keep count / scan / write / offsets."""
    import bench

    assert bench.nms_launch_count(1) == 1 + 3 + 4 * 3 + 1 + 1 * 3 + 1 + 2 + 4
    assert bench.nms_launch_count(64) == 1 + 3 + 5 * 3 + 1 + 2 * 3 + 1 + 2 + 4 == 33
    assert bench.nms_launch_count(300) == 1 + 3 + 6 * 3 + 1 + 3 * 3 + 1 + 2 + 4
