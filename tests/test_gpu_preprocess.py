"""GPU parity of the letterbox pre-processing kernel (yolo_letterbox_u8 through preprocess.letterbox_batch) against the
oracle restatement of the reference's albumentations/OpenCV pipeline (config.py:101-113) and against the OpenCV-made
golden vectors: BIT-EXACT (integer fixed-point resize, one fp32 multiply)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "preprocess.npz")


def _expected(canvas_u8):
    """A.Normalize(mean 0, std 1, max 255) + ToTensorV2 on OpenCV's letterboxed uint8 canvas."""
    return np.ascontiguousarray((canvas_u8.astype(np.float32) * np.float32(1.0 / 255.0)).transpose(2, 0, 1))


def test_letterbox_matches_opencv_golden():
    from yolo_for_turbines_b200.preprocess import letterbox_batch

    z = np.load(GOLD)
    for i in range(int(z["n"])):
        got = letterbox_batch([z[f"c{i}/img"]], int(z[f"c{i}/size"]))
        torch.cuda.synchronize()
        assert np.array_equal(got[0].cpu().numpy(), _expected(z[f"c{i}/canvas"])), i


@pytest.mark.parametrize("size", [416, 608, 320])
def test_letterbox_mixed_batch_matches_oracle(size):
    from oracle import preprocess_oracle as po
    from yolo_for_turbines_b200.preprocess import letterbox_batch, unletterbox_boxes

    rng = np.random.default_rng(size)
    shapes = [(480, 640), (640, 480), (size, size), (1080, 1920), (37, 23), (size, size // 2), (1, 9), (2000, 1500)]
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in shapes]
    got = letterbox_batch(imgs, size)
    torch.cuda.synchronize()
    assert got.shape == (len(imgs), 3, size, size) and got.dtype == torch.float32
    for i, im in enumerate(imgs):
        assert np.array_equal(got[i].cpu().numpy(), po.letterbox(im, size)), shapes[i]
    rows = [[0.5, 0.4, 0.2, 0.1, 0.9, 1.0], [0.1, 0.9, 0.05, 0.3, 0.6, 0.0]]
    ref = po.unletterbox_boxes(rows, 480, 640, size)
    dev = unletterbox_boxes(torch.tensor(rows, device="cuda"), 480, 640, size).cpu()
    assert torch.allclose(dev, torch.tensor(ref), rtol=1e-6, atol=1e-7)
    assert unletterbox_boxes(rows, 480, 640, size) == ref


def test_letterbox_feeds_the_detector():
    """demo.predict's chain (demo.py:30-55) on device: uint8 photo -> letterbox -> model -> decode -> NMS."""
    from oracle import yolo_oracle as orc
    from yolo_for_turbines_b200.model import YOLOv3
    from yolo_for_turbines_b200.preprocess import letterbox_batch
    from yolo_for_turbines_b200.utils import Detector

    torch.manual_seed(0)
    m = YOLOv3(num_classes=2).eval().cuda()
    rng = np.random.default_rng(5)
    x = letterbox_batch([rng.integers(0, 256, (300, 200, 3), dtype=np.uint8), rng.integers(0, 256, (90, 160, 3), dtype=np.uint8)], 96)
    det = Detector(m, orc.TURBINE_ANCHORS, 0.45, 0.5, "center")
    res, plan = det(x)
    plan.check_status()
    assert len(res.to_lists()) == 2


@pytest.mark.gpu
@pytest.mark.parametrize("h,w,size", [(120, 160, 96), (96, 96, 96), (50, 33, 64)])
def test_letterbox_plan_equals_letterbox_batch_and_oracle(h, w, size):
    """The static serving-loop variant (one uint8 H2D copy + one launch per batch) is the same arithmetic."""
    from oracle import preprocess_oracle as po
    from yolo_for_turbines_b200.preprocess import LetterboxPlan, letterbox_batch

    rng = np.random.default_rng(11)
    frames = rng.integers(0, 256, (3, h, w, 3), dtype=np.uint8)
    lp = LetterboxPlan(3, h, w, size)
    got = lp.run(torch.from_numpy(frames).pin_memory()).cpu().numpy()
    ref = letterbox_batch([f for f in frames], size).cpu().numpy()
    assert np.array_equal(got, ref)
    for i in range(3):
        assert np.array_equal(got[i], po.letterbox(frames[i], size))
    with pytest.raises(Exception):
        lp.run(torch.zeros(3, h, w, 3))   # fp32 frames are refused, not converted
