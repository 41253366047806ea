"""Pins oracle/preprocess_oracle.py (restatement of albumentations LongestMaxSize/PadIfNeeded/Normalize/ToTensorV2 over
cv2.resize INTER_LINEAR, config.py:101-113) against vectors produced by OpenCV itself (tests/golden/preprocess.npz,
oracle/gen_golden_preprocess.py), and -- when cv2 is importable -- against cv2 live.  CPU only."""
import os

import numpy as np
import pytest

from oracle import preprocess_oracle as po

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "preprocess.npz")


def _expected(canvas_u8):
    """A.Normalize(mean 0, std 1, max 255) + ToTensorV2 on OpenCV's letterboxed uint8 canvas."""
    return np.ascontiguousarray((canvas_u8.astype(np.float32) * np.float32(1.0 / 255.0)).transpose(2, 0, 1))


def test_letterbox_oracle_matches_opencv_golden():
    z = np.load(GOLD)
    for i in range(int(z["n"])):
        got = po.letterbox(z[f"c{i}/img"], int(z[f"c{i}/size"]))
        assert got.dtype == np.float32 and np.array_equal(got, _expected(z[f"c{i}/canvas"])), i


def test_resize_matches_cv2_live():
    cv2 = pytest.importorskip("cv2")
    rng = np.random.default_rng(3)
    for _ in range(20):
        h, w = int(rng.integers(8, 500)), int(rng.integers(8, 500))
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        for size in (64, 416):
            nh, nw = po.longest_max_size_shape(h, w, size)
            ref = cv2.resize(img, (nw, nh), interpolation=cv2.INTER_LINEAR)
            assert np.array_equal(po.resize_linear_u8(img, nh, nw), ref), (h, w, size)


def test_geometry_rules():
    assert po.letterbox_geometry(375, 500, 416) == (312, 416, 52, 0)
    assert po.letterbox_geometry(500, 375, 416) == (416, 312, 0, 52)
    assert po.letterbox_geometry(100, 250, 416) == (166, 416, 125, 0)       # 166.4 -> 166
    assert po.letterbox_geometry(101, 202, 101) == (50, 101, 25, 0)         # 50.5 -> 50 (half to even), odd padding: 25 | 26
    assert po.letterbox_geometry(416, 416, 416) == (416, 416, 0, 0)


def test_unletterbox_matches_plot_original_arithmetic():
    # utils.py:475-501 on a 375x500 image letterboxed to 416: scale .832, new 416x312, pad (0, 52)
    b = po.unletterbox_boxes([[0.5, 0.5, 0.25, 0.25, 0.9, 3.0]], 375, 500, 416)[0]
    assert b[0] == pytest.approx(0.5) and b[1] == pytest.approx((0.5 * 416 - 52) / 312)
    assert b[2] == pytest.approx(0.25) and b[3] == pytest.approx(0.25 * 416 / 312) and b[4:] == [0.9, 3.0]
