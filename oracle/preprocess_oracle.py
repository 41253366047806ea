"""CPU oracle of the reference's inference pre-processing -- TEST INFRASTRUCTURE ONLY.

Reference call sites: code/config.py:101-113 `set_only_image_transforms` (= `test_transforms` :88-99 without boxes),
used by code/demo.py:37-39 and the loaders; code/utils.py:475-501 `plot_original` (the inverse box mapping).

    A.LongestMaxSize(max_size=S)  ->  A.PadIfNeeded(S, S, border_mode=cv2.BORDER_CONSTANT, value=0)
    ->  A.Normalize(mean=0, std=1, max_pixel_value=255)  ->  ToTensorV2()

The arithmetic lives in third-party packages that are NOT part of the reference checkout and are unpinned in its
requirements.txt (`albumentations`, `opencv-python`): restated here from their published behaviour --
  * LongestMaxSize: scale = S / max(h, w); new (h, w) = round-half-even(dim * scale); cv2.resize(..., INTER_LINEAR);
  * cv2.resize INTER_LINEAR on uint8 (OpenCV resize.cpp, generic path): source coordinate
    fx = (float)((dx + 0.5) * scale - 0.5) with scale = 1 / (dst / src) in double, borders clamped, 11-bit fixed-point
    weights saturate_cast<short>(w * 2048), horizontal pass in int32, vertical pass
    ((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2;
  * PadIfNeeded (centre): top = int((S - h) / 2), bottom = S - h - top, same for left / right, constant 0;
  * Normalize: float32(u8) * float32(1 / 255);  ToTensorV2: HWC -> CHW.
Pinned against OpenCV itself (cv2 4.13 in this image) by tests/test_oracle_preprocess.py: bit-exact on the reference's
example photographs and on random images; albumentations is not installed here, so its two integer rules above are
pinned only by the published source and by `plot_original`'s own padding arithmetic (utils.py:483-488).
"""
from __future__ import annotations

import numpy as np

COEF_BITS = 11
COEF_SCALE = 1 << COEF_BITS


def longest_max_size_shape(h: int, w: int, size: int):
    scale = size / float(max(h, w))
    if scale == 1.0:
        return h, w
    return int(round(h * scale)), int(round(w * scale))   # Python round = half to even (albumentations py3round)


def _axis_coeffs(src: int, dst: int, vertical: bool = False):
    """Per destination index: the two source indices and the two 11-bit weights (OpenCV resize.cpp, linear, ksize 2).
    Horizontally a coordinate beyond the border snaps to the border pixel with weight 1; vertically OpenCV only clips
    the ROW INDICES and keeps the fractional weights (both rows are then the border row, but the two truncated
    products differ from one by up to 1 -- visible when up-scaling)."""
    inv_scale = float(dst) / float(src)
    scale = 1.0 / inv_scale
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if not vertical:
        lo = s < 0
        f[lo], s[lo] = 0.0, 0
        hi = s >= src - 1
        f[hi], s[hi] = 0.0, src - 1
    w1 = np.rint(f.astype(np.float32) * np.float32(COEF_SCALE)).astype(np.int32)       # saturate_cast<short>: round to nearest even
    w0 = np.rint((np.float32(1.0) - f) * np.float32(COEF_SCALE)).astype(np.int32)
    return np.clip(s, 0, src - 1), np.clip(s + 1, 0, src - 1), w0, w1


def resize_linear_u8(img: np.ndarray, new_h: int, new_w: int) -> np.ndarray:
    """cv2.resize(img, (new_w, new_h), interpolation=cv2.INTER_LINEAR) for uint8 HWC images, bit for bit."""
    h, w = img.shape[:2]
    if (h, w) == (new_h, new_w):
        return img.copy()
    x0, x1, a0, a1 = _axis_coeffs(w, new_w)
    y0, y1, b0, b1 = _axis_coeffs(h, new_h, vertical=True)
    src = img.astype(np.int32)
    rows = src[:, x0] * a0[None, :, None] + src[:, x1] * a1[None, :, None]             # horizontal pass, scale 2^11
    s0, s1 = rows[y0], rows[y1]
    out = (((b0[:, None, None] * (s0 >> 4)) >> 16) + ((b1[:, None, None] * (s1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8)


def letterbox_geometry(h: int, w: int, size: int):
    nh, nw = longest_max_size_shape(h, w, size)
    top = int((size - nh) / 2.0) if nh < size else 0
    left = int((size - nw) / 2.0) if nw < size else 0
    return nh, nw, top, left


def letterbox(img: np.ndarray, size: int) -> np.ndarray:
    """uint8 HWC image -> float32 (3, size, size), the tensor demo.py:38-39 feeds the model."""
    h, w = img.shape[:2]
    nh, nw, top, left = letterbox_geometry(h, w, size)
    small = resize_linear_u8(img, nh, nw)
    canvas = np.zeros((size, size, img.shape[2]), dtype=np.uint8)
    canvas[top:top + nh, left:left + nw] = small
    out = canvas.astype(np.float32) * np.float32(1.0 / 255.0)
    return np.ascontiguousarray(out.transpose(2, 0, 1))


def unletterbox_boxes(boxes, orig_h: int, orig_w: int, size: int):
    """plot_original's box mapping (utils.py:475-501): rows [cx, cy, w, h, score, cls] normalised to the letterboxed
    square -> normalised to the original image.  Python floats, like the reference."""
    scale = min(size / orig_w, size / orig_h)
    new_w, new_h = int(orig_w * scale), int(orig_h * scale)
    pad_w, pad_h = (size - new_w) // 2, (size - new_h) // 2
    return [[(b[0] * size - pad_w) / new_w, (b[1] * size - pad_h) / new_h, (b[2] * size) / new_w, (b[3] * size) / new_h, b[4], b[5]]
            for b in boxes]
