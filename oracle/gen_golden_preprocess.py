"""Generates tests/golden/preprocess.npz with OpenCV itself (the third-party package the reference's pre-processing
calls through albumentations, config.py:101-113): cv2.resize(INTER_LINEAR) to albumentations' LongestMaxSize size,
centred zero padding, /255, HWC -> CHW.  Inputs: crops of the reference's own example photographs + random images.

    python -m oracle.gen_golden_preprocess      # needs cv2 and /root/reference/examples -- not the GPU box
"""
import glob
import os
import sys

import cv2
import numpy as np
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import preprocess_oracle as po  # noqa: E402  (only for the two integer rules: size and padding)


def cv2_pipeline(img, size):
    nh, nw, top, left = po.letterbox_geometry(img.shape[0], img.shape[1], size)
    small = img if (nh, nw) == img.shape[:2] else cv2.resize(img, (nw, nh), interpolation=cv2.INTER_LINEAR)
    bottom, right = size - nh - top, size - nw - left
    # stored as the uint8 canvas (small fixture); Normalize + ToTensorV2 = float32(canvas) * float32(1/255), HWC -> CHW
    return cv2.copyMakeBorder(small, top, bottom, left, right, cv2.BORDER_CONSTANT, value=0)


def main():
    out, rng, n = {}, np.random.default_rng(7), 0
    for f in sorted(glob.glob("/root/reference/examples/*.jpg"))[:3]:
        im = np.array(Image.open(f).convert("RGB"))
        y, x = im.shape[0] // 3, im.shape[1] // 4
        for (h, w, size) in ((150, 200, 96), (70, 45, 96), (96, 60, 64)):
            crop = np.ascontiguousarray(im[y:y + h, x:x + w])
            out[f"c{n}/img"], out[f"c{n}/size"], out[f"c{n}/canvas"] = crop, np.int32(size), cv2_pipeline(crop, size)
            n += 1
    for (h, w, size) in ((33, 97, 64), (128, 128, 64), (64, 64, 64), (20, 31, 96), (201, 77, 128), (5, 300, 32)):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        out[f"c{n}/img"], out[f"c{n}/size"], out[f"c{n}/canvas"] = img, np.int32(size), cv2_pipeline(img, size)
        n += 1
    out["n"] = np.int32(n)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "preprocess.npz"), **out)
    print("cases", n)


if __name__ == "__main__":
    main()
