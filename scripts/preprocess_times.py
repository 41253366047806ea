"""Times the letterbox pre-processing kernel and the target encoder with CUDA events."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import preprocess_oracle as po  # noqa: E402  (CPU baseline beside the GPU number)
from oracle import target_oracle as to  # noqa: E402
from oracle import yolo_oracle as orc  # noqa: E402
from oracle.gen_golden_targets import synth_boxes  # noqa: E402
from yolo_for_turbines_b200.dataset import encode_targets  # noqa: E402
from yolo_for_turbines_b200.preprocess import letterbox_batch  # noqa: E402

dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
for (h, w, S, B) in ((1080, 1920, 416, 64), (480, 640, 416, 64), (1080, 1920, 608, 32)):
    imgs = [torch.from_numpy(rng.integers(0, 256, (h, w, 3), dtype=np.uint8)).to(dev) for _ in range(B)]
    out = torch.empty(B, 3, S, S, device=dev)
    for _ in range(3):
        letterbox_batch(imgs, S, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        letterbox_batch(imgs, S, out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    nh, nw, _, _ = po.letterbox_geometry(h, w, S)
    # algorithmic bytes: every source row the resize touches is read once (the whole image when down-scaling), fp32 out
    byts = B * (h * w * 3 + 3 * S * S * 4)
    t0 = time.time()
    po.letterbox(imgs[0].cpu().numpy(), S)
    cpu_ms = (time.time() - t0) * 1e3
    print(f"letterbox {B} x {h}x{w} -> {S}: {ms:.3f} ms/batch = {B / ms * 1e3:.0f} img/s, {byts / ms / 1e6:.0f} GB/s algorithmic "
          f"(incl. {B} host-side descriptor rows + table upload per call); numpy oracle {cpu_ms:.1f} ms/image")
for B, n in ((32, 40), (64, 10), (32, 200)):
    batch = [synth_boxes(n, 2, rng, orc.TURBINE_ANCHORS) for _ in range(B)]
    for _ in range(3):
        encode_targets(batch, orc.TURBINE_ANCHORS, image_size=416)
    torch.cuda.synchronize()
    t0 = time.time()
    for _ in range(10):
        encode_targets(batch, orc.TURBINE_ANCHORS, image_size=416)
    torch.cuda.synchronize()
    ms = (time.time() - t0) * 100
    t0 = time.time()
    to.encode_batch(batch[:4], orc.TURBINE_ANCHORS, [13, 26, 52])
    cpu_ms = (time.time() - t0) * 1e3 / 4 * B
    print(f"encode_targets B={B} x {n} boxes @416: {ms:.3f} ms/batch wall incl. host packing + upload; oracle (reference algorithm, CPU) {cpu_ms:.1f} ms/batch")
