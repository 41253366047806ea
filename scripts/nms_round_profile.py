"""Dev tool: per-phase clock sums of k_nms_segments' register path.  Needs a library built with
    make -C yolo_for_turbines_b200/csrc EXTRA_NVCCFLAGS=-DYB_NMS_PROFILE   (touch nms.cu first)
python scripts/nms_round_profile.py [B] [S] [conf]"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_for_turbines_b200 import _lib, config as cfg  # noqa: E402
from yolo_for_turbines_b200.model import YOLOv3  # noqa: E402
from yolo_for_turbines_b200.utils import Detector  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = int(sys.argv[2]) if len(sys.argv) > 2 else 416
conf = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = YOLOv3(num_classes=80).eval().to(dev)
det = Detector(m, cfg.ANCHORS, 0.45, conf, "center")
det.use_graph = False
x = torch.rand(B, 3, S, S, device=dev, generator=torch.Generator(device=dev).manual_seed(1234))
for _ in range(2):
    res, plan = det(x)
torch.cuda.synchronize()
fn = _lib.lib.raw("yolo_debug_nms_prof")
out = (C.c_ulonglong * 8)()
fn(out, 1)
res, plan = det(x)
torch.cuda.synchronize()
fn(out, 0)
v = list(out)
names = ["(0) window publish + sync", "(0b,0c) select + member load + sync", "(a) pair matrix + bins + sync", "(b) serial resolve + keep store + sync",
         "(c) apply to owned boxes"]
rounds, segs = v[5], v[6]
print(f"B={B} S={S} conf={conf}: {segs} segments >= 1024 boxes, {rounds} rounds in total ({rounds / max(segs, 1):.1f} per segment)")
tot = sum(v[:5])
for n, c in zip(names, v[:5]):
    print(f"  {n:45s} {c / max(rounds, 1):9.0f} clk per round  ({100.0 * c / max(tot, 1):4.1f} %)")
print(f"  total {tot / max(rounds, 1):.0f} clk per round")
