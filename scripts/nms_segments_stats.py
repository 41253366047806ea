"""Segment statistics of the NMS stage on the detector's own candidates: sizes of the (image, class) groups that pass the
threshold, survivors per group, and the stage time (graph replay).  python scripts/nms_segments_stats.py [B] [S] [conf]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_for_turbines_b200 import config as cfg  # noqa: E402
from yolo_for_turbines_b200.model import YOLOv3  # noqa: E402
from yolo_for_turbines_b200.utils import Detector, batched_nms  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = int(sys.argv[2]) if len(sys.argv) > 2 else 416
conf = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = YOLOv3(num_classes=80).eval().to(dev)
det = Detector(m, cfg.ANCHORS, 0.45, conf, "center")
x = torch.rand(B, 3, S, S, device=dev, generator=torch.Generator(device=dev).manual_seed(1234))
for _ in range(3):
    res, plan = det(x)
torch.cuda.synchronize()
cand = res.boxes.view(B, -1, 6)
n = cand.shape[1]
passing = cand[..., 4].double() > conf
img = torch.arange(B, device=dev).view(B, 1).expand(B, n)
grp = (img * 80 + cand[..., 5].long())[passing]
cnt = torch.bincount(grp, minlength=B * 80)
keep = res.keep_idx[: int(res.keep_off[-1])].long()
kflat = cand.view(-1, 6)[keep]
kgrp = (keep // n) * 80 + kflat[:, 5].long()
kcnt = torch.bincount(kgrp, minlength=B * 80)
top = torch.argsort(cnt, descending=True)[:12]
print(f"B={B} S={S} conf={conf}: {int(passing.sum())} pass of {B * n}, kept {keep.numel()}, non-empty groups {(cnt > 0).sum().item()}")
print("largest groups (size, survivors):", [(int(cnt[i]), int(kcnt[i])) for i in top])
per_img = cnt.view(B, 80)
print("per image: max group", per_img.max(1).values[:8].tolist(), " classes with >0:", (per_img > 0).sum(1)[:8].tolist())
hist_edges = [1, 32, 128, 512, 2048, 8192, 1 << 30]
sizes = cnt[cnt > 0]
print("group size histogram:", {f"<{e}": int(((sizes < e) & (sizes >= (hist_edges[i - 1] if i else 0))).sum()) for i, e in enumerate(hist_edges)})
st = det._get_state(B, [h.H for h in plan.heads], dev)
fn = lambda: batched_nms(cand.reshape(-1, 6), st["off"], 0.45, conf, "center", workspace=st["ws"], class_bits=8)
fn(); torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    fn()
g.replay(); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    g.replay()
b.record(); torch.cuda.synchronize()
print(f"nms stage (graph replay): {a.elapsed_time(b) / 10:.3f} ms")
