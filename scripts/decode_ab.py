"""A/B of the Detector step with the anchor decode fused into the head convs vs yolo_decode on stored heads."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_for_turbines_b200 import config as cfg
from yolo_for_turbines_b200.model import YOLOv3
from yolo_for_turbines_b200.utils import Detector
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = YOLOv3(num_classes=80).eval().to(dev)
xs = [torch.rand(64, 3, 416, 416, device=dev) for _ in range(3)]
for fuse in (True, False, True, False):
    det = Detector(m, cfg.ANCHORS, 0.45, 0.5, "center")
    det.fuse_decode = fuse
    for i in range(6):
        det(xs[i % 3])
    torch.cuda.synchronize()
    ts = []
    for rep in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(20):
            det(xs[i % 3])
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / 20)
    print(f"fuse_decode={fuse}: ms/step {sorted(ts)[2]:.3f} (min {min(ts):.3f})")
