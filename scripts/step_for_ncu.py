"""One steady-state detection step bracketed by cudaProfilerStart/Stop, for
    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum ...
(warm-up steps, weight packing and graph capture stay outside the profiled range)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_for_turbines_b200 import config as cfg  # noqa: E402
from yolo_for_turbines_b200.model import YOLOv3  # noqa: E402
from yolo_for_turbines_b200.utils import Detector  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
S = int(sys.argv[2]) if len(sys.argv) > 2 else 416
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = YOLOv3(num_classes=80).eval().to(dev)
det = Detector(m, cfg.ANCHORS, 0.45, float(sys.argv[3]) if len(sys.argv) > 3 else 0.5, "center")
xs = [torch.rand(B, 3, S, S, device=dev) for _ in range(3)]
for i in range(4):
    res, plan = det(xs[i % 3])
torch.cuda.synchronize()
torch.cuda.profiler.start()
res, plan = det(xs[1])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
plan.check_status()
print("kept", int(res.keep_off[-1]))
