"""One steady-state training step bracketed by cudaProfilerStart/Stop (ncu --profile-from-start off)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import yolo_oracle as orc  # noqa: E402  (synthetic targets only)
from yolo_for_turbines_b200.model import YOLOv3  # noqa: E402
from yolo_for_turbines_b200.train import Trainer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
S = int(sys.argv[2]) if len(sys.argv) > 2 else 416
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = YOLOv3(num_classes=2, activation="mish").to(dev).train()
tr = Trainer(m, orc.TURBINE_ANCHORS, lr=1e-4, momentum=0.9, weight_decay=5e-4)
x = torch.rand(B, 3, S, S, device=dev)
tg = [t.to(dev) for t in orc.synth_targets(B, S, 2, 1)]
for _ in range(3):
    tr.step(x, tg)
torch.cuda.synchronize()
torch.cuda.profiler.start()
losses = tr.step(x, tg)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("losses", losses.tolist())
