"""Times the training step (SURVEY config #4: nc=2 turbine model, TURBINE_ANCHORS, batch 32/GPU) with CUDA events.
    python scripts/train_times.py [batch] [size] [activation] [steps]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import yolo_oracle as orc  # noqa: E402  (synthetic targets only)
from yolo_for_turbines_b200.model import YOLOv3  # noqa: E402
from yolo_for_turbines_b200.train import Trainer  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
S = int(sys.argv[2]) if len(sys.argv) > 2 else 416
act = sys.argv[3] if len(sys.argv) > 3 else "leaky_relu"
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 10
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = YOLOv3(num_classes=2, activation=act).to(dev).train()
tr = Trainer(m, orc.TURBINE_ANCHORS, lr=1e-4, momentum=0.9, weight_decay=5e-4)
x = torch.rand(B, 3, S, S, device=dev)
tg = [t.to(dev) for t in orc.synth_targets(B, S, 2, 1)]
t0 = time.time()
for _ in range(3):
    losses = tr.step(x, tg)
torch.cuda.synchronize()
print(f"warm-up {time.time() - t0:.2f} s, losses {losses.tolist()}")
plan = tr.plan(B, S, S)
print(f"plan arena {plan.total_bytes / 1e9:.2f} GB, launches/step {tr.launches_per_step(plan)}")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]


def timed(fn, n=steps):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


t_step = timed(lambda: tr.step(x, tg))
for _ in range(2):
    tr.step(x, tg, graph=True)
t_graph = timed(lambda: tr.step(x, tg, graph=True))
t_fwd = timed(lambda: plan.forward(x))
t_loss = timed(lambda: tr._loss_and_head_grads(plan, tg))
t_bwd = timed(lambda: plan.backward())
t_rep = timed(lambda: tr.repack())
gf = {416: 65.297, 320: 38.637}.get(S, 65.297 * (S / 416.0) ** 2) * B * 3
print(f"B={B} S={S} {act}: step {t_step:.2f} ms -> {B / t_step * 1e3:.1f} img/s, {gf / t_step:.1f} TFLOP/s (3x fwd FLOPs); graph replay {t_graph:.2f} ms")
print(f"  forward {t_fwd:.2f} ms | loss fwd+bwd {t_loss:.3f} ms | backward {t_bwd:.2f} ms | repack {t_rep:.2f} ms")
print(f"  losses {tr.losses.tolist()}")
