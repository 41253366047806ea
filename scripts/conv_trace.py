"""Per-CTA timeline of conv launches (yolo_conv_fwd_trace): where a launch spends its time -- prologue, wait for the
previous launch, pipeline fill, MMA main loop, epilogue drain.  Dev tool, GPU box only.

    python scripts/conv_trace.py [--batch 64] [--size 416] [--layers layers.8.layers.0.0,layers.10.layers.0.1]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_for_turbines_b200 import config as cfg  # noqa: E402
from yolo_for_turbines_b200._lib import lib, ptr, stream_ptr  # noqa: E402
from yolo_for_turbines_b200.model import YOLOv3  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--size", type=int, default=416)
ap.add_argument("--layers", default="")
ap.add_argument("--box", type=int, default=0, help="epilogue box of the first tile that gets the fine stamps")
args = ap.parse_args()
dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = YOLOv3(num_classes=80).eval().to(dev)
x = torch.rand(args.batch, 3, args.size, args.size, device=dev)
plan, _ = m.forward_async(x)
plan.check_status()
want = set(args.layers.split(",")) if args.layers else None
st, sp = stream_ptr(dev), ptr(plan.status)
print(f"{'layer':30s} {'ctas':>4s} {'tiles':>5s} | per-CTA medians in us: prologue, dep wait, first load->landed, "
      f"MMA loop, last MMA->drained, exit | launch span (first entry -> last exit)")
ops = plan.ops[1:] if plan.stem_direct else plan.ops
for i, op in enumerate(ops):
    if want is not None and op.name not in want:
        continue
    tr = torch.zeros(148 * 32, dtype=torch.int64, device=dev)
    prev = ops[i - 1] if i > 0 else None
    for _ in range(3):
        tr.zero_()
        if prev is not None:
            lib.yolo_conv_fwd(prev.plan_ptr, sp, st)      # the producer launch: PDL overlap + L2 state as in the graph
        lib.yolo_conv_fwd_trace(op.plan_ptr, sp, ptr(tr), args.box, st)
        torch.cuda.synchronize()
    t = tr.view(148, 32).cpu()
    t = t[t[:, 0] > 0]
    lead = t[t[:, 5] > 0]
    med = lambda v: float(v.double().median()) / 1e3  # noqa: E731
    span = (int(t[:, 8].max()) - int(t[:, 0].min())) / 1e3
    e = lead[lead[:, 14] > 0]
    epi = (f" | epi box: tmem {med(e[:, 10] - e[:, 6]):5.2f} res {med(e[:, 11] - e[:, 10]):5.2f} [bn_act {med(e[:, 19] - e[:, 11]):5.2f} +res {med(e[:, 20] - e[:, 19]):5.2f} +sts {med(e[:, 21] - e[:, 20]):5.2f} half1 {med(e[:, 22] - e[:, 21]):5.2f} fence {med(e[:, 12] - e[:, 22]):5.2f}] "
           f"store {med(e[:, 13] - e[:, 12]):5.2f} tile {med(e[:, 14] - e[:, 6]):5.2f}") if e.shape[0] else ""
    print(f"{op.name:30s} {t.shape[0]:4d} {int(t[:, 9].max()):5d} | {med(t[:, 1] - t[:, 0]):6.2f} {med(t[:, 2] - t[:, 1]):6.2f} "
          f"{med(lead[:, 4] - lead[:, 3]):6.2f} {med(lead[:, 5] - lead[:, 4]):7.2f} {med(lead[:, 7] - lead[:, 5]):6.2f} "
          f"{med(t[:, 8] - t[:, 7]):6.2f} | {span:7.2f}  entry spread {(int(t[:, 0].max()) - int(t[:, 0].min())) / 1e3:6.2f} "
          f"exit spread {(int(t[:, 8].max()) - int(t[:, 8].min())) / 1e3:6.2f}{epi}"
          f" | waits: MMA on operands {med(lead[:, 16]):6.2f} on accumulator {med(lead[:, 17]):6.2f} producer on ring {med(t[:, 18]):6.2f}")
