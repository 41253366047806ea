"""Dev check: per-step loss terms of eager vs eager (run-to-run noise) vs graphed training steps from the same init."""
import copy
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import yolo_oracle as orc  # noqa: E402
from yolo_for_turbines_b200.model import YOLOv3  # noqa: E402
from yolo_for_turbines_b200.train import Trainer  # noqa: E402

torch.manual_seed(3)
base = YOLOv3(num_classes=2, activation=sys.argv[1] if len(sys.argv) > 1 else "mish")
B, S = 4, 96
lr = float(sys.argv[2]) if len(sys.argv) > 2 else 1e-3
xs = [torch.rand(B, 3, S, S, generator=torch.Generator().manual_seed(10 + i)).cuda() for i in range(3)]
tgs = [[t.cuda() for t in orc.synth_targets(B, S, 2, 20 + i)] for i in range(3)]
out = {}
for mode in ("eager", "eager2", "graph"):
    m = copy.deepcopy(base).cuda().train()
    tr = Trainer(m, orc.TURBINE_ANCHORS, lr=lr, momentum=0.9, weight_decay=5e-4)
    ls = []
    for i in range(6):
        ls.append(tr.step(xs[i % 3], tgs[i % 3], graph=(mode == "graph")).clone())
    torch.cuda.synchronize()
    out[mode] = (torch.stack(ls).cpu(), tr.flat_p[: tr.n_trainable].clone().cpu())
for i in range(6):
    print(i, "eager", [round(v, 4) for v in out["eager"][0][i].tolist()], "eager2", [round(v, 4) for v in out["eager2"][0][i].tolist()],
          "graph", [round(v, 4) for v in out["graph"][0][i].tolist()])
pe, p2, pg = out["eager"][1], out["eager2"][1], out["graph"][1]
print("max |p_eager - p_eager2|", float((pe - p2).abs().max()), " max |p_eager - p_graph|", float((pe - pg).abs().max()))
