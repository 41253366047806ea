"""Layer-by-layer train-mode forward comparison vs the fp32 oracle."""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth, yolo_oracle as orc
from yolo_for_turbines_b200.model import YOLOv3
from yolo_for_turbines_b200.train import Trainer

nc, act, size, bsz, seed = 2, "leaky_relu", 64, 2, 5
if len(sys.argv) > 1:
    act, size, bsz, seed = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
m = YOLOv3(num_classes=nc, activation=act)
sd = synth.synth_state_dict(m.state_dict(), seed=seed)
m.load_state_dict(sd)
x = torch.rand(bsz, 3, size, size, generator=torch.Generator().manual_seed(40 + seed))
cap = []
orig = orc._cnn_block
def hook(sd_, prefix, xx, k, stride, activation, bn_act=True):
    pad = 1 if k == 3 else 0
    z = F.conv2d(xx, sd_[prefix + "conv.weight"], sd_.get(prefix + "conv.bias") if not bn_act else None, stride, pad)
    y = orig(sd_, prefix, xx, k, stride, activation, bn_act)
    cap.append((prefix, z.detach(), y.detach()))
    return y
orc._cnn_block = hook
orc.forward({k: v.clone() for k, v in sd.items()}, x, nc, act, training=True)
orc._cnn_block = orig
m = m.cuda().train()
tr = Trainer(m, orc.TURBINE_ANCHORS, lr=0.0)
plan = tr.plan(bsz, size, size)
plan.forward(x.cuda())
torch.cuda.synchronize()
cs = lambda a, b: float(F.cosine_similarity(a.flatten(), b.flatten(), dim=0))
for op, (prefix, zr, yr) in zip(plan.ops, cap):
    if op.head:
        continue
    B, C, H, W = zr.shape
    z = op.z.view(B, H, W, -1)[..., :C].float().cpu().permute(0, 3, 1, 2)
    zc = zr - zr.mean(dim=(0, 2, 3), keepdim=True)
    zz = z - z.mean(dim=(0, 2, 3), keepdim=True)
    ratio = float((zr.mean(dim=(0, 2, 3)).abs() / (zr.std(dim=(0, 2, 3)) + 1e-12)).max())
    print(f"{op.name:28s} z cos {cs(z, zr):.6f} centred cos {cs(zz, zc):.6f} max|mean|/std {ratio:8.2f} n={B*H*W}")
