"""Diagnostics: train-mode forward heads and parameter gradients vs the fp32 oracle."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth, yolo_oracle as orc
from yolo_for_turbines_b200.model import YOLOv3
from yolo_for_turbines_b200.train import Trainer

for nc, act, size, bsz, seed in [(2, "leaky_relu", 64, 2, 5), (2, "leaky_relu", 128, 4, 8), (2, "mish", 96, 2, 6), (2, "leaky_relu", 128, 8, 11), (80, "mish", 128, 8, 12)]:
    m = YOLOv3(num_classes=nc, activation=act)
    sd = synth.synth_state_dict(m.state_dict(), seed=seed)
    m.load_state_dict(sd)
    x = torch.rand(bsz, 3, size, size, generator=torch.Generator().manual_seed(40 + seed))
    tg = orc.synth_targets(bsz, size, nc, 50 + seed)
    sd2 = {k: v.clone() for k, v in sd.items()}
    ref_heads = [o.detach() for o in orc.forward({k: v.clone() for k, v in sd.items()}, x, nc, act, training=True)]
    ref_terms, ref_grads = orc.train_step_grads(sd2, x, tg, orc.TURBINE_ANCHORS, nc, act)
    sim_terms, sim_grads = orc.train_step_grads({k: v.clone() for k, v in sd.items()}, x, tg, orc.TURBINE_ANCHORS, nc, act, bf16_sim=True)
    sim_heads = [o.detach() for o in orc.forward({k: v.clone() for k, v in sd.items()}, x, nc, act, training=True, bf16_sim=True)]
    m = m.cuda().train()
    tr = Trainer(m, orc.TURBINE_ANCHORS, lr=0.0, momentum=0.0, weight_decay=0.0)
    plan = tr.plan(bsz, size, size)
    plan.forward(x.cuda())
    heads = [h.float().cpu() for h in plan.head_views()]
    for h, r, sm in zip(heads, ref_heads, sim_heads):
        cos = torch.nn.functional.cosine_similarity(h.flatten(), r.flatten(), dim=0)
        cos2 = torch.nn.functional.cosine_similarity(h.flatten(), sm.flatten(), dim=0)
        print(f"  head {tuple(r.shape)} cos {float(cos):.6f} maxabs {float((h - r).abs().max()):.4f} refmax {float(r.abs().max()):.3f} | vs bf16-sim cos {float(cos2):.6f} maxabs {float((h - sm).abs().max()):.4f}")
    terms = tr.step(x.cuda(), [t.cuda() for t in tg]).cpu().tolist()
    print(f"{nc}/{act}/{size}/B{bsz}: terms {terms} ref {ref_terms}")
    rows = []
    for k, p in m.named_parameters():
        g, r = p.grad.detach().float().cpu().flatten(), ref_grads[k].flatten()
        cos = float(torch.nn.functional.cosine_similarity(g, r, dim=0))
        rows.append((cos, float(g.norm() / (r.norm() + 1e-30)), k))
    rows.sort()
    print("  worst:", [(f"{c:.4f}", f"{n:.3f}", k) for c, n, k in rows[:8]])
    print(f"  mean cos {sum(r[0] for r in rows) / len(rows):.5f}; norm ratio range {min(r[1] for r in rows):.3f}..{max(r[1] for r in rows):.3f}")
    rows = []
    for k, p in m.named_parameters():
        g, r = p.grad.detach().float().cpu().flatten(), sim_grads[k].flatten()
        rows.append((float(torch.nn.functional.cosine_similarity(g, r, dim=0)), float(g.norm() / (r.norm() + 1e-30)), k))
    rows.sort()
    print(f"  vs bf16-sim: terms {sim_terms}; mean cos {sum(r[0] for r in rows) / len(rows):.5f}; worst", [(f"{c:.4f}", f"{n:.3f}", k) for c, n, k in rows[:5]])
