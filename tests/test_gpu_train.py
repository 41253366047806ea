"""GPU parity of one full training step (Trainer.step: train-mode forward, YOLOLoss x3, backward, SGD) against the
oracle's fp32 autograd restatement of code/train.py:53-69 (oracle.train_step_grads, pinned bit-for-bit to the
unmodified reference by tests/test_oracle_train.py) on the same seeded weights, images and targets.

Tolerances (stated, and why).  Activations, raw conv outputs and their gradients are stored in bf16 (the reference
trains under fp16 autocast, train.py:53); parameter gradients accumulate in fp32.  The kernels themselves are pinned
tightly in tests/test_gpu_train_kernels.py (wgrad/dgrad 2e-3, BN 2^-7).  End to end, through 75 conv + batch-stat
BatchNorm layers, bf16 rounding noise is amplified by the network itself: `oracle.train_step_grads(bf16_sim=True)` -- the
reference's arithmetic on the CPU with nothing changed but bf16 rounding at the same storage points -- deviates from the
fp32 reference exactly as much as this CUDA path does (profiles/r1_train_parity.txt).  With the smooth Mish activation
(what the reference trains with, train.py:299) gradients keep cosine >= 0.93 per tensor / >= 0.965 on average against the
fp32 oracle.  LeakyReLU's derivative jumps from 0.1 to 1 at zero, so every forward value that rounding moves across zero
flips a gradient factor; two bf16 evaluations of the SAME network then only correlate at ~0.75 per tensor (CPU bf16-sim
vs fp32: 0.42-0.77), and the leaky bounds below are that wide for that reason.  Loss terms must match to 8 % (observed <= 3.5 %; the object term of a 96x96 / batch-2 case moves by 1 % when only
the summation order of the BatchNorm statistics changes); gradient
norms to the stated ratio; BatchNorm running statistics to 1e-2."""
import pytest
import torch

pytestmark = pytest.mark.gpu

BOUNDS = {  # activation -> (min cosine per tensor, min mean cosine, allowed norm ratio range)
    "mish": (0.93, 0.965, (0.85, 1.15)),
    # End to end the LeakyReLU gradient is chaotic under bf16 storage (the CPU bf16-sim of the reference itself only
    # reaches cosine 0.35-0.45 per tensor / 0.67 on average against fp32, on synthetic AND default-init weights), so
    # this bound can only say "not anti-correlated".  What pins the LeakyReLU backward is the teacher-forced test
    # below (test_teacher_forced_layer_backward): every layer's dz / dgamma / dbeta / dW against fp32 arithmetic on
    # the SAME saved tensors, cosine >= 0.995.
    "leaky_relu": (0.0, 0.55, (0.4, 2.5)),
}


def _setup(nc, act, size, bsz, seed):
    from oracle import synth
    from oracle import yolo_oracle as orc
    from yolo_for_turbines_b200.model import YOLOv3

    m = YOLOv3(num_classes=nc, activation=act)
    sd = synth.synth_state_dict(m.state_dict(), seed=seed)
    m.load_state_dict(sd)
    x = torch.rand(bsz, 3, size, size, generator=torch.Generator().manual_seed(40 + seed))
    tg = orc.synth_targets(bsz, size, nc, 50 + seed)
    return m, {k: v.clone() for k, v in sd.items()}, x, tg


@pytest.mark.parametrize("nc,act,size,bsz,seed", [(2, "mish", 96, 4, 6), (80, "mish", 128, 8, 12), (2, "leaky_relu", 128, 8, 11)])
def test_train_step_matches_oracle(nc, act, size, bsz, seed):
    from oracle import yolo_oracle as orc
    from yolo_for_turbines_b200.train import Trainer

    m, sd, x, tg = _setup(nc, act, size, bsz, seed)
    ref_terms, ref_grads = orc.train_step_grads(sd, x, tg, orc.TURBINE_ANCHORS, nc, act)

    m = m.cuda().train()
    lr, mu, wd = 1e-3, 0.9, 5e-4
    before = {k: p.detach().clone() for k, p in m.named_parameters()}
    tr = Trainer(m, orc.TURBINE_ANCHORS, lr=lr, momentum=mu, weight_decay=wd)
    terms = tr.step(x.cuda(), [t.cuda() for t in tg])
    torch.cuda.synchronize()
    got_terms = terms.cpu().tolist()
    for a, b in zip(got_terms, ref_terms):
        assert abs(a - b) <= 0.08 * max(1.0, abs(b)), (got_terms, ref_terms)

    COS_MIN, COS_MIN_MEAN, (R_LO, R_HI) = BOUNDS[act]
    coss, worst = [], (2.0, None)
    for k, p in m.named_parameters():
        g, r = p.grad.detach().float().cpu().flatten(), ref_grads[k].flatten()
        cos = float(torch.nn.functional.cosine_similarity(g, r, dim=0))
        ratio = float(g.norm() / (r.norm() + 1e-30))
        coss.append(cos)
        if cos < worst[0]:
            worst = (cos, k)
        assert cos >= COS_MIN and R_LO <= ratio <= R_HI, (k, cos, ratio)
        # SGD, first step: p -= lr * (g + wd * p)
        exp = before[k].cpu() - lr * (p.grad.detach().cpu() + wd * before[k].cpu())
        assert torch.allclose(p.detach().cpu(), exp, rtol=1e-5, atol=1e-7), k
    print(f"train step {nc}/{act}/{size}: mean cosine {sum(coss) / len(coss):.5f}, worst {worst}")
    assert sum(coss) / len(coss) >= COS_MIN_MEAN

    msd = m.state_dict()
    for k in ("layers.0.batch_norm.running_mean", "layers.0.batch_norm.running_var",
              "layers.29.pred_block.0.batch_norm.running_mean", "layers.29.pred_block.0.batch_norm.running_var"):
        assert torch.allclose(msd[k].cpu(), sd[k], atol=1e-2, rtol=1e-2), k
    assert int(msd["layers.0.batch_norm.num_batches_tracked"]) == 1


def test_second_step_uses_momentum_and_new_weights():
    from oracle import yolo_oracle as orc
    from yolo_for_turbines_b200.train import Trainer

    m, sd, x, tg = _setup(2, "leaky_relu", 64, 2, 9)
    m = m.cuda().train()
    tr = Trainer(m, orc.TURBINE_ANCHORS, lr=1e-3, momentum=0.9, weight_decay=0.0)
    xs, ts = x.cuda(), [t.cuda() for t in tg]
    l0 = tr.step(xs, ts).cpu()
    p1 = {k: p.detach().clone() for k, p in m.named_parameters()}
    g1 = {k: p.grad.detach().clone() for k, p in m.named_parameters()}
    l1 = tr.step(xs, ts).cpu()
    torch.cuda.synchronize()
    k = "layers.29.pred_block.1.conv.bias"
    p = dict(m.named_parameters())[k]
    exp = p1[k] - 1e-3 * (0.9 * g1[k] + p.grad)      # buf = mu * buf + g
    assert torch.allclose(p.detach(), exp, rtol=1e-4, atol=1e-7)
    assert torch.isfinite(l0).all() and torch.isfinite(l1).all()
    # eval-mode inference after training picks up the new weights and running statistics
    m.eval()
    outs = m(xs)
    assert all(torch.isfinite(o).all() for o in outs)


def test_dropin_training_loop_matches_fused_trainer():
    """The reference's loop body (train.py:42-69) with the drop-in modules -- model.train(), `out = model(x)`,
    three YOLOLoss calls, GradScaler.scale(loss).backward(), scaler.step(torch SGD) -- against Trainer.step on a
    copy of the same model.  Both run the same kernels (the 2^16 loss scale is exact in bf16/fp32); the only
    difference is the order of fp32 atomics, which the network amplifies (see the module docstring), so the
    comparison is made on the first step: loss within 1 %, per-tensor gradient cosine >= 0.98, and each path's own
    update must be exactly SGD's."""
    import copy

    from oracle import yolo_oracle as orc
    from yolo_for_turbines_b200.loss import YOLOLoss
    from yolo_for_turbines_b200.train import Trainer

    m, sd, x, tg = _setup(2, "mish", 128, 8, 21)
    m2 = copy.deepcopy(m)
    m, m2 = m.cuda().train(), m2.cuda().train()
    xs, ts = x.cuda(), [t.cuda() for t in tg]
    lr, mu, wd = 1e-3, 0.9, 5e-4
    before = {k: p.detach().clone() for k, p in m.named_parameters()}

    opt = torch.optim.SGD(m.parameters(), lr=lr, momentum=mu, weight_decay=wd)
    scaler = torch.amp.GradScaler()
    loss_fn = YOLOLoss()
    scaled_anchors = [torch.tensor(a) * s for a, s in zip(orc.TURBINE_ANCHORS, (4, 8, 16))]
    opt.zero_grad()
    with torch.amp.autocast(device_type="cuda"):
        out = m(xs)
        terms = [loss_fn(o, t.clone(), a.cuda()) for o, t, a in zip(out, ts, scaled_anchors)]
        loss = sum(sum(t) for t in terms)
    scaler.scale(loss).backward()
    scaler.step(opt)
    scaler.update()
    torch.cuda.synchronize()
    assert out[0].shape == (8, 3, 4, 4, 7) and torch.isfinite(loss)

    tr = Trainer(m2, orc.TURBINE_ANCHORS, lr=lr, momentum=mu, weight_decay=wd)
    fused = tr.step(xs, ts)
    torch.cuda.synchronize()
    assert abs(float(fused.sum()) - float(loss)) <= 1e-2 * abs(float(loss))
    for (k, p), (_, q) in zip(m.named_parameters(), m2.named_parameters()):
        g1, g2 = p.grad.float(), q.grad    # GradScaler.step() has already unscaled p.grad in place
        cos = float(torch.nn.functional.cosine_similarity(g1.flatten(), g2.flatten(), dim=0))
        assert cos >= 0.98, (k, cos)
        for pp, gg in ((p, g1), (q, g2)):   # first SGD step: p -= lr * (g + wd * p)
            exp = before[k] - lr * (gg + wd * before[k])
            assert torch.allclose(pp.detach(), exp, rtol=1e-5, atol=1e-7), k
    sd1, sd2 = m.state_dict(), m2.state_dict()
    for k in sd1:
        if "running" in k:
            assert torch.allclose(sd1[k], sd2[k], rtol=5e-2, atol=2e-2), k
    assert int(sd1["layers.0.batch_norm.num_batches_tracked"]) == 1
    # a second pass through both loops keeps working (momentum buffers, repacked weights)
    opt.zero_grad()
    out = m(xs)
    sum(sum(loss_fn(o, t.clone(), a.cuda())) for o, t, a in zip(out, ts, scaled_anchors)).backward()
    opt.step()
    assert torch.isfinite(tr.step(xs, ts)).all()


def test_frozen_parameters_stay_put():
    """`freeze=True` in the reference marks loaded tensors requires_grad=False (model.py:306-309, 330-334) and
    torch.optim.SGD then skips them: frozen tensors must not move, trainable ones must."""
    from oracle import yolo_oracle as orc
    from yolo_for_turbines_b200.train import Trainer

    m, sd, x, tg = _setup(2, "leaky_relu", 64, 2, 31)
    m = m.cuda().train()
    named = list(m.named_parameters())
    frozen = [k for k, _ in named[:120]]
    for k, p in named[:120]:
        p.requires_grad = False
    before = {k: p.detach().clone() for k, p in named}
    tr = Trainer(m, orc.TURBINE_ANCHORS, lr=1e-2, momentum=0.9, weight_decay=5e-4)
    tr.step(x.cuda(), [t.cuda() for t in tg])
    torch.cuda.synchronize()
    moved = 0
    for k, p in m.named_parameters():
        if k in frozen:
            assert torch.equal(p.detach(), before[k]), k
            assert p.grad is None
        else:
            moved += int(not torch.equal(p.detach(), before[k]))
    assert moved == len(named) - len(frozen)


def test_multi_scale_steps_rebuild_plans():
    """Multi-scale training (config.py:43-45, dataset.py:113-117 change the image size every 10 batches): plans are
    cached per size (at most `max_plans`), evicted plans are rebuilt, and every step keeps training the same weights."""
    from oracle import yolo_oracle as orc
    from yolo_for_turbines_b200.dataset import encode_targets
    from yolo_for_turbines_b200.train import Trainer

    m, sd, x, tg = _setup(2, "mish", 64, 2, 41)
    m = m.cuda().train()
    tr = Trainer(m, orc.TURBINE_ANCHORS, lr=1e-3, momentum=0.9, weight_decay=5e-4, max_plans=2)
    g = torch.Generator().manual_seed(0)
    prev = {k: p.detach().clone() for k, p in m.named_parameters()}
    for size in (64, 96, 128, 64, 96):
        xb = torch.rand(2, 3, size, size, generator=g).cuda()
        boxes = [[[0.3, 0.4, 0.2, 0.3, 1.0], [0.7, 0.6, 0.4, 0.2, 0.0]], [[0.5, 0.5, 0.6, 0.5, 1.0]]]
        tb = encode_targets(boxes, orc.TURBINE_ANCHORS, image_size=size)
        losses = tr.step(xb, tb)
        torch.cuda.synchronize()
        assert torch.isfinite(losses).all(), (size, losses)
        k = "layers.29.pred_block.1.conv.bias"
        now = dict(m.named_parameters())[k].detach()
        assert not torch.equal(now, prev[k]), size
        prev[k] = now.clone()
    assert len(tr.plans) <= 2 and tr.steps_done == 5


def test_full_size_step_properties():
    """BASELINE configs[3] at full size (batch 32, 416x416, 2 classes, Mish), through properties that need no CPU
    reference: (1) a fixed batch is over-fitted -- the summed loss falls over 6 SGD steps; (2) every parameter,
    gradient, running statistic and loss term stays finite; (3) the data-gradient chain reaches the stem (its
    weight gradient is non-zero); (4) BatchNorm running statistics move towards the batch statistics."""
    import numpy as np

    from yolo_for_turbines_b200 import config as cfg
    from yolo_for_turbines_b200.dataset import encode_targets
    from yolo_for_turbines_b200.model import YOLOv3
    from yolo_for_turbines_b200.train import Trainer

    torch.manual_seed(0)
    B, S = 32, 416
    m = YOLOv3(num_classes=2, activation="mish").cuda().train()
    rm0 = m.layers[0].batch_norm.running_mean.clone()
    tr = Trainer(m, cfg.TURBINE_ANCHORS, lr=1e-3, momentum=0.9, weight_decay=5e-4)
    x = torch.rand(B, 3, S, S, device="cuda")
    rng = np.random.default_rng(0)
    boxes = []
    for _ in range(B):
        wh = rng.uniform(0.05, 0.5, (6, 2))
        xy = rng.uniform(wh / 2, 1 - wh / 2)
        boxes.append(np.concatenate([xy, wh, rng.integers(0, 2, (6, 1)).astype(np.float64)], axis=1))
    tg = encode_targets(boxes, cfg.TURBINE_ANCHORS, image_size=S)
    hist = []
    for _ in range(6):
        hist.append(float(tr.step(x, tg).sum()))
    torch.cuda.synchronize()
    assert all(np.isfinite(hist)) and hist[-1] < 0.9 * hist[0], hist
    assert all(bool(torch.isfinite(p).all()) and bool(torch.isfinite(p.grad).all()) for p in m.parameters())
    assert float(m.layers[0].conv.weight.grad.abs().max()) > 0
    sd = m.state_dict()
    assert all(bool(torch.isfinite(v).all()) for v in sd.values())
    assert not torch.equal(m.layers[0].batch_norm.running_mean, rm0)
    assert int(sd["layers.0.batch_norm.num_batches_tracked"]) == 6


@pytest.mark.parametrize("act", ["leaky_relu", "mish"])
def test_teacher_forced_layer_backward(act):
    """Every layer of the backward pass, teacher-forced: after one Trainer.step the plan still holds each layer's
    raw conv output z, batch statistics, input activation, the complete gradient dA of its output and the dz it
    produced.  For every batch-normalised layer the fp32 arithmetic of nn.BatchNorm2d + activation backward is applied
    to the CUDA path's OWN saved z / dA, and the weight gradient is recomputed from its own input and dz -- so bf16
    noise cannot accumulate across layers and the bound bites: cosine >= 0.995 for dz, dgamma, dbeta and dW of all
    72 layers (a wrong negative slope, a mis-scaled BatchNorm backward or a mis-indexed layer fails it)."""
    import torch.nn.functional as F
    from oracle import yolo_oracle as orc
    from yolo_for_turbines_b200.train import Trainer

    m, sd, x, tg = _setup(2, act, 64, 4, 21)
    m = m.cuda().train()
    tr = Trainer(m, orc.TURBINE_ANCHORS, lr=0.0, momentum=0.0, weight_decay=0.0)   # lr 0: parameters stay put
    tr.step(x.cuda(), [t.cuda() for t in tg])
    torch.cuda.synchronize()
    plan = tr.plan(4, 64, 64)
    B, checked, worst = plan.B, 0, (2.0, None)

    def cos(a, b):
        return float(F.cosine_similarity(a.flatten().double(), b.flatten().double(), dim=0))

    for op in plan.ops:
        if op.head or op.upsample:
            continue
        pc, blk = op.pc, op.block
        C_, cp, P = pc.c_out, pc.c_out_pad, op.P
        z = op.z.view(P, cp)[:, :C_].float()
        gbuf, goff, gpitch = plan._final_grad(op.dst)
        dA = gbuf.view(-1, gpitch)[:P, goff:goff + C_].float()
        mean, rstd = op.bn["mean"][:C_].float(), op.bn["rstd"][:C_].float()
        gamma, beta = blk.batch_norm.weight.detach().float(), blk.batch_norm.bias.detach().float()
        xhat = (z - mean) * rstd
        y = (xhat * gamma + beta).requires_grad_(True)
        out = F.leaky_relu(y, 0.1) if act == "leaky_relu" else F.mish(y)
        (dy,) = torch.autograd.grad(out, y, dA)
        dbeta, dgamma = dy.sum(0), (dy * xhat).sum(0)
        dz_ref = gamma * rstd * (dy - dbeta / P - xhat * (dgamma / P))
        got_dz = op.dz.view(P, cp)[:, :C_].float()
        res = {"dz": cos(got_dz, dz_ref), "dgamma": cos(blk.batch_norm.weight.grad, dgamma),
               "dbeta": cos(blk.batch_norm.bias.grad, dbeta)}
        if op.index > 0:   # weight gradient from the layer's own input and its own dz (the stem reads a patch matrix)
            sroot, soff = op.src.resolve()
            a = sroot.buf.view(B, op.src.H, op.src.W, sroot.C)[..., soff:soff + pc.c_in].float().permute(0, 3, 1, 2)
            go = got_dz.view(B, op.ho, op.wo, C_).permute(0, 3, 1, 2)
            dw_ref = torch.nn.grad.conv2d_weight(a, blk.conv.weight.shape, go, stride=pc.stride, padding=pc.pad)
            res["dW"] = cos(blk.conv.weight.grad, dw_ref)
        for k, v in res.items():
            if v < worst[0]:
                worst = (v, f"{op.name}.{k}")
            assert v >= 0.995, (op.name, k, v)
        checked += 1
    print(f"teacher-forced backward ({act}): {checked} layers, worst cosine {worst}")
    assert checked >= 68


def test_load_weights_after_first_forward_and_after_trainer(tmp_path):
    """Darknet weights loaded AFTER the packs were built (first forward / Trainer construction -- the
    reference's order is model -> optimizer -> load) must reach the kernels: load_weights() writes through the
    tensors (version bump), invalidates both caches, and Trainer.step re-validates its packs."""
    import numpy as np
    from oracle import yolo_oracle as orc
    from yolo_for_turbines_b200.model import YOLOv3
    from yolo_for_turbines_b200.train import Trainer

    import torch.nn as nn

    torch.manual_seed(3)
    ref = YOLOv3(num_classes=2)
    path = str(tmp_path / "synthetic.weights")
    rng = np.random.default_rng(5)
    chunks = []
    for mod in ref._darknet_modules():   # the file order: per block beta, gamma, mean, var, then the conv weights
        if isinstance(mod, nn.BatchNorm2d):
            c = mod.num_features
            chunks += [0.1 * rng.standard_normal(c), 1.0 + 0.1 * rng.standard_normal(c), 0.1 * rng.standard_normal(c),
                       1.0 + 0.2 * rng.random(c)]
        elif isinstance(mod, nn.Conv2d):
            if mod.bias is not None:
                chunks.append(0.1 * rng.standard_normal(mod.out_channels))
            fan_in = mod.in_channels * mod.kernel_size[0] * mod.kernel_size[1]
            chunks.append(rng.standard_normal(mod.weight.numel()) * (2.0 / fan_in) ** 0.5)
    with open(path, "wb") as f:
        np.zeros(5, dtype=np.int32).tofile(f)
        np.concatenate(chunks).astype(np.float32).tofile(f)
    x = torch.rand(2, 3, 64, 64, generator=torch.Generator().manual_seed(1)).cuda()

    a = YOLOv3(num_classes=2, weights_path=path)
    a.load_weights()
    a = a.cuda().eval()
    want = [o.clone() for o in a(x)]

    b = YOLOv3(num_classes=2, weights_path=path).cuda().eval()
    first = [o.clone() for o in b(x)]                  # packs built from the random init
    b.load_weights()                                   # ... then the file is loaded
    got = b(x)
    assert any(not torch.equal(u, v) for u, v in zip(first, got))
    for u, v in zip(want, got):
        assert torch.equal(u, v)

    tg = [t.cuda() for t in orc.synth_targets(2, 64, 2, 9)]
    c = YOLOv3(num_classes=2, weights_path=path).cuda().train()
    tr_c = Trainer(c, orc.TURBINE_ANCHORS, lr=0.0)     # packs built from the random init
    c.load_weights()
    d = YOLOv3(num_classes=2, weights_path=path)
    d.load_weights()
    d = d.cuda().train()
    tr_d = Trainer(d, orc.TURBINE_ANCHORS, lr=0.0)     # packs built from the loaded weights
    lc, ld = tr_c.step(x, tg).cpu(), tr_d.step(x, tg).cpu()   # step() re-validates the packs before its forward
    torch.cuda.synchronize()
    assert torch.isfinite(lc).all() and torch.isfinite(ld).all()
    # lr = 0: the parameters did not move, so both trainers must now hold bit-identical bf16 operand packs
    for bc, bd in zip(tr_c.blocks, tr_d.blocks):
        assert torch.equal(tr_c.engine.packed[id(bc)].w, tr_d.engine.packed[id(bd)].w)
        if id(bc) in tr_c.wT:
            assert torch.equal(tr_c.wT[id(bc)], tr_d.wT[id(bd)])
    # (the loss terms themselves only agree loosely: batch-2 BatchNorm over a 2x2 grid amplifies the order of fp32 atomics)
    assert torch.allclose(lc, ld, rtol=0.3, atol=1e-3), (lc, ld)


def test_autograd_forward_keeps_one_outstanding_graph():
    """the autograd wrapper keeps its activations in the plan's static buffers; backward() through a
    graph whose buffers a later forward has reused must raise instead of returning wrong gradients."""
    from oracle import yolo_oracle as orc
    from yolo_for_turbines_b200.loss import YOLOLoss

    m, sd, x, tg = _setup(2, "mish", 64, 2, 4)
    m = m.cuda().train()
    xs, ts = x.cuda(), [t.cuda() for t in tg]
    crit = YOLOLoss()

    def loss_of(outs):
        return sum(sum(crit(o, t.clone(), torch.tensor(orc.TURBINE_ANCHORS[i]) * o.shape[2])) for i, (o, t) in enumerate(zip(outs, ts)))

    l1 = loss_of(m(xs))
    l2 = loss_of(m(xs))                # same shape: overwrites the buffers l1's graph points at
    with pytest.raises(RuntimeError, match="ONE outstanding"):
        l1.backward()
    l2.backward()                      # the latest graph is intact
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters() if p.requires_grad)


def test_trainer_state_dict_round_trip():
    """the fused trainer's optimizer state (momentum buffers, step count) survives save / restore in the
    torch.optim.SGD layout utils.save_checkpoint stores: a restored run continues on the same trajectory."""
    from oracle import yolo_oracle as orc
    from yolo_for_turbines_b200.model import YOLOv3
    from yolo_for_turbines_b200.train import Trainer

    m, sd, x, tg = _setup(2, "mish", 64, 2, 8)
    xs, ts = x.cuda(), [t.cuda() for t in tg]
    a = m.cuda().train()
    tr_a = Trainer(a, orc.TURBINE_ANCHORS, lr=1e-3, momentum=0.9, weight_decay=5e-4)
    tr_a.step(xs, ts)
    tr_a.step(xs, ts)
    torch.cuda.synchronize()
    model_sd = {k: v.detach().clone() for k, v in a.state_dict().items()}
    opt_sd = tr_a.state_dict()
    ref_opt = torch.optim.SGD([p for p in a.parameters() if p.requires_grad], lr=1e-3, momentum=0.9, weight_decay=5e-4)
    assert set(opt_sd["param_groups"][0]) >= set(ref_opt.state_dict()["param_groups"][0])   # torch.optim.SGD layout
    ref_opt.load_state_dict({k: v for k, v in opt_sd.items() if k != "steps_done"})         # torch accepts it
    m_saved = tr_a.flat_m.clone()
    tr_a.step(xs, ts)
    torch.cuda.synchronize()
    want = {k: p.detach().clone() for k, p in a.named_parameters()}

    b = YOLOv3(num_classes=2, activation="mish")
    b.load_state_dict(model_sd)
    b = b.cuda().train()
    tr_b = Trainer(b, orc.TURBINE_ANCHORS, lr=0.5, momentum=0.0)     # wrong hyper-parameters: the state dict fixes them
    tr_b.load_state_dict(opt_sd)
    assert tr_b.steps_done == 2 and tr_b.momentum == 0.9 and tr_b.lr == 1e-3 and tr_b.weight_decay == 5e-4
    assert torch.equal(tr_b.flat_m[: tr_b.n_trainable], m_saved[: tr_a.n_trainable])   # momentum restored bit for bit
    before = {k: p.detach().clone() for k, p in b.named_parameters()}
    tr_b.step(xs, ts)
    torch.cuda.synchronize()
    # Same kernels, same inputs: the third step of both runs applies p -= lr (mu buf + g + wd p) with the same buf.
    # Gradients of a batch-2 / 64x64 step are only reproducible up to the order of fp32 atomics amplified through
    # batch-statistics BatchNorm, so the comparison is per tensor on the UPDATE, loosely -- a lost momentum buffer
    # (update = lr g instead of lr (0.9 buf + g)) or a reset step count changes it by far more.
    for k in ("layers.29.pred_block.1.conv.bias", "layers.22.pred_block.1.conv.bias", "layers.15.pred_block.1.conv.bias"):
        ub = (dict(b.named_parameters())[k].detach() - before[k]).flatten()
        ua = (want[k] - before[k]).flatten()
        cos = float(torch.nn.functional.cosine_similarity(ua, ub, dim=0))
        assert cos > 0.98 and 0.8 < float(ub.norm() / ua.norm()) < 1.25, (k, cos)


@pytest.mark.gpu
def test_graphed_step_follows_the_eager_trajectory():
    """Trainer.step(graph=True): the step captured once (three streams) and replayed follows the eager trajectory up to
    the run-to-run noise of the atomically accumulated sums (two EAGER runs differ by the same amount:
    scripts/graph_step_debug.py); static input buffers pick up new batches; an eager step after graphed ones sees
    current weight packs (the "mixed" run)."""
    import copy

    from oracle import yolo_oracle as orc
    from yolo_for_turbines_b200.model import YOLOv3
    from yolo_for_turbines_b200.train import Trainer

    torch.manual_seed(3)
    base = YOLOv3(num_classes=2, activation="mish")
    B, S = 4, 96
    xs = [torch.rand(B, 3, S, S, generator=torch.Generator().manual_seed(10 + i)).cuda() for i in range(3)]
    tgs = [[t.cuda() for t in orc.synth_targets(B, S, 2, 20 + i)] for i in range(3)]
    runs = {}
    for mode in ("eager", "graph", "mixed"):
        m = copy.deepcopy(base).cuda().train()
        tr = Trainer(m, orc.TURBINE_ANCHORS, lr=1e-5, momentum=0.9, weight_decay=5e-4)
        p0 = tr.flat_p[: tr.n_trainable].clone()
        losses = []
        for i in range(6):
            g = mode == "graph" or (mode == "mixed" and i in (2, 3))
            lr_i = 1e-5 * (0.5 + 0.25 * i)       # a per-iteration schedule (train.py:71-74): no new capture per value
            losses.append(tr.step(xs[i % 3], tgs[i % 3], lr=lr_i, graph=g).clone())
        torch.cuda.synchronize()
        runs[mode] = (torch.stack(losses).cpu(), (tr.flat_p[: tr.n_trainable] - p0).cpu(), len(tr._graphs))
    assert runs["eager"][2] == 0 and runs["graph"][2] == 1 and runs["mixed"][2] == 1
    le, de = runs["eager"][0], runs["eager"][1]
    assert float(de.abs().max()) > 1e-5     # the steps did move the parameters
    for mode in ("graph", "mixed"):
        lg, dg = runs[mode][0], runs[mode][1]
        assert torch.isfinite(lg).all()
        assert torch.allclose(lg, le, rtol=0.15, atol=0.05), (mode, lg, le)
        cos = torch.nn.functional.cosine_similarity(dg.flatten(), de.flatten(), dim=0)
        assert cos > 0.9, (mode, float(cos))             # same accumulated update (two eager runs: ~0.99)
        assert float((dg - de).abs().max()) < 1e-3, mode
