"""Training step of the reference (`train_one_epoch`'s loop body, code/train.py:42-69) on the sm_100a kernels.

    trainer = Trainer(model, anchors, lr=..., momentum=..., weight_decay=...)
    losses = trainer.step(x, (y0, y1, y2))          # forward (batch-stat BN) + YOLOLoss x3 + backward + SGD

What the reference does with autograd over ATen/cuDNN ops is planned once per (batch, H, W) as an explicit
launch list:

  forward   per CNNBlock (model.py:80-86, train mode): tcgen05 conv -> raw bf16 z; batch statistics
            (nn.BatchNorm2d semantics incl. the running-stat update); BN + activation (+ residual, + the 2x2
            replicated store that stands for nn.Upsample + torch.cat, model.py:189-191, :222)
  loss      YOLOLoss per scale (loss.py:29-81): fused forward sums, then the gradient w.r.t. the head logits
            written straight into the head convs' bf16 dz
  backward  per CNNBlock in reverse: BN + activation backward (dz, dgamma, dbeta); weight gradient on tcgen05
            (csrc/wgrad.cu); data gradient = the forward conv kernel on dz with the transposed / flipped weight
            pack (stride-2 layers: on the zero-stuffed dz).  Skip connections, the concat and the fan-out at
            the scale heads accumulate through the conv kernel's residual operand, so no add kernel exists.
  update    torch.optim.SGD(momentum, weight_decay) semantics (train.py:171-172) in one kernel over the flat
            fp32 parameter buffer, then the bf16 weight repack.

Data-parallel: one process per GPU; gradients live in ONE flat fp32 buffer that is all-reduced (NCCL, summed and
divided by the world size in the SGD kernel) in buckets on a side stream while the backward pass is still running.
The reference's fp16 autocast + GradScaler (train.py:53,67-69) becomes bf16 activations with fp32 statistics,
gradients of parameters and optimizer state; no loss scaling is needed with bf16's exponent range.

Parameters stay ordinary `nn.Parameter`s of the drop-in modules (their `.data` / `.grad` become views of the flat
buffers), so `state_dict()`, checkpoints and `model.eval()` inference keep working between steps.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn as nn

from . import engine as E
from ._lib import ACT_CODES, BnFinalizeDesc, ConvDesc, YoloB200Error, lib, ptr, require_cuda, stream_ptr


def _p(t: torch.Tensor, byte_off: int = 0) -> C.c_void_p:
    return C.c_void_p(t.data_ptr() + byte_off)


class _Grad:
    """Where the gradient of one activation tensor lives: an own dense bf16 buffer that data-gradient launches
    write (and accumulate into through their residual operand), and/or an alias = a channel slice of another
    gradient buffer that contributes to it (skip connection, route into the concat)."""
    __slots__ = ("buf", "alias", "init")

    def __init__(self):
        self.buf, self.alias, self.init = None, None, False


class _TrainOp:
    __slots__ = ("pc", "block", "src", "dst", "res", "upsample", "head", "name", "z", "P", "ho", "wo", "fwd_plan",
                 "dgrad_plan", "wgrad_plan", "dz", "bn", "g_w", "g_b", "g_gamma", "g_beta", "stuffed", "index")

    def __init__(self, **kw):
        for k in self.__slots__:
            setattr(self, k, kw.get(k))


class TrainPlan:
    """Static buffers, tensor maps and launch lists of one training step for one (batch, H, W)."""

    def __init__(self, trainer: "Trainer", B: int, H: int, W: int):
        from .model import CNNBlock, ResidualBlock, ScalePredictionBlock

        if H % 32 or W % 32:
            raise YoloB200Error(f"input size {H}x{W} must be a multiple of 32")
        self.trainer, self.B, self.H, self.W = trainer, B, H, W
        model, dev, eng = trainer.model, trainer.device, trainer.engine
        self.ops: List[_TrainOp] = []
        self.heads: List[_TrainOp] = []
        self.head_meta = []
        layers = list(model.layers)
        first_pc = eng.packed[id(layers[0])]
        if not first_pc.stem:
            raise YoloB200Error("the training path expects the 3-channel 3x3 stem (model.py:21)")
        cur = E._Act(first_pc.c_in_eff, H, W)
        self.input_act = cur
        routes: List[E._Act] = []

        def conv(block, src, res=None, head=False, name=""):
            pc = eng.packed[id(block)]
            ho = (src.H + 2 * pc.pad_eff - pc.k_eff) // pc.stride_eff + 1
            wo = (src.W + 2 * pc.pad_eff - pc.k_eff) // pc.stride_eff + 1
            dst = E._Act(pc.c_out_pad if head else pc.c_out, ho, wo, head)
            if not head and pc.c_out % 32:
                raise YoloB200Error(f"{name}: intermediate channel count {pc.c_out} must be a multiple of 32")
            if src.C != pc.c_in_eff:
                raise YoloB200Error(f"{name}: expects {pc.c_in_eff} input channels, got {src.C}")
            op = _TrainOp(pc=pc, block=block, src=src, dst=dst, res=res, upsample=False, head=head, name=name,
                          P=B * ho * wo, ho=ho, wo=wo)
            self.ops.append(op)
            return dst

        for li, layer in enumerate(layers):
            if isinstance(layer, ScalePredictionBlock):
                t = conv(layer.pred_block[0], cur, name=f"layers.{li}.pred_block.0")
                conv(layer.pred_block[1], t, head=True, name=f"layers.{li}.pred_block.1")
                self.heads.append(self.ops[-1])
                self.head_meta.append((layer.anchors_per_scale, layer.num_classes))
            elif isinstance(layer, CNNBlock):
                cur = conv(layer, cur, name=f"layers.{li}")
            elif isinstance(layer, ResidualBlock):
                for ri, seq in enumerate(layer.layers):
                    t = conv(seq[0], cur, name=f"layers.{li}.layers.{ri}.0")
                    cur = conv(seq[1], t, res=cur if layer.use_residual else None, name=f"layers.{li}.layers.{ri}.1")
                if layer.num_blocks == 8:
                    routes.append(cur)
            elif isinstance(layer, nn.Upsample):
                prod = self.ops[-1]
                if prod.dst is not cur or not routes:
                    raise YoloB200Error("Upsample must directly follow a conv and have a pending route")
                route = routes.pop()
                cat = E._Act(cur.C + route.C, 2 * cur.H, 2 * cur.W)
                up = E._Act(cur.C, cat.H, cat.W)
                up.root, up.ch_off = cat, 0
                route.root, route.ch_off = cat, cur.C
                prod.dst, prod.upsample = up, True
                cur = cat
            else:
                raise YoloB200Error(f"unsupported layer type {type(layer).__name__}")

        # ---- buffers: every activation and every raw conv output is kept for the backward pass -------------
        self.total_bytes = 0

        def alloc(n, dtype):
            t = torch.empty(n, dtype=dtype, device=dev)
            self.total_bytes += t.numel() * t.element_size()
            return t

        roots = {}
        for op in self.ops:
            for t in (op.src, op.dst, op.res):
                if t is not None:
                    r = t.resolve()[0]
                    if id(r) not in roots:
                        roots[id(r)] = r
                        r.buf = alloc(B * r.H * r.W * r.C, torch.float32 if r.fp32 else torch.bfloat16)
        cmax = max(op.pc.c_out_pad for op in self.ops)
        self.ones = torch.ones(max(cmax, 1024), dtype=torch.float32, device=dev)
        self.zeros = torch.zeros(max(cmax, 1024), dtype=torch.float32, device=dev)
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        n_bn = sum(op.pc.c_out_pad for op in self.ops)
        self.sums = torch.zeros(2 * n_bn * 2, dtype=torch.float64, device=dev)      # fwd stats | bwd sums
        self.bnf = torch.empty(6 * n_bn, dtype=torch.float32, device=dev)           # mean rstd scale bias m1 m2
        self.loss_sums = torch.zeros(18, dtype=torch.float64, device=dev)
        self.counters = torch.zeros(2 * len(self.ops), dtype=torch.int32, device=dev)   # last-block tickets (self-resetting)
        self.ev_dz = [torch.cuda.Event() for _ in self.ops]
        self.ev_branch = [torch.cuda.Event() for _ in self.ops]
        self.stuffed = None   # stride-2 data gradients run as row-parity sub-convolutions over dz: no zero-stuffed copy

        # gradient records
        self.grads: Dict[int, _Grad] = {}
        boff = 0
        for i, op in enumerate(self.ops):
            cp = op.pc.c_out_pad
            op.index = i
            op.bn = dict(sums=self.sums[2 * boff: 2 * boff + 2 * cp], bsums=self.sums[2 * n_bn + 2 * boff: 2 * n_bn + 2 * boff + 2 * cp],
                         mean=self.bnf[boff:boff + cp], rstd=self.bnf[n_bn + boff:n_bn + boff + cp],
                         scale=self.bnf[2 * n_bn + boff:2 * n_bn + boff + cp], bias=self.bnf[3 * n_bn + boff:3 * n_bn + boff + cp],
                         m1m2=self.bnf[4 * n_bn + 2 * boff:4 * n_bn + 2 * boff + 2 * cp])
            boff += cp
            op.z = None if op.head else alloc(op.P * cp, torch.bfloat16)
            op.dz = alloc(op.P * cp, torch.bfloat16)
            if op.head:
                op.dz.zero_()  # the padding channels (255 -> 256, 21 -> 32) are never written again
            op.stuffed = None

        def grad_of(t) -> _Grad:
            g = self.grads.get(id(t))
            if g is None:
                g = self.grads[id(t)] = _Grad()
            return g

        self._grad_of = grad_of
        for op in self.ops[1:]:  # every conv input except the image gets a dense gradient buffer
            g = grad_of(op.src)
            if g.buf is None:
                g.buf = alloc(B * op.src.H * op.src.W * op.src.C, torch.bfloat16)

        # ---- plans ------------------------------------------------------------------------------------
        for i, op in enumerate(self.ops):
            pc = op.pc
            sroot, soff = op.src.resolve()
            d = ConvDesc()
            d.batch, d.h_in, d.w_in, d.c_in, d.in_pitch = B, op.src.H, op.src.W, pc.c_in_eff, sroot.C
            d.c_out, d.c_out_pad, d.out_pitch = pc.c_out, pc.c_out_pad, pc.c_out_pad
            d.ksize, d.stride, d.pad = pc.k_eff, pc.stride_eff, pc.pad_eff
            d.act, d.out_fp32, d.check_nan = 0, int(op.head), 0
            d.want_stats = int(not op.head)
            d.pdl_hint = 1 if "f" in os.environ.get("YOLO_B200_TRAIN_NO_PDL", "") else 0   # A/B: "f" forward, "b" backward convs
            x_ptr = C.c_void_p(sroot.buf.data_ptr() + soff * 2)
            if op.head:
                droot, _ = op.dst.resolve()
                d.out_pitch = droot.C
                op.fwd_plan = E.make_conv_plan(d, x_ptr, ptr(pc.w), ptr(pc.scale), ptr(pc.bias), None, ptr(droot.buf))
            else:
                op.fwd_plan = E.make_conv_plan(d, x_ptr, ptr(pc.w), ptr(self.ones), ptr(self.zeros), None, ptr(op.z))
            # weight gradient: same geometry, dz as the second operand
            raw, blob = E._aligned_blob(int(lib.yolo_wgrad_plan_bytes()))
            lib.yolo_wgrad_plan_init(blob, lib.yolo_wgrad_plan_bytes(), C.byref(d), x_ptr, ptr(op.dz), pc.c_out_pad,
                                     _p(trainer.dw_packed, 4 * trainer.dw_off[id(op.block)]), 0)
            op.wgrad_plan = (raw, blob)
        # data gradients, in EXECUTION order (reverse): the first launch that targets a tensor's gradient writes it
        # (absorbing the skip / route contribution through the residual operand), later ones accumulate in place
        for i in range(len(self.ops) - 1, 0, -1):
            op = self.ops[i]
            pc = op.pc
            g = grad_of(op.src)
            s2 = pc.stride_eff == 2
            res_ptr, res_pitch = None, 0
            if g.init:
                res_ptr, res_pitch = ptr(g.buf), op.src.C
            else:
                alias = self._pending_alias(op.src)
                if alias is not None:
                    abuf, aoff, apitch = alias
                    res_ptr, res_pitch = _p(abuf, 2 * aoff), apitch
            plans = []
            for r in ((0, 1) if s2 else (None,)):
                dd = ConvDesc()
                dd.batch, dd.h_in, dd.w_in = B, op.ho, op.wo
                dd.c_in, dd.in_pitch = pc.c_out_pad, pc.c_out_pad
                dd.stride, dd.out_pitch = 1, op.src.C
                dd.pdl_hint = 1 if "b" in os.environ.get("YOLO_B200_TRAIN_NO_PDL", "") else 0
                if s2:   # dx[2a+r][2b+t] from dz[a+da][b+db]: stride-1 sub-convolution per output row parity
                    dd.c_out = dd.c_out_pad = 2 * pc.c_in_eff
                    dd.ksize, dd.pad = r + 1, 0
                    dd.ksize_w, dd.stride_w, dd.pad_w_hi_plus1, dd.pad_h_hi_plus1 = 2, 1, 2, r + 1
                    dd.s2_parity, dd.s2_cin = r + 1, pc.c_in_eff
                    wts = trainer.wT_s2[id(op.block)][r]
                else:
                    dd.c_out, dd.c_out_pad = pc.c_in_eff, pc.c_in_eff
                    dd.ksize, dd.pad = pc.k_eff, pc.pad_eff
                    wts = trainer.wT[id(op.block)]
                if res_ptr is not None:
                    dd.has_residual, dd.res_pitch = 1, res_pitch
                plans.append(E.make_conv_plan(dd, ptr(op.dz), ptr(wts), ptr(self.ones), ptr(self.zeros), res_ptr, ptr(g.buf)))
            op.dgrad_plan = plans
            g.init = True

    # the alias contribution to the gradient of tensor t, known statically from the graph:
    #   * t is the skip input of a residual pair  -> the gradient of the pair's output
    #   * t is a route tensor inside a concat     -> its channel slice of the concat's gradient
    def _pending_alias(self, t):
        for op in self.ops:
            if op.res is t:
                return self._final_grad(op.dst)
        if t.root is not None:
            cat = t.root
            return (self._grad_of(cat).buf, t.ch_off, cat.C)
        return None

    def _final_grad(self, t):
        """(buffer, channel offset, pitch) of the complete gradient of tensor t."""
        g = self.grads.get(id(t))
        if g is not None and g.buf is not None:
            return (g.buf, 0, t.C)
        if t.root is not None:   # the upsampled half of a concat: read through the concat's gradient
            cat = t.root
            return (self._grad_of(cat).buf, t.ch_off, cat.C)
        raise YoloB200Error("internal: activation without a gradient")

    # ----------------------------------------------------------------------------------------------------
    def head_views(self):
        outs = []
        for op, (na, nc) in zip(self.heads, self.head_meta):
            root, _ = op.dst.resolve()
            ch = nc + 5
            outs.append(torch.as_strided(root.buf, (self.B, na, op.ho, op.wo, ch),
                                         (op.ho * op.wo * root.C, ch, op.wo * root.C, root.C, 1)))
        return outs

    def head_grad_strides(self, op, nc):
        cp = op.pc.c_out_pad
        return (op.ho * op.wo * cp, nc + 5, op.wo * cp, cp, 1)

    def forward(self, x: torch.Tensor):
        tr, dev = self.trainer, self.trainer.device
        st = stream_ptr(dev)
        sp = ptr(self.status)
        self.sums.zero_()
        self.status.zero_()
        lib.yolo_input_patchify(ptr(x), self.B, x.shape[1], self.H, self.W, ptr(self.input_act.resolve()[0].buf), sp, st)
        fused_stats = tr.fused_stats
        pending = tr._packs_pending
        main = torch.cuda.current_stream(dev)
        def run_op(op, st):
            if op.head:
                lib.yolo_conv_fwd(op.fwd_plan[1], sp, st)
                return
            pc, bn, blk = op.pc, op.bn, op.block.batch_norm
            C_ = pc.c_out
            mom = float(blk.momentum if blk.momentum is not None else 0.1)
            if fused_stats:   # batch statistics accumulated by the conv epilogue itself; its last CTA finalises them
                fin = BnFinalizeDesc(op.P, blk.weight.data_ptr(), blk.bias.data_ptr(), float(blk.eps), mom,
                                     blk.running_mean.data_ptr(), blk.running_var.data_ptr(), bn["mean"].data_ptr(),
                                     bn["rstd"].data_ptr(), bn["scale"].data_ptr(), bn["bias"].data_ptr(),
                                     self.counters.data_ptr() + 8 * op.index)
                lib.yolo_conv_fwd_stats(op.fwd_plan[1], sp, ptr(bn["sums"]), C.byref(fin), st)
            else:
                lib.yolo_conv_fwd(op.fwd_plan[1], sp, st)
                lib.yolo_bn_stats_finalize(ptr(op.z), op.P, C_, pc.c_out_pad, ptr(bn["sums"]), _p(self.counters, 8 * op.index),
                                           ptr(blk.weight), ptr(blk.bias), float(blk.eps), mom, ptr(blk.running_mean),
                                           ptr(blk.running_var), ptr(bn["mean"]), ptr(bn["rstd"]), ptr(bn["scale"]),
                                           ptr(bn["bias"]), st)
            droot, doff = op.dst.resolve()
            res_ptr, res_pitch = None, 0
            if op.res is not None:
                rroot, roff = op.res.resolve()
                res_ptr, res_pitch = _p(rroot.buf, 2 * roff), rroot.C
            lib.yolo_bn_act_fwd(ptr(op.z), op.P, C_, pc.c_out_pad, ptr(bn["scale"]), ptr(bn["bias"]), ACT_CODES[pc.act],
                                res_ptr, res_pitch, _p(droot.buf, 2 * doff), droot.C, int(op.upsample), op.ho, op.wo, st)

        # The scale heads (ScalePredictionBlock, model.py:123-148) branch off the trunk and only the loss reads them:
        # they run on the side stream, beside the trunk layers that follow (small 13x13 / 26x26 kernels either way).
        side = tr.wgrad_stream
        branched = False
        for op in self.ops:
            if pending:
                main.wait_event(tr.ev_pack[id(op.block)])   # this layer's operand packs (side stream) are ready
            if side is not None and ".pred_block." in op.name:
                if op.name.endswith(".pred_block.0"):
                    self.ev_branch[op.index].record(main)   # the trunk tensor the head reads is complete
                    side.wait_event(self.ev_branch[op.index])
                with torch.cuda.stream(side):
                    run_op(op, stream_ptr(dev))
                branched = True
            else:
                run_op(op, st)
        if branched:
            main.wait_stream(side)
        tr._packs_pending = False
        torch._foreach_add_(tr.bn_counters, 1)   # nn.BatchNorm2d.num_batches_tracked (state_dict parity)

    def backward(self, on_op_done=None):
        """Runs after the head convs' dz have been written (Trainer.step does that with yolo_loss_bwd).

        The weight gradients (tensor-core / L2-bound, almost no DRAM traffic) go to a side stream: nothing downstream
        in the backward pass depends on them, so they overlap with the DRAM-bound BatchNorm / activation passes and
        with the data-gradient convs of the layers below instead of serialising behind them."""
        tr, dev = self.trainer, self.trainer.device
        st = stream_ptr(dev)
        sp = ptr(self.status)
        main = torch.cuda.current_stream(dev)
        side = tr.wgrad_stream
        for i in range(len(self.ops) - 1, -1, -1):
            op = self.ops[i]
            pc, bn = op.pc, op.bn
            if op.head:
                if op.g_b is not None:
                    lib.yolo_bias_grad(ptr(op.dz), op.P, pc.c_out_pad, pc.c_out_pad, pc.c_out, ptr(bn["bsums"]), ptr(op.g_b), st)
            else:
                gbuf, goff, gpitch = self._final_grad(op.dst)
                lib.yolo_bn_act_bwd(_p(gbuf, 2 * goff), gpitch, int(op.upsample), ptr(op.z), pc.c_out_pad, op.P, pc.c_out,
                                    op.ho, op.wo, ptr(bn["scale"]), ptr(bn["bias"]), ptr(bn["mean"]), ptr(bn["rstd"]),
                                    ACT_CODES[pc.act], ptr(bn["bsums"]), _p(self.counters, 8 * op.index + 4), ptr(op.g_gamma),
                                    ptr(op.g_beta), ptr(bn["m1m2"]),
                                    ptr(op.dz), pc.c_out_pad, ptr(op.stuffed) if op.stuffed is not None else None,
                                    pc.c_out_pad, st)
            if op.g_w is not None:
                if side is not None:
                    self.ev_dz[i].record(main)          # dz of this op (and everything before it on main) is ready
                    side.wait_event(self.ev_dz[i])
                    with torch.cuda.stream(side):
                        sst = stream_ptr(dev)
                        lib.yolo_wgrad(op.wgrad_plan[1], sst)
                        lib.yolo_unpack_wgrad(_p(tr.dw_packed, 4 * tr.dw_off[id(op.block)]), pc.c_out, pc.c_in, pc.ksize,
                                              pc.c_in_eff, int(pc.stem), ptr(op.g_w), sst)
                else:
                    lib.yolo_wgrad(op.wgrad_plan[1], st)
                    lib.yolo_unpack_wgrad(_p(tr.dw_packed, 4 * tr.dw_off[id(op.block)]), pc.c_out, pc.c_in, pc.ksize,
                                          pc.c_in_eff, int(pc.stem), ptr(op.g_w), st)
            if i > 0:
                for dp in op.dgrad_plan:
                    lib.yolo_conv_fwd(dp[1], sp, st)
            if on_op_done is not None:
                on_op_done(i)
        if side is not None:
            main.wait_stream(side)


def make_buckets(firsts, n_trainable: int, bucket_elems: int):
    """All-reduce schedule over the flat gradient buffer.  `firsts` = (offset of the op's first parameter, op index)
    for every op that owns trainable parameters; parameters are laid out in module order = forward op order, and the
    backward pass finishes ops in DESCENDING index order.  Returns [(first_op, lo, hi)] in firing order: the slice
    [lo, hi) is complete -- and is all-reduced on the side stream -- as soon as op `first_op` has run its backward.
    Slices are disjoint, cover [0, n_trainable) and hold at least `bucket_elems` elements (except the last one)."""
    firsts = sorted(firsts)
    buckets, hi_idx, hi = [], len(firsts), n_trainable
    while hi_idx > 0:
        j = hi_idx - 1
        while j > 0 and hi - firsts[j][0] < bucket_elems:
            j -= 1
        lo = firsts[j][0] if j > 0 else 0
        buckets.append((min(f[1] for f in firsts[j:hi_idx]), lo, hi))
        hi, hi_idx = lo, j
    return buckets


class _TrainForwardFn(torch.autograd.Function):
    """Train-mode forward of the whole network as ONE autograd node: forward = TrainPlan.forward, backward = the head
    gradients converted to the head convs' bf16 dz + TrainPlan.backward; parameter gradients are returned to autograd."""

    @staticmethod
    def forward(ctx, trainer, x, *params):
        require_cuda(x, "YOLOv3 input")
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.float().contiguous()
        with torch.cuda.device(trainer.device):
            trainer.repack_if_changed()
            plan = trainer.plan(x.shape[0], x.shape[2], x.shape[3])
            plan.forward(x)
        # The activations the backward pass needs live in the plan's static buffers (11 GB at batch 32, 416^2), not
        # in this node: ONE forward of a given shape may be outstanding.  Every forward / backward that touches the
        # buffers bumps the plan's generation; backward() refuses to run on buffers a later forward has overwritten.
        plan.generation = getattr(plan, "generation", 0) + 1
        ctx.generation = plan.generation
        ctx.trainer, ctx.plan = trainer, plan
        ctx.n_params = len(params)
        eng = trainer.model.__dict__.get("_yb_engine")
        if eng is not None:
            eng._sig = None   # running statistics moved: the eval-mode engine must re-fold BatchNorm
        return tuple(v.clone(memory_format=torch.preserve_format) for v in plan.head_views())

    @staticmethod
    def backward(ctx, *gheads):
        tr, plan = ctx.trainer, ctx.plan
        if ctx.generation != getattr(plan, "generation", 0):
            raise RuntimeError(
                "backward() of a train-mode forward whose saved activations are gone: this path keeps ONE outstanding "
                f"forward per input shape ({plan.B}x{plan.H}x{plan.W}) in static buffers, and a later forward (or an earlier "
                "backward through the same graph) has reused them.  Call backward() before the next model(x) of this "
                "shape; for gradient accumulation run forward+backward per micro-batch.")
        plan.generation += 1   # this backward consumes the buffers (dz is rewritten in place): no second backward
        with torch.cuda.device(tr.device):
            tr.dw_packed.zero_()
            plan.sums[plan.sums.numel() // 2:].zero_()
            for op, (na, nc), g in zip(plan.heads, plan.head_meta, gheads):
                dzv = op.dz.view(plan.B, op.ho, op.wo, op.pc.c_out_pad)[..., : na * (nc + 5)]
                if g is None:
                    dzv.zero_()
                else:   # (B, 3, S, S, 5+nc) -> NHWC channel order a*(5+nc)+c, bf16
                    dzv.copy_(g.permute(0, 2, 3, 1, 4).reshape(plan.B, op.ho, op.wo, na * (nc + 5)))
            plan.backward()
        grads = [tr.grad_view.get(id(p)) for p in tr.all_params]
        return (None, None) + tuple(g.clone() if g is not None else None for g in grads)


class Trainer:
    """SGD training of a drop-in `YOLOv3` (or any module built from its blocks) on the sm_100a path."""

    def __init__(self, model, anchors, lr: float, momentum: float = 0.0, weight_decay: float = 0.0,
                 process_group=None, bucket_mb: float = 32.0, max_plans: int = 2, own_params: bool = True,
                 data_parallel: bool = True):
        from .model import CNNBlock

        p0 = next(model.parameters())
        require_cuda(p0, "model parameters")
        self.model, self.device = model, p0.device
        self.lr, self.momentum, self.weight_decay = float(lr), float(momentum), float(weight_decay)
        self.anchors = [[tuple(map(float, a)) for a in scale] for scale in anchors]   # fractions of the image, config.py:47-57
        self.steps_done = 0
        self.max_plans = max_plans
        # BatchNorm statistics from the conv epilogue (default) or from a separate pass over z (YOLO_B200_BN_STATS_PASS=1)
        self.fused_stats = os.environ.get("YOLO_B200_BN_STATS_PASS") != "1"
        self.pg = process_group
        self.world = 1
        if data_parallel and (process_group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized())):
            self.world = torch.distributed.get_world_size(process_group)
        dev = self.device
        with torch.cuda.device(dev):
            # a dedicated engine: no pixel-pair folding (the backward kernels see the plain NHWC tensors)
            self.engine = E.Engine(model, dev)
            self.engine.allow_fold = False
            first = model.layers[0]
            self.engine.packed = {}
            self.blocks = [m for m in model.modules() if isinstance(m, CNNBlock)]
            for b in self.blocks:
                self.engine.packed[id(b)] = E.PackedConv(b, dev, as_stem=b is first, allow_fold=False)
            self.bn_modules = [b.batch_norm for b in self.blocks if b.batch_norm_act]
            self.bn_counters = [m.num_batches_tracked for m in self.bn_modules if m.num_batches_tracked is not None]

            # ---- flat fp32 parameter / gradient / momentum buffers (trainable first) --------------------
            params = [p for p in model.parameters()]
            trainable = [p for p in params if p.requires_grad]
            frozen = [p for p in params if not p.requires_grad]
            offs, n = {}, 0
            for p in trainable:
                offs[id(p)] = n
                n += (p.numel() + 3) // 4 * 4
            self.n_trainable = n
            for p in frozen:
                offs[id(p)] = n
                n += (p.numel() + 3) // 4 * 4
            # own_params=False (the autograd wrapper of model.YOLOv3.forward): parameters and optimizer stay the
            # caller's; only the flat gradient buffer exists and its views are handed to autograd.
            self.own_params = bool(own_params)
            self.flat_g = torch.zeros(n, dtype=torch.float32, device=dev)
            self.grad_view = {}
            if self.own_params:
                self.flat_p = torch.zeros(n, dtype=torch.float32, device=dev)
                self.flat_m = torch.zeros(max(self.n_trainable, 4), dtype=torch.float32, device=dev)
            for p in trainable + frozen:
                o = offs[id(p)]
                if self.own_params:
                    view = self.flat_p[o:o + p.numel()].view_as(p)
                    view.copy_(p.data)
                    p.data = view
                if p.requires_grad:
                    self.grad_view[id(p)] = self.flat_g[o:o + p.numel()].view_as(p)
                    if self.own_params:
                        p.grad = self.grad_view[id(p)]
            self.param_off = offs
            self.all_params = params
            self._packed_sig = None

            # ---- packed weight-gradient arena + transposed weight packs --------------------------------
            self.dw_off, tot = {}, 0
            self.wT = {}
            for b in self.blocks:
                pc = self.engine.packed[id(b)]
                self.dw_off[id(b)] = tot
                tot += pc.c_out_pad * pc.k_eff * pc.k_eff * pc.c_in_eff
                if b is not first:
                    self.wT[id(b)] = torch.empty(pc.c_in_eff * pc.ksize * pc.ksize * pc.c_out_pad, dtype=torch.bfloat16, device=dev)
            self.wT_s2 = {}   # stride-2 layers: one data-gradient pack per output row parity
            for b in self.blocks:
                pc = self.engine.packed[id(b)]
                if b is not first and pc.stride_eff == 2:
                    self.wT_s2[id(b)] = [torch.empty(2 * pc.c_in_eff * (r + 1) * 2 * pc.c_out_pad, dtype=torch.bfloat16, device=dev)
                                         for r in (0, 1)]
            self.dw_packed = torch.zeros(tot, dtype=torch.float32, device=dev)
            self.losses = torch.zeros(4, dtype=torch.float32, device=dev)
        self.plans: Dict[tuple, TrainPlan] = {}
        self.comm_stream = torch.cuda.Stream(device=dev) if self.world > 1 else None
        # weight gradients on their own stream (YOLO_B200_WGRAD_STREAM=0: same stream, for A/B timing)
        self.wgrad_stream = torch.cuda.Stream(device=dev) if os.environ.get("YOLO_B200_WGRAD_STREAM") != "0" else None
        self.debug_local_grads = None   # set to a tensor like flat_g to capture the pre-all-reduce gradients (tests)
        self.bucket_elems = int(bucket_mb * 1e6 / 4)
        self.ev_pack = {id(b): torch.cuda.Event() for b in self.blocks}
        self._ev_sgd = torch.cuda.Event()
        self._packs_pending = False
        self._graphs: Dict[tuple, dict] = {}
        self._lr_dev = torch.zeros(1, dtype=torch.float32, device=dev)   # learning rate of the captured step
        self.repack(full=True)
        self._packed_sig = self._param_sig()

    # ------------------------------------------------------------------------------------------------------
    def repack_async(self):
        """After the SGD update: repack on the side stream, one event per block; the next forward waits per layer
        (TrainPlan.forward), so the packs of the deep layers overlap with the first convs of the next step."""
        dev = self.device
        if self.wgrad_stream is None:
            self.repack()
            self._packs_pending = False
            return
        main = torch.cuda.current_stream(dev)
        self._ev_sgd.record(main)
        self.wgrad_stream.wait_event(self._ev_sgd)
        with torch.cuda.stream(self.wgrad_stream):
            self.repack(events=True)
        self._packs_pending = True

    def repack(self, full: bool = False, events: bool = False):
        """bf16 operand packs of the current fp32 weights: forward [Cout][k][k][Cin] and data-gradient
        [Cin][k][k][Cout] (flipped).  full=True also writes the zero padding (once, at construction); afterwards
        one launch per layer rewrites the real entries of both packs."""
        dev = self.device
        with torch.cuda.device(dev):
            st = stream_ptr(dev)
            for b in self.blocks:
                pc = self.engine.packed[id(b)]
                w = b.conv.weight
                if pc.stem:
                    lib.yolo_pack_stem_weights(ptr(w), pc.c_out, pc.c_in, pc.c_out_pad, ptr(pc.w), st)
                elif full:
                    lib.yolo_pack_weights(ptr(w), pc.c_out, pc.c_in, pc.ksize, pc.c_out_pad, pc.c_in_eff, ptr(pc.w), st)
                    lib.yolo_pack_weights_dgrad(ptr(w), pc.c_out, pc.c_in, pc.ksize, pc.c_in_eff, pc.c_out_pad,
                                                ptr(self.wT[id(b)]), st)
                else:
                    lib.yolo_pack_weights_train(ptr(w), pc.c_out, pc.c_in, pc.ksize, pc.c_in_eff, pc.c_out_pad, ptr(pc.w),
                                                ptr(self.wT[id(b)]), st)
                if id(b) in self.wT_s2:
                    for r in (0, 1):
                        lib.yolo_pack_weights_dgrad_s2(ptr(w), pc.c_out, pc.c_in, r, pc.c_in_eff, pc.c_out_pad,
                                                       ptr(self.wT_s2[id(b)][r]), st)
                if not b.batch_norm_act:  # head conv: scale 1, bias = conv bias
                    lib.yolo_fold_bn(None, None, None, None, ptr(b.conv.bias), 0.0, pc.c_out, pc.c_out_pad, ptr(pc.scale),
                                     ptr(pc.bias), st)
                if events:
                    self.ev_pack[id(b)].record(torch.cuda.current_stream(dev))

    def _param_sig(self):
        return tuple((p._version, p.data_ptr()) for p in self.all_params)

    def _repack_if_written_externally(self):
        """Fused mode: the SGD kernel updates the flat buffer directly (no version bump) and repacks itself, so the
        version counters only move when someone ELSE wrote the parameters -- load_weights(), load_state_dict(),
        load_checkpoint() after the Trainer was built (the reference's order: model -> optimizer -> load).  Then the
        bf16 packs are rebuilt before the forward uses them."""
        sig = self._param_sig()
        if sig != self._packed_sig:
            if self._packs_pending and self.wgrad_stream is not None:
                torch.cuda.current_stream(self.device).wait_stream(self.wgrad_stream)
            self.repack()
            self._packs_pending = False
            self._packed_sig = sig

    def repack_if_changed(self):
        """Autograd mode: the caller's optimizer updates the parameters in place between forwards."""
        sig = tuple((p._version, p.data_ptr()) for p in self.all_params)
        if sig != self._packed_sig:
            self.repack()
            self._packed_sig = sig

    # ---- autograd wrapper (model.YOLOv3.forward in train mode) -------------------------------------------
    def autograd_forward(self, x: torch.Tensor):
        """`out = model(x)` of train.py:54 with model.train(): returns the three head tensors attached to the autograd
        graph, so that the reference's own `loss.backward()` / optimizer / GradScaler code runs unchanged."""
        outs = _TrainForwardFn.apply(self, x, *self.all_params)
        return list(outs)

    def plan(self, B, H, W) -> TrainPlan:
        key = (B, H, W)
        p = self.plans.get(key)
        if p is None:
            while len(self.plans) >= self.max_plans:   # multi-scale training (config.py:43-45): bound the arenas kept
                old = next(iter(self.plans))
                self.plans.pop(old)
                for gk in [k for k in self._graphs if (k[0][0], k[0][2], k[0][3]) == old]:
                    self._graphs.pop(gk)   # a captured step points into the evicted plan's buffers
            with torch.cuda.device(self.device):
                p = TrainPlan(self, B, H, W)
                for op in p.ops:
                    blk = op.block
                    gv = lambda t: self.grad_view.get(id(t)) if t is not None else None  # noqa: E731
                    op.g_w = gv(blk.conv.weight)
                    op.g_b = gv(blk.conv.bias) if blk.conv.bias is not None else None
                    if blk.batch_norm_act:
                        op.g_gamma, op.g_beta = gv(blk.batch_norm.weight), gv(blk.batch_norm.bias)
                        if op.g_gamma is None or op.g_beta is None:   # frozen BN affine: scratch sinks
                            sink = torch.empty(2 * op.pc.c_out_pad, dtype=torch.float32, device=self.device)
                            op.g_gamma = op.g_gamma if op.g_gamma is not None else sink[:op.pc.c_out_pad]
                            op.g_beta = op.g_beta if op.g_beta is not None else sink[op.pc.c_out_pad:]
            self.plans[key] = p
        return p

    # ------------------------------------------------------------------------------------------------------
    def _check_targets(self, plan: TrainPlan, targets: Sequence[torch.Tensor]):
        """Shape / dtype / device of the three target tensors (also called before a step is captured: an error must not
        surface in the middle of a stream capture)."""
        if len(targets) != len(plan.heads):
            raise YoloB200Error(f"expected {len(plan.heads)} target tensors (one per scale), got {len(targets)}")
        for s, (op, tgt) in enumerate(zip(plan.heads, targets)):
            if tuple(tgt.shape) != (plan.B, 3, op.ho, op.wo, 6) or tgt.dtype != torch.float32 or tgt.device != self.device:
                raise YoloB200Error(f"target {s}: expected fp32 {(plan.B, 3, op.ho, op.wo, 6)} on {self.device}, got {tuple(tgt.shape)}")
            if op.ho != op.wo:
                raise YoloB200Error("YOLOLoss works on square grids (loss.py:29-81 with the reference's (B,3,S,S,6) targets)")

    def _loss_and_head_grads(self, plan: TrainPlan, targets: Sequence[torch.Tensor]):
        dev = self.device
        st = stream_ptr(dev)
        plan.loss_sums.zero_()
        views = plan.head_views()
        calls = []
        self._check_targets(plan, targets)
        for s, (op, (na, nc), pred, tgt) in enumerate(zip(plan.heads, plan.head_meta, views, targets)):
            S = op.ho
            anc = (C.c_float * 6)(*[v * S for a in self.anchors[s] for v in a])   # train.py:195-197 scaled anchors
            ps, ts = (C.c_int64 * 5)(*pred.stride()), (C.c_int64 * 5)(*tgt.stride())
            sums = plan.loss_sums[6 * s:6 * s + 6]
            lib.yolo_loss_fwd(ptr(pred), ps, ptr(tgt), ts, plan.B, S, nc, anc, 0, ptr(sums), st)
            calls.append((op, nc, pred, tgt, ps, ts, anc, sums, S))
        for op, nc, pred, tgt, ps, ts, anc, sums, S in calls:
            ds = (C.c_int64 * 5)(*plan.head_grad_strides(op, nc))
            lib.yolo_loss_bwd(ptr(pred), ps, ptr(tgt), ts, plan.B, S, nc, anc, ptr(sums), 1.0, None, ptr(op.dz), ds, 1, st)
        # [box, object, no-object, class] summed over the three scales, weighted as loss.py:78-81 (device, no sync)
        s = plan.loss_sums.view(3, 6)
        n_obj, n_no = s[:, 5], s[:, 1]
        has = n_obj > 0
        z = torch.zeros_like(n_obj)
        box = torch.where(has, s[:, 3] / (4 * n_obj), z)
        obj = torch.where(has, s[:, 2] / n_obj, z)
        noobj = s[:, 0] / n_no
        cls = torch.where(has, s[:, 4] / n_obj, z)
        return torch.stack([5 * box.sum(), obj.sum(), 0.5 * noobj.sum(), cls.sum()]).to(torch.float32)

    def _buckets(self, plan: TrainPlan):
        if getattr(plan, "_bk", None) is None:
            firsts = []
            for i, op in enumerate(plan.ops):
                ps = [p for p in op.block.parameters() if p.requires_grad]
                if ps:
                    firsts.append((min(self.param_off[id(p)] for p in ps), i))
            plan._bk = make_buckets(firsts, self.n_trainable, self.bucket_elems)
        return plan._bk

    def step(self, x: torch.Tensor, targets: Sequence[torch.Tensor], lr: Optional[float] = None, graph: bool = False) -> torch.Tensor:
        """One optimisation step (train.py:42-69).  Returns the device tensor [box, object, no_object, class] of
        the loss terms summed over the three scales (their sum is the reference's `loss`); no host sync.
        graph=True replays the step as one CUDA graph (see _step_graphed)."""
        require_cuda(x, "Trainer.step input")
        if not self.own_params:
            raise YoloB200Error("Trainer.step needs own_params=True (the autograd wrapper leaves the update to the caller)")
        if x.dim() != 4 or x.shape[1] != self.model.in_channels:
            raise YoloB200Error(f"expected (B,{self.model.in_channels},H,W) input, got {tuple(x.shape)}")
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.float().contiguous()
        if graph:
            return self._step_graphed(x, targets, float(self.lr if lr is None else lr))
        dev = self.device
        with torch.cuda.device(dev):
            self._repack_if_written_externally()
            self._step_body(x, targets, float(self.lr if lr is None else lr), repack_first=False)
        self.steps_done += 1
        eng = self.model.__dict__.get("_yb_engine")
        if eng is not None:
            eng._sig = None  # the inference engine's packed weights are stale now
        return self.losses

    def _step_body(self, x, targets, lr: float, repack_first: bool):
        """The launches of one step on the current stream (+ the side / communication streams forked from and joined
        back into it).  repack_first=False (eager): forward .. SGD, then the operand repack on the side stream, which the
        NEXT call's forward waits for layer by layer.  repack_first=True (the CUDA-graph order): the same repack opens the
        step instead -- it packs the parameters the previous replay's SGD wrote -- so every stream has rejoined the
        origin when the step ends, as stream capture requires; the overlap with the first forward layers is the same."""
        dev = self.device
        plan = self.plan(x.shape[0], x.shape[2], x.shape[3])
        st = stream_ptr(dev)
        main = torch.cuda.current_stream(dev)
        if repack_first:
            self.repack_async()
        self.dw_packed.zero_()
        plan.forward(x)
        plan.generation = getattr(plan, "generation", 0) + 1
        self.losses = self._loss_and_head_grads(plan, targets)
        if self.world > 1:
            pending = list(self._buckets(plan))

            def on_op_done(i):
                while pending and pending[0][0] >= i:
                    _, lo, hi = pending.pop(0)
                    ev = torch.cuda.Event()
                    ev.record(main)
                    self.comm_stream.wait_event(ev)
                    if self.wgrad_stream is not None:   # the bucket's weight gradients are written on the side stream
                        ev2 = torch.cuda.Event()
                        ev2.record(self.wgrad_stream)
                        self.comm_stream.wait_event(ev2)
                    with torch.cuda.stream(self.comm_stream):
                        if self.debug_local_grads is not None:   # tests: this rank's gradient before the exchange
                            self.debug_local_grads[lo:hi].copy_(self.flat_g[lo:hi])
                        torch.distributed.all_reduce(self.flat_g[lo:hi], group=self.pg)
            plan.backward(on_op_done)
            main.wait_stream(self.comm_stream)
        else:
            plan.backward()
        if repack_first:   # captured step: the learning rate is a device scalar, set before every replay
            lib.yolo_sgd_step_dev(ptr(self.flat_p), ptr(self.flat_g), ptr(self.flat_m), self.n_trainable, ptr(self._lr_dev),
                                  self.momentum, self.weight_decay, 1.0 / self.world, int(self.steps_done == 0), st)
        else:
            lib.yolo_sgd_step(ptr(self.flat_p), ptr(self.flat_g), ptr(self.flat_m), self.n_trainable,
                              lr, self.momentum, self.weight_decay, 1.0 / self.world, int(self.steps_done == 0), st)
        if repack_first:
            if self.wgrad_stream is not None:
                main.wait_stream(self.wgrad_stream)
        else:
            self.repack_async()
        plan._ran = True

    def _step_graphed(self, x, targets, lr: float):
        """`step(..., graph=True)`: the ~700 launches of a step (three streams, the NCCL all-reduce included) captured
        once per (input shape, world size) and replayed (the learning rate is a device scalar: per-iteration schedules replay) -- no host work per launch, which matters when several
        ranks share few host cores (8 ranks on 16 cores: the eager step loses ~1 ms to launch contention).  The first
        step of a Trainer runs eagerly (it seeds the momentum buffer), inputs are copied into static buffers, and the
        returned loss tensor is the graph's own output buffer (rewritten by every replay)."""
        dev = self.device
        plan = self.plans.get((x.shape[0], x.shape[2], x.shape[3]))
        if self.steps_done == 0 or plan is None or not getattr(plan, "_ran", False):
            return self.step(x, targets, lr=lr, graph=False)   # a shape's first step builds its plan and runs eagerly
        self._check_targets(plan, targets)
        key = (tuple(x.shape), tuple(tuple(t.shape) for t in targets), self.world)   # not the learning rate: a device scalar
        st = self._graphs.get(key)
        with torch.cuda.device(dev):
            if st is None:
                while len(self._graphs) >= self.max_plans:
                    self._graphs.pop(next(iter(self._graphs)))
                gx = torch.empty_like(x)
                gts = [torch.empty_like(t) for t in targets]
                if self._packs_pending and self.wgrad_stream is not None:
                    torch.cuda.current_stream(dev).wait_stream(self.wgrad_stream)
                torch.cuda.synchronize(dev)
                g = torch.cuda.CUDAGraph()
                # thread_local: NCCL's watchdog thread queries events while this thread captures
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    self._step_body(gx, gts, lr, repack_first=True)
                    losses = self.losses
                st = dict(g=g, gx=gx, gts=gts, losses=losses)
                self._graphs[key] = st
            self._lr_dev.fill_(lr)
            st["gx"].copy_(x, non_blocking=True)
            for d, t in zip(st["gts"], targets):
                d.copy_(t, non_blocking=True)
            st["g"].replay()
        self.losses = st["losses"]
        self._packs_pending = False
        self._packed_sig = None   # an eager step after this one repacks first (the graph packs at its own start)
        self.steps_done += 1
        eng = self.model.__dict__.get("_yb_engine")
        if eng is not None:
            eng._sig = None
        return self.losses

    # ---- optimizer state (utils.py:383-416 save_checkpoint / load_checkpoint store {state_dict, optimizer}) --------
    def state_dict(self) -> dict:
        """A torch.optim.SGD-compatible state dict: one param group over the trainable parameters in module order,
        per-parameter `momentum_buffer` copies of the flat momentum buffer (absent before the first step, as in
        torch.optim.SGD)."""
        trainable = [p for p in self.all_params if p.requires_grad]
        state = {}
        if self.steps_done > 0 and self.momentum != 0.0:
            for i, p in enumerate(trainable):
                o = self.param_off[id(p)]
                state[i] = {"momentum_buffer": self.flat_m[o:o + p.numel()].view_as(p).clone()}
        group = {"lr": self.lr, "momentum": self.momentum, "dampening": 0, "weight_decay": self.weight_decay,
                 "nesterov": False, "maximize": False, "foreach": None, "differentiable": False, "fused": None,
                 "params": list(range(len(trainable)))}
        return {"state": state, "param_groups": [group], "steps_done": self.steps_done}

    def load_state_dict(self, sd: dict):
        """Accepts what state_dict() emits or a torch.optim.SGD state dict of the same parameters (same order)."""
        if not self.own_params:
            raise YoloB200Error("load_state_dict needs own_params=True (the fused trainer owns the optimizer state)")
        trainable = [p for p in self.all_params if p.requires_grad]
        groups = sd.get("param_groups", [])
        idx = [i for g in groups for i in g["params"]]
        if len(idx) != len(trainable):
            raise YoloB200Error(f"optimizer state holds {len(idx)} parameters, the model has {len(trainable)} trainable ones")
        if groups:
            g = groups[0]
            self.lr, self.momentum = float(g.get("lr", self.lr)), float(g.get("momentum", self.momentum))
            self.weight_decay = float(g.get("weight_decay", self.weight_decay))
        state = sd.get("state", {})
        have = 0
        self.flat_m.zero_()
        for pos, (key, p) in enumerate(zip(idx, trainable)):
            ent = state.get(key, state.get(str(key)))
            buf = None if ent is None else ent.get("momentum_buffer")
            if buf is None:
                continue
            if tuple(buf.shape) != tuple(p.shape):
                raise YoloB200Error(f"momentum buffer {pos}: shape {tuple(buf.shape)} does not match {tuple(p.shape)}")
            o = self.param_off[id(p)]
            self.flat_m[o:o + p.numel()].view_as(p).copy_(buf.to(device=self.device, dtype=torch.float32))
            have += 1
        # torch.optim.SGD seeds the buffer with the first gradient; after a restore with buffers the next step must NOT
        self.steps_done = int(sd.get("steps_done", 1 if have else 0))
        if have and self.steps_done == 0:
            self.steps_done = 1

    def launches_per_step(self, plan: TrainPlan) -> int:
        n_bn = sum(1 for op in plan.ops if not op.head)
        n_head = len(plan.heads)
        fwd = 1 + len(plan.ops) + 2 * n_bn
        loss = 2 * n_head
        n_s2 = sum(1 for op in plan.ops[1:] if op.pc.stride_eff == 2)
        bwd = 2 * n_bn + 2 * n_head + 2 * len(plan.ops) + (len(plan.ops) - 1) + n_s2
        upd = 1 + len(plan.ops) + n_head + 2 * n_s2
        return fwd + loss + bwd + upd
