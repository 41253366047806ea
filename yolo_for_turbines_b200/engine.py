"""Host-side executor of the YOLOv3 forward pass on the sm_100a conv kernels.

What the reference does with one ATen/cuDNN call per op (code/model.py:172-193) is planned here
once per (batch, H, W): every CNNBlock becomes ONE fused launch (conv + folded BN + activation
[+ residual]) on NHWC bf16 buffers; nn.Upsample is a 2x2 replicated store of the producing conv
and torch.cat is two producers writing disjoint channel ranges of one buffer (model.py:189-191,
:222), so neither exists as a kernel.  The launch list is captured in a CUDA graph and replayed.
Buffers, packed weights and TMA tensor maps are static per plan; torch owns all memory.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from ._lib import ACT_CODES, STATUS_NAN_INPUT, STATUS_NAN_LAYER, ConvDesc, YoloB200Error, lib, ptr, stream_ptr


def _round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def _act_name(block) -> str:
    if not block.batch_norm_act:
        return "none"
    if isinstance(block.activation, nn.LeakyReLU):
        if abs(block.activation.negative_slope - 0.1) > 1e-12:
            raise YoloB200Error("only LeakyReLU(0.1) is implemented (code/model.py:64)")
        return "leaky_relu"
    if isinstance(block.activation, nn.Mish):
        return "mish"
    raise YoloB200Error(f"unsupported activation module {type(block.activation).__name__}")


class PackedConv:
    """bf16 K-major weights + folded-BN scale/bias of one CNNBlock, refreshed in place so that
    plans and CUDA graphs that hold their addresses stay valid across weight updates."""

    def __init__(self, block, device, as_stem: bool = False, allow_fold: bool = False):
        conv = block.conv
        k, s, p = conv.kernel_size, conv.stride, conv.padding
        if k[0] != k[1] or s[0] != s[1] or p[0] != p[1] or k[0] not in (1, 3) or s[0] not in (1, 2) \
                or p[0] != (1 if k[0] == 3 else 0) or conv.groups != 1 or conv.dilation != (1, 1):
            raise YoloB200Error(f"conv {k}/{s}/{p} is outside the YOLOv3 layer set (1x1/p0 or 3x3/p1, stride 1|2)")
        self.block = block
        self.ksize, self.stride, self.pad = k[0], s[0], p[0]
        self.c_in, self.c_out = conv.in_channels, conv.out_channels
        self.c_out_pad = _round_up(self.c_out, 32)
        self.act = _act_name(block)
        self.stem = bool(as_stem)
        if self.stem:  # runs as a 1x1 conv over the 32-wide patch matrix (yolo_input_patchify)
            assert self.ksize == 3 and self.stride == 1 and 9 * self.c_in <= 32
            self.k_eff, self.stride_eff, self.pad_eff, self.c_in_eff = 1, 1, 0, 32
        else:
            self.k_eff, self.stride_eff, self.pad_eff = self.ksize, self.stride, self.pad
            self.c_in_eff = _round_up(self.c_in, 32)
        # Pixel-pair folding: a 32-channel NHWC pixel is a 64-byte row, which wastes half of every TMA/L2
        # request.  Two horizontally adjacent pixels are contiguous, so (B,H,W,32) IS (B,H,W/2,64): a stride-1
        # conv over pixel pairs with block-structured weights (half of them zero) computes the same outputs
        # from 128-byte rows.  Tensor-core work doubles, which is irrelevant for these HBM/TMA-bound layers.
        self.fold = bool(allow_fold and self.stride_eff == 1 and (self.c_in_eff == 32 or self.c_out == 32)
                         and self.c_out % 32 == 0 and block.batch_norm_act)
        # Stride-2 variant (the 32->64 down-sampling conv): pairs on the INPUT side only.  Output column q reads
        # input columns 2q-1, 2q, 2q+1 = second half of pair q-1 and both halves of pair q: a 3x2 filter over
        # pairs with stride (2, 1) and padding one pair on the left, none on the right.
        self.fold_s2 = bool(allow_fold and not self.fold and self.ksize == 3 and self.stride == 2
                            and self.c_in_eff == 32 and block.batch_norm_act and not self.stem)
        self.c_in_run = self.c_in_eff * (2 if (self.fold or self.fold_s2) else 1)       # what the kernel sees
        self.c_out_run = self.c_out * (2 if self.fold else 1)
        self.c_out_pad_run = self.c_out_pad * (2 if self.fold else 1)
        self.kw_run = 2 if self.fold_s2 else self.k_eff
        # fused stem: the kernel gathers the 3x3x3 taps from the NCHW image itself (needs the folded layout)
        self.stem_direct = bool(self.stem and self.fold and self.c_in == 3 and self.c_out == 32)
        self.w = torch.empty(self.c_out_pad_run * self.k_eff * self.kw_run * self.c_in_run, dtype=torch.bfloat16,
                             device=device)
        self.scale = torch.empty(self.c_out_pad_run, dtype=torch.float32, device=device)
        self.bias = torch.empty(self.c_out_pad_run, dtype=torch.float32, device=device)

    def _refresh_folded(self, w, st):
        b, dev = self.block, self.w.device
        if self.stem:  # K index of the patch matrix is (kh*3+kw)*C + c, zero padded to 32 (yolo_input_patchify)
            wk = torch.zeros(self.c_out, 32, 1, 1, device=dev)
            wk[:, : 9 * self.c_in, 0, 0] = w.permute(0, 2, 3, 1).reshape(self.c_out, 9 * self.c_in)
        else:
            wk = w
            if self.c_in_eff != self.c_in:
                wk = torch.zeros(self.c_out, self.c_in_eff, self.ksize, self.ksize, device=dev)
                wk[:, : self.c_in] = w
        w2 = _fold_weight_pairs(wk).contiguous()
        lib.yolo_pack_weights(ptr(w2), w2.shape[0], w2.shape[1], self.k_eff, self.c_out_pad_run, self.c_in_run,
                              ptr(self.w), st)
        bn = b.batch_norm
        f32 = lambda t: t.detach().to(device=dev, dtype=torch.float32).contiguous()  # noqa: E731
        g, be, mu, var = f32(bn.weight), f32(bn.bias), f32(bn.running_mean), f32(bn.running_var)
        sc = torch.empty(self.c_out_pad, dtype=torch.float32, device=dev)
        bi = torch.empty(self.c_out_pad, dtype=torch.float32, device=dev)
        lib.yolo_fold_bn(ptr(g), ptr(be), ptr(mu), ptr(var), None, float(bn.eps), self.c_out, self.c_out_pad, ptr(sc),
                         ptr(bi), st)
        self.scale.copy_(sc.repeat(2))
        self.bias.copy_(bi.repeat(2))

    def tensors(self):
        b = self.block
        ts = [b.conv.weight]
        if b.conv.bias is not None:
            ts.append(b.conv.bias)
        if b.batch_norm_act:
            ts += [b.batch_norm.weight, b.batch_norm.bias, b.batch_norm.running_mean, b.batch_norm.running_var]
        return ts

    def refresh(self):
        b = self.block
        st = stream_ptr(self.w.device)
        f32 = lambda t: t.detach().to(device=self.w.device, dtype=torch.float32).contiguous()  # noqa: E731
        w = f32(b.conv.weight)
        if self.fold:
            self._refresh_folded(w, st)
            return
        if self.fold_s2:
            O, I = self.c_out, self.c_in_eff
            wk = torch.zeros(O, I, 3, 3, device=self.w.device)
            wk[:, : self.c_in] = w
            w2 = torch.zeros(self.c_out_pad, 2 * I, 3, 2, device=self.w.device)
            w2[:O, I:, :, 0] = wk[:, :, :, 0]   # pair q-1, second half  <- tap s = 0 (column 2q-1)
            w2[:O, :I, :, 1] = wk[:, :, :, 1]   # pair q,   first half   <- tap s = 1 (column 2q)
            w2[:O, I:, :, 1] = wk[:, :, :, 2]   # pair q,   second half  <- tap s = 2 (column 2q+1)
            self.w.copy_(w2.permute(0, 2, 3, 1).reshape(-1).to(torch.bfloat16))  # [Cout][kh][kw][Cin']
            bn = b.batch_norm
            g, be, mu, var = f32(bn.weight), f32(bn.bias), f32(bn.running_mean), f32(bn.running_var)
            lib.yolo_fold_bn(ptr(g), ptr(be), ptr(mu), ptr(var), None, float(bn.eps), self.c_out, self.c_out_pad,
                             ptr(self.scale), ptr(self.bias), st)
            return
        if self.stem:
            lib.yolo_pack_stem_weights(ptr(w), self.c_out, self.c_in, self.c_out_pad, ptr(self.w), st)
        else:
            lib.yolo_pack_weights(ptr(w), self.c_out, self.c_in, self.ksize, self.c_out_pad, self.c_in_eff, ptr(self.w), st)
        if b.batch_norm_act:
            bn = b.batch_norm
            g, be, mu, var = f32(bn.weight), f32(bn.bias), f32(bn.running_mean), f32(bn.running_var)
            lib.yolo_fold_bn(ptr(g), ptr(be), ptr(mu), ptr(var), None, float(bn.eps), self.c_out, self.c_out_pad,
                             ptr(self.scale), ptr(self.bias), st)
        else:
            cb = f32(b.conv.bias) if b.conv.bias is not None else None
            lib.yolo_fold_bn(None, None, None, None, ptr(cb), 0.0, self.c_out, self.c_out_pad, ptr(self.scale),
                             ptr(self.bias), st)


def _fold_weight_pairs(w: torch.Tensor) -> torch.Tensor:
    """(O, I, k, k) weights of a stride-1 'same' conv -> (2O, 2I, k, k) weights of the equivalent conv over
    horizontally adjacent pixel PAIRS: output half oh / input half ih of pair-tap s' touch input column offset
    2*(s' - k//2) + ih - oh, which is original tap s = offset + k//2 when that lies inside the filter."""
    O, I, k, _ = w.shape
    c = k // 2
    w2 = torch.zeros(2 * O, 2 * I, k, k, dtype=w.dtype, device=w.device)
    for oh in range(2):
        for ih in range(2):
            for sp in range(k):
                s = 2 * (sp - c) + ih - oh + c
                if 0 <= s < k:
                    w2[oh * O:(oh + 1) * O, ih * I:(ih + 1) * I, :, sp] = w[:, :, :, s]
    return w2


class _Act:
    """Symbolic NHWC activation; `root`/`ch_off` let it live inside a wider (concat) buffer."""
    __slots__ = ("C", "H", "W", "fp32", "root", "ch_off", "first", "last", "buf")

    def __init__(self, C_, H, W, fp32=False):
        self.C, self.H, self.W, self.fp32 = C_, H, W, fp32
        self.root, self.ch_off = None, 0
        self.first, self.last, self.buf = None, None, None

    def resolve(self):
        t, off = self, 0
        while t.root is not None:
            off += t.ch_off
            t = t.root
        return t, off


class _ConvOp:
    __slots__ = ("pc", "src", "dst", "res", "upsample", "check_nan", "plan", "plan_ptr", "name", "desc", "ptrs", "plan_dec",
                 "plan_dec_ptr")

    def __init__(self, pc, src, dst, res=None, check_nan=True, name=""):
        self.pc, self.src, self.dst, self.res = pc, src, dst, res
        self.upsample, self.check_nan, self.name = False, check_nan, name
        self.plan = self.plan_ptr = self.desc = self.ptrs = self.plan_dec = self.plan_dec_ptr = None


def _aligned_blob(nbytes: int, align: int = 64):
    raw = (C.c_uint8 * (nbytes + align))()
    addr = C.addressof(raw)
    off = (-addr) % align
    return raw, C.c_void_p(addr + off)


def make_conv_plan(desc: ConvDesc, x, w, scale, bias, residual, y):
    raw, p = _aligned_blob(int(lib.yolo_conv_plan_bytes()))
    lib.yolo_conv_plan_init(p, lib.yolo_conv_plan_bytes(), C.byref(desc), x, w, scale, bias, residual, y)
    return raw, p


class ForwardPlan:
    """Everything static for one (batch, H, W): buffers, tensor maps, launch list, CUDA graph."""

    def __init__(self, engine: "Engine", B: int, H: int, W: int, use_graph: bool = True):
        from .model import CNNBlock, ResidualBlock, ScalePredictionBlock  # cycle-free at call time

        self.engine, self.B, self.H, self.W = engine, B, H, W
        model, dev = engine.model, engine.device
        if H % 32 or W % 32:
            raise YoloB200Error(f"input size {H}x{W} must be a multiple of 32 (five stride-2 stages)")
        self.ops: List[_ConvOp] = []
        self.heads: List[_Act] = []
        layers = list(model.layers)
        if not layers or not isinstance(layers[0], CNNBlock):
            raise YoloB200Error("first layer must be a CNNBlock")
        first_pc = engine.packed[id(layers[0])]
        cur = _Act(first_pc.c_in_eff, H, W)
        self.input_act, self.stem = cur, first_pc.stem
        routes: List[_Act] = []

        def conv(block, src, res=None, fp32=False, check_nan=True, name=""):
            pc = engine.packed[id(block)]
            ho = (src.H + 2 * pc.pad_eff - pc.k_eff) // pc.stride_eff + 1
            wo = (src.W + 2 * pc.pad_eff - pc.k_eff) // pc.stride_eff + 1
            dst = _Act(pc.c_out_pad if fp32 else pc.c_out, ho, wo, fp32)
            if not fp32 and pc.c_out % 32:
                raise YoloB200Error(f"{name}: intermediate channel count {pc.c_out} must be a multiple of 32")
            if src.C != pc.c_in_eff:
                raise YoloB200Error(f"{name}: expects {pc.c_in_eff} input channels, got {src.C}")
            self.ops.append(_ConvOp(pc, src, dst, res, check_nan, name))
            return dst

        for li, layer in enumerate(layers):
            if isinstance(layer, ScalePredictionBlock):  # model.py:177-179: does not advance x
                t = conv(layer.pred_block[0], cur, check_nan=False, name=f"layers.{li}.pred_block.0")
                self.heads.append(conv(layer.pred_block[1], t, fp32=True, check_nan=False, name=f"layers.{li}.pred_block.1"))
                self.head_meta = getattr(self, "head_meta", []) + [(layer.anchors_per_scale, layer.num_classes)]
            elif isinstance(layer, CNNBlock):
                cur = conv(layer, cur, name=f"layers.{li}")
            elif isinstance(layer, ResidualBlock):
                for ri, seq in enumerate(layer.layers):
                    t = conv(seq[0], cur, name=f"layers.{li}.layers.{ri}.0")
                    cur = conv(seq[1], t, res=cur if layer.use_residual else None, name=f"layers.{li}.layers.{ri}.1")
                if layer.num_blocks == 8:  # model.py:186-187
                    routes.append(cur)
            elif isinstance(layer, nn.Upsample):
                sf = layer.scale_factor
                if float(sf if not isinstance(sf, (tuple, list)) else sf[0]) != 2.0 or layer.mode != "nearest":
                    raise YoloB200Error("only nn.Upsample(scale_factor=2, mode='nearest') is implemented")
                prod = self.ops[-1]
                if prod.dst is not cur or not routes:
                    raise YoloB200Error("Upsample must directly follow a conv and have a pending route")
                route = routes.pop()
                cat = _Act(cur.C + route.C, 2 * cur.H, 2 * cur.W)
                if (route.H, route.W) != (cat.H, cat.W):
                    raise YoloB200Error("route / upsample size mismatch")
                up = _Act(cur.C, cat.H, cat.W)
                up.root, up.ch_off = cat, 0              # model.py:190: upsampled channels first
                route.root, route.ch_off = cat, cur.C
                prod.dst, prod.upsample = up, True
                cur = cat
            else:
                raise YoloB200Error(f"unsupported layer type {type(layer).__name__}")

        # NaNs never vanish downstream (conv, leaky/mish, residual add and bf16 rounding all propagate them), so
        # "any top-level layer output holds a NaN" (model.py:183) == "the LAST top-level layer output holds
        # one": only that launch pays for the check.  ScalePredictionBlock outputs are unchecked, as upstream.
        chain = [op for op in self.ops if op.check_nan]
        for op in chain[:-1]:
            op.check_nan = False

        # ---- liveness on root buffers, then greedy reuse -------------------------------------
        self.input_root = self.input_act.resolve()[0]
        self.input_root.first = -1
        for i, op in enumerate(self.ops):
            for t in (op.src, op.dst, op.res):
                if t is None:
                    continue
                r = t.resolve()[0]
                r.first = i if r.first is None else r.first
                r.last = i
        for h in self.heads:
            h.resolve()[0].last = len(self.ops) + 1  # head buffers outlive the forward
        pool = []  # [tensor(uint8), free_from_op]
        self.total_bytes = 0
        roots = sorted({id(t.resolve()[0]): t.resolve()[0] for op in self.ops for t in (op.src, op.dst, op.res) if t}.values(),
                       key=lambda r: r.first)
        self.stem_direct = bool(first_pc.stem_direct and engine.stem_direct)
        for r in roots:
            need = B * r.H * r.W * r.C * (4 if r.fp32 else 2)
            if self.stem_direct and r is self.input_root:
                need = 256  # no patch matrix: the stem kernel reads the image directly
            best = None
            for ent in pool:
                if ent[1] <= r.first and ent[0].numel() >= need and (best is None or ent[0].numel() < best[0].numel()):
                    best = ent
            if best is None:
                best = [torch.empty(_round_up(need, 256), dtype=torch.uint8, device=dev), 0]
                pool.append(best)
                self.total_bytes += best[0].numel()
            best[1] = r.last + 1
            r.buf = best[0]
        self.pool = pool
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)

        # ---- plans ---------------------------------------------------------------------------
        for op in self.ops:
            pc = op.pc
            sroot, soff = op.src.resolve()
            droot, doff = op.dst.resolve()
            d = ConvDesc()
            f = 1
            if pc.fold:  # run on pixel pairs: needs dense tensors (a pair must be one contiguous row)
                ok = (op.src.W % 2 == 0 and sroot.C == op.src.C and droot.C == op.dst.C and not op.upsample
                      and not op.dst.fp32 and (op.res is None or op.res.resolve()[0].C == op.res.C))
                if not ok:
                    raise YoloB200Error(f"{op.name}: pixel-pair folding needs dense, even-width tensors")
                f = 2
            fi = f
            if pc.fold_s2:
                if op.src.W % 2 or sroot.C != op.src.C:
                    raise YoloB200Error(f"{op.name}: pixel-pair folding needs a dense, even-width input")
                fi = 2
                d.ksize_w, d.stride_w, d.pad_w_hi_plus1 = 2, 1, 1  # 3x2 filter, stride (2,1), no right padding
            d.batch, d.h_in, d.w_in, d.c_in, d.in_pitch = B, op.src.H, op.src.W // fi, pc.c_in_run, sroot.C * fi
            d.c_out, d.c_out_pad, d.out_pitch = pc.c_out_run, pc.c_out_pad_run, droot.C * f
            d.ksize, d.stride, d.pad = pc.k_eff, pc.stride_eff, pc.pad_eff
            d.act = ACT_CODES[pc.act]
            d.upsample2x, d.out_fp32, d.check_nan = int(op.upsample), int(op.dst.fp32), int(op.check_nan)
            d.a_mode, d.block_n_hint, d.stages_hint = 0, engine.block_n_hint, engine.stages_hint
            d.impl_hint, d.cta_pair_hint = engine.impl_hint, engine.cta_pair_hint
            d.pdl_hint, d.tail_split_hint, d.row_hint = engine.pdl_hint, engine.tail_split_hint, engine.row_hint
            d.mc_hint = engine.mc_hint
            x_ptr = sroot.buf.data_ptr() + soff * 2
            if self.stem_direct and op is self.ops[0]:
                d.stem_c, x_ptr = 3, 0
            y_ptr = droot.buf.data_ptr() + doff * (4 if op.dst.fp32 else 2)
            r_ptr = None
            if op.res is not None:
                rroot, roff = op.res.resolve()
                d.has_residual, d.res_pitch = 1, rroot.C * f
                r_ptr = C.c_void_p(rroot.buf.data_ptr() + roff * 2)
            op.desc, op.ptrs = d, (C.c_void_p(x_ptr), ptr(pc.w), ptr(pc.scale), ptr(pc.bias), r_ptr)
            op.plan, op.plan_ptr = make_conv_plan(d, C.c_void_p(x_ptr), ptr(pc.w), ptr(pc.scale), ptr(pc.bias), r_ptr,
                                                  C.c_void_p(y_ptr))
        self.graph: Optional[torch.cuda.CUDAGraph] = None
        self.use_graph = use_graph
        self.cand: Optional[torch.Tensor] = None      # decode_plans(): (B, sum 3 S^2, 6) candidate rows written by the head convs
        self._dec_key = None
        self.launches_per_forward = len(self.ops) + (0 if self.stem_direct else 1)

    # ------------------------------------------------------------------------------------------
    def head_views(self):
        """(B,3,S,S,5+nc) fp32 views of the NHWC head buffers -- the same non-contiguous layout the
        reference's reshape+permute yields (model.py:147-148)."""
        outs = []
        for h, (na, nc) in zip(self.heads, self.head_meta):
            root, _ = h.resolve()
            ch = nc + 5
            flat = root.buf[: self.B * h.H * h.W * root.C * 4].view(torch.float32)
            outs.append(torch.as_strided(flat, (self.B, na, h.H, h.W, ch),
                                         (h.H * h.W * root.C, ch, h.W * root.C, root.C, 1)))
        return outs

    def decode_plans(self, anchors) -> bool:
        """Head convs with the anchor decode in their epilogue (yolo_conv_desc.decode_mode): builds, once per anchor
        set, a candidate tensor `self.cand` (B, 3*(S1^2+S2^2+S3^2), 6) in the reference's order (utils.py:300-309) and a
        second plan per head conv that writes it.  Returns False when the heads do not fit the fused form (more than
        80 classes: 3*(5+nc) must fit one 256-column tile), in which case the caller decodes the stored heads."""
        import struct

        key = tuple(float(v) for s_ in anchors for a in s_ for v in a)
        if self._dec_key == key:
            return self.cand is not None
        self._dec_key, self.cand = key, None
        head_ops = [op for op in self.ops if op.dst.fp32]
        if len(head_ops) != len(self.head_meta) or len(anchors) < len(head_ops):
            return False
        for op, (na, nc) in zip(head_ops, self.head_meta):
            if na != 3 or op.pc.c_out_pad not in (32, 256) or 3 * (5 + nc) > op.pc.c_out_pad or op.dst.H != op.dst.W:
                return False
        n = sum(3 * op.dst.H * op.dst.W for op in head_ops)
        with torch.cuda.device(self.engine.device):
            cand = torch.zeros(self.B, n, 6, dtype=torch.float32, device=self.engine.device)
            off = 0
            for i, (op, (na, nc)) in enumerate(zip(head_ops, self.head_meta)):
                S = op.dst.H
                d = ConvDesc.from_buffer_copy(op.desc)
                d.decode_mode, d.dec_nc, d.dec_rows_per_image, d.dec_row_offset = 1, nc, n, off
                a = anchors[i] if torch.is_tensor(anchors) else torch.tensor([*anchors[i]])
                scaled = (a.detach().to(device="cpu", dtype=torch.float32) * S).reshape(-1).tolist()   # utils.py:303, fp32
                for k in range(6):
                    d.dec_anchor_bits[k] = struct.unpack("<i", struct.pack("<f", scaled[k]))[0]
                x_ptr, w, sc, bi, _ = op.ptrs
                op.plan_dec, op.plan_dec_ptr = make_conv_plan(d, x_ptr, w, sc, bi, None, ptr(cand))
                off += 3 * S * S
        self.cand = cand
        return True

    def _launch_convs(self, decode: bool = False):
        """decode=True: the head convs write candidate rows into self.cand (decode_plans) instead of fp32 heads."""
        st = stream_ptr(self.engine.device)
        sp = ptr(self.status)
        for op in (self.ops[1:] if self.stem_direct else self.ops):
            lib.yolo_conv_fwd(op.plan_dec_ptr if (decode and op.plan_dec_ptr is not None) else op.plan_ptr, sp, st)

    def _launch_input(self, x):
        """Everything that reads the caller's tensor (its address changes per call, so it stays out of the CUDA
        graph): the fused stem conv, or the patchify / NHWC conversion of the non-fused paths."""
        st = stream_ptr(self.engine.device)
        self.status.zero_()
        dst = ptr(self.input_root.buf)
        if self.stem_direct:
            lib.yolo_conv_fwd_stem(self.ops[0].plan_ptr, ptr(x), ptr(self.status), st)
        elif self.stem:
            lib.yolo_input_patchify(ptr(x), self.B, x.shape[1], self.H, self.W, dst, ptr(self.status), st)
        else:
            lib.yolo_nchw_to_nhwc_bf16(ptr(x), self.B, x.shape[1], self.H, self.W, self.input_act.C, self.input_root.C,
                                       dst, ptr(self.status), st)

    def run(self, x: torch.Tensor):
        """Enqueues the forward on the current stream; no host sync."""
        self._launch_input(x)
        if not self.use_graph:
            self._launch_convs()
            return
        if self.graph is None:
            self._launch_convs()  # warm-up: sets the dynamic-smem attributes outside capture
            torch.cuda.current_stream(self.engine.device).synchronize()
            self._launch_input(x)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._launch_convs()
            self.graph = g
        self.graph.replay()

    def check_status(self):
        if getattr(self, "done_event", None) is not None:
            self.done_event.synchronize()  # the forward ran on a Detector lane stream
        s = int(self.status.item())  # the ONE host sync of a forward (the reference has 28)
        if s & STATUS_NAN_INPUT:
            raise AssertionError("NaN in the input tensor")  # model.py:175
        if s & STATUS_NAN_LAYER:
            raise ValueError("Nan in layer")  # model.py:184


class Engine:
    """Per-model cache: packed weights (refreshed when parameters change) + plans per input shape."""

    def __init__(self, model, device):
        from .model import CNNBlock

        self.model, self.device = model, torch.device(device)
        self.block_n_hint, self.stages_hint = 0, 0
        self.impl_hint, self.cta_pair_hint = 0, 0  # 0 = library defaults (include/yolo_b200.h)
        self.pdl_hint, self.tail_split_hint = 0, 0   # 1 switches the feature off (A/B runs, scripts/layer_times.py)
        self.row_hint = 0                            # include/yolo_b200.h: 0 auto | 1 off | 2 on, base_offset variant
        # weight-tile multicast across two CTA pairs: 2 = on (measured slower in isolation; YOLO_B200_MC=2 for A/B runs)
        self.mc_hint = int(os.environ.get("YOLO_B200_MC", "0"))
        # fused stem (default): the first conv reads the NCHW fp32 image itself (TMA windows -> bf16 taps in shared memory),
        # no patch matrix in HBM.  YOLO_B200_FUSED_STEM=0 restores yolo_input_patchify + a K=64 GEMM (the A/B baseline).
        self.stem_direct = os.environ.get("YOLO_B200_FUSED_STEM") != "0"
        self.allow_fold = hasattr(model, "layers") and hasattr(model, "num_classes") and os.environ.get("YOLO_B200_NO_FOLD") != "1"
        self.packed: Dict[int, PackedConv] = {}
        blocks = [m for m in model.modules() if isinstance(m, CNNBlock)]
        first = model.layers[0] if hasattr(model, "layers") and len(model.layers) else None
        with torch.cuda.device(self.device):
            for b in blocks:
                stem = b is first and b.conv.kernel_size == (3, 3) and b.conv.stride == (1, 1) and 9 * b.conv.in_channels <= 32
                self.packed[id(b)] = PackedConv(b, self.device, as_stem=stem, allow_fold=self.allow_fold)
        self.plans: Dict[tuple, ForwardPlan] = {}
        self._sig = None

    def _signature(self):
        return tuple((t._version, t.data_ptr()) for pc in self.packed.values() for t in pc.tensors())

    def refresh_if_needed(self):
        sig = self._signature()
        if sig != self._sig:
            with torch.cuda.device(self.device):
                for pc in self.packed.values():
                    pc.refresh()
            self._sig = sig

    def plan(self, B, H, W, lane: int = 0) -> ForwardPlan:
        """`lane` > 0 gives an independent plan (own buffers, own CUDA graph) for the same shape, so that several
        batches can be in flight on different streams (utils.Detector(lanes=2))."""
        key = (B, H, W) if lane == 0 else (B, H, W, lane)
        p = self.plans.get(key)
        if p is None:
            with torch.cuda.device(self.device):
                p = ForwardPlan(self, B, H, W)
            self.plans[key] = p
        return p
