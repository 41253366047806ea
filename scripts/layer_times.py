"""Per-launch timing of the forward plan (CUDA events, eager launches, L2 flushed between reps).
Writes a table to stdout: one row per conv launch + the non-conv stages.  Dev tool, GPU box only."""
import argparse
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_for_turbines_b200 import config as cfg  # noqa: E402
from yolo_for_turbines_b200._lib import lib, ptr, stream_ptr  # noqa: E402
from yolo_for_turbines_b200.model import YOLOv3  # noqa: E402
from yolo_for_turbines_b200.utils import Detector  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--size", type=int, default=416)
ap.add_argument("--classes", type=int, default=80)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--block-n", type=int, default=0)
ap.add_argument("--stages", type=int, default=0)
ap.add_argument("--impl", type=int, default=0)
ap.add_argument("--pair", type=int, default=0)
ap.add_argument("--conf", type=float, default=0.5)
ap.add_argument("--no-pdl", action="store_true")
ap.add_argument("--no-split", action="store_true")
ap.add_argument("--no-row", action="store_true")
ap.add_argument("--mc", action="store_true", help="weight-tile multicast across two CTA pairs (opt-in)")
ap.add_argument("--warm", action="store_true", help="no L2 flush: run the producing layer right before the timed one (in-graph cache state)")
args = ap.parse_args()

dev = torch.device("cuda", 0)
torch.manual_seed(0)
m = YOLOv3(num_classes=args.classes).eval().to(dev)
eng = m._engine(dev)
eng.block_n_hint, eng.stages_hint = args.block_n, args.stages
eng.impl_hint, eng.cta_pair_hint = args.impl, args.pair
eng.pdl_hint, eng.tail_split_hint, eng.row_hint = int(args.no_pdl), int(args.no_split), int(args.no_row)
eng.mc_hint = 2 if args.mc else 0
x = torch.rand(args.batch, 3, args.size, args.size, device=dev)
det = Detector(m, cfg.ANCHORS, 0.45, args.conf, "center")
res, plan = det(x)
plan.check_status()
plan.run(x)   # model.forward's path: conv-only graph with fp32 heads (the Detector fuses the decode into the head convs)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
st, sp = stream_ptr(dev), ptr(plan.status)


def timed(fn, reps=args.reps, do_flush=True):
    ts = []
    for _ in range(reps):
        if do_flush:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


rows, tot_ms, tot_gf = [], 0.0, 0.0
info = (C.c_int32 * 8)()
prev = None
for op in plan.ops:
    pc = op.pc
    lib.yolo_conv_plan_info(op.plan_ptr, info)
    if args.warm:
        ts = []
        for _ in range(args.reps):
            if prev is not None and not (plan.stem_direct and prev is plan.ops[0]):
                lib.yolo_conv_fwd(prev.plan_ptr, sp, st)  # leaves this layer's input in L2 as the graph does
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            if plan.stem_direct and op is plan.ops[0]:
                lib.yolo_conv_fwd_stem(op.plan_ptr, ptr(x), sp, st)
            else:
                lib.yolo_conv_fwd(op.plan_ptr, sp, st)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = sorted(ts)[len(ts) // 2]
    elif plan.stem_direct and op is plan.ops[0]:
        ms = timed(lambda: lib.yolo_conv_fwd_stem(op.plan_ptr, ptr(x), sp, st))
    else:
        ms = timed(lambda: lib.yolo_conv_fwd(op.plan_ptr, sp, st))
    prev = op
    M = args.batch * op.dst.H * op.dst.W // (4 if op.upsample else 1)
    gf = 2.0 * M * pc.c_out * (pc.c_in if not pc.stem else 27) * (pc.ksize ** 2 if not pc.stem else 1) / 1e9
    tot_ms += ms
    tot_gf += gf
    rows.append((op.name, pc.c_in, pc.c_out, pc.ksize, pc.stride, op.src.H, info[0], info[1], info[2], info[7],
                 ms, gf / ms))
print(f"{'layer':34s} {'cin':>5s} {'cout':>5s} k s {'H':>4s} {'BN':>4s} {'KC':>3s} st {'ctas':>6s} {'ms':>8s} {'TFLOP/s':>8s}")
for r in rows:
    print(f"{r[0]:34s} {r[1]:5d} {r[2]:5d} {r[3]} {r[4]} {r[5]:4d} {r[6]:4d} {r[7]:3d} {r[8]:2d} {r[9]:6d} {r[10]:8.4f} {r[11]:8.1f}")
if plan.decode_plans(cfg.ANCHORS):
    for op in plan.ops:
        if op.plan_dec_ptr is not None:
            ms = timed(lambda: lib.yolo_conv_fwd(op.plan_dec_ptr, sp, st))
            print(f"{op.name + ' +decode':34s} head conv with the anchor decode in its epilogue: {ms:8.4f} ms")
print(f"conv total (isolated, L2 flushed): {tot_ms:.3f} ms  {tot_gf:.1f} GFLOP  {tot_gf / tot_ms:.1f} TFLOP/s")
ms_in = timed(lambda: plan._launch_input(x))
print(f"input patchify: {ms_in:.4f} ms  ({x.numel() * 4 / 1e6:.0f} MB in, {args.batch * args.size ** 2 * 64 / 1e6:.0f} MB out)")
ms_graph = timed(lambda: (plan._launch_input(x) if plan.stem_direct else None, plan.graph.replay()), do_flush=False)
print(f"conv graph replay (back to back): {ms_graph:.3f} ms -> {tot_gf / ms_graph:.1f} TFLOP/s, {args.batch / ms_graph * 1e3:.0f} img/s conv-only")
from yolo_for_turbines_b200.utils import batched_nms, decode_boxes_multi  # noqa: E402
heads = plan.head_views()
stt = det._get_state(args.batch, [h.shape[2] for h in heads], dev)


def dec():   # the three stored heads in one launch (yolo_decode_multi)
    decode_boxes_multi(heads, [torch.tensor(cfg.ANCHORS[i]) * h.shape[2] for i, h in enumerate(heads)], stt["cand"])


ms_dec = timed(dec)
n = stt["cand"].shape[1]
print(f"decode (3 scales, one launch): {ms_dec:.4f} ms  ({args.batch * n} candidates, {args.batch * n * ((5 + args.classes) * 4 + 24) / ms_dec / 1e6:.1f} GB/s algorithmic)")
ms_nms = timed(lambda: batched_nms(stt["cand"].view(-1, 6), stt["off"], 0.45, args.conf, "center", workspace=stt["ws"]))
kept = int(stt["ws"].keep_off[-1].item())
print(f"nms pipeline: {ms_nms:.4f} ms  ({args.batch * n / ms_nms / 1e3:.2f} M candidates/s, kept {kept})")
ms_all = timed(lambda: det(x), do_flush=False)
print(f"detect() end to end on device: {ms_all:.3f} ms -> {args.batch / ms_all * 1e3:.0f} img/s")
