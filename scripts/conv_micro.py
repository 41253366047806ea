"""Micro-benchmark of single conv launches (CUDA events, L2 flushed): compares A-operand TMA modes and tile configs."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_for_turbines_b200._lib import ConvDesc, lib, ptr, stream_ptr  # noqa: E402
from yolo_for_turbines_b200.engine import make_conv_plan  # noqa: E402

dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def bench(B, H, cin, cout, k, stride, a_mode=0, impl=2, pair=1, bn=0, stages=0, reps=7):
    pad = 1 if k == 3 else 0
    x = torch.randn(B, H, H, cin, device=dev).bfloat16()
    w = torch.randn(cout, k * k, cin, device=dev).bfloat16()
    sc, bi = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
    Ho = (H + 2 * pad - k) // stride + 1
    y = torch.empty(B, Ho, Ho, cout, dtype=torch.bfloat16, device=dev)
    st = torch.zeros(1, dtype=torch.int32, device=dev)
    d = ConvDesc()
    d.batch, d.h_in, d.w_in, d.c_in, d.in_pitch, d.c_out, d.c_out_pad, d.out_pitch = B, H, H, cin, cin, cout, cout, cout
    d.ksize, d.stride, d.pad, d.act = k, stride, pad, 1
    d.a_mode, d.block_n_hint, d.stages_hint, d.impl_hint, d.cta_pair_hint = a_mode, bn, stages, impl, pair
    plan = make_conv_plan(d, ptr(x), ptr(w), ptr(sc), ptr(bi), None, ptr(y))
    info = (C.c_int32 * 8)()
    lib.yolo_conv_plan_info(plan[1], info)
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        lib.yolo_conv_fwd(plan[1], ptr(st), stream_ptr())
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    ms = ts[len(ts) // 2]
    gf = 2.0 * B * Ho * Ho * cout * cin * k * k / 1e9
    gb = (x.numel() + y.numel()) * 2 / 1e9
    print(f"B{B} H{H} {cin:4d}->{cout:4d} k{k} s{stride} a_mode {a_mode} impl {impl} pair {pair} bn {info[0]:3d} kc {info[1]} "
          f"st {info[2]} ctas {info[7]:4d}: {ms:7.4f} ms {gf / ms:7.1f} TFLOP/s {gb / ms * 1e3:7.1f} GB/s(min traffic)")


for args in sys.argv[1:] or ["all"]:
    pass
print("--- 1x1 tiled vs im2col A loads")
for pair in (1, 2):
    for a_mode in (1, 2):
        bench(64, 52, 256, 256, 1, 1, a_mode=a_mode, pair=pair)
        bench(64, 26, 512, 512, 1, 1, a_mode=a_mode, pair=pair)
        bench(64, 208, 64, 64, 1, 1, a_mode=a_mode, pair=pair)
print("--- 3x3")
for pair in (1, 2):
    bench(64, 26, 256, 512, 3, 1, pair=pair)
    bench(64, 52, 128, 256, 3, 1, pair=pair)
    bench(64, 208, 32, 64, 3, 1, pair=pair)
    bench(64, 104, 64, 128, 3, 1, pair=pair)
print("--- stages / tile sweeps on 256->512 3x3 @26")
for pair, bn, stg in ((2, 256, 3), (2, 256, 4), (2, 256, 5), (2, 128, 6), (1, 256, 3), (1, 128, 5)):
    bench(64, 26, 256, 512, 3, 1, pair=pair, bn=bn, stages=stg)
