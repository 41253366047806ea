"""Times yolo_wgrad on the main YOLOv3-416 layer shapes (batch 32) for several split-K factors."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_for_turbines_b200._lib import ConvDesc, lib, ptr, stream_ptr  # noqa: E402
from yolo_for_turbines_b200.engine import _aligned_blob  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda", 0)
shapes = [(32, 64, 3, 2, 416), (32, 64, 3, 1, 208), (64, 128, 3, 1, 104), (128, 256, 3, 1, 52), (256, 512, 3, 1, 26), (512, 1024, 3, 1, 13),
          (256, 128, 1, 1, 52), (512, 256, 1, 1, 26), (1024, 512, 1, 1, 13), (128, 256, 3, 2, 104)]
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
for cin, cout, k, s, H in shapes:
    pad = 1 if k == 3 else 0
    Ho = (H + 2 * pad - k) // s + 1
    x = torch.randn(B, H, H, cin, device=dev).bfloat16()
    dz = torch.randn(B, Ho, Ho, cout, device=dev).bfloat16()
    dw = torch.zeros(cout, k * k, cin, dtype=torch.float32, device=dev)
    d = ConvDesc()
    d.batch, d.h_in, d.w_in, d.c_in, d.in_pitch = B, H, H, cin, cin
    d.c_out, d.c_out_pad, d.out_pitch = cout, cout, cout
    d.ksize, d.stride, d.pad = k, s, pad
    gf = 2.0 * B * Ho * Ho * cout * cin * k * k / 1e9
    line = f"{cin:5d}->{cout:5d} k{k} s{s} H{H:4d}: {gf:7.1f} GF |"
    for splits in (0, -1, -2, -4):
        raw, plan = _aligned_blob(int(lib.yolo_wgrad_plan_bytes()))
        lib.yolo_wgrad_plan_init(plan, lib.yolo_wgrad_plan_bytes(), C.byref(d), ptr(x), ptr(dz), cout, ptr(dw), 0)
        info = (C.c_int32 * 6)()
        lib.yolo_wgrad_plan_info(plan, info)
        sp = info[2]
        if splits < 0:
            sp = max(1, info[2] // (-splits * 2) * 1 if splits != -1 else info[2] // 2)
            lib.yolo_wgrad_plan_init(plan, lib.yolo_wgrad_plan_bytes(), C.byref(d), ptr(x), ptr(dz), cout, ptr(dw), sp)
            lib.yolo_wgrad_plan_info(plan, info)
        st = stream_ptr(dev)
        ts = []
        for rep in range(5):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            lib.yolo_wgrad(plan, st)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = sorted(ts)[1]
        line += f" nt{info[0] % 1000} tp{info[0] // 1000} st{info[1]} sp{info[2]:3d} ctas{info[5]:4d}: {t * 1e3:7.1f} us {gf / t:6.0f} TF |"
    print(line, flush=True)
