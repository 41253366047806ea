"""Runs ONE weight-gradient layer a few times (for ncu captures): python scripts/one_wgrad.py B H Cin Cout k stride"""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_for_turbines_b200._lib import ConvDesc, lib, ptr, stream_ptr  # noqa: E402
from yolo_for_turbines_b200.engine import _aligned_blob  # noqa: E402

B, H, cin, cout, k, s = (int(v) for v in sys.argv[1:7])
dev = torch.device("cuda", 0)
pad = 1 if k == 3 else 0
Ho = (H + 2 * pad - k) // s + 1
x = torch.randn(B, H, H, cin, device=dev).bfloat16()
dz = torch.randn(B, Ho, Ho, cout, device=dev).bfloat16()
dw = torch.zeros(cout, k * k, cin, dtype=torch.float32, device=dev)
d = ConvDesc()
d.batch, d.h_in, d.w_in, d.c_in, d.in_pitch = B, H, H, cin, cin
d.c_out, d.c_out_pad, d.out_pitch = cout, cout, cout
d.ksize, d.stride, d.pad = k, s, pad
raw, plan = _aligned_blob(int(lib.yolo_wgrad_plan_bytes()))
lib.yolo_wgrad_plan_init(plan, lib.yolo_wgrad_plan_bytes(), C.byref(d), ptr(x), ptr(dz), cout, ptr(dw), 0)
for _ in range(5):
    lib.yolo_wgrad(plan, stream_ptr(dev))
torch.cuda.synchronize()
print("ok")
