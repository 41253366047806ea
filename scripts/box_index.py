"""Speed index of the GPU box this call landed on (boxes of the pool differ by ~10 % in sustained clocks): burst cuBLAS
bf16 GEMM, a device copy, and the co-resident cluster counts of the conv kernel.  Printed beside every A/B timing."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_for_turbines_b200._lib import lib  # noqa: E402

dev = torch.device("cuda", 0)
a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
b = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
best = 1e9
for _ in range(12):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); torch.matmul(a, b); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
src = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
dst = torch.empty_like(src)
bc = 1e9
for _ in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); dst.copy_(src); e1.record(); torch.cuda.synchronize()
    bc = min(bc, e0.elapsed_time(e1))
out = C.c_int(0)
cl = []
for cs in (1, 2, 4, 8):
    lib.yolo_conv_max_clusters(cs, C.byref(out))
    cl.append(f"{cs}:{out.value}")
print(f"box index: cuBLAS bf16 8192^3 {2 * 8192 ** 3 / best / 1e9:.0f} TFLOP/s burst, copy {2 * (1 << 30) / bc / 1e6:.0f} GB/s, "
      f"max co-resident clusters by size {' '.join(cl)}")
