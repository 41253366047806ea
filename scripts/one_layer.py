"""Runs ONE conv layer a few times (for ncu captures): python scripts/one_layer.py B H Cin Cout k stride [pair] [bn]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from yolo_for_turbines_b200._lib import ConvDesc, lib, ptr, stream_ptr  # noqa: E402
from yolo_for_turbines_b200.engine import make_conv_plan  # noqa: E402

B, H, cin, cout, k, stride = (int(v) for v in sys.argv[1:7])
pair = int(sys.argv[7]) if len(sys.argv) > 7 else 0
bn = int(sys.argv[8]) if len(sys.argv) > 8 else 0
res = int(sys.argv[9]) if len(sys.argv) > 9 else 0
pad = 1 if k == 3 else 0
dev = "cuda"
x = torch.randn(B, H, H, cin, device=dev).bfloat16()
w = torch.randn(cout, k * k, cin, device=dev).bfloat16()
sc, bi = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
Ho = (H + 2 * pad - k) // stride + 1
y = torch.empty(B, Ho, Ho, cout, dtype=torch.bfloat16, device=dev)
r = torch.randn(B, Ho, Ho, cout, device=dev).bfloat16() if res else None
st = torch.zeros(1, dtype=torch.int32, device=dev)
d = ConvDesc()
d.batch, d.h_in, d.w_in, d.c_in, d.in_pitch, d.c_out, d.c_out_pad, d.out_pitch = B, H, H, cin, cin, cout, cout, cout
d.ksize, d.stride, d.pad, d.act, d.cta_pair_hint, d.block_n_hint = k, stride, pad, 1, pair, bn
d.has_residual, d.res_pitch = res, cout
plan = make_conv_plan(d, ptr(x), ptr(w), ptr(sc), ptr(bi), ptr(r), ptr(y))
for _ in range(5):
    lib.yolo_conv_fwd(plan[1], ptr(st), stream_ptr())
torch.cuda.synchronize()
print("ok")
