"""Data-parallel training check, run under torchrun with N >= 2 ranks (one per GPU, NCCL):
  1. the bucketed, overlapped all-reduce leaves in flat_g exactly the SUM of the ranks' local gradients (captured
     per bucket right before the exchange), every element is covered by exactly one bucket, and the local gradients
     agree with a second, non-distributed Trainer on the same weights and shard;
  2. after the steps every rank holds bit-identical parameters.
Prints one line per check on rank 0 and exits non-zero on failure."""
import copy
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import yolo_oracle as orc  # noqa: E402  (synthetic targets only)
from yolo_for_turbines_b200.model import YOLOv3  # noqa: E402
from yolo_for_turbines_b200.train import Trainer  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B, S = 4, 128
torch.manual_seed(0)
m = YOLOv3(num_classes=2, activation="mish").to(dev).train()
m_local = copy.deepcopy(m)
x = torch.rand(B, 3, S, S, generator=torch.Generator().manual_seed(10 + rank)).to(dev)
tg = [t.to(dev) for t in orc.synth_targets(B, S, 2, 20 + rank)]

tr = Trainer(m, orc.TURBINE_ANCHORS, lr=1e-3, momentum=0.9, weight_decay=5e-4, bucket_mb=8.0)
assert tr.world == world
# (1) exactness of the exchange: every bucket's pre-all-reduce content is captured on the communication stream; summing
#     those captures over the ranks must reproduce flat_g bit for bit, and every element must have been covered.
tr.debug_local_grads = torch.full_like(tr.flat_g, float("nan"))
tr.step(x, tg)
torch.cuda.synchronize()
local = tr.debug_local_grads.clone()
tr.debug_local_grads = None
covered = bool(torch.isfinite(local[: tr.n_trainable]).all())
expect = local[: tr.n_trainable].clone()
dist.all_reduce(expect)
ok1 = covered and bool(torch.equal(expect, tr.flat_g[: tr.n_trainable]))
err = float((tr.flat_g[: tr.n_trainable] - expect).abs().max())
# (1b) the local gradients are what a non-distributed trainer computes on the same shard (up to the run-to-run
#      noise of fp32 atomics that the network amplifies: cosine, not equality)
tr_local = Trainer(m_local, orc.TURBINE_ANCHORS, lr=1e-3, momentum=0.9, weight_decay=5e-4, data_parallel=False)
assert tr_local.world == 1
tr_local.step(x, tg)
torch.cuda.synchronize()
cos = float(torch.nn.functional.cosine_similarity(local[: tr.n_trainable], tr_local.flat_g[: tr.n_trainable], dim=0))
ok1 = ok1 and cos > 0.97
n_buckets = len(tr._buckets(tr.plan(B, S, S)))
for _ in range(2):
    tr.step(x, tg)
torch.cuda.synchronize()
ref = tr.flat_p.clone()
dist.broadcast(ref, 0)
ok2 = bool(torch.equal(ref, tr.flat_p))
flags = torch.tensor([int(ok1), int(ok2)], device=dev)
dist.all_reduce(flags, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"dp{world}: all-reduced gradients == sum of the ranks' local gradients: max abs err {err:.1e}, every element covered "
          f"({n_buckets} buckets); local vs single-GPU trainer cosine {cos:.4f} -> {'ok' if flags[0] else 'FAIL'}")
    print(f"dp{world}: parameters identical on all ranks after 3 steps -> {'ok' if flags[1] else 'FAIL'}")
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if int(flags.min()) == 1 else 1)
