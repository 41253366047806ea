"""BASELINE config 5: standalone NMS (+ mAP matching) sweep, 10^4..10^6 candidate boxes per batch of 64 images.
GPU: yolo_nms through utils.batched_nms (CUDA events).  CPU: the reference algorithm (oracle port, torch ops per kept box)
timed per image on a subset and scaled to the batch (labelled extrapolated), plus the C restatement for context."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import yolo_oracle as orc  # noqa: E402  (CPU baseline leg only)
from yolo_for_turbines_b200.utils import NmsWorkspace, batched_nms, map_match  # noqa: E402

B = 64
WORLD, RANK, LOCAL = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
torch.cuda.set_device(LOCAL)
dev = torch.device("cuda", LOCAL)
dist = None
if WORLD > 1:   # torchrun: every rank sweeps its own seeded boxes (weak scaling, no collective on the data path);
    import torch.distributed as dist   # NCCL only carries the barrier and the max-over-ranks of the device times
    _fd = os.dup(1)
    os.dup2(2, 1)   # NCCL's banner goes to stderr
    dist.init_process_group("nccl", device_id=dev)
    dist.barrier()
    sys.stdout.flush()
    os.dup2(_fd, 1)


def max_over_ranks(ms):
    if dist is None:
        return ms
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def say(*a, **k):
    if RANK == 0:
        print(*a, **k)


def make(total, nc, seed=42):
    g = torch.Generator(device=dev).manual_seed(seed + RANK)
    b = torch.rand(total, 6, generator=g, device=dev)
    b[:, 2:4] = 0.02 + 0.28 * b[:, 2:4]
    b[:, 5] = torch.floor(b[:, 5] * nc)
    n_tie = total // 100
    b[torch.randperm(total, generator=g, device=dev)[:n_tie], 4] = 0.7311  # 1 % tie-score subset
    return b


def gpu_time(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if dist is not None:
            dist.barrier()
        a, c = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        c.record()
        torch.cuda.synchronize()
        ts.append(max_over_ranks(a.elapsed_time(c)))
    return sorted(ts)[len(ts) // 2]


say(f"{WORLD} GPU(s); N/batch is PER GPU, Mbox/s is the aggregate over all ranks (max-over-ranks device time)")
def graph_of(fn):
    """The same launches captured once and replayed (what utils.Detector does per shape): no host launch cost."""
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    return g.replay


say(f"{'N/batch':>9s} {'nc':>3s} {'conf':>5s} {'kept':>8s} {'eager ms':>8s} {'graph ms':>8s} {'GPU Mbox/s':>11s} {'CPU ref s (extrap.)':>20s} {'CPU kbox/s':>10s} {'speedup':>9s}")
for total in (10_000, 30_000, 100_000, 300_000, 1_000_000):
    per = total // B
    total = per * B
    for nc in (80, 2):
        boxes = make(total, nc)
        off = (torch.arange(B + 1, dtype=torch.int32, device=dev) * per)
        ws = NmsWorkspace(total, B, dev)
        for conf in (0.5, 0.01):
            run = lambda: batched_nms(boxes, off, 0.45, conf, "center", workspace=ws, class_bits=8)
            ms_eager = gpu_time(run)
            ms = gpu_time(graph_of(run))    # the rate column uses the replayed graph
            kept = int(ws.keep_off[-1])
            # CPU: the reference's algorithm on image 0 (and 1 more when cheap), scaled to 64 images
            n_img = 2 if per <= 2000 else 1
            cpu = 0.0
            if per <= 5000 and WORLD == 1:
                for i in range(n_img):
                    rows = boxes[i * per:(i + 1) * per].cpu().tolist()
                    t0 = time.perf_counter()
                    orc.non_max_suppression(rows, 0.45, conf, "center")
                    cpu += time.perf_counter() - t0
                cpu = cpu / n_img * B
                cpu_s = f"{cpu:20.2f}"
                rate = f"{total / cpu / 1e3:10.1f}"
                sp = f"{cpu * 1e3 / ms:9.0f}"
            else:
                cpu_s, rate, sp = f"{'(skipped: > minutes)':>20s}", f"{'-':>10s}", f"{'-':>9s}"
            say(f"{total:9d} {nc:3d} {conf:5.2f} {kept:8d} {ms_eager:8.3f} {ms:8.3f} {WORLD * total / ms / 1e3:11.1f} {cpu_s} {rate} {sp}", flush=True)

# mAP matching at evaluation scale: D detections vs G ground truths
say()
say(f"{'D dets':>9s} {'G gts':>7s} {'images':>7s} {'GPU ms (map_match)':>20s}")
for D, G, n_img in ((10_000, 1_000, 64), (100_000, 10_000, 640), (1_000_000, 35_000, 5000)):
    g = torch.Generator(device=dev).manual_seed(7)
    gts = torch.rand(G, 7, generator=g, device=dev)
    gts[:, 0] = torch.floor(gts[:, 0] * n_img)
    gts[:, 3:5] = 0.05 + 0.35 * gts[:, 3:5]
    gts[:, 6] = torch.floor(gts[:, 6] * 80)
    dets = gts[torch.randint(0, G, (D,), generator=g, device=dev)].clone()
    dets[:, 1:5] += 0.03 * torch.randn(D, 4, generator=g, device=dev)
    dets[:, 5] = torch.rand(D, generator=g, device=dev)
    ms = gpu_time(lambda: map_match(dets, gts, 0.5, "center"), reps=3)
    say(f"{D:9d} {G:7d} {n_img:7d} {ms:20.3f}")
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
