"""world_size-2 gloo test (CPU) of the data-parallel host logic: image sharding and the rank-ordered
detection gather that feeds the distributed mAP."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from yolo_for_turbines_b200.parallel import gather_rows, shard_range

    n_img = 7
    lo, hi = shard_range(n_img, rank, world)
    # rank r holds (img, k) rows for its images: image i contributes i % 3 detections
    rows = [[float(i), float(k), 0, 0, 0, 0.5, 1.0] for i in range(lo, hi) for k in range(i % 3)]
    local = torch.tensor(rows, dtype=torch.float32).reshape(-1, 7)
    allrows = gather_rows(local)
    q.put((rank, (lo, hi), allrows.tolist()))
    dist.destroy_process_group()


def test_shard_and_gather_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [o[1] for o in out] == [(0, 4), (4, 7)]
    expect = [[float(i), float(k), 0, 0, 0, 0.5, 1.0] for i in range(7) for k in range(i % 3)]
    assert out[0][2] == expect and out[1][2] == expect  # identical, globally image-ordered, on every rank


def test_shard_range_covers_everything():
    from yolo_for_turbines_b200.parallel import shard_range

    for n in (0, 1, 7, 64, 65):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def _bucket_worker(rank, world, port, q):
    import os

    import torch
    import torch.distributed as dist

    from yolo_for_turbines_b200.train import make_buckets

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # 12 "ops" with parameter blocks of uneven size laid out in op order (as Trainer lays out flat_g)
        sizes = [40, 8, 8, 120, 16, 300, 16, 16, 64, 500, 12, 100]
        offs, n = [], 0
        for s_ in sizes:
            offs.append(n)
            n += s_
        buckets = make_buckets([(o, i) for i, o in enumerate(offs)], n, 256)
        g = torch.Generator().manual_seed(100 + rank)
        local = torch.randn(n, generator=g)
        flat = torch.full((n,), float("nan"))
        pending = list(buckets)
        for i in range(len(sizes) - 1, -1, -1):           # the backward pass: op i writes its gradient slice ...
            flat[offs[i]:offs[i] + sizes[i]] = local[offs[i]:offs[i] + sizes[i]]
            while pending and pending[0][0] >= i:         # ... and every bucket that just became complete is reduced
                _, lo, hi = pending.pop(0)
                dist.all_reduce(flat[lo:hi])
        assert not pending
        total = local.clone()
        dist.all_reduce(total)
        ok = bool(torch.allclose(flat, total))
        cover = sorted((lo, hi) for _, lo, hi in buckets)
        ok = ok and cover[0][0] == 0 and cover[-1][1] == n and all(a[1] == b[0] for a, b in zip(cover, cover[1:]))
        ok = ok and all(hi - lo >= 256 for _, lo, hi in buckets[:-1])
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_gradient_bucket_schedule_gloo_world2():
    """Trainer's bucketed gradient all-reduce (flat buffer, buckets fired as the backward pass completes them) on the
    gloo backend with two CPU ranks: the result equals one all-reduce of the whole buffer."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_bucket_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok in res), res
