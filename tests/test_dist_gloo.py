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
