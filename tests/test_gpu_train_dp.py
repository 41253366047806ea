"""Data-parallel training step on 2 GPUs (NCCL): gradient all-reduce == sum of local gradients, replicas stay
identical.  Needs two visible GPUs; skipped otherwise (the gloo world-size-2 test in test_dist_gloo.py covers the
bucket schedule on CPU)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_dp2_gradient_allreduce_and_replica_consistency():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29671", os.path.join(ROOT, "scripts", "train_dp_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("-> ok") == 2, r.stdout
