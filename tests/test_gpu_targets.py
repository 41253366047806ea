"""GPU parity of the training-target encoder (yolo_encode_targets through dataset.encode_targets) against targets made
by the UNMODIFIED reference `YOLODataset.__getitem__` (tests/golden/targets.npz) and against the oracle on larger
random batches: BIT-EXACT."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_targets_match_reference_golden():
    from test_oracle_targets import golden_cases
    from yolo_for_turbines_b200.dataset import encode_targets

    by_cfg = {}
    for c, boxes, anchors, grid, dense in golden_cases():
        got = encode_targets([boxes], anchors, grid_sizes=grid)
        torch.cuda.synchronize()
        for s in range(3):
            assert torch.equal(got[s][0].cpu(), dense[s]), (c, s)
        by_cfg.setdefault((id(anchors), tuple(grid)), (anchors, grid, []))[2].append((boxes, dense))
    # the same images as ONE batch per configuration (collate_fn's stacking, utils.py:699)
    for anchors, grid, items in by_cfg.values():
        got = encode_targets([b for b, _ in items], anchors, grid_sizes=grid)
        for s in range(3):
            assert torch.equal(got[s].cpu(), torch.stack([d[s] for _, d in items]))


def test_targets_match_oracle_on_random_batches_and_feed_the_trainer():
    from oracle import target_oracle as to
    from oracle import yolo_oracle as orc
    from oracle.gen_golden_targets import synth_boxes
    from yolo_for_turbines_b200.dataset import encode_targets

    rng = np.random.default_rng(2)
    for size in (320, 416, 608):
        grid = [size // 32, size // 16, size // 8]
        batch = [synth_boxes(int(rng.integers(0, 80)), 2, rng, orc.TURBINE_ANCHORS) for _ in range(16)]
        ref = to.encode_batch(batch, orc.TURBINE_ANCHORS, grid)
        got = encode_targets(batch, orc.TURBINE_ANCHORS, image_size=size)
        torch.cuda.synchronize()
        for s in range(3):
            assert torch.equal(got[s].cpu(), ref[s]), (size, s)
