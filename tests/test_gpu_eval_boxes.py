"""GPU parity of the callers around the hot path: get_eval_boxes (utils.py:276-332) and the mode
changes of check_model_accuracy (utils.py:334-381), against lists produced by the UNMODIFIED reference
(tests/golden/eval_boxes.npz, oracle/gen_golden_eval.py).

Bar: same number of rows in the same order; image index and class exact; box floats within 1e-5 (decode tolerance,
fp32); true boxes within 1e-6; the model sees eval() then train(), as the reference calls them."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


class _FakeModel:
    """What the reference needs from a model in these two functions: eval(), train(), __call__ -> 3 head tensors."""

    def __init__(self, batches, device):
        self.batches, self.i, self.calls, self.device = batches, 0, [], device

    def eval(self):
        self.calls.append("eval")

    def train(self):
        self.calls.append("train")

    def __call__(self, x):
        out = [o.clone().to(self.device) for o in self.batches[self.i]]
        self.i += 1
        return out


def _batches(e):
    return [([torch.from_numpy(e[f"b{bi}_head{si}"]) for si in range(3)], [torch.from_numpy(e[f"b{bi}_tgt{si}"]) for si in range(3)])
            for bi in range(2)]


def test_get_eval_boxes_matches_reference_lists(gold):
    from yolo_for_turbines_b200.utils import get_eval_boxes

    e = gold.eval_boxes
    batches = _batches(e)
    model = _FakeModel([h for h, _ in batches], "cuda")
    loader = [(torch.zeros(len(h[0]), 3, 64, 64), [t.clone() for t in tg]) for h, tg in batches]
    preds, trues = get_eval_boxes(loader, model, float(e["iou_thr"]), e["anchors"].tolist(), float(e["obj_thr"]),
                                  box_format="center", device="cuda")
    assert model.calls == e["calls"].tolist() == ["eval", "train"]
    got_p = np.asarray(preds, dtype=np.float64).reshape(-1, 7)
    got_t = np.asarray(trues, dtype=np.float64).reshape(-1, 7)
    assert got_p.shape == e["preds"].shape and got_t.shape == e["trues"].shape
    assert np.array_equal(got_p[:, 0], e["preds"][:, 0]) and np.array_equal(got_p[:, 6], e["preds"][:, 6])
    assert np.allclose(got_p[:, 1:6], e["preds"][:, 1:6], rtol=1e-5, atol=1e-5)
    assert np.array_equal(got_t[:, [0, 6]], e["trues"][:, [0, 6]])
    assert np.allclose(got_t[:, 1:6], e["trues"][:, 1:6], rtol=1e-6, atol=1e-6)


def test_get_eval_boxes_native_model_equals_generic_path():
    """The planned Detector path (a model of this package) and the generic path (any callable returning heads) give
    the same lists on the same network; training mode is switched back on unconditionally (utils.py:331)."""
    from oracle import synth
    from oracle import yolo_oracle as orc
    from yolo_for_turbines_b200.model import YOLOv3
    from yolo_for_turbines_b200.utils import get_eval_boxes

    m = YOLOv3(num_classes=2).eval()
    m.load_state_dict(synth.synth_state_dict(m.state_dict(), seed=3))
    m = m.cuda()
    g = torch.Generator().manual_seed(11)
    loader = []
    for bsz in (2, 1):
        x = torch.rand(bsz, 3, 64, 64, generator=g)
        tg = orc.synth_targets(bsz, 64, 2, seed=bsz)
        loader.append((x, tg))
    assert not m.training
    p1, t1 = get_eval_boxes(loader, m, 0.45, orc.TURBINE_ANCHORS, 0.3, "center", device="cuda")
    assert m.training                       # the reference leaves the model in train mode, whatever it was before
    m.eval()

    class Wrap:                             # hides _prepare: forces the generic path on the same network
        def __init__(self, inner):
            self.inner = inner

        def eval(self):
            self.inner.eval()

        def train(self):
            self.inner.train()

        def __call__(self, x):
            return self.inner(x)

    p2, t2 = get_eval_boxes(loader, Wrap(m), 0.45, orc.TURBINE_ANCHORS, 0.3, "center", device="cuda")
    assert len(p1) > 0 and p1 == p2 and t1 == t2


def test_check_model_accuracy_mode_calls(gold, capsys):
    from yolo_for_turbines_b200.utils import check_model_accuracy

    a = gold.accuracy
    outs = [torch.from_numpy(a[f"out{i}"]) for i in range(3)]
    tgts = [torch.from_numpy(a[f"tgt{i}"]) for i in range(3)]
    model = _FakeModel([outs], "cuda")
    res = check_model_accuracy(model, [(torch.zeros(3, 3, 64, 64), tgts)], float(a["thr"]))
    assert model.calls == ["eval", "train"]          # utils.py:344, :380
    assert np.allclose([float(r) for r in res], a["result"], rtol=1e-6)
    assert "Class accuracy is" in capsys.readouterr().out
