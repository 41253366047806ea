"""CPU oracle for the YOLOv3 detection hot path -- TEST INFRASTRUCTURE ONLY.

A restatement, in plain torch-CPU fp32 / Python, of the algorithms of the
reference (GabeTsai/YOLO-For-Turbines, paths relative to its checkout):

  forward(...)              code/model.py:150-225  (YOLOv3.forward and its blocks :47-148)
  cells_to_boxes(...)       code/utils.py:86-148
  calc_iou / iou_aligned    code/utils.py:38-84 / :22-36
  non_max_suppression(...)  code/utils.py:150-191
  calc_mAP(...)             code/utils.py:193-274
  check_model_accuracy(...) code/utils.py:334-381 (the per-batch reductions)
  yolo_loss(...)            code/loss.py:29-81
  read_darknet_weights(...) code/model.py:162-170, 227-337

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module; the product package never does (it fails
loudly when the CUDA library is missing instead of falling back to this).

Parity pin: `oracle/gen_golden.py` imports the UNMODIFIED reference from
/root/reference (with stub modules for the absent matplotlib/albumentations),
runs it on seeded inputs and stores inputs + outputs under tests/golden/;
tests/test_oracle_golden.py checks every function here against those vectors
bit-for-bit (NMS/mAP/decode) or to 1e-6 (forward).  The arithmetic that lives in
a third-party dependency (conv / batch-norm / activations / sigmoid / exp) is
PyTorch's (the reference pins torch==2.4.0 in requirements.txt:9; this image has
2.11.0) and is called here exactly where the reference calls it.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Sequence

import numpy as np
import torch
import torch.nn.functional as F

# --- architecture (semantics of code/model.py:20-45, own notation) -------------
# ("c", out, k, stride) conv block | ("r", repeats) residual stage | "s" scale head | "u" upsample
ARCH = [
    ("c", 32, 3, 1), ("c", 64, 3, 2), ("r", 1), ("c", 128, 3, 2), ("r", 2), ("c", 256, 3, 2), ("r", 8),
    ("c", 512, 3, 2), ("r", 8), ("c", 1024, 3, 2), ("r", 4),
    ("c", 512, 1, 1), ("c", 1024, 3, 1), "s", ("c", 256, 1, 1), "u",
    ("c", 256, 1, 1), ("c", 512, 3, 1), "s", ("c", 128, 1, 1), "u",
    ("c", 128, 1, 1), ("c", 256, 3, 1), "s",
]

# reference constants used as test inputs (code/config.py:18-20, 47-57)
CONF_THRESHOLD = 0.5
NMS_IOU_THRESHOLD = 0.45
MAP_IOU_THRESHOLD = 0.5
ANCHORS = [
    [(0.28, 0.22), (0.38, 0.48), (0.9, 0.78)],
    [(0.07, 0.15), (0.15, 0.11), (0.14, 0.29)],
    [(0.02, 0.03), (0.04, 0.07), (0.08, 0.06)],
]
TURBINE_ANCHORS = [
    [(0.215, 0.461), (0.992, 0.349), (0.436, 0.952)],
    [(0.06, 0.143), (0.143, 0.189), (0.408, 0.181)],
    [(0.016, 0.0349), (0.0408, 0.0598), (0.110, 0.0777)],
]


def _act(x: torch.Tensor, activation: str) -> torch.Tensor:
    if activation == "leaky_relu":
        return F.leaky_relu(x, 0.1)  # model.py:64
    if activation == "mish":
        return F.mish(x)  # model.py:66
    raise ValueError(f"Unsupported activation: {activation}")  # model.py:68


_TRAINING = False  # set by forward(training=True): nn.BatchNorm2d in train mode (batch statistics)
_BF16_SIM = False  # set by forward(bf16_sim=True): round where the CUDA path stores bf16 (see train_step_grads)


class _RoundBoth(torch.autograd.Function):
    """value -> bf16 -> fp32 in forward AND on the gradient in backward: a tensor the CUDA path keeps in bf16
    together with its gradient (activations a / dA, raw conv outputs z / dz)."""

    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


class _RoundFwd(torch.autograd.Function):
    """bf16 value, fp32 gradient: conv weights (bf16 operand packs, fp32 weight gradients)."""

    @staticmethod
    def forward(ctx, x):
        return x.bfloat16().float()

    @staticmethod
    def backward(ctx, g):
        return g


class _RoundBwd(torch.autograd.Function):
    """fp32 value, bf16 gradient: the head logits (fp32 out, loss gradient written as the head conv's bf16 dz)."""

    @staticmethod
    def forward(ctx, x):
        return x.clone()

    @staticmethod
    def backward(ctx, g):
        return g.bfloat16().float()


def _rb(x):
    return _RoundBoth.apply(x) if _BF16_SIM else x


def _cnn_block(sd: Dict[str, torch.Tensor], prefix: str, x: torch.Tensor, k: int, stride: int,
               activation: str, bn_act: bool = True) -> torch.Tensor:
    """model.py:80-86.  Eval mode: BatchNorm uses the running statistics; train mode (model.train(), train.py:38):
    batch statistics, and the running ones in `sd` are updated in place as nn.BatchNorm2d does."""
    pad = 1 if k == 3 else 0  # model.py:201
    w = sd[prefix + "conv.weight"]
    if _BF16_SIM:
        w = _RoundFwd.apply(w)
    if not bn_act:
        y = F.conv2d(x, w, sd[prefix + "conv.bias"], stride, pad)
        return _RoundBwd.apply(y) if _BF16_SIM else y
    y = _rb(F.conv2d(x, w, None, stride, pad))
    y = F.batch_norm(y, sd[prefix + "batch_norm.running_mean"], sd[prefix + "batch_norm.running_var"],
                     sd[prefix + "batch_norm.weight"], sd[prefix + "batch_norm.bias"], _TRAINING, 0.1, 1e-5)
    if _TRAINING and prefix + "batch_norm.num_batches_tracked" in sd:
        sd[prefix + "batch_norm.num_batches_tracked"] += 1
    return _act(y, activation)


def _residual_stage(sd, prefix, x, repeats, activation, use_residual=True):
    """model.py:115-121."""
    for r in range(repeats):
        y = _rb(_cnn_block(sd, f"{prefix}layers.{r}.0.", x, 1, 1, activation))
        y = _cnn_block(sd, f"{prefix}layers.{r}.1.", y, 3, 1, activation)
        x = _rb(x + y if use_residual else y)
    return x


def forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, num_classes: int = 80,
            activation: str = "leaky_relu", training: bool = False, bf16_sim: bool = False) -> List[torch.Tensor]:
    """YOLOv3.forward (model.py:172-193) from a reference-keyed state_dict; eval mode unless training=True.
    bf16_sim=True keeps the reference's op sequence but rounds values (and, under autograd, gradients) to bf16
    at the points where the CUDA training path stores bf16 tensors: the image, conv weights, every raw conv
    output and every block output.  With the flag off this is the reference's fp32 arithmetic exactly."""
    global _TRAINING, _BF16_SIM
    _TRAINING, _BF16_SIM = bool(training), bool(bf16_sim)
    try:
        return _forward(sd, _RoundFwd.apply(x) if bf16_sim else x, num_classes, activation)
    finally:
        _TRAINING = _BF16_SIM = False


def _forward(sd, x, num_classes, activation):
    assert torch.sum(torch.isnan(x)) == 0  # model.py:175
    outs, routes = [], []
    i = 0  # index into the reference's nn.ModuleList
    for item in ARCH:
        if item == "s":  # model.py:213-219: three modules
            x = _residual_stage(sd, f"layers.{i}.", x, 1, activation, use_residual=False)
            _nan_guard(x)
            x = _rb(_cnn_block(sd, f"layers.{i + 1}.", x, 1, 1, activation))
            _nan_guard(x)
            p = _rb(_cnn_block(sd, f"layers.{i + 2}.pred_block.0.", x, 3, 1, activation))
            p = _cnn_block(sd, f"layers.{i + 2}.pred_block.1.", p, 1, 1, activation, bn_act=False)
            b, _, h, w = p.shape  # model.py:147-148
            outs.append(p.reshape(b, 3, num_classes + 5, h, w).permute(0, 1, 3, 4, 2))
            i += 3
        elif item == "u":  # model.py:222, 189-191: upsampled channels first
            x = F.interpolate(x, scale_factor=2, mode="nearest")
            _nan_guard(x)
            x = torch.cat([x, routes.pop()], dim=1)
            i += 1
        elif item[0] == "c":
            x = _rb(_cnn_block(sd, f"layers.{i}.", x, item[2], item[3], activation))
            _nan_guard(x)
            i += 1
        else:
            x = _residual_stage(sd, f"layers.{i}.", x, item[1], activation)
            _nan_guard(x)
            if item[1] == 8:  # model.py:186-187
                routes.append(x)
            i += 1
    return outs


def _nan_guard(x):
    if torch.sum(torch.isnan(x)) > 0:  # model.py:183-184
        raise ValueError("Nan in layer")


# --- box utilities --------------------------------------------------------------
def iou_aligned(box1: torch.Tensor, box2: torch.Tensor) -> torch.Tensor:
    """utils.py:22-36."""
    inter = torch.min(box1[..., 0], box2[..., 0]) * torch.min(box1[..., 1], box2[..., 1])
    union = box1[..., 0] * box1[..., 1] + box2[..., 0] * box2[..., 1] - inter
    return inter / union


def calc_iou(boxes1: torch.Tensor, boxes2: torch.Tensor, box_format: str = "center") -> torch.Tensor:
    """utils.py:38-84.  Any box_format other than "center" means top-left x,y + w,h."""
    if boxes1.dim() == 1:
        boxes1 = boxes1.unsqueeze(0)
    if boxes2.dim() == 1:
        boxes2 = boxes2.unsqueeze(0)
    if box_format == "center":
        x1, y1 = boxes1[..., 0] - boxes1[..., 2] / 2, boxes1[..., 1] - boxes1[..., 3] / 2
        x2, y2 = boxes2[..., 0] - boxes2[..., 2] / 2, boxes2[..., 1] - boxes2[..., 3] / 2
    else:
        x1, y1, x2, y2 = boxes1[..., 0], boxes1[..., 1], boxes2[..., 0], boxes2[..., 1]
    w1, h1, w2, h2 = boxes1[..., 2], boxes1[..., 3], boxes2[..., 2], boxes2[..., 3]
    xa, ya = torch.max(x1, x2), torch.max(y1, y2)
    xb, yb = torch.min(x1 + w1, x2 + w2), torch.min(y1 + h1, y2 + h2)
    inter = torch.clamp(xb - xa, min=0) * torch.clamp(yb - ya, min=0)
    union = w1 * h1 + w2 * h2 - inter
    return inter / (union + 1e-6)


def cells_to_boxes(predictions: torch.Tensor, anchors: torch.Tensor, grid_size: int,
                   is_pred: bool = True) -> list:
    """utils.py:86-148, including the in-place mutation of predictions[..., :4]."""
    b = predictions.shape[0]
    na = len(anchors)
    box = predictions[..., :4]  # a view: writes below reach the caller's tensor (utils.py:102)
    if is_pred:
        box[..., 0:2] = torch.sigmoid(box[..., 0:2])
        box[..., 2:] = torch.exp(box[..., 2:]) * anchors.reshape(1, na, 1, 1, 2)
        obj = torch.sigmoid(predictions[..., 4:5])
        cls = torch.argmax(predictions[..., 5:], dim=-1).unsqueeze(-1)
    else:
        obj = predictions[..., 4:5]
        cls = predictions[..., 5:]
    idx = torch.arange(grid_size).repeat(b, 3, grid_size, 1).unsqueeze(-1).to(predictions.device)
    cx = 1 / grid_size * (box[..., 0:1] + idx)
    cy = 1 / grid_size * (box[..., 1:2] + idx.permute(0, 1, 3, 2, 4))
    wh = 1 / grid_size * box[..., 2:]
    out = torch.cat((cx, cy, wh, obj, cls), dim=-1).reshape(b, na * grid_size * grid_size, 6)
    return out.tolist()


def non_max_suppression(boxes: list, iou_threshold: float, obj_threshold: float,
                        box_format: str = "corners") -> list:
    """utils.py:150-191: filter (Python float compare), stable descending sort, greedy loop."""
    cand = [bx for bx in boxes if bx[4] > obj_threshold]
    cand = torch.tensor(sorted(cand, key=lambda bx: bx[4], reverse=True))
    kept = []
    while cand.size(0) > 0:
        top, cand = cand[0], cand[1:]
        ious = calc_iou(top[:4].unsqueeze(0), cand[:, :4], box_format)
        cand = cand[(cand[:, 5] != top[5]) | (ious < iou_threshold)]
        kept.append(top)
    return torch.stack(kept).tolist() if kept else []


def nms_keep_indices(boxes: torch.Tensor, iou_threshold: float, obj_threshold: float,
                     box_format: str = "corners") -> List[int]:
    """Same algorithm as non_max_suppression but returns the ROW INDICES of the survivors (in the
    reference's output order) so that GPU parity can be asserted on indices, not on float rows."""
    rows = boxes.tolist()
    order = [i for i, bx in enumerate(rows) if bx[4] > obj_threshold]
    order.sort(key=lambda i: rows[i][4], reverse=True)  # stable, like sorted(..., reverse=True)
    cand = boxes[order] if order else boxes[:0]
    ids = torch.tensor(order, dtype=torch.long)
    kept: List[int] = []
    while cand.size(0) > 0:
        top, cand, tid, ids = cand[0], cand[1:], int(ids[0]), ids[1:]
        ious = calc_iou(top[:4].unsqueeze(0), cand[:, :4], box_format)
        m = (cand[:, 5] != top[5]) | (ious < iou_threshold)
        cand, ids = cand[m], ids[m]
        kept.append(tid)
    return kept


def calc_mAP(pred_boxes: list, true_boxes: list, iou_threshold: float = 0.5,
             box_format: str = "center", num_classes: int = 20, return_tp: bool = False):
    """utils.py:193-274.  Rows are [image, cx, cy, w, h, score, class]."""
    aps = []
    tp_rows = {}  # row index in pred_boxes -> TP flag (for matching-step parity)
    for c in range(num_classes):
        dets = [(i, d) for i, d in enumerate(pred_boxes) if d[-1] == c]
        gts = [g for g in true_boxes if g[-1] == c]
        if not gts:
            continue
        claimed = {}
        for g in gts:
            claimed[g[0]] = claimed.get(g[0], 0) + 1
        claimed = {k: torch.zeros(v) for k, v in claimed.items()}
        dets.sort(key=lambda t: t[1][5], reverse=True)
        tp, fp = torch.zeros(len(dets)), torch.zeros(len(dets))
        for di, (row, d) in enumerate(dets):
            img_gts = [g for g in gts if g[0] == d[0]]
            best, best_j = 0, 0
            for j, g in enumerate(img_gts):
                iou = calc_iou(torch.tensor(d[1:5]), torch.tensor(g[1:5]), box_format=box_format)
                if iou > best:
                    best, best_j = iou, j
            if best > iou_threshold and claimed[d[0]][best_j] == 0:
                tp[di] = 1
                claimed[d[0]][best_j] = 1
            else:
                fp[di] = 1
            tp_rows[row] = float(tp[di])
        ctp, cfp = torch.cumsum(tp, 0), torch.cumsum(fp, 0)
        prec = torch.cat((torch.tensor([1]), ctp / (ctp + cfp)))
        rec = torch.cat((torch.tensor([0]), ctp / len(gts)))
        aps.append(torch.trapz(prec, rec))
    result = sum(aps) / len(aps)  # ZeroDivisionError when no class has ground truth, as the reference
    return (result, tp_rows) if return_tp else result


def check_model_accuracy(outs, targets, object_threshold):
    """Reductions of utils.py:356-381 for one batch: returns (class_acc, noobj_acc, obj_acc) and the six counts."""
    cc = tc = co = to = cn = tn = 0
    for out, tgt in zip(outs, targets):
        obj, noobj = tgt[..., 4] == 1, tgt[..., 4] == 0
        cc += torch.sum(torch.argmax(out[..., 5:][obj], dim=-1) == tgt[..., 5][obj])
        tc += torch.sum(obj)
        pred = torch.sigmoid(out[..., 4]) > object_threshold
        co += torch.sum(pred[obj] == tgt[..., 4][obj])
        to += torch.sum(obj)
        cn += torch.sum(pred[noobj] == tgt[..., 4][noobj])
        tn += torch.sum(noobj)
    counts = [int(v) for v in (cc, tc, co, to, cn, tn)]
    return (cc / (tc + 1e-16), cn / (tn + 1e-16), co / (to + 1e-16)), counts


def yolo_loss(predictions: torch.Tensor, targets: torch.Tensor, anchors: torch.Tensor) -> list:
    """YOLOLoss.forward (loss.py:29-81) for one scale, including its in-place updates of predictions[..., 1:3]
    and targets[..., 2:4] and the index quirk (sigmoid on entries 1:3, i.e. ty and tw)."""
    obj, noobj = targets[..., 4] == 1, targets[..., 4] == 0
    anchors = anchors.reshape(1, 3, 1, 1, 2)
    zero = torch.tensor(0.0)
    object_loss = box_loss = class_loss = zero
    no_obj_loss = F.binary_cross_entropy_with_logits(predictions[..., 4][noobj], targets[..., 4][noobj])
    if obj.any():
        boxes = torch.cat([torch.sigmoid(predictions[..., :2]), torch.exp(predictions[..., 2:4]) * anchors], dim=-1)
        ious = calc_iou(boxes[obj], targets[..., :4][obj]).unsqueeze(1).detach()
        object_loss = F.mse_loss(predictions[..., 4:5][obj], ious * targets[..., 4:5][obj])
        predictions[..., 1:3] = torch.sigmoid(predictions[..., 1:3])
        targets[..., 2:4] = torch.log(1e-16 + targets[..., 2:4] / anchors)
        box_loss = F.mse_loss(predictions[..., :4][obj], targets[..., :4][obj])
        class_loss = F.cross_entropy(predictions[..., 5:][obj], targets[..., 5][obj].long())
    return [5 * box_loss, 1 * object_loss, 0.5 * no_obj_loss, 1 * class_loss]


def train_step_grads(sd: Dict[str, torch.Tensor], x: torch.Tensor, targets: Sequence[torch.Tensor], anchors,
                     num_classes: int, activation: str = "leaky_relu", bf16_sim: bool = False):
    """The autograd part of one training step (train.py:53-67) in fp32: model.train() forward, YOLOLoss on the three
    scales with anchors scaled by the grid size (train.py:195-197), loss = sum of all twelve terms, backward.
    `sd` is modified like the module would be (running statistics, num_batches_tracked).  Returns
    ([box, object, no_object, class] summed over scales, {parameter key: gradient})."""
    params = {k: v for k, v in sd.items() if k.endswith((".weight", ".bias"))}
    for v in params.values():
        v.requires_grad_(True)
        v.grad = None
    outs = forward(sd, x, num_classes, activation, training=True, bf16_sim=bf16_sim)
    terms = [torch.zeros(()) for _ in range(4)]
    for o, t, a in zip(outs, targets, anchors):
        S = o.shape[2]
        per = yolo_loss(o, t.clone(), torch.tensor(a, dtype=torch.float32) * S)
        terms = [acc + v for acc, v in zip(terms, per)]
    sum(terms).backward()
    grads = {k: v.grad.detach().clone() for k, v in params.items()}
    for v in params.values():
        v.requires_grad_(False)
        v.grad = None
    return [float(v.detach()) for v in terms], grads


def synth_targets(batch: int, size: int, num_classes: int, seed: int, per_scale: int = 4, ignore: int = 1):
    """Synthetic YOLO targets (B,3,S,S,6) for S = size/32, /16, /8: `per_scale` object cells
    per image and scale with x,y in (0,1), w,h in (0.5,4) grid units, obj 1, a class label; `ignore` cells of -1."""
    g = torch.Generator().manual_seed(seed)
    out = []
    for S in (size // 32, size // 16, size // 8):
        t = torch.zeros(batch, 3, S, S, 6)
        for b in range(batch):
            cells = torch.randperm(3 * S * S, generator=g)[: per_scale + ignore]
            for n, c in enumerate(cells.tolist()):
                a, r = divmod(c, S * S)
                i, j = divmod(r, S)
                if n < per_scale:
                    t[b, a, i, j, 0:2] = torch.rand(2, generator=g)
                    t[b, a, i, j, 2:4] = 0.5 + 3.5 * torch.rand(2, generator=g)
                    t[b, a, i, j, 4] = 1.0
                    t[b, a, i, j, 5] = float(torch.randint(0, num_classes, (1,), generator=g))
                else:
                    t[b, a, i, j, 4] = -1.0
        out.append(t)
    return out


# --- Darknet weight file (model.py:162-170, 227-337) ------------------------------
def darknet_param_order(sd_keys: Sequence[str], num_classes: int = 80):
    """Yields (kind, prefix) in the order the reference consumes the flat fp32 file:
    per BN block: beta, gamma, running_mean, running_var, then conv weight (model.py:241-244,
    311-328, 301-305); per bias conv: bias then weight (model.py:294-305).  `layer_id` advances
    once per Conv2d, once per BatchNorm2d and once per Upsample (model.py:336)."""
    i = 0
    for item in ARCH:
        if item == "s":
            yield ("bn_conv", f"layers.{i}.layers.0.0.")
            yield ("bn_conv", f"layers.{i}.layers.0.1.")
            yield ("bn_conv", f"layers.{i + 1}.")
            yield ("bn_conv", f"layers.{i + 2}.pred_block.0.")
            yield ("bias_conv", f"layers.{i + 2}.pred_block.1.")
            i += 3
        elif item == "u":
            yield ("upsample", f"layers.{i}.")
            i += 1
        elif item[0] == "c":
            yield ("bn_conv", f"layers.{i}.")
            i += 1
        else:
            for r in range(item[1]):
                yield ("bn_conv", f"layers.{i}.layers.{r}.0.")
                yield ("bn_conv", f"layers.{i}.layers.{r}.1.")
            i += 1


def read_darknet_weights(path: str, sd: Dict[str, torch.Tensor]) -> Dict[str, int]:
    """Fills `sd` (reference-keyed tensors, modified in place) from a Darknet file the way
    YOLOv3.load_weights does, including the `.conv.N` cutoff quirk (model.py:167-170, 277-291)."""
    with open(path, "rb") as f:
        np.fromfile(f, dtype=np.int32, count=5)
        flat = np.fromfile(f, dtype=np.float32)
    name = os.path.basename(path)
    cutoff = int(name.split(".")[-1]) if ".conv" in name else None
    pos, layer_id, loaded = 0, 0, 0

    def take(key):
        nonlocal pos
        n = sd[key].numel()
        if cutoff is None or layer_id < cutoff:
            sd[key].copy_(torch.from_numpy(flat[pos:pos + n]).view_as(sd[key]))
        pos += n

    for kind, pre in darknet_param_order(list(sd.keys())):
        if kind == "upsample":
            layer_id += 1
            continue
        if kind == "bn_conv":
            for k in ("batch_norm.bias", "batch_norm.weight", "batch_norm.running_mean",
                      "batch_norm.running_var"):
                take(pre + k)
            layer_id += 1
            if cutoff is None or layer_id < cutoff:
                loaded += 1
            take(pre + "conv.weight")
            layer_id += 1
        else:
            take(pre + "conv.bias")
            take(pre + "conv.weight")
            layer_id += 1
    return {"param_idx": pos, "layer_id": layer_id, "n_floats": int(flat.size), "loaded_bn_convs": loaded}


# --- end-to-end detection as the reference's callers do it (utils.py:296-321, demo.py:30-55) ----
def get_eval_boxes(batches, anchors, iou_threshold, obj_threshold, box_format="center"):
    """utils.py:276-332 restated on precomputed head tensors: `batches` yields (heads[3], targets[3]) per batch, i.e.
    what `model(x)` and the loader hand the reference.  Per image: the three scales' decoded boxes concatenated in
    scale order (:300-309), NMS (:317-321), image index prepended (:323-324); true boxes are the rows of the LAST
    scale's targets (anchors of the last scale, :313-315) whose objectness exceeds obj_threshold (:326-328)."""
    data_idx, preds, trues = 0, [], []
    for heads, targets in batches:
        bsz = heads[0].shape[0]
        per_image = [[] for _ in range(bsz)]
        for i, o in enumerate(heads):
            s = o.shape[2]
            a = torch.tensor([*anchors[i]]) * s
            for b, rows in enumerate(cells_to_boxes(o.clone(), a, s, is_pred=True)):
                per_image[b] += rows
        true_rows = cells_to_boxes(targets[2].clone(), a, s, is_pred=False)
        for b in range(bsz):
            for row in non_max_suppression(per_image[b], iou_threshold, obj_threshold, box_format):
                preds.append([data_idx] + row)
            for row in true_rows[b]:
                if row[4] > obj_threshold:
                    trues.append([data_idx] + row)
            data_idx += 1
    return preds, trues


def detect(sd, x, anchors=ANCHORS, iou_threshold=NMS_IOU_THRESHOLD, obj_threshold=CONF_THRESHOLD,
           num_classes=80, activation="leaky_relu", box_format="center"):
    with torch.no_grad():
        outs = forward(sd, x, num_classes, activation)
        per_image = [[] for _ in range(x.shape[0])]
        for i, o in enumerate(outs):
            s = o.shape[2]
            a = torch.tensor(anchors[i]) * s
            for b, rows in enumerate(cells_to_boxes(o.clone(), a, s, is_pred=True)):
                per_image[b] += rows
        return [non_max_suppression(rows, iou_threshold, obj_threshold, box_format) for rows in per_image]
