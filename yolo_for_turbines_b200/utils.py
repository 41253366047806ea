"""Drop-in mirror of the hot-path functions of the reference's `code/utils.py`, on sm_100a kernels.

Reference signatures kept verbatim (paths relative to the reference checkout):
    iou_aligned(box1, box2)                                              utils.py:22
    calc_iou(boxes1, boxes2, box_format="center")                        utils.py:38
    cells_to_boxes(predictions, anchors, grid_size, is_pred=True)        utils.py:86
    non_max_suppression(boxes, iou_threshold, obj_threshold, box_format="corners")   utils.py:150
    calc_mAP(pred_boxes, true_boxes, iou_threshold=0.5, box_format="center", num_classes=20)  utils.py:193
    get_eval_boxes(loader, model, iou_threshold, anchors, obj_threshold, box_format="center", device=...)  utils.py:276
plus the names BASELINE.json's north_star uses (cells_to_bboxes, intersection_over_union,
mean_average_precision) as aliases, and tensor-in/tensor-out variants (decode_boxes, batched_nms,
detect, map_match) that skip the Python-list materialisation the reference API forces.

Everything runs through libyolo_b200.so; there is no CPU fallback -- inputs are moved to the CUDA
device (lists/CPU tensors are accepted because the reference API passes Python lists) and a missing
library or GPU raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from . import config
from ._lib import BOX_CENTER, BOX_CORNERS, YoloB200Error, lib, ptr, stream_ptr


def _device(device=None) -> torch.device:
    if device is not None:
        device = torch.device(device)
        if device.type != "cuda":
            raise YoloB200Error("this package is the sm_100a path only: a CUDA device is required (no CPU fallback)")
        return device
    if not torch.cuda.is_available():
        raise YoloB200Error("no CUDA device visible: the sm_100a kernels cannot run and there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _fmt(box_format: str) -> int:
    # utils.py:57-67: only the literal "center" converts; every other string means top-left x,y + w,h
    return BOX_CENTER if box_format == "center" else BOX_CORNERS


def _as_f32_cuda(t, device=None) -> torch.Tensor:
    if not torch.is_tensor(t):
        t = torch.tensor(t, dtype=torch.float32)
    dev = t.device if t.is_cuda else _device(device)
    return t.to(device=dev, dtype=torch.float32).contiguous()


# ---------------------------------------------------------------------------------------- IoU
def calc_iou(boxes1, boxes2, box_format="center"):
    """Element-wise IoU with the reference's exact fp32 op order and +1e-6 (utils.py:38-84)."""
    b1, b2 = _as_f32_cuda(boxes1), _as_f32_cuda(boxes2)
    b2 = b2.to(b1.device)
    if b1.dim() == 1:
        b1 = b1.unsqueeze(0)
    if b2.dim() == 1:
        b2 = b2.unsqueeze(0)
    lead = torch.broadcast_shapes(b1.shape[:-1], b2.shape[:-1])
    b1f = b1.expand(*lead, b1.shape[-1]).reshape(-1, b1.shape[-1]).contiguous()
    b2f = b2.expand(*lead, b2.shape[-1]).reshape(-1, b2.shape[-1]).contiguous()
    n = b1f.shape[0]
    out = torch.empty(n, dtype=torch.float32, device=b1.device)
    with torch.cuda.device(b1.device):
        lib.yolo_iou(ptr(b1f), n, b1f.shape[1], ptr(b2f), n, b2f.shape[1], _fmt(box_format), 0, ptr(out),
                     stream_ptr(b1.device))
    return out.reshape(lead)


def iou_aligned(box1, box2):
    """IoU of [w, h] pairs sharing a centre (utils.py:22-36) -- used by anchor assignment."""
    b1, b2 = _as_f32_cuda(box1), _as_f32_cuda(box2)
    b2 = b2.to(b1.device)
    lead = torch.broadcast_shapes(b1.shape[:-1], b2.shape[:-1])
    b1f = b1.expand(*lead, 2).reshape(-1, 2).contiguous()
    b2f = b2.expand(*lead, 2).reshape(-1, 2).contiguous()
    n = b1f.shape[0]
    out = torch.empty(n, dtype=torch.float32, device=b1.device)
    with torch.cuda.device(b1.device):
        lib.yolo_iou(ptr(b1f), n, 2, ptr(b2f), n, 2, 0, 1, ptr(out), stream_ptr(b1.device))
    return out.reshape(lead)


# ------------------------------------------------------------------------------------- decode
def decode_boxes(predictions: torch.Tensor, anchors, grid_size: int, is_pred: bool = True,
                 out: Optional[torch.Tensor] = None, out_offset: int = 0, writeback: bool = False) -> torch.Tensor:
    """Tensor form of cells_to_boxes: (B,3,S,S,5+nc) fp32 (any strides) -> rows [cx,cy,w,h,obj,cls].
    With `out` (B, N, 6) given, writes rows [out_offset, out_offset + 3*S*S) of every image so that
    the three scales land in one candidate tensor in the reference's order (utils.py:300-309)."""
    if not predictions.is_cuda:
        raise YoloB200Error("decode_boxes needs a CUDA tensor (no CPU fallback)")
    if predictions.dtype != torch.float32:
        if writeback:
            raise YoloB200Error("writeback needs an fp32 predictions tensor")
        predictions = predictions.float()
    B, A, S, S2, Cc = predictions.shape
    if A != 3 or S != grid_size or S2 != grid_size:
        raise YoloB200Error(f"predictions {tuple(predictions.shape)} do not match grid_size {grid_size} / 3 anchors")
    anc = torch.as_tensor(anchors, dtype=torch.float32).reshape(-1).cpu()
    if anc.numel() != 6:
        raise YoloB200Error("anchors must hold 3 (w, h) pairs")
    n = 3 * S * S
    if out is None:
        out = torch.empty(B, n, 6, dtype=torch.float32, device=predictions.device)
        out_offset = 0
    nc = Cc - 5
    strides = (C.c_int64 * 5)(*predictions.stride())
    anc_c = (C.c_float * 6)(*anc.tolist())
    with torch.cuda.device(predictions.device):
        lib.yolo_decode(ptr(predictions), strides, B, S, nc, anc_c, int(bool(is_pred)), int(bool(writeback)), ptr(out),
                        out.shape[1], out_offset, stream_ptr(predictions.device))
    return out


def _dense_head(h: torch.Tensor) -> bool:
    """(B,3,S,S,C) view of the model's own [B*S*S pixels][pitch] fp32 head buffer (what yolo_decode_multi reads)."""
    if not (h.is_cuda and h.dtype == torch.float32 and h.dim() == 5 and h.shape[1] == 3 and h.shape[2] == h.shape[3]):
        return False
    S, Cc = h.shape[2], h.shape[4]
    st = h.stride()
    pitch = st[3]
    return st[4] == 1 and st[1] == Cc and 3 * Cc <= pitch <= 384 and st[2] == S * pitch and st[0] == S * S * pitch


def decode_boxes_multi(heads: Sequence[torch.Tensor], anchors_per_scale, out: torch.Tensor) -> torch.Tensor:
    """decode_boxes for all scales of a detector, concatenated per image in the given order (utils.py:300-309), into
    `out` (B, sum 3*S*S, 6).  One launch when every head is the model's dense layout (yolo_decode_multi: the small
    scales are launch-latency bound on their own); otherwise one yolo_decode call per scale."""
    heads = list(heads)
    if 1 <= len(heads) <= 4 and all(_dense_head(h) for h in heads) and len({h.shape[-1] for h in heads}) == 1 \
            and len({h.device for h in heads}) == 1 and out.data_ptr() % 8 == 0:
        n = len(heads)
        B, nc = heads[0].shape[0], heads[0].shape[-1] - 5
        ptrs = (C.c_void_p * n)(*[h.data_ptr() for h in heads])
        strides = (C.c_int64 * (5 * n))(*[v for h in heads for v in h.stride()])
        sizes = (C.c_int32 * n)(*[h.shape[2] for h in heads])
        anc = []
        for a in anchors_per_scale:
            a = torch.as_tensor(a, dtype=torch.float32).reshape(-1).cpu()
            if a.numel() != 6:
                raise YoloB200Error("anchors must hold 3 (w, h) pairs per scale")
            anc += a.tolist()
        anc_c = (C.c_float * (6 * n))(*anc)
        with torch.cuda.device(heads[0].device):
            lib.yolo_decode_multi(ptrs, strides, B, sizes, nc, anc_c, n, ptr(out), out.shape[1], stream_ptr(heads[0].device))
        return out
    off = 0
    for h, a in zip(heads, anchors_per_scale):
        s = h.shape[2]
        decode_boxes(h, a, s, True, out=out, out_offset=off)
        off += 3 * s * s
    return out


def cells_to_boxes(predictions, anchors, grid_size, is_pred=True):
    """utils.py:86-148: returns the nested list B x 3*S*S x [cx,cy,w,h,obj,cls] and, like the
    reference, rewrites predictions[..., 0:4] in place when is_pred (utils.py:102-110)."""
    if not torch.is_tensor(predictions):
        raise TypeError("predictions must be a tensor")
    if predictions.is_cuda and predictions.dtype == torch.float32:
        return decode_boxes(predictions, anchors, grid_size, is_pred, writeback=bool(is_pred)).tolist()
    dev = _device()
    work = predictions.to(device=dev, dtype=torch.float32)
    out = decode_boxes(work, anchors, grid_size, is_pred, writeback=bool(is_pred))
    if is_pred:  # propagate the reference's input mutation back to the caller's tensor
        predictions[..., :4] = work[..., :4].to(device=predictions.device, dtype=predictions.dtype)
    return out.tolist()


# ---------------------------------------------------------------------------------------- NMS
@dataclass
class NmsResult:
    boxes: torch.Tensor      # [total, 6] candidates as given
    keep_idx: torch.Tensor   # [total] int32: row indices of survivors, image-major, reference order
    keep_off: torch.Tensor   # [B+1] int32: image b owns keep_idx[keep_off[b]:keep_off[b+1]]
    ready: Optional[torch.cuda.Event] = None   # set when the result was produced on another stream (Detector lanes)

    def wait(self):
        """Makes the current stream wait for the result (no-op for results produced on the current stream)."""
        if self.ready is not None:
            torch.cuda.current_stream(self.boxes.device).wait_event(self.ready)
        return self

    def kept_rows(self) -> List[torch.Tensor]:
        self.wait()
        off = self.keep_off.tolist()
        rows = self.boxes[self.keep_idx[: off[-1]].long()]
        return [rows[off[b]:off[b + 1]] for b in range(len(off) - 1)]

    def to_lists(self) -> List[list]:
        return [r.tolist() for r in self.kept_rows()]


class NmsWorkspace:
    """Caller-owned scratch for yolo_nms (sort ping-pong buffers, masks, segment lists)."""

    def __init__(self, total: int, batch: int, device):
        self.total, self.batch = total, batch
        self.nbytes = int(lib.yolo_nms_workspace_bytes(total, batch))
        self.buf = torch.empty(self.nbytes, dtype=torch.uint8, device=device)
        self.keep_idx = torch.empty(max(total, 1), dtype=torch.int32, device=device)
        self.keep_off = torch.zeros(batch + 1, dtype=torch.int32, device=device)


def batched_nms(boxes: torch.Tensor, img_offsets: torch.Tensor, iou_threshold: float, obj_threshold: float,
                box_format: str = "corners", workspace: Optional[NmsWorkspace] = None,
                class_bits: int = 0) -> NmsResult:
    """Class-aware greedy NMS of a whole batch in one pipeline (K4 compaction, K5 sort, K6 NMS).
    boxes [total,6] fp32 CUDA rows [x,y,w,h,score,cls]; img_offsets [B+1] int32 CUDA."""
    if not boxes.is_cuda:
        raise YoloB200Error("batched_nms needs CUDA tensors (no CPU fallback)")
    boxes = boxes.to(torch.float32).contiguous()
    total, B = boxes.shape[0], img_offsets.numel() - 1
    img_offsets = img_offsets.to(device=boxes.device, dtype=torch.int32).contiguous()
    ws = workspace
    if ws is None or ws.total < total or ws.batch != B or ws.buf.device != boxes.device:
        ws = NmsWorkspace(total, B, boxes.device)
    thr32 = float(torch.tensor(iou_threshold, dtype=torch.float32))  # utils.py:179 compares in fp32
    with torch.cuda.device(boxes.device):
        lib.yolo_nms(ptr(boxes), ptr(img_offsets), B, total, thr32, float(obj_threshold), _fmt(box_format),
                     int(class_bits), ptr(ws.keep_idx), ptr(ws.keep_off), ptr(ws.buf), ws.nbytes, stream_ptr(boxes.device))
    return NmsResult(boxes, ws.keep_idx, ws.keep_off)


def non_max_suppression(boxes, iou_threshold, obj_threshold, box_format="corners"):
    """utils.py:150-191 for ONE image: list (or tensor) of [x,y,w,h,obj,cls] -> surviving rows as a
    list of lists in descending-score order; [] when nothing passes."""
    if not torch.is_tensor(boxes):
        if len(boxes) == 0:
            return []
        boxes = torch.tensor(boxes, dtype=torch.float32)
    if boxes.numel() == 0:
        return []
    dev = boxes.device if boxes.is_cuda else _device()
    b = boxes.to(device=dev, dtype=torch.float32).reshape(-1, 6).contiguous()
    off = torch.tensor([0, b.shape[0]], dtype=torch.int32, device=dev)
    return batched_nms(b, off, iou_threshold, obj_threshold, box_format).to_lists()[0]


# ---------------------------------------------------------------------------------------- mAP
def map_match(dets: torch.Tensor, gts: torch.Tensor, iou_threshold: float = 0.5, box_format: str = "center"):
    """K7 on device.  dets [D,7], gts [G,7] rows [img,cx,cy,w,h,score,cls] (fp32, CUDA).
    Returns (order, tp_sorted, cls_sorted, score_sorted): detections in the reference's evaluation
    order (class ascending, then stable descending score, utils.py:210,229) with their TP flags."""
    dev = dets.device
    D, G = dets.shape[0], gts.shape[0]
    dets = dets.to(torch.float32).contiguous()
    gts = gts.to(device=dev, dtype=torch.float32).contiguous()
    # evaluation order: stable sort by score desc, then stable by class  == per-class stable score order
    o1 = torch.sort(dets[:, 5], descending=True, stable=True).indices
    o2 = torch.sort(dets[o1, 6], stable=True).indices
    order = o1[o2]
    rank = torch.empty(D, dtype=torch.int32, device=dev)
    rank[order] = torch.arange(D, dtype=torch.int32, device=dev)
    # group ground truths by image, keeping their original relative order (utils.py:236-238)
    gorder = torch.sort(gts[:, 0], stable=True).indices if G else torch.zeros(0, dtype=torch.long, device=dev)
    gsorted = gts[gorder].contiguous()
    gimg = gsorted[:, 0].contiguous()
    lo = torch.searchsorted(gimg, dets[:, 0].contiguous(), right=False).to(torch.int32)
    hi = torch.searchsorted(gimg, dets[:, 0].contiguous(), right=True).to(torch.int32)
    tp = torch.zeros(D, dtype=torch.float32, device=dev)
    best_iou = torch.zeros(D, dtype=torch.float32, device=dev)
    best_gt = torch.full((D,), -1, dtype=torch.int32, device=dev)
    claim = torch.empty(max(G, 1), dtype=torch.int32, device=dev)
    thr32 = float(torch.tensor(iou_threshold, dtype=torch.float32))
    with torch.cuda.device(dev):
        lib.yolo_map_match(ptr(dets), D, ptr(gsorted), G, ptr(lo), ptr(hi), ptr(rank), thr32, _fmt(box_format), ptr(tp),
                           ptr(best_iou), ptr(best_gt), ptr(claim), stream_ptr(dev))
    return order, tp[order], dets[order, 6], dets[order, 5], tp


def calc_mAP(pred_boxes, true_boxes, iou_threshold=0.5, box_format="center", num_classes=20):
    """utils.py:193-274.  Rows [img, cx, cy, w, h, score, cls]; returns a 0-dim fp32 tensor.
    The O(D*G) matching (yolo_map_match) and the per-class cumsum / precision / recall / trapz tail (yolo_map_ap, one
    CTA per class) run on device; the host reads back one scalar."""
    dev = pred_boxes.device if torch.is_tensor(pred_boxes) and pred_boxes.is_cuda else _device()
    dets = torch.as_tensor(pred_boxes, dtype=torch.float32).reshape(-1, 7).to(dev)
    gts = torch.as_tensor(true_boxes, dtype=torch.float32).reshape(-1, 7).to(dev)
    classes = torch.arange(num_classes, dtype=torch.float32, device=dev)
    n_gt = (gts[:, 6].unsqueeze(0) == classes.unsqueeze(1)).sum(1)  # utils.py:211-213
    valid = n_gt > 0
    n_valid = int(valid.sum())
    if n_valid == 0:
        raise ZeroDivisionError("division by zero")  # utils.py:274 with no ground truth at all
    if dets.shape[0] == 0:
        return torch.zeros((), dtype=torch.float32)
    order, tp_s, cls_s, _, _ = map_match(dets, gts, iou_threshold, box_format)
    # per-class AP on device (one CTA per class, csrc/map.cu k_map_ap): no per-class host loop, one read-back
    cls_c = cls_s.contiguous()
    start = torch.searchsorted(cls_c, classes, right=False).to(torch.int32)   # first det of each class
    end = torch.searchsorted(cls_c, classes, right=True).to(torch.int32)
    ap = torch.empty(num_classes, dtype=torch.float32, device=dev)
    tp_c, n_gt32 = tp_s.contiguous(), n_gt.to(torch.int32).contiguous()   # named: they must outlive the argument evaluation
    with torch.cuda.device(dev):
        lib.yolo_map_ap(ptr(tp_c), ptr(start), ptr(end), ptr(n_gt32), int(num_classes), ptr(ap), stream_ptr(dev))
    # mean over the classes that have ground truth (utils.py:274): sum(aps) / len(aps), accumulated in class order
    aps = ap[valid]
    return (aps.double().sum() / float(n_valid)).to(torch.float32).cpu()


# ------------------------------------------------------------------- fused detection pipeline
def _scaled_anchors(anchors, i: int, grid_size: int) -> torch.Tensor:
    """utils.py:303 / demo.py:33-35: fp32 anchors times the grid size, rounded in fp32."""
    a = anchors[i] if torch.is_tensor(anchors) else torch.tensor([*anchors[i]])
    return a.detach().to(device="cpu", dtype=torch.float32) * grid_size


class Detector:
    """forward -> decode x3 -> batched NMS with every intermediate resident in HBM and static per
    (batch, H, W): the device pipeline behind get_eval_boxes (utils.py:276-332) and demo.predict
    (demo.py:30-55).  No host sync happens until results are read."""

    def __init__(self, model, anchors=None, iou_threshold=config.NMS_IOU_THRESHOLD,
                 obj_threshold=config.CONF_THRESHOLD, box_format="center", lanes: int = 1):
        """lanes > 1: consecutive calls alternate between `lanes` independent pipelines (own activation buffers, CUDA
        graph and stream), so the DRAM- / latency-bound stages of one batch (input conversion, decode, NMS) and the
        tails of its conv kernels overlap with the conv kernels of the next batch.  Results then carry a `ready`
        event: `NmsResult.wait()` / `kept_rows()` / `plan.check_status()` order the consumer after the lane."""
        self.model = model
        self.lanes = max(1, int(lanes))
        self._next_lane = 0
        self._lane_streams = {}
        self.anchors = config.ANCHORS if anchors is None else anchors
        self.iou_threshold, self.obj_threshold, self.box_format = iou_threshold, obj_threshold, box_format
        self.use_graph = True
        self.fuse_decode = True    # head convs decode in their epilogue when the plan supports it (engine.decode_plans)
        self._state = {}

    def _get_state(self, B, grids, device, lane=0):
        key = (B, tuple(grids), str(device), lane)
        st = self._state.get(key)
        if st is None:
            n = sum(3 * s * s for s in grids)
            st = dict(n=n, cand=torch.empty(B, n, 6, dtype=torch.float32, device=device),
                      off=(torch.arange(B + 1, dtype=torch.int32, device=device) * n).contiguous(),
                      ws=NmsWorkspace(B * n, B, device))
            self._state[key] = st
        return st

    def _post(self, plan, st, fused: bool = False):
        """decode x3 (unless the head convs already wrote plan.cand) + batched NMS, all launches on the current stream."""
        if fused:
            cand = plan.cand
            nc = max(m[1] for m in plan.head_meta)
        else:
            heads = plan.head_views()
            cand = st["cand"]
            decode_boxes_multi(heads, [_scaled_anchors(self.anchors, i, h.shape[2]) for i, h in enumerate(heads)], cand)
            nc = max(h.shape[-1] - 5 for h in heads)
        # classes come from the decode's argmax: integers < num_classes, so the grouping sort needs one pass
        return batched_nms(cand.view(-1, 6), st["off"], self.iou_threshold, self.obj_threshold, self.box_format,
                           workspace=st["ws"], class_bits=8 if nc <= 256 else (16 if nc <= 65536 else 0))

    def __call__(self, x: torch.Tensor):
        """Returns (NmsResult, plan); plan.check_status() reports NaN inputs/layers after the fact.
        Call 1 of a shape runs eagerly (warm-up), call 2 captures convs + decode + NMS (~115 launches) in ONE
        CUDA graph, later calls replay it; only the input conversion (its source pointer changes) stays eager."""
        if self.lanes == 1:
            return self._run(x, 0)
        lane = self._next_lane
        self._next_lane = (lane + 1) % self.lanes
        dev = x.device
        key = (str(dev), lane)
        ls = self._lane_streams.get(key)
        if ls is None:
            ls = self._lane_streams[key] = (torch.cuda.Stream(device=dev), torch.cuda.Event(), torch.cuda.Event())
        stream, ev_in, ev_done = ls
        cur = torch.cuda.current_stream(dev)
        ev_in.record(cur)            # x (and everything the caller enqueued before) is ready
        stream.wait_event(ev_in)
        x.record_stream(stream)
        with torch.cuda.stream(stream):
            res, plan = self._run(x, lane)
            ev_done.record(stream)
        res.ready = ev_done
        plan.done_event = ev_done
        return res, plan

    def join(self, device=None):
        """Makes the current stream wait for everything enqueued on the lane streams so far."""
        for (dev, _), (stream, _, _) in self._lane_streams.items():
            if device is None or str(device) == dev:
                torch.cuda.current_stream(stream.device).wait_stream(stream)

    def _run(self, x: torch.Tensor, lane: int):
        plan, x = self.model._prepare(x, lane)
        with torch.cuda.device(x.device):
            st = self._get_state(plan.B, [h.H for h in plan.heads], x.device, lane)
            if st.get("plan") is not plan:  # the model re-planned (new engine): drop the stale graph
                st.update(plan=plan, graph=None, calls=0, res=None)
            fused = bool(self.fuse_decode and plan.decode_plans(self.anchors))
            if st.get("fused") is not fused:
                st.update(graph=None, calls=0, res=None, fused=fused)
            if st["graph"] is not None:
                plan._launch_input(x)
                st["graph"].replay()
            elif st["calls"] == 0 or not self.use_graph:
                plan._launch_input(x)
                plan._launch_convs(decode=fused)   # eager: also sets the dynamic-smem attributes outside capture
                st["res"] = self._post(plan, st, fused)
            else:
                plan._launch_input(x)
                torch.cuda.current_stream(x.device).synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    plan._launch_convs(decode=fused)
                    st["res"] = self._post(plan, st, fused)
                st["graph"] = g
                g.replay()
            st["calls"] += 1
        return st["res"], plan


def detect(model, x, anchors=None, iou_threshold=config.NMS_IOU_THRESHOLD, obj_threshold=config.CONF_THRESHOLD,
           box_format="center") -> List[list]:
    """One call for demo.predict's body (demo.py:41-55): per image, the NMS survivors as lists."""
    det = model.__dict__.get("_yb_detector")
    if det is None or (det.iou_threshold, det.obj_threshold, det.box_format) != (iou_threshold, obj_threshold, box_format) \
            or (anchors is not None and det.anchors is not anchors):
        det = Detector(model, anchors, iou_threshold, obj_threshold, box_format)
        model.__dict__["_yb_detector"] = det
    res, plan = det(x)
    out = res.to_lists()
    plan.check_status()
    return out


def detect_from_heads(heads, anchors, iou_threshold, obj_threshold, box_format="center") -> NmsResult:
    """decode x3 + batched NMS on head tensors produced by ANY model: (B,3,S,S,5+nc) per scale, CUDA, any strides.
    The three scales are concatenated per image in the caller's order, as utils.py:300-309 does."""
    heads = [h if h.dtype == torch.float32 else h.float() for h in heads]
    dev, B = heads[0].device, heads[0].shape[0]
    n = sum(3 * h.shape[2] * h.shape[2] for h in heads)
    cand = torch.empty(B, n, 6, dtype=torch.float32, device=dev)
    decode_boxes_multi(heads, [_scaled_anchors(anchors, i, h.shape[2]) for i, h in enumerate(heads)], cand)
    img_off = (torch.arange(B + 1, dtype=torch.int32, device=dev) * n).contiguous()
    return batched_nms(cand.view(-1, 6), img_off, iou_threshold, obj_threshold, box_format)


def get_eval_boxes(loader, model, iou_threshold, anchors, obj_threshold, box_format="center", device=None):
    """utils.py:276-332 on the fused device pipeline: returns (all_box_predictions, all_true_boxes)
    as lists of [image_idx, cx, cy, w, h, obj, cls], image indices counted across batches.
    Like the reference it calls model.eval() first (:295) and model.train() at the end UNCONDITIONALLY (:331).
    A model of this package runs through the planned Detector (no head tensor leaves the device pipeline); any other
    callable returning the three head tensors is evaluated with the same decode + NMS kernels on its outputs."""
    device = _device(device)
    model.eval()
    native = hasattr(model, "_prepare")
    det = Detector(model, anchors, iou_threshold, obj_threshold, box_format) if native else None
    preds, trues, data_idx = [], [], 0
    for x, targets in loader:
        x = x.to(device)
        plan = None
        if native:
            res, plan = det(x)
        else:
            with torch.no_grad():
                heads = [h.to(device) for h in model(x)]
            res = detect_from_heads(heads, anchors, iou_threshold, obj_threshold, box_format)
        t2 = targets[2].to(device=device, dtype=torch.float32)
        s = t2.shape[2]
        anc = _scaled_anchors(anchors, 2, s)
        true_rows = decode_boxes(t2, anc, s, is_pred=False)  # utils.py:313-315
        kept = res.kept_rows()
        if plan is not None:
            plan.check_status()
        for b in range(x.shape[0]):
            for row in kept[b].tolist():
                preds.append([data_idx] + row)
            tb = true_rows[b]
            for row in tb[(tb[:, 4].double() > obj_threshold)].tolist():  # utils.py:326-328
                trues.append([data_idx] + row)
            data_idx += 1
    model.train()
    return preds, trues


def accuracy_counts(outs, targets, object_threshold) -> torch.Tensor:
    """Six int64 counts (correct_class, total_class, correct_obj, total_obj, correct_noobj, total_noobj) summed
    over the three scales: the reductions of utils.py:356-371 on device."""
    dev = outs[0].device
    counts = torch.zeros(6, dtype=torch.int64, device=dev)
    thr32 = float(torch.tensor(object_threshold, dtype=torch.float32))
    with torch.cuda.device(dev):
        for o, t in zip(outs, targets):
            o = o if o.dtype == torch.float32 else o.float()
            t = t.to(device=dev, dtype=torch.float32)
            B, A, S, _, Cc = o.shape
            if A != 3 or tuple(t.shape[:4]) != (B, 3, S, S):
                raise YoloB200Error(f"head {tuple(o.shape)} and target {tuple(t.shape)} do not match")
            lib.yolo_accuracy_counts(ptr(o), (C.c_int64 * 5)(*o.stride()), ptr(t), (C.c_int64 * 5)(*t.stride()), B, S,
                                     Cc - 5, thr32, ptr(counts), stream_ptr(dev))
            counts.record_stream(torch.cuda.current_stream(dev))
    return counts


def check_model_accuracy(model, loader, object_threshold):
    """utils.py:334-381: class / no-object / object accuracy over a loader; same prints, same return order, and the
    same mode changes: model.eval() first (:344), model.train() at the end unconditionally (:380)."""
    model.eval()
    native = hasattr(model, "forward_async")
    params = list(model.parameters()) if hasattr(model, "parameters") else []
    dev = params[0].device if params and params[0].is_cuda else _device()
    total = torch.zeros(6, dtype=torch.int64, device=dev)
    for x, target in loader:
        x = x.to(dev)
        if native:
            plan, heads = model.forward_async(x)
        else:
            with torch.no_grad():
                plan, heads = None, [h.to(dev) for h in model(x)]
        total += accuracy_counts(heads, target, object_threshold)
        if plan is not None:
            plan.check_status()
    c = total.to(torch.float32).cpu()
    class_accuracy = c[0] / (c[1] + 1e-16)
    noobj_accuracy = c[4] / (c[5] + 1e-16)
    obj_accuracy = c[2] / (c[3] + 1e-16)
    print(f"Class accuracy is: {(class_accuracy)*100:2f}%")
    print(f"No obj accuracy is: {(noobj_accuracy)*100:2f}%")
    print(f"Obj accuracy is: {(obj_accuracy)*100:2f}%")
    model.train()
    return class_accuracy, noobj_accuracy, obj_accuracy


# names used by BASELINE.json's north_star (upstream Aladdin-Persson spelling)
cells_to_bboxes = cells_to_boxes
intersection_over_union = calc_iou
mean_average_precision = calc_mAP
