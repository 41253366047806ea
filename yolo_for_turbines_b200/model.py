"""Drop-in mirror of the reference's `code/model.py` module surface on the sm_100a kernels.

Same class names, constructor signatures, attribute names and `state_dict()` keys as the reference
(CNNBlock model.py:47-86, ResidualBlock :88-121, ScalePredictionBlock :123-148, YOLOv3 :150-337), so
checkpoints and Darknet `.weights` files load unchanged.  The nn.Conv2d / nn.BatchNorm2d children
are kept as PARAMETER CONTAINERS only: `forward` never calls them.  It runs the fused NHWC bf16
tcgen05 kernels through `engine.Engine` and fails loudly (no CPU or ATen fallback) when the input is
not on a CUDA device or libyolo_b200.so is missing.

Deviations, all deliberate:
  * train mode (`model.train()`): the whole network is one autograd node (train._TrainForwardFn) running the
    batch-statistics forward and the hand-written backward.  Standalone blocks called in train mode (the
    reference's own unit tests do, model_tests.py:16-45) normalise with batch statistics and update the running
    ones like nn.BatchNorm2d, but their output carries no autograd graph: the training entry points are
    YOLOv3.forward and train.Trainer.
  * the reference's 28 per-layer NaN syncs (model.py:175,183) are one device status word checked
    once per forward; the same AssertionError / ValueError("Nan in layer") are raised.
  * compute is bf16 with fp32 accumulation; head outputs are returned as fp32.
"""
from __future__ import annotations

import os

import numpy as np
import torch
import torch.nn as nn

from ._lib import YoloB200Error, lib, ptr, require_cuda, stream_ptr

# (filters, kernel, stride) | ["B", repeats] | "S" scale head | "U" upsample+concat -- the reference's
# architecture table (model.py:20-45); Darknet-53 ends after ["B", 4].
layer_config = [
    (32, 3, 1), (64, 3, 2), ["B", 1], (128, 3, 2), ["B", 2], (256, 3, 2), ["B", 8], (512, 3, 2), ["B", 8],
    (1024, 3, 2), ["B", 4],
    (512, 1, 1), (1024, 3, 1), "S", (256, 1, 1), "U", (256, 1, 1), (512, 3, 1), "S", (128, 1, 1), "U",
    (128, 1, 1), (256, 3, 1), "S",
]

_ACTIVATIONS = {"leaky_relu": lambda: nn.LeakyReLU(0.1), "mish": nn.Mish}


def _run_blocks_standalone(owner: nn.Module, x: torch.Tensor, program):
    """Runs a small list of CNNBlocks on NCHW fp32 input (module-level drop-in: the reference's
    unit tests call blocks directly, model_tests.py:16-45).  `program(conv)` receives a callable
    conv(block, act, residual=None) working on NHWC bf16 activations and returns the result."""
    from . import engine as E

    require_cuda(x, f"{type(owner).__name__} input")
    eng = getattr(owner, "_yb_engine", None)
    if eng is None or eng.device != x.device:
        eng = E.Engine(owner, x.device)
        object.__setattr__(owner, "_yb_engine", eng)
    eng.refresh_if_needed()
    B, Cin, H, W = x.shape
    x = x.contiguous().float()
    st = stream_ptr(x.device)
    status = torch.zeros(1, dtype=torch.int32, device=x.device)
    cpad = E._round_up(Cin, 32)
    a0 = torch.empty(B * H * W * cpad, dtype=torch.bfloat16, device=x.device)
    lib.yolo_nchw_to_nhwc_bf16(ptr(x), B, Cin, H, W, cpad, cpad, ptr(a0), ptr(status), st)
    keep = []  # plans/buffers must outlive the async launches

    def conv(block, act, residual=None, fp32=False):
        t, C_, h, w = act
        pc = eng.packed[id(block)]
        if pc.stem:  # standalone blocks take the generic route: repack without the stem trick
            pc = E.PackedConv(block, x.device, as_stem=False)
            pc.refresh()
        if C_ != pc.c_in_eff:
            raise YoloB200Error(f"expected {pc.c_in_eff} (padded) input channels, got {C_}")
        ho = (h + 2 * pc.pad_eff - pc.k_eff) // pc.stride_eff + 1
        wo = (w + 2 * pc.pad_eff - pc.k_eff) // pc.stride_eff + 1
        y = torch.empty(B * ho * wo * pc.c_out_pad, dtype=torch.float32 if fp32 else torch.bfloat16, device=x.device)
        bn_train = bool(block.batch_norm_act and block.batch_norm.training)
        d = E.ConvDesc()
        d.batch, d.h_in, d.w_in, d.c_in, d.in_pitch = B, h, w, pc.c_in_eff, C_
        d.c_out, d.c_out_pad, d.out_pitch = pc.c_out, pc.c_out_pad, pc.c_out_pad
        d.ksize, d.stride, d.pad, d.act = pc.k_eff, pc.stride_eff, pc.pad_eff, E.ACT_CODES[pc.act]
        d.out_fp32, d.check_nan = int(fp32), 0
        if residual is not None and not bn_train:
            d.has_residual, d.res_pitch = 1, residual[1]
        if not bn_train:
            plan = E.make_conv_plan(d, ptr(t), ptr(pc.w), ptr(pc.scale), ptr(pc.bias),
                                    ptr(residual[0]) if residual is not None else None, ptr(y))
            lib.yolo_conv_fwd(plan[1], ptr(status), st)
            keep.append((plan, pc, t, y))
            return (y, pc.c_out_pad, ho, wo), pc.c_out
        # train mode (nn.BatchNorm2d with batch statistics, model.py:61 under .train()): raw conv output z, its batch
        # statistics (running ones updated with the module's momentum), then BN + activation (+ residual)
        bn, Cp, P = block.batch_norm, pc.c_out_pad, B * ho * wo
        d.act = 0
        dev = x.device
        ones, zeros = torch.ones(Cp, device=dev), torch.zeros(Cp, device=dev)
        z = torch.empty(P * Cp, dtype=torch.bfloat16, device=dev)
        plan = E.make_conv_plan(d, ptr(t), ptr(pc.w), ptr(ones), ptr(zeros), None, ptr(z))
        lib.yolo_conv_fwd(plan[1], ptr(status), st)
        sums = torch.zeros(2 * Cp, dtype=torch.float64, device=dev)
        counter = torch.zeros(1, dtype=torch.int32, device=dev)
        f32 = lambda v: v.detach().to(device=dev, dtype=torch.float32).contiguous()  # noqa: E731
        gamma, beta = f32(bn.weight), f32(bn.bias)
        mean, rstd, sc, bi = (torch.empty(Cp, device=dev) for _ in range(4))
        momentum = 0.1 if bn.momentum is None else float(bn.momentum)
        track = bn.track_running_stats and bn.running_mean is not None
        rm = bn.running_mean if track else None
        rv = bn.running_var if track else None
        lib.yolo_bn_stats_finalize(ptr(z), P, pc.c_out, Cp, ptr(sums), ptr(counter), ptr(gamma), ptr(beta), float(bn.eps),
                                   momentum, ptr(rm), ptr(rv), ptr(mean), ptr(rstd), ptr(sc), ptr(bi), st)
        if track:
            bn.num_batches_tracked += 1
        lib.yolo_bn_act_fwd(ptr(z), P, pc.c_out, Cp, ptr(sc), ptr(bi), E.ACT_CODES[pc.act],
                            ptr(residual[0]) if residual is not None else None, residual[1] if residual is not None else 0,
                            ptr(y), Cp, 0, ho, wo, st)
        keep.append((plan, pc, t, y, z, sums, counter, gamma, beta, mean, rstd, sc, bi, ones, zeros))
        return (y, pc.c_out_pad, ho, wo), pc.c_out

    (y, cp, ho, wo), c_out, fp32 = program(conv, (a0, cpad, H, W))
    out = torch.empty(B, c_out, ho, wo, dtype=torch.float32, device=x.device)
    lib.yolo_nhwc_to_nchw_f32(ptr(y), int(fp32), B, c_out, ho, wo, cp, ptr(out), st)
    if int(status.item()) & 1:
        raise AssertionError("NaN in the input tensor")
    del keep
    return out


class _EngineHolder:
    """The cached engine holds ctypes plan blobs and CUDA graphs: never pickled / deep-copied."""

    def __getstate__(self):
        d = self.__dict__.copy()
        d.pop("_yb_engine", None)
        d.pop("_yb_train", None)
        return d


class CNNBlock(_EngineHolder, nn.Module):
    """Conv2d -> BatchNorm2d -> LeakyReLU(0.1)|Mish, or a bare biased Conv2d (model.py:47-86)."""

    def __init__(self, in_channels, out_channels, batch_norm_act=True, activation="leaky_relu", **kwargs):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, bias=not batch_norm_act, **kwargs)
        self.batch_norm = nn.BatchNorm2d(out_channels) if batch_norm_act else None
        self.activation = None
        if batch_norm_act:
            if activation not in _ACTIVATIONS:
                raise ValueError(f"Unsupported activation: {activation}")
            self.activation = _ACTIVATIONS[activation]()
        self.batch_norm_act = batch_norm_act

    def set_layers(self, layers):
        self.conv = layers[0]
        if self.batch_norm_act:
            self.batch_norm, self.activation = layers[1], layers[2]

    def forward(self, x):
        def program(conv, a):
            r, c_out = conv(self, a)
            return r, c_out, False
        return _run_blocks_standalone(self, x, program)


class ResidualBlock(_EngineHolder, nn.Module):
    """num_blocks x [1x1 C->C/2, 3x3 C/2->C] with optional skip add (model.py:88-121)."""

    def __init__(self, in_channels, activation="leaky_relu", use_residual=True, num_blocks=1):
        super().__init__()
        self.layers = nn.ModuleList(
            nn.Sequential(CNNBlock(in_channels, in_channels // 2, activation=activation, kernel_size=1),
                          CNNBlock(in_channels // 2, in_channels, activation=activation, kernel_size=3, padding=1))
            for _ in range(num_blocks))
        self.use_residual = use_residual
        self.num_blocks = num_blocks

    def set_layers(self, layers):
        self.layers = layers

    def forward(self, x):
        def program(conv, a):
            c_out = a[1]
            for seq in self.layers:
                t, _ = conv(seq[0], a)
                a, c_out = conv(seq[1], t, residual=(a[0], a[1]) if self.use_residual else None)
            return a, c_out, False
        return _run_blocks_standalone(self, x, program)


class ScalePredictionBlock(_EngineHolder, nn.Module):
    """3x3 C->2C then 1x1 2C->3*(nc+5) with bias; output (B, 3, S, S, nc+5) (model.py:123-148)."""

    def __init__(self, in_channels, num_classes, activation="leaky_relu", anchors_per_scale=3):
        super().__init__()
        self.pred_block = nn.Sequential(
            CNNBlock(in_channels, in_channels * 2, activation=activation, kernel_size=3, padding=1),
            CNNBlock(2 * in_channels, (num_classes + 5) * anchors_per_scale, activation=activation,
                     batch_norm_act=False, kernel_size=1))
        self.num_classes = num_classes
        self.anchors_per_scale = anchors_per_scale

    def set_layers(self, layers):
        self.pred_block = layers

    def forward(self, x):
        def program(conv, a):
            t, _ = conv(self.pred_block[0], a)
            r, c_out = conv(self.pred_block[1], t, fp32=True)
            return r, c_out, True
        y = _run_blocks_standalone(self, x, program)  # (B, 3*(nc+5), S, S)
        y = y.reshape(y.shape[0], self.anchors_per_scale, self.num_classes + 5, y.shape[2], y.shape[3])
        return y.permute(0, 1, 3, 4, 2)


class YOLOv3(_EngineHolder, nn.Module):
    """Darknet-53 backbone + 3-scale head (model.py:150-337)."""

    def __init__(self, in_channels=3, num_classes=80, activation="leaky_relu", weights_path=None, freeze=False):
        super().__init__()
        self.in_channels = in_channels
        self.num_classes = num_classes
        self.activation = activation
        self.layers = self._create_model_layers()
        self.param_idx = 0
        self.layer_id = 0
        self.weights_path = None
        self.freeze = freeze
        if weights_path:  # model.py:162-170
            self.weights_path = weights_path
            with open(weights_path, "rb") as f:
                np.fromfile(f, dtype=np.int32, count=5)  # major, minor, revision, seen(2 x int32)
                self.weights = np.fromfile(f, dtype=np.float32)
            name = os.path.basename(weights_path)
            self.cutoff = int(name.split(".")[-1]) if ".conv" in name else None

    # -- construction ---------------------------------------------------------------------------
    def _create_model_layers(self):
        layers = nn.ModuleList()
        c = self.in_channels
        act = self.activation
        for item in layer_config:
            if isinstance(item, tuple):
                out_c, k, s = item
                layers.append(CNNBlock(c, out_c, activation=act, kernel_size=k, stride=s, padding=1 if k == 3 else 0))
                c = out_c
            elif isinstance(item, list):
                layers.append(ResidualBlock(c, activation=act, num_blocks=item[1]))
            elif item == "S":
                layers.append(ResidualBlock(c, activation=act, use_residual=False, num_blocks=1))
                layers.append(CNNBlock(c, c // 2, activation=act, kernel_size=1))
                layers.append(ScalePredictionBlock(c // 2, num_classes=self.num_classes, activation=act))
                c //= 2
            elif item == "U":
                layers.append(nn.Upsample(scale_factor=2))
                c *= 3  # concatenated with a route that has twice the channels (model.py:223)
        return layers

    # -- forward --------------------------------------------------------------------------------
    def _engine(self, device):
        from .engine import Engine

        eng = self.__dict__.get("_yb_engine")
        if eng is None or eng.device != device:
            p = next(self.parameters())
            if p.device != device:
                raise YoloB200Error(f"model parameters are on {p.device} but the input is on {device}")
            eng = Engine(self, device)
            self.__dict__["_yb_engine"] = eng
        return eng

    def _prepare(self, x, lane: int = 0):
        """Input checks + (cached) plan lookup shared by forward_async and utils.Detector."""
        require_cuda(x, "YOLOv3 input")
        if self.training:
            raise YoloB200Error("the planned inference pipeline (forward_async / utils.Detector) needs model.eval(); "
                                "in train mode call model(x) (autograd) or train.Trainer.step")
        if x.dim() != 4 or x.shape[1] != self.in_channels:
            raise YoloB200Error(f"expected (B,{self.in_channels},H,W) input, got {tuple(x.shape)}")
        eng = self._engine(x.device)
        eng.refresh_if_needed()
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.float().contiguous()
        return eng.plan(x.shape[0], x.shape[2], x.shape[3], lane), x

    def forward_async(self, x):
        """Enqueues the whole forward on the current stream and returns (plan, head views) WITHOUT
        the host sync; callers must eventually call plan.check_status().  Views alias static buffers
        that the next forward of the same shape overwrites."""
        plan, x = self._prepare(x)
        with torch.cuda.device(x.device):
            plan.run(x)
        return plan, plan.head_views()

    def _train_session(self, device):
        from .train import Trainer

        sess = self.__dict__.get("_yb_train")
        if sess is None or sess.device != device:
            from . import config as cfg
            sess = Trainer(self, cfg.ANCHORS, lr=0.0, own_params=False)
            self.__dict__["_yb_train"] = sess
        return sess

    def forward(self, x):
        if self.training:
            # model.train() (train.py:38): batch-statistics BatchNorm, outputs attached to autograd so that the
            # reference's loss.backward() / optimizer.step() loop runs unchanged (see train._TrainForwardFn)
            require_cuda(x, "YOLOv3 input")
            if x.dim() != 4 or x.shape[1] != self.in_channels:
                raise YoloB200Error(f"expected (B,{self.in_channels},H,W) input, got {tuple(x.shape)}")
            return self._train_session(x.device).autograd_forward(x)
        plan, views = self.forward_async(x)
        outs = [v.clone(memory_format=torch.preserve_format) for v in views]
        plan.check_status()  # AssertionError on NaN input / ValueError("Nan in layer"), model.py:175,184
        return outs

    # -- Darknet weights (model.py:227-337) --------------------------------------------------------
    def _darknet_modules(self):
        """nn.BatchNorm2d / nn.Conv2d / nn.Upsample in the order the flat file stores them."""
        for layer in self.layers:
            if isinstance(layer, nn.Upsample):
                yield layer
                continue
            for blk in layer.modules():
                if isinstance(blk, CNNBlock):
                    if blk.batch_norm_act:
                        yield blk.batch_norm  # beta, gamma, mean, var come BEFORE the conv weights
                    yield blk.conv

    def load_weights(self):
        flat = self.weights  # AttributeError when constructed without weights_path, like the reference
        for m in self._darknet_modules():
            # `layer_id` counts Conv2d, BatchNorm2d AND Upsample (model.py:336), so "darknet53.conv.74"
            # stops after 37 conv blocks; skipped tensors still advance the file cursor (model.py:277-291).
            skip = self.cutoff is not None and self.layer_id >= self.cutoff
            if isinstance(m, nn.BatchNorm2d):
                slots = [m.bias, m.weight, m.running_mean, m.running_var]
            elif isinstance(m, nn.Conv2d):
                slots = ([m.bias] if m.bias is not None else []) + [m.weight]
            else:
                slots = []
            for t in slots:
                n = t.numel()
                if not skip:
                    with torch.no_grad():  # an in-place copy THROUGH the tensor: bumps its version counter, which
                        # is what Engine.refresh_if_needed / Trainer.repack_if_changed watch (t.data.copy_ does not)
                        t.copy_(torch.from_numpy(flat[self.param_idx:self.param_idx + n]).view_as(t))
                    if self.freeze:
                        t.requires_grad = False
                self.param_idx += n
            self.layer_id += 1
        self.invalidate_packed_weights()
        print(f"Weights from {self.weights_path} loaded successfully.")

    def invalidate_packed_weights(self):
        """Forces the bf16 weight packs / folded BatchNorm vectors (inference engine and training session) to be
        rebuilt on the next forward.  Needed after writes that bypass the tensors' version counters (`p.data...`)."""
        eng = self.__dict__.get("_yb_engine")
        if eng is not None:
            eng._sig = None
        sess = self.__dict__.get("_yb_train")
        if sess is not None and hasattr(sess, "_packed_sig"):
            sess._packed_sig = None
