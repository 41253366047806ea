"""GPU parity: K1/K2 (tcgen05 implicit-GEMM conv + folded BN + activation + residual + upsample store)
through the C-ABI, against (a) torch-CPU fp32 conv2d on the same bf16-rounded operands (the oracle
arithmetic) and (b) the test-only SIMT kernel of the same library.

Tolerance (stated, bf16 path): outputs are bf16, accumulation fp32.  |err| <= 2^-7 * max(1, |ref|)
(one bf16 ulp is 2^-8 relative) and cosine similarity >= 0.9999."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

REL = 2.0 ** -7


def _run_case(B, H, cin, cout, k, stride, act="leaky_relu", residual=False, upsample=False, fp32=False,
              a_mode=0, block_n=0, stages=0, in_pitch=None, out_pitch=None, seed=0, also_simt=True, impl=0, pair=0,
              pdl=0, split=0, launches=1, row=0, W=None, expect_impl=None, mc=0):
    from yolo_for_turbines_b200._lib import ACT_CODES, ConvDesc, lib, ptr, stream_ptr
    from yolo_for_turbines_b200.engine import make_conv_plan

    g = torch.Generator().manual_seed(seed)
    pad = 1 if k == 3 else 0
    cpad = (cout + 31) // 32 * 32
    in_pitch = in_pitch or cin
    out_pitch = out_pitch or cpad
    W = W or H
    x = torch.randn(B, H, W, in_pitch, generator=g).bfloat16()
    w = (torch.randn(cout, cin, k, k, generator=g) * (1.0 / (cin * k * k)) ** 0.5)
    scale = 0.5 + torch.rand(cout, generator=g)
    bias = 0.2 * torch.randn(cout, generator=g)
    Ho = (H + 2 * pad - k) // stride + 1
    Wo = (W + 2 * pad - k) // stride + 1
    res = torch.randn(B, Ho, Wo, cpad, generator=g).bfloat16() if residual else None

    # oracle arithmetic: fp32 conv on the bf16-rounded operands
    xr = x[..., :cin].float().permute(0, 3, 1, 2)
    wr = w.bfloat16().float()
    y = F.conv2d(xr, wr, None, stride, pad) * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)
    y = F.leaky_relu(y, 0.1) if act == "leaky_relu" else (F.mish(y) if act == "mish" else y)
    if residual:
        y = y + res[..., :cout].float().permute(0, 3, 1, 2)
    if upsample:
        y = F.interpolate(y, scale_factor=2, mode="nearest")
    ref = y.permute(0, 2, 3, 1).contiguous()  # NHWC

    dev = "cuda"
    xd = x.to(dev).contiguous()
    wpk = torch.zeros(cpad, k * k, cin, dtype=torch.bfloat16)
    wpk[:cout] = w.permute(0, 2, 3, 1).reshape(cout, k * k, cin).bfloat16()
    wd = wpk.to(dev).contiguous()
    sc = torch.zeros(cpad); sc[:cout] = scale
    bi = torch.zeros(cpad); bi[:cout] = bias
    sd, bd = sc.to(dev), bi.to(dev)
    rd = res.to(dev).contiguous() if residual else None
    Hy = Ho * (2 if upsample else 1)
    Wy = Wo * (2 if upsample else 1)
    odt = torch.float32 if fp32 else torch.bfloat16
    status = torch.zeros(1, dtype=torch.int32, device=dev)

    d = ConvDesc()
    d.batch, d.h_in, d.w_in, d.c_in, d.in_pitch = B, H, W, cin, in_pitch
    d.c_out, d.c_out_pad, d.out_pitch = cout, cpad, out_pitch
    d.ksize, d.stride, d.pad, d.act = k, stride, pad, ACT_CODES[act]
    d.has_residual, d.res_pitch = int(residual), cpad
    d.upsample2x, d.out_fp32, d.check_nan = int(upsample), int(fp32), 1
    d.a_mode, d.block_n_hint, d.stages_hint = a_mode, block_n, stages
    d.impl_hint, d.cta_pair_hint = impl, pair
    d.pdl_hint, d.tail_split_hint, d.row_hint, d.mc_hint = pdl, split, row, mc

    outs = {}
    yd = torch.full((B, Hy, Wy, out_pitch), 7.0, dtype=odt, device=dev)
    plan = make_conv_plan(d, ptr(xd), ptr(wd), ptr(sd), ptr(bd), ptr(rd), ptr(yd))
    if expect_impl is not None:
        info = (C.c_int32 * 8)()
        lib.yolo_conv_plan_info(plan[1], info)
        assert info[5] == expect_impl, f"plan chose impl {info[5]}, expected {expect_impl}" 
    for _ in range(launches):   # back-to-back launches of one plan: the programmatic-dependent-launch chain
        lib.yolo_conv_fwd(plan[1], ptr(status), stream_ptr())
    torch.cuda.synchronize()
    outs["tcgen05"] = yd.float().cpu()
    if also_simt:
        ys = torch.full((B, Hy, Wy, out_pitch), 7.0, dtype=odt, device=dev)
        lib.yolo_conv_fwd_simt(C.byref(d), ptr(xd), ptr(wd), ptr(sd), ptr(bd), ptr(rd), ptr(ys), ptr(status), stream_ptr())
        torch.cuda.synchronize()
        outs["simt"] = ys.float().cpu()
    assert int(status.item()) == 0
    for name, o in outs.items():
        got = o[..., :cout]
        err = (got - ref).abs()
        tol = REL * torch.clamp(ref.abs(), min=1.0)
        bad = err > tol
        cos = F.cosine_similarity(got.flatten(), ref.flatten(), dim=0)
        if bool(bad.any()) or cos < 0.9999:
            idx = torch.nonzero(bad)
            rows = torch.unique(idx[:, 0] * Hy * Wy + idx[:, 1] * Wy + idx[:, 2]) if idx.numel() else idx
            raise AssertionError(
                f"{name}: max err {float(err.max()):.4g}, cos {float(cos):.6f}, bad {int(bad.sum())}/{bad.numel()}, "
                f"first bad idx {idx[:5].tolist()}, bad pixel rows%128 {sorted(set((rows % 128).tolist()))[:16]}, "
                f"bad channels {sorted(set(idx[:, 3].tolist()))[:16]}")
        if out_pitch > cpad:  # the kernel must not touch the neighbouring channels of a concat buffer
            assert bool((o[..., cpad:] == 7.0).all()), name + " wrote outside its channel range"
    return outs


def test_conv1x1_tiled_single_kblock():
    _run_case(B=1, H=16, cin=64, cout=64, k=1, stride=1)          # M=256: 2 tiles, K=64: 1 k-block, N=64


def test_conv1x1_tiled_multi_kblock_and_m_tail():
    _run_case(B=2, H=13, cin=256, cout=128, k=1, stride=1)        # M=338 (tail), 4 k-blocks, 128-wide tile


def test_conv1x1_kc32_swizzle64():
    _run_case(B=1, H=16, cin=32, cout=32, k=1, stride=1)          # the stem GEMM shape (K=32, N=32)
    _run_case(B=1, H=24, cin=96, cout=64, k=1, stride=1)          # K=96: 3 k-blocks of 32


def test_conv1x1_im2col_mode_equals_tiled():
    _run_case(B=2, H=13, cin=128, cout=64, k=1, stride=1, a_mode=2)


def test_conv3x3_s1_im2col():
    _run_case(B=2, H=13, cin=64, cout=128, k=3, stride=1)         # padding via TMA zero fill, tiles cross images
    _run_case(B=1, H=26, cin=128, cout=256, k=3, stride=1, block_n=256)


def test_conv3x3_s2_im2col():
    _run_case(B=2, H=26, cin=64, cout=128, k=3, stride=2)
    _run_case(B=1, H=32, cin=32, cout=64, k=3, stride=2)          # kc=32 path


def test_conv_residual_mish_fp32_head():
    _run_case(B=2, H=13, cin=64, cout=128, k=3, stride=1, residual=True)
    _run_case(B=1, H=13, cin=64, cout=64, k=1, stride=1, act="mish")
    _run_case(B=2, H=13, cin=256, cout=255, k=1, stride=1, act="none", fp32=True)   # head: N padded 255->256
    _run_case(B=1, H=13, cin=128, cout=21, k=1, stride=1, act="none", fp32=True)    # nc=2 head: 21->32


def test_conv_upsample_store_and_concat_pitch():
    _run_case(B=2, H=13, cin=128, cout=64, k=1, stride=1, upsample=True, out_pitch=192)  # writes ch [0,64) of 192
    _run_case(B=1, H=26, cin=64, cout=64, k=3, stride=1, in_pitch=160, out_pitch=96)     # reads a pitched slice


def test_conv_pipeline_depths_and_tiles():
    for bn, st in ((32, 2), (64, 3), (128, 1), (128, 6), (256, 4)):
        _run_case(B=1, H=20, cin=128, cout=256, k=3, stride=1, block_n=bn, stages=st, also_simt=False)


def test_conv_darknet_shapes_batch4():
    """Every distinct (Cin, Cout, k, stride, H) of YOLOv3-416 at H/4 resolution to keep the CPU oracle fast."""
    shapes = [(32, 64, 3, 2, 104), (64, 32, 1, 1, 52), (32, 64, 3, 1, 52), (64, 128, 3, 2, 52), (128, 64, 1, 1, 26),
              (64, 128, 3, 1, 26), (128, 256, 3, 2, 26), (256, 128, 1, 1, 13), (128, 256, 3, 1, 13), (256, 512, 3, 2, 14),
              (512, 256, 1, 1, 7), (256, 512, 3, 1, 7), (512, 1024, 3, 2, 8), (1024, 512, 1, 1, 4), (512, 1024, 3, 1, 4),
              (768, 256, 1, 1, 8), (384, 128, 1, 1, 13), (1024, 255, 1, 1, 4)]
    for cin, cout, k, s, h in shapes:
        _run_case(B=4, H=h, cin=cin, cout=cout, k=k, stride=s, act="none" if cout == 255 else "leaky_relu",
                  fp32=(cout == 255), also_simt=False, seed=cin + cout)


def test_tail_split_and_pdl_switches():
    """Tail splitting (half-width tiles in a last round that is at most half full) and programmatic dependent
    launch, each on and off, on tile counts with whole rounds plus a remainder: 86 tiles on 74 CTA pairs
    (74 whole + 12 split), 170 tiles on 148 single CTAs, and a 3x3 layer with a residual."""
    for pdl, split in ((0, 0), (1, 0), (0, 1), (1, 1)):
        _run_case(B=64, H=13, cin=128, cout=512, k=1, stride=1, pdl=pdl, split=split, also_simt=False, launches=3)
        _run_case(B=64, H=13, cin=64, cout=256, k=1, stride=1, pair=1, pdl=pdl, split=split, also_simt=False)
    _run_case(B=40, H=13, cin=64, cout=512, k=3, stride=1, residual=True, also_simt=False, launches=2)
    _run_case(B=40, H=13, cin=64, cout=512, k=3, stride=1, residual=True, also_simt=False, split=1)
    _run_case(B=24, H=26, cin=64, cout=128, k=3, stride=1, act="mish", also_simt=False)   # bn 128 -> 64-wide halves


ROW_CASES = [  # B, H, W, cin, cout, stride, residual  -- shapes the row-window mode takes (w_out >= 64, weights resident)
    (2, 12, 104, 64, 128, 1, True),     # the 64->128 layers at 104^2: one 104-pixel segment per row
    (1, 10, 208, 64, 128, 1, False),    # two segments of 104 per row
    (3, 9, 152, 64, 64, 1, True),       # 608-geometry: two segments of 76; N = 64 (one epilogue box)
    (2, 16, 80, 64, 128, 2, False),     # stride 2 along H only is not a real layer shape, but exercises h0 = 2 ho - 1 ... (W stride must be 1)
    (1, 7, 128, 128, 64, 1, False),     # two 64-channel chunks per window, a full 128-pixel segment
    (5, 3, 64, 64, 128, 1, True),       # odd number of segments: the last pair is ragged
]


def test_row_window_mode(row=0):
    """Row-window mode (resident weights, one TMA window per filter row, column taps as shifted shared-memory views)
    against the fp32 oracle arithmetic, and against the im2col mode of the same kernel (row_hint = 1).  The column tap
    is a descriptor whose start address is shifted by tap * 128 B: measured on B200, tcgen05 swizzles on ABSOLUTE
    shared-memory address bits, so the plain shifted start is right and the descriptor's base_offset field must stay 0
    (row_hint = 2, which sets it, gives cosine 0.83 -- kept only as an A/B switch).  See DESIGN.md."""
    for B, H, W, cin, cout, stride, res in ROW_CASES:
        if stride == 2:
            continue   # plain 3x3/s2 has stride 2 along W: not a row-window shape (covered by the folded case below)
        a = _run_case(B=B, H=H, W=W, cin=cin, cout=cout, k=3, stride=1, residual=res, row=row, also_simt=False, expect_impl=3,
                      launches=2)
        b = _run_case(B=B, H=H, W=W, cin=cin, cout=cout, k=3, stride=1, residual=res, row=1, also_simt=False, expect_impl=2)
        assert torch.equal(a["tcgen05"], b["tcgen05"]) or float((a["tcgen05"] - b["tcgen05"]).abs().max()) <= 2.0 ** -6


def test_weight_tile_multicast_across_two_cta_pairs():
    """Clusters of two CTA pairs sharing the weight tile by TMA multicast (mc_hint = 2; opt-in, measured slower on B200
    because only 132 of 148 SMs fit clusters of 4) against the oracle arithmetic and against the plain pair kernel: odd and even numbers
    of M tiles (the second pair of the last cluster idles), 1x1 / 3x3 / stride 2, residual, tail-split tiles."""
    cases = [dict(B=64, H=13, cin=128, cout=512, k=1, stride=1),                       # 43 M tiles x 2 N tiles, split tail
             dict(B=40, H=13, cin=64, cout=256, k=3, stride=1, residual=True),         # 27 M tiles x 1
             dict(B=16, H=26, cin=64, cout=512, k=3, stride=2),                        # 11 M tiles x 2, im2col stride 2
             dict(B=24, H=26, cin=128, cout=256, k=3, stride=1, residual=True, act="mish")]   # 64 M tiles (even)
    for c in cases:
        a = _run_case(**c, also_simt=False, mc=2, launches=2)
        b = _run_case(**c, also_simt=False, mc=0)
        assert torch.equal(a["tcgen05"], b["tcgen05"]), c


def test_nan_layer_flag():
    from yolo_for_turbines_b200._lib import ConvDesc, lib, ptr, stream_ptr
    from yolo_for_turbines_b200.engine import make_conv_plan

    x = torch.zeros(1, 8, 8, 64, dtype=torch.bfloat16, device="cuda")
    x[0, 3, 3, 5] = float("nan")
    w = torch.ones(64, 1, 64, dtype=torch.bfloat16, device="cuda")
    sc, bi = torch.ones(64, device="cuda"), torch.zeros(64, device="cuda")
    y = torch.empty(1, 8, 8, 64, dtype=torch.bfloat16, device="cuda")
    st = torch.zeros(1, dtype=torch.int32, device="cuda")
    d = ConvDesc()
    d.batch, d.h_in, d.w_in, d.c_in, d.in_pitch, d.c_out, d.c_out_pad, d.out_pitch = 1, 8, 8, 64, 64, 64, 64, 64
    d.ksize, d.stride, d.pad, d.act, d.check_nan = 1, 1, 0, 1, 1
    plan = make_conv_plan(d, ptr(x), ptr(w), ptr(sc), ptr(bi), None, ptr(y))
    lib.yolo_conv_fwd(plan[1], ptr(st), stream_ptr())
    assert int(st.item()) == 2  # YB_STATUS_NAN_LAYER


def test_conv_large_grid_matches_simt():
    """The stem GEMM at the bench shape: M = 64*416*416 = 11 075 584 pixels = 86 528 M-tiles (> 65 535, so the
    grid must be 1-D).  Too big for the CPU oracle: checked against the SIMT kernel on the same operands."""
    from yolo_for_turbines_b200._lib import ConvDesc, lib, ptr, stream_ptr
    from yolo_for_turbines_b200.engine import make_conv_plan

    B, H, cin, cout = 64, 416, 32, 32
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(B, H, H, cin, generator=g, device="cuda").bfloat16()
    w = (torch.randn(cout, 1, cin, generator=g, device="cuda") * cin ** -0.5).bfloat16()
    sc, bi = torch.ones(cout, device="cuda"), torch.zeros(cout, device="cuda")
    st = torch.zeros(1, dtype=torch.int32, device="cuda")
    d = ConvDesc()
    d.batch, d.h_in, d.w_in, d.c_in, d.in_pitch, d.c_out, d.c_out_pad, d.out_pitch = B, H, H, cin, cin, cout, cout, cout
    d.ksize, d.stride, d.pad, d.act, d.check_nan = 1, 1, 0, 1, 1
    y1 = torch.empty(B, H, H, cout, dtype=torch.bfloat16, device="cuda")
    y2 = torch.empty_like(y1)
    plan = make_conv_plan(d, ptr(x), ptr(w), ptr(sc), ptr(bi), None, ptr(y1))
    lib.yolo_conv_fwd(plan[1], ptr(st), stream_ptr())
    lib.yolo_conv_fwd_simt(C.byref(d), ptr(x), ptr(w), ptr(sc), ptr(bi), None, ptr(y2), ptr(st), stream_ptr())
    torch.cuda.synchronize()
    assert int(st.item()) == 0
    diff = (y1.float() - y2.float()).abs()
    assert float(diff.max()) <= 2.0 ** -6 * max(1.0, float(y2.float().abs().max()))
    assert float((diff > 0).float().mean()) < 0.02  # only accumulation-order ulps differ


# ---- the same cases on every kernel variant: v1 (one tile per CTA), v2 persistent, v2 with a CTA pair ----
VARIANTS = [pytest.param(1, 0, id="v1"), pytest.param(2, 1, id="v2"), pytest.param(2, 2, id="v2pair")]


@pytest.mark.parametrize("impl,pair", VARIANTS)
def test_variants_1x1(impl, pair):
    _run_case(B=2, H=13, cin=256, cout=128, k=1, stride=1, impl=impl, pair=pair, also_simt=False)       # M tail 338
    _run_case(B=1, H=16, cin=32, cout=32, k=1, stride=1, impl=impl, pair=pair, also_simt=False)         # kc 32, N 32
    _run_case(B=4, H=26, cin=512, cout=256, k=1, stride=1, impl=impl, pair=pair, also_simt=False)       # many tiles per CTA
    _run_case(B=3, H=13, cin=128, cout=512, k=1, stride=1, impl=impl, pair=pair, also_simt=False, block_n=256)


@pytest.mark.parametrize("impl,pair", VARIANTS)
def test_variants_3x3(impl, pair):
    _run_case(B=2, H=13, cin=64, cout=128, k=3, stride=1, residual=True, impl=impl, pair=pair, also_simt=False)
    _run_case(B=2, H=26, cin=64, cout=128, k=3, stride=2, impl=impl, pair=pair, also_simt=False)
    _run_case(B=1, H=32, cin=32, cout=64, k=3, stride=2, impl=impl, pair=pair, also_simt=False)
    _run_case(B=5, H=13, cin=256, cout=512, k=3, stride=1, residual=True, impl=impl, pair=pair, also_simt=False)


@pytest.mark.parametrize("impl,pair", VARIANTS)
def test_variants_direct_store_paths(impl, pair):
    _run_case(B=2, H=13, cin=256, cout=255, k=1, stride=1, act="none", fp32=True, impl=impl, pair=pair, also_simt=False)
    _run_case(B=2, H=13, cin=128, cout=64, k=1, stride=1, upsample=True, out_pitch=192, impl=impl, pair=pair,
              also_simt=False)
    _run_case(B=1, H=26, cin=64, cout=64, k=3, stride=1, in_pitch=160, out_pitch=96, impl=impl, pair=pair,
              also_simt=False)
    _run_case(B=1, H=13, cin=64, cout=64, k=1, stride=1, act="mish", residual=True, impl=impl, pair=pair, also_simt=False)


@pytest.mark.parametrize("B,H,W", [(2, 16, 24), (1, 8, 256), (2, 6, 416)])
def test_rectangular_geometry_pair_folded_stride2(B, H, W):
    """The engine runs the 32->64 3x3/s2 layer on input pixel PAIRS: a 3x2 filter, stride (2,1), left pad 1, right
    pad 0, Cin' = 64.  Checked against the plain 3x3/s2/p1 convolution on the unfolded tensor.  The wide cases take
    the row-window mode of the persistent kernel (one segment of 128 pairs / two of 104), the narrow one im2col."""
    from yolo_for_turbines_b200._lib import ConvDesc, lib, ptr, stream_ptr
    from yolo_for_turbines_b200.engine import make_conv_plan

    g = torch.Generator().manual_seed(5)
    I, O = 32, 64
    x = torch.randn(B, H, W, I, generator=g).bfloat16()
    w = (torch.randn(O, I, 3, 3, generator=g) * (I * 9) ** -0.5).bfloat16().float()
    ref = F.leaky_relu(F.conv2d(x.float().permute(0, 3, 1, 2), w, None, 2, 1), 0.1).permute(0, 2, 3, 1)
    w2 = torch.zeros(O, 2 * I, 3, 2)
    w2[:, I:, :, 0] = w[:, :, :, 0]
    w2[:, :I, :, 1] = w[:, :, :, 1]
    w2[:, I:, :, 1] = w[:, :, :, 2]
    wd = w2.permute(0, 2, 3, 1).contiguous().bfloat16().cuda()
    xd = x.cuda()
    sc, bi = torch.ones(O, device="cuda"), torch.zeros(O, device="cuda")
    st = torch.zeros(1, dtype=torch.int32, device="cuda")
    wide = W // 2 >= 64
    for impl, row in ((1, 0), (2, 1), (2, 0)):
        y = torch.zeros(B, H // 2, W // 2, O, dtype=torch.bfloat16, device="cuda")
        d = ConvDesc()
        d.batch, d.h_in, d.w_in, d.c_in, d.in_pitch, d.c_out, d.c_out_pad, d.out_pitch = B, H, W // 2, 2 * I, 2 * I, O, O, O
        d.ksize, d.stride, d.pad, d.act, d.impl_hint, d.row_hint = 3, 2, 1, 1, impl, row
        d.ksize_w, d.stride_w, d.pad_w_hi_plus1 = 2, 1, 1
        plan = make_conv_plan(d, ptr(xd), ptr(wd), ptr(sc), ptr(bi), None, ptr(y))
        info = (C.c_int32 * 8)()
        lib.yolo_conv_plan_info(plan[1], info)
        assert info[5] == (3 if (impl == 2 and row == 0 and wide) else impl), (impl, row, info[5])
        lib.yolo_conv_fwd(plan[1], ptr(st), stream_ptr())
        torch.cuda.synchronize()
        err = (y.float().cpu() - ref).abs()
        assert float(err.max()) <= REL * max(1.0, float(ref.abs().max())), (impl, row, float(err.max()))
