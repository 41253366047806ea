"""GPU parity of the training-step kernels through the C-ABI, each against plain PyTorch fp32 autograd of the
same op on the same bf16-rounded operands (what the reference's `loss.backward()` computes, train.py:67):

  * yolo_wgrad (tcgen05, MN-major operands, split-K)   vs  d conv2d / d weight
  * data gradient = yolo_conv_fwd on dz with transposed/flipped weights (stride 2: zero-stuffed dz)
                                                        vs  d conv2d / d input
  * yolo_bn_stats/finalize/act_fwd and yolo_bn_act_bwd  vs  F.batch_norm(training=True) + activation autograd
  * yolo_loss_bwd                                       vs  autograd of the oracle restatement of loss.py
  * yolo_sgd_step                                       vs  torch.optim.SGD

Tolerances (stated): fp32 accumulation of bf16 products, so wgrad / dgrad match to 2e-3 relative of the tensor's
max (wgrad outputs stay fp32; dgrad outputs are bf16-rounded: 2^-7 relative); BN outputs are bf16: 2^-7 relative."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _blob(nbytes):
    from yolo_for_turbines_b200.engine import _aligned_blob
    return _aligned_blob(int(nbytes))


def _desc(B, H, W, cin, in_pitch, cout_pad, k, stride):
    from yolo_for_turbines_b200._lib import ConvDesc
    d = ConvDesc()
    d.batch, d.h_in, d.w_in, d.c_in, d.in_pitch = B, H, W, cin, in_pitch
    d.c_out, d.c_out_pad, d.out_pitch = cout_pad, cout_pad, cout_pad
    d.ksize, d.stride, d.pad = k, stride, 1 if k == 3 else 0
    return d


@pytest.mark.parametrize("B,H,cin,cout,k,stride,in_pitch,splits", [
    (2, 16, 64, 128, 3, 1, None, 0),
    (3, 16, 128, 256, 3, 2, None, 0),
    (2, 8, 256, 128, 1, 1, None, 0),
    (2, 32, 32, 64, 3, 2, None, 0),      # 64-byte x rows (SWIZZLE_64B)
    (2, 16, 32, 32, 1, 1, None, 3),      # the stem GEMM: both operands 32 wide
    (2, 8, 384, 128, 1, 1, None, 0),     # concat consumer
    (2, 8, 128, 256, 1, 1, 384, 0),      # channel slice of a wider buffer
    (4, 4, 1024, 255, 1, 1, None, 0),    # head conv, 255 -> 256 rows
    (2, 4, 512, 21, 1, 1, None, 0),      # turbine head, 21 -> 32 rows
    (2, 13, 512, 1024, 3, 1, None, 0),   # 13x13 rows: chunks cross image boundaries
    (1, 5, 64, 64, 3, 1, None, 1),       # ragged last chunk (25 pixels)
    (2, 12, 64, 128, 3, 2, None, 5),
])
def test_wgrad_matches_autograd(B, H, cin, cout, k, stride, in_pitch, splits):
    from yolo_for_turbines_b200._lib import lib, ptr, stream_ptr

    g = torch.Generator().manual_seed(cin * 7 + cout + k + stride)
    pad = 1 if k == 3 else 0
    cpad = (cout + 31) // 32 * 32
    in_pitch = in_pitch or cin
    Ho = (H + 2 * pad - k) // stride + 1
    x = torch.randn(B, H, H, in_pitch, generator=g).bfloat16()
    dz = torch.zeros(B, Ho, Ho, cpad).bfloat16()
    dz[..., :cout] = (torch.randn(B, Ho, Ho, cout, generator=g) * 0.5).bfloat16()

    w = torch.zeros(cout, cin, k, k, requires_grad=True)
    y = F.conv2d(x[..., :cin].float().permute(0, 3, 1, 2), w, None, stride, pad)
    y.backward(dz[..., :cout].float().permute(0, 3, 1, 2))
    ref = w.grad

    dev = "cuda"
    xd, dzd = x.to(dev).contiguous(), dz.to(dev).contiguous()
    packed = torch.zeros(cpad, k * k, cin, dtype=torch.float32, device=dev)
    d = _desc(B, H, H, cin, in_pitch, cpad, k, stride)
    raw, plan = _blob(lib.yolo_wgrad_plan_bytes())
    lib.yolo_wgrad_plan_init(plan, lib.yolo_wgrad_plan_bytes(), C.byref(d), ptr(xd), ptr(dzd), cpad, ptr(packed), splits)
    st = stream_ptr(torch.device(dev))
    lib.yolo_wgrad(plan, st)
    grad = torch.empty(cout, cin, k, k, dtype=torch.float32, device=dev)
    lib.yolo_unpack_wgrad(ptr(packed), cout, cin, k, cin, 0, ptr(grad), st)
    torch.cuda.synchronize()
    got = grad.cpu()
    tol = 2e-3 * max(1.0, float(ref.abs().max()))
    assert float((got - ref).abs().max()) <= tol, (float((got - ref).abs().max()), tol)
    # rows beyond c_out of the packed gradient only ever see zero dz
    assert float(packed[cout:].abs().max() if cpad > cout else 0.0) == 0.0
    # accumulation semantics: a second launch doubles the result
    lib.yolo_wgrad(plan, st)
    lib.yolo_unpack_wgrad(ptr(packed), cout, cin, k, cin, 0, ptr(grad), st)
    torch.cuda.synchronize()
    assert float((grad.cpu() - 2 * ref).abs().max()) <= 2 * tol


@pytest.mark.parametrize("B,H,cin,cout,k,stride", [
    (2, 16, 64, 128, 3, 1), (2, 8, 256, 128, 1, 1), (2, 16, 64, 128, 3, 2), (2, 8, 512, 255, 1, 1), (3, 12, 32, 64, 3, 2),
])
def test_dgrad_via_forward_kernel(B, H, cin, cout, k, stride):
    """dX = conv(dz, W^T flipped): the forward tcgen05 kernel with the transposed weight pack; stride-2 layers
    run on the zero-stuffed dz (written here with torch; the product writes it in yolo_bn_act_bwd)."""
    from yolo_for_turbines_b200._lib import lib, ptr, stream_ptr
    from yolo_for_turbines_b200.engine import make_conv_plan

    g = torch.Generator().manual_seed(11 + cin + cout + stride)
    pad = 1 if k == 3 else 0
    cpad = (cout + 31) // 32 * 32
    Ho = (H + 2 * pad - k) // stride + 1
    w = (torch.randn(cout, cin, k, k, generator=g) * (1.0 / (cout * k * k)) ** 0.5).bfloat16().float()
    dz = torch.zeros(B, Ho, Ho, cpad).bfloat16()
    dz[..., :cout] = torch.randn(B, Ho, Ho, cout, generator=g).bfloat16()
    x = torch.zeros(B, cin, H, H, requires_grad=True)
    y = F.conv2d(x, w, None, stride, pad)
    y.backward(dz[..., :cout].float().permute(0, 3, 1, 2))
    ref = x.grad.permute(0, 2, 3, 1).contiguous()

    dev = torch.device("cuda")
    src = dz
    if stride == 2:
        src = torch.zeros(B, H, H, cpad).bfloat16()
        src[:, ::2, ::2] = dz
    srcd = src.to(dev).contiguous()
    wpk = torch.empty(cin * k * k * cpad, dtype=torch.bfloat16, device=dev)
    st = stream_ptr(dev)
    wdev = w.to(dev).contiguous()
    lib.yolo_pack_weights_dgrad(ptr(wdev), cout, cin, k, cin, cpad, ptr(wpk), st)
    ones, zeros = torch.ones(cin, device=dev), torch.zeros(cin, device=dev)
    out = torch.empty(B, H, H, cin, dtype=torch.bfloat16, device=dev)
    d = _desc(B, src.shape[1], src.shape[2], cpad, cpad, cin, k, 1)
    d.c_out = cin
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    plan = make_conv_plan(d, ptr(srcd), ptr(wpk), ptr(ones), ptr(zeros), None, ptr(out))
    lib.yolo_conv_fwd(plan[1], ptr(status), st)
    torch.cuda.synchronize()
    got = out.float().cpu()
    tol = 2.0 ** -7 * max(1.0, float(ref.abs().max()))
    assert float((got - ref).abs().max()) <= tol


@pytest.mark.parametrize("act", ["leaky_relu", "mish"])
@pytest.mark.parametrize("B,h,C,residual,up2x", [(2, 6, 64, False, False), (3, 5, 384, True, False), (2, 4, 256, False, True),
                                                 (4, 13, 1024, True, False), (2, 8, 32, False, False)])
def test_bn_act_forward_backward(act, B, h, C, residual, up2x):
    from yolo_for_turbines_b200._lib import ACT_CODES, lib, ptr, stream_ptr

    g = torch.Generator().manual_seed(C + h)
    P = B * h * h
    z = (torch.randn(P, C, generator=g) * 1.5 + 0.3).bfloat16()
    gamma = 0.5 + torch.rand(C, generator=g)
    beta = 0.3 * torch.randn(C, generator=g)
    res = torch.randn(P, C, generator=g).bfloat16() if residual else None
    oh = 2 * h if up2x else h
    dA = torch.randn(B * oh * oh, C, generator=g).bfloat16()
    rm0, rv0 = torch.zeros(C), torch.ones(C)

    # reference: nn.BatchNorm2d training semantics + activation (+ residual, + upsample) under autograd
    zr = z.float().view(B, h, h, C).permute(0, 3, 1, 2).clone().requires_grad_(True)
    gr, br = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    rm, rv = rm0.clone(), rv0.clone()
    y = F.batch_norm(zr, rm, rv, gr, br, True, 0.1, 1e-5)
    y = F.leaky_relu(y, 0.1) if act == "leaky_relu" else F.mish(y)
    if residual:
        y = y + res.float().view(B, h, h, C).permute(0, 3, 1, 2)
    if up2x:
        y = F.interpolate(y, scale_factor=2, mode="nearest")
    y.backward(dA.float().view(B, oh, oh, C).permute(0, 3, 1, 2))
    ref_y = y.detach().permute(0, 2, 3, 1).reshape(-1, C)
    ref_dz = zr.grad.permute(0, 2, 3, 1).reshape(-1, C)

    dev = torch.device("cuda")
    st = stream_ptr(dev)
    f = lambda t: t.to(dev).contiguous()  # noqa: E731
    zd, gd, bd, dAd = f(z), f(gamma), f(beta), f(dA)
    rmd, rvd = f(rm0), f(rv0)
    resd = f(res) if residual else None
    sums = torch.zeros(2 * C, dtype=torch.float64, device=dev)
    mean, rstd, scale, bias = (torch.empty(C, device=dev) for _ in range(4))
    yd = torch.empty(B * oh * oh, C, dtype=torch.bfloat16, device=dev)
    counter = torch.zeros(2, dtype=torch.int32, device=dev)
    if C % 64:   # both forms of the statistics pass: two launches / one launch with the last-block finalize
        lib.yolo_bn_stats(ptr(zd), P, C, C, ptr(sums), st)
        lib.yolo_bn_finalize(ptr(sums), P, C, ptr(gd), ptr(bd), 1e-5, 0.1, ptr(rmd), ptr(rvd), ptr(mean), ptr(rstd), ptr(scale),
                             ptr(bias), st)
    else:
        lib.yolo_bn_stats_finalize(ptr(zd), P, C, C, ptr(sums), ptr(counter), ptr(gd), ptr(bd), 1e-5, 0.1, ptr(rmd), ptr(rvd),
                                   ptr(mean), ptr(rstd), ptr(scale), ptr(bias), st)
    lib.yolo_bn_act_fwd(ptr(zd), P, C, C, ptr(scale), ptr(bias), ACT_CODES[act], ptr(resd), C, ptr(yd), C, int(up2x), h, h, st)
    sums2 = torch.zeros(2 * C, dtype=torch.float64, device=dev)
    dgam, dbet = torch.empty(C, device=dev), torch.empty(C, device=dev)
    m1m2 = torch.empty(2 * C, device=dev)
    dz = torch.empty(P, C, dtype=torch.bfloat16, device=dev)
    stuffed = torch.full((B * 4 * h * h, C), 7.0, dtype=torch.bfloat16, device=dev)
    lib.yolo_bn_act_bwd(ptr(dAd), C, int(up2x), ptr(zd), C, P, C, h, h, ptr(scale), ptr(bias), ptr(mean), ptr(rstd),
                        ACT_CODES[act], ptr(sums2), ptr(counter[1:]), ptr(dgam), ptr(dbet), ptr(m1m2), ptr(dz), C,
                        ptr(stuffed), C, st)
    torch.cuda.synchronize()
    assert int(counter.abs().sum()) == 0   # tickets reset themselves
    rel = 2.0 ** -7
    assert torch.allclose(rmd.cpu(), rm, atol=1e-5) and torch.allclose(rvd.cpu(), rv, rtol=1e-4, atol=1e-5)
    assert float((yd.float().cpu() - ref_y).abs().max()) <= rel * max(1.0, float(ref_y.abs().max()))
    assert float((dz.float().cpu() - ref_dz).abs().max()) <= 2 * rel * max(1.0, float(ref_dz.abs().max()))
    assert torch.allclose(dgam.cpu(), gr.grad, rtol=2e-3, atol=2e-3 * float(gr.grad.abs().max()))
    assert torch.allclose(dbet.cpu(), br.grad, rtol=2e-3, atol=2e-3 * float(br.grad.abs().max()))
    s = stuffed.view(B, 2 * h, 2 * h, C)
    assert torch.equal(s[:, ::2, ::2].reshape(-1, C), dz) and float(s[:, 1::2].abs().max()) == 0.0 \
        and float(s[:, :, 1::2].abs().max()) == 0.0


@pytest.mark.parametrize("nc,S,B,bf16", [(2, 13, 4, False), (80, 8, 2, False), (2, 19, 3, True)])
def test_loss_backward_matches_autograd(nc, S, B, bf16):
    from oracle import yolo_oracle as orc
    from yolo_for_turbines_b200._lib import lib, ptr, stream_ptr

    g = torch.Generator().manual_seed(nc + S)
    pred = torch.randn(B, 3, S, S, 5 + nc, generator=g)
    tgt = torch.zeros(B, 3, S, S, 6)
    sel = torch.rand(B, 3, S, S, generator=g)
    tgt[..., 4] = torch.where(sel < 0.05, 1.0, torch.where(sel < 0.08, -1.0, 0.0))
    tgt[..., 0:2] = torch.rand(B, 3, S, S, 2, generator=g)
    tgt[..., 2:4] = 0.5 + 3.5 * torch.rand(B, 3, S, S, 2, generator=g)
    tgt[..., 5] = torch.randint(0, nc, (B, 3, S, S), generator=g).float()
    anchors = torch.tensor(orc.TURBINE_ANCHORS[1]) * S

    p = pred.clone().requires_grad_(True)
    terms = orc.yolo_loss(p * 1.0, tgt.clone(), anchors)  # `* 1.0`: the restatement mutates its input like loss.py:71
    sum(terms).backward()
    ref = p.grad

    dev = torch.device("cuda")
    st = stream_ptr(dev)
    pd, td = pred.to(dev), tgt.to(dev)
    sums = torch.zeros(6, dtype=torch.float64, device=dev)
    anc = (C.c_float * 6)(*anchors.reshape(-1).tolist())
    ps, ts = (C.c_int64 * 5)(*pd.stride()), (C.c_int64 * 5)(*td.stride())
    lib.yolo_loss_fwd(ptr(pd), ps, ptr(td), ts, B, S, nc, anc, 0, ptr(sums), st)
    dp = torch.empty(pd.shape, dtype=torch.bfloat16 if bf16 else torch.float32, device=dev)
    lib.yolo_loss_bwd(ptr(pd), ps, ptr(td), ts, B, S, nc, anc, ptr(sums), 1.0, None, ptr(dp), (C.c_int64 * 5)(*dp.stride()), int(bf16), st)
    torch.cuda.synchronize()
    got = dp.float().cpu()
    tol = (2.0 ** -8 if bf16 else 1e-5) * max(1e-3, float(ref.abs().max()))
    assert float((got - ref).abs().max()) <= tol, float((got - ref).abs().max())


def test_sgd_matches_torch():
    from yolo_for_turbines_b200._lib import lib, ptr, stream_ptr

    g = torch.Generator().manual_seed(3)
    n = 10007
    p0 = torch.randn(n, generator=g)
    ref_p = p0.clone().requires_grad_(True)
    opt = torch.optim.SGD([ref_p], lr=0.01, momentum=0.9, weight_decay=5e-4)
    dev = torch.device("cuda")
    pad = (n + 3) // 4 * 4
    p = torch.zeros(pad, device=dev); p[:n] = p0.to(dev)
    buf = torch.zeros(pad, device=dev)
    for step in range(3):
        gr = torch.randn(n, generator=g)
        ref_p.grad = gr.clone()
        opt.step()
        gd = torch.zeros(pad, device=dev); gd[:n] = gr.to(dev)
        lib.yolo_sgd_step(ptr(p), ptr(gd), ptr(buf), n, 0.01, 0.9, 5e-4, 1.0, int(step == 0), stream_ptr(dev))
    torch.cuda.synchronize()
    assert torch.allclose(p[:n].cpu(), ref_p.detach(), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("B,H,cin,cout,k,stride", [(2, 16, 64, 128, 3, 1), (3, 13, 512, 256, 1, 1), (2, 20, 32, 64, 3, 2),
                                                   (1, 9, 128, 32, 1, 1), (4, 26, 128, 256, 3, 1)])
def test_conv_epilogue_statistics(B, H, cin, cout, k, stride):
    """yolo_conv_fwd_stats: the per-channel sum / sum of squares accumulated by the conv epilogue equal those of
    the bf16 tensor it stored (what yolo_bn_stats computes in a separate pass)."""
    from yolo_for_turbines_b200._lib import lib, ptr, stream_ptr
    from yolo_for_turbines_b200.engine import make_conv_plan

    g = torch.Generator().manual_seed(cin + cout + H)
    dev = torch.device("cuda")
    pad = 1 if k == 3 else 0
    Ho = (H + 2 * pad - k) // stride + 1
    x = torch.randn(B, H, H, cin, generator=g).bfloat16().to(dev)
    w = (torch.randn(cout, cin, k, k, generator=g) * (1.0 / (cin * k * k)) ** 0.5).to(dev)
    wpk = torch.empty(cout * k * k * cin, dtype=torch.bfloat16, device=dev)
    st = stream_ptr(dev)
    lib.yolo_pack_weights(ptr(w), cout, cin, k, cout, cin, ptr(wpk), st)
    ones, zeros = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
    z = torch.empty(B * Ho * Ho, cout, dtype=torch.bfloat16, device=dev)
    d = _desc(B, H, H, cin, cin, cout, k, stride)
    d.want_stats = 1
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    plan = make_conv_plan(d, ptr(x), ptr(wpk), ptr(ones), ptr(zeros), None, ptr(z))
    sums = torch.zeros(2 * cout, dtype=torch.float64, device=dev)
    lib.yolo_conv_fwd_stats(plan[1], ptr(status), ptr(sums), None, st)
    torch.cuda.synchronize()
    zd = z.double()
    ref = torch.stack([zd.sum(0), (zd * zd).sum(0)], dim=1).reshape(-1)
    assert torch.allclose(sums, ref, rtol=1e-5, atol=1e-4 * float(ref.abs().max())), float((sums - ref).abs().max())
    # with a finalize descriptor the conv's last CTA also produces what yolo_bn_finalize produces
    from yolo_for_turbines_b200._lib import BnFinalizeDesc
    gamma, beta = (0.5 + torch.rand(cout, generator=g)).to(dev), (0.2 * torch.randn(cout, generator=g)).to(dev)
    outs = {k: torch.empty(cout, device=dev) for k in ("mean", "rstd", "scale", "bias")}
    refs = {k: torch.empty(cout, device=dev) for k in ("mean", "rstd", "scale", "bias")}
    rm, rv, rm2, rv2 = (torch.zeros(cout, device=dev), torch.ones(cout, device=dev), torch.zeros(cout, device=dev),
                        torch.ones(cout, device=dev))
    counter = torch.zeros(1, dtype=torch.int32, device=dev)
    P = B * Ho * Ho
    fin = BnFinalizeDesc(P, gamma.data_ptr(), beta.data_ptr(), 1e-5, 0.1, rm.data_ptr(), rv.data_ptr(), outs["mean"].data_ptr(),
                         outs["rstd"].data_ptr(), outs["scale"].data_ptr(), outs["bias"].data_ptr(), counter.data_ptr())
    sums2 = torch.zeros_like(sums)
    lib.yolo_conv_fwd_stats(plan[1], ptr(status), ptr(sums2), C.byref(fin), st)
    lib.yolo_bn_finalize(ptr(sums), P, cout, ptr(gamma), ptr(beta), 1e-5, 0.1, ptr(rm2), ptr(rv2), ptr(refs["mean"]),
                         ptr(refs["rstd"]), ptr(refs["scale"]), ptr(refs["bias"]), st)
    torch.cuda.synchronize()
    assert int(counter) == 0
    for k in outs:
        assert torch.allclose(outs[k], refs[k], rtol=1e-4, atol=1e-5), k
    assert torch.allclose(rm, rm2, rtol=1e-4, atol=1e-6) and torch.allclose(rv, rv2, rtol=1e-4, atol=1e-6)


@pytest.mark.parametrize("cout,cin,k", [(64, 32, 3), (255, 1024, 1), (21, 256, 1), (1024, 512, 3), (128, 384, 1)])
def test_fused_weight_pack_equals_the_two_reference_packs(cout, cin, k):
    """yolo_pack_weights_train (one pass, tiled) writes exactly what yolo_pack_weights + yolo_pack_weights_dgrad write."""
    from yolo_for_turbines_b200._lib import lib, ptr, stream_ptr

    dev = torch.device("cuda")
    st = stream_ptr(dev)
    w = torch.randn(cout, cin, k, k, generator=torch.Generator().manual_seed(cout + cin)).to(dev)
    cpad = (cout + 31) // 32 * 32
    f_ref = torch.empty(cpad * k * k * cin, dtype=torch.bfloat16, device=dev)
    b_ref = torch.empty(cin * k * k * cpad, dtype=torch.bfloat16, device=dev)
    lib.yolo_pack_weights(ptr(w), cout, cin, k, cpad, cin, ptr(f_ref), st)
    lib.yolo_pack_weights_dgrad(ptr(w), cout, cin, k, cin, cpad, ptr(b_ref), st)
    f, b = torch.zeros_like(f_ref), torch.zeros_like(b_ref)
    lib.yolo_pack_weights_train(ptr(w), cout, cin, k, cin, cpad, ptr(f), ptr(b), st)
    torch.cuda.synchronize()
    assert torch.equal(f, f_ref) and torch.equal(b, b_ref)


@pytest.mark.parametrize("B,H,cin,cout,residual", [(2, 16, 64, 128, False), (3, 12, 32, 64, True), (2, 8, 256, 512, True),
                                                   (1, 26, 128, 256, False)])
def test_stride2_dgrad_parity_subconvs(B, H, cin, cout, residual):
    """Data gradient of a 3x3/s2/p1 conv as two stride-1 sub-convolutions over dz (yolo_conv_desc.s2_parity), with the
    accumulate-through-residual path, vs autograd."""
    from yolo_for_turbines_b200._lib import ConvDesc, lib, ptr, stream_ptr
    from yolo_for_turbines_b200.engine import make_conv_plan

    g = torch.Generator().manual_seed(5 + cin + cout)
    Ho = H // 2
    w = (torch.randn(cout, cin, 3, 3, generator=g) * (1.0 / (cout * 9)) ** 0.5).bfloat16().float()
    dz = torch.randn(B, Ho, Ho, cout, generator=g).bfloat16()
    acc = torch.randn(B, H, H, cin, generator=g).bfloat16() if residual else None
    x = torch.zeros(B, cin, H, H, requires_grad=True)
    F.conv2d(x, w, None, 2, 1).backward(dz.float().permute(0, 3, 1, 2))
    ref = x.grad.permute(0, 2, 3, 1).contiguous()
    if residual:
        ref = ref + acc.float()

    dev = torch.device("cuda")
    st = stream_ptr(dev)
    wdev, dzd = w.to(dev).contiguous(), dz.to(dev).contiguous()
    out = torch.full((B, H, H, cin), float("nan"), dtype=torch.bfloat16, device=dev)
    accd = acc.to(dev).contiguous() if residual else None
    ones, zeros = torch.ones(2 * cin, device=dev), torch.zeros(2 * cin, device=dev)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    keep = []
    for r in (0, 1):
        wpk = torch.empty(2 * cin * (r + 1) * 2 * cout, dtype=torch.bfloat16, device=dev)
        lib.yolo_pack_weights_dgrad_s2(ptr(wdev), cout, cin, r, cin, cout, ptr(wpk), st)
        d = ConvDesc()
        d.batch, d.h_in, d.w_in, d.c_in, d.in_pitch = B, Ho, Ho, cout, cout
        d.c_out, d.c_out_pad, d.out_pitch = 2 * cin, 2 * cin, cin
        d.ksize, d.stride, d.pad = r + 1, 1, 0
        d.ksize_w, d.stride_w, d.pad_w_hi_plus1, d.pad_h_hi_plus1 = 2, 1, 2, r + 1
        d.s2_parity, d.s2_cin = r + 1, cin
        if residual:
            d.has_residual, d.res_pitch = 1, cin
        plan = make_conv_plan(d, ptr(dzd), ptr(wpk), ptr(ones), ptr(zeros), ptr(accd), ptr(out))
        lib.yolo_conv_fwd(plan[1], ptr(status), st)
        keep.append((plan, wpk))
    torch.cuda.synchronize()
    got = out.float().cpu()
    assert torch.isfinite(got).all()          # every pixel of dx is written by exactly one of the two launches
    tol = 2.0 ** -7 * max(1.0, float(ref.abs().max()))
    assert float((got - ref).abs().max()) <= tol
