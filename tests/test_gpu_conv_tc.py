"""GPU: the tcgen05 implicit-GEMM convolution (bf16) against torch fp32 on bf16-rounded operands.

Inputs and weights are rounded to bf16 first, so the only differences are fp32 accumulation order and the
final bf16 rounding of the output: tolerance = 2^-8 relative to the output scale (one bf16 ulp) + a small
absolute term.  The parity-phase variants (up-sampled sources, transposed conv) pre-sum weights per output
parity and round the SUM to bf16, which costs up to one more bf16 ulp of the summed weight."""
import ctypes as C
import os
import sys

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _bf(x):
    return x.to(torch.bfloat16).float()


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def _run_tc(kind, stride, srcs, weights, bias, cout, Ho, Wo, act=0, slope=0.0, post=None, res=None, head=0, expect_tc=True,
            host_params=True):
    """srcs: list of (NCHW fp32 tensor, up2, is_gray).  host_params: also pass the optional host copies of the epilogue
    parameters (narrow layers then read them from the constant bank)."""
    from disentangledcolorization_b200 import _lib
    h = _lib.Handle.get(0)
    B = srcs[0][0].shape[0]
    blocks, offs, off = [], [], 0
    for w in weights:
        blk = (w.permute(2, 3, 0, 1) if kind == _lib.DECONV4 else w.permute(2, 3, 1, 0)).reshape(-1)
        offs.append(off)
        off += blk.numel()
        blocks.append(blk)
    w32_host = torch.cat(blocks).contiguous()
    w32_dev = w32_host.cuda()
    d = _lib.ConvDesc()
    d.kind, d.stride, d.dtype, d.batch, d.Ho, d.Wo, d.Cout, d.n_src = kind, stride, _lib.BF16, B, Ho, Wo, cout, len(srcs)
    keep = []
    for i, (t, up2, gray) in enumerate(srcs):
        tt = t.cuda().contiguous() if gray else _nhwc(t).cuda().to(torch.bfloat16)
        keep.append(tt)
        d.src[i].ptr = tt.data_ptr()
        d.src[i].H, d.src[i].W, d.src[i].C = t.shape[2], t.shape[3], t.shape[1]
        d.src[i].up2, d.src[i].is_f32, d.src[i].w_off = up2, int(gray), offs[i]
    bias_d = bias.cuda().contiguous()
    d.bias = bias_d.data_ptr()
    if post is not None:
        ps, pb = post[0].cuda().contiguous(), post[1].cuda().contiguous()
        keep += [ps, pb]
        d.post_scale, d.post_shift = ps.data_ptr(), pb.data_ptr()
    if res is not None:
        r = _nhwc(res).cuda().to(torch.bfloat16)
        keep.append(r)
        d.residual = r.data_ptr()
    d.act, d.slope, d.head = act, slope, head
    if host_params:
        bh = bias.float().contiguous()
        keep.append(bh)
        d.bias_host = bh.data_ptr()
        if post is not None:
            psh, pbh = post[0].float().contiguous(), post[1].float().contiguous()
            keep += [psh, pbh]
            d.post_scale_host, d.post_shift_host = psh.data_ptr(), pbh.data_ptr()
    out = torch.empty(B, cout, Ho, Wo, device="cuda") if head else torch.empty(B, Ho, Wo, cout, device="cuda", dtype=torch.bfloat16)
    d.out = out.data_ptr()
    d.weights = w32_dev.data_ptr()
    supported = bool(h.lib.disco_conv_tc_supported(h.h, C.byref(d)))
    assert supported == expect_tc
    if supported:
        n = int(h.lib.disco_conv_tc_weight_elems(C.byref(d)))
        w16 = torch.empty(n, dtype=torch.int16)
        _lib.check(h.lib.disco_conv_tc_pack_weights(C.byref(d), C.c_void_p(w32_host.data_ptr()), C.c_void_p(w16.data_ptr())), "pack")
        w16d = w16.cuda()
        d.weights = w16d.data_ptr()
        for i in range(d.n_src):
            if d.src[i].is_f32:
                d.gray_weights = w32_dev.data_ptr() + 4 * offs[i]
                if host_params:
                    d.gray_weights_host = w32_host.data_ptr() + 4 * offs[i]
    _lib.check(h.lib.disco_conv(h.h, C.byref(d), C.c_void_p(torch.cuda.current_stream().cuda_stream)), "disco_conv")
    torch.cuda.synchronize()
    o = out.float().cpu()
    return o if head else o.permute(0, 3, 1, 2)


def _check(out, ref, ulps=1.0):
    scale = float(ref.abs().max())
    err = float((out - ref).abs().max())
    tol = ulps * scale * 2 ** -8 + 1e-3 * scale
    assert err <= tol, f"max err {err:.4g} > tol {tol:.4g} (scale {scale:.3g})"


PLAIN = [
    # cin, cout, B, H, W, stride
    (64, 64, 2, 32, 32, 1), (128, 128, 1, 16, 48, 1), (512, 512, 2, 32, 32, 1), (256, 512, 2, 16, 16, 2),
    (64, 128, 1, 64, 32, 2), (16, 16, 1, 64, 64, 1), (16, 32, 2, 32, 32, 2), (32, 32, 1, 32, 32, 1),
    (32, 64, 1, 32, 64, 2), (256, 256, 3, 4, 4, 1), (128, 256, 5, 8, 8, 2), (64, 64, 1, 40, 24, 1),
    (256, 256, 1, 30, 40, 1), (64, 48, 1, 16, 16, 1),
    # CTA-pair (cta_group::2) variant: wide N, even number of tile columns; ragged rows; stride 2; 256-pixel sub-tiles
    (256, 256, 2, 16, 64, 1), (128, 256, 1, 32, 64, 2), (128, 128, 2, 64, 64, 1), (64, 128, 1, 64, 64, 2),
    (256, 256, 1, 20, 32, 1), (512, 512, 1, 8, 96, 1), (128, 128, 3, 40, 32, 1),
    # stride-2 layers through the grouped kernel (parity-plane halo boxes): ragged rows, two n-tiles, odd/even tile columns
    (64, 128, 2, 48, 32, 2), (256, 512, 1, 32, 32, 2), (128, 128, 1, 32, 96, 2), (64, 256, 1, 64, 48, 2),
]


@pytest.mark.parametrize("case", PLAIN, ids=str)
def test_tc_conv3_plain_and_strided(case):
    from disentangledcolorization_b200 import _lib
    cin, cout, B, H, W, stride = case
    g = torch.Generator().manual_seed(sum(case))
    x = _bf(torch.randn(B, cin, H, W, generator=g))
    w = _bf(torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5)
    b = torch.randn(cout, generator=g) * 0.1
    ps, pb = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1
    ref = F.leaky_relu(F.conv2d(x, w, b, stride=stride, padding=1), 0.2) * ps.view(1, -1, 1, 1) + pb.view(1, -1, 1, 1)
    out = _run_tc(_lib.CONV3, stride, [(x, 0, False)], [w], b, cout, H // stride, W // stride, act=_lib.ACT_LRELU, slope=0.2,
                  post=(ps, pb))
    _check(out, ref)


@pytest.mark.parametrize("case", [(128, 64, 2, 16, 16), (512, 256, 1, 8, 16), (32, 16, 1, 32, 32), (256, 128, 2, 4, 4),
                                  (256, 128, 1, 32, 32), (512, 256, 1, 16, 32)], ids=str)
def test_tc_upsample_conv_parity_phases(case):
    from disentangledcolorization_b200 import _lib
    cin, cout, B, H, W = case
    g = torch.Generator().manual_seed(sum(case))
    x = _bf(torch.randn(B, cin, H, W, generator=g))
    w = _bf(torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5)
    b = torch.randn(cout, generator=g) * 0.1
    ref = F.relu(F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, b, padding=1))
    out = _run_tc(_lib.CONV3, 1, [(x, 1, False)], [w], b, cout, 2 * H, 2 * W, act=_lib.ACT_RELU)
    _check(out, ref, ulps=3.0)


def test_tc_two_sources_up2_plus_direct_with_residual():
    """conv8up(up2(f7)) + conv3short8(f3) and the HourGlass2 up-block combine: parity phases + stride-2-sampled skip."""
    from disentangledcolorization_b200 import _lib
    g = torch.Generator().manual_seed(11)
    a, s = _bf(torch.randn(2, 128, 8, 8, generator=g)), _bf(torch.randn(2, 64, 16, 16, generator=g))
    wa = _bf(torch.randn(64, 128, 3, 3, generator=g) / (128 * 9) ** 0.5)
    ws = _bf(torch.randn(64, 64, 3, 3, generator=g) / (64 * 9) ** 0.5)
    b = torch.randn(64, generator=g) * 0.1
    res = _bf(torch.randn(2, 64, 16, 16, generator=g))
    ref = F.relu(F.conv2d(F.interpolate(a, scale_factor=2, mode="nearest"), wa, None, padding=1) + F.conv2d(s, ws, b, padding=1) + res)
    out = _run_tc(_lib.CONV3, 1, [(a, 1, False), (s, 0, False)], [wa, ws], b, 64, 16, 16, act=_lib.ACT_RELU, res=res)
    _check(out, ref, ulps=3.0)


def test_tc_two_sources_wide_pair_variant():
    """HourGlass2 up2.combine shape family (Cout = 128): CTA-pair kernel with parity phases + stride-2-sampled skip."""
    from disentangledcolorization_b200 import _lib
    g = torch.Generator().manual_seed(21)
    a, s = _bf(torch.randn(2, 256, 16, 32, generator=g)), _bf(torch.randn(2, 128, 32, 64, generator=g))
    wa = _bf(torch.randn(128, 256, 3, 3, generator=g) / (256 * 9) ** 0.5)
    ws = _bf(torch.randn(128, 128, 3, 3, generator=g) / (128 * 9) ** 0.5)
    b = torch.randn(128, generator=g) * 0.1
    res = _bf(torch.randn(2, 128, 32, 64, generator=g))
    ref = F.relu(F.conv2d(F.interpolate(a, scale_factor=2, mode="nearest"), wa, None, padding=1) + F.conv2d(s, ws, b, padding=1) + res)
    out = _run_tc(_lib.CONV3, 1, [(a, 1, False), (s, 0, False)], [wa, ws], b, 128, 32, 64, act=_lib.ACT_RELU, res=res)
    _check(out, ref, ulps=3.0)


def test_tc_concat_two_direct_sources():
    from disentangledcolorization_b200 import _lib
    g = torch.Generator().manual_seed(12)
    a, s = _bf(torch.randn(1, 128, 16, 16, generator=g)), _bf(torch.randn(1, 128, 16, 16, generator=g))
    w = _bf(torch.randn(128, 256, 3, 3, generator=g) / (256 * 9) ** 0.5)
    b = torch.randn(128, generator=g) * 0.1
    ref = F.leaky_relu(F.conv2d(torch.cat((a, s), 1), w, b, padding=1), 0.1)
    out = _run_tc(_lib.CONV3, 1, [(a, 0, False), (s, 0, False)], [w[:, :128], w[:, 128:]], b, 128, 16, 16, act=_lib.ACT_LRELU, slope=0.1)
    _check(out, ref)


@pytest.mark.parametrize("case", [(256, 128, 2, 16, 16), (32, 16, 1, 32, 48), (64, 32, 1, 4, 4)], ids=str)
def test_tc_deconv4(case):
    from disentangledcolorization_b200 import _lib
    cin, cout, B, H, W = case
    g = torch.Generator().manual_seed(sum(case))
    x = _bf(torch.randn(B, cin, H, W, generator=g))
    w = _bf(torch.randn(cin, cout, 4, 4, generator=g) / (cin * 4) ** 0.5)
    b = torch.randn(cout, generator=g) * 0.1
    ref = F.leaky_relu(F.conv_transpose2d(x, w, b, stride=2, padding=1), 0.1)
    out = _run_tc(_lib.DECONV4, 1, [(x, 0, False)], [w], b, cout, 2 * H, 2 * W, act=_lib.ACT_LRELU, slope=0.1)
    _check(out, ref)


def test_tc_heads_and_gray_side_input():
    from disentangledcolorization_b200 import _lib
    g = torch.Generator().manual_seed(13)
    x = _bf(torch.randn(2, 16, 32, 32, generator=g))
    w9, b9 = _bf(torch.randn(9, 16, 3, 3, generator=g) * 0.2), torch.randn(9, generator=g)
    out = _run_tc(_lib.CONV3, 1, [(x, 0, False)], [w9], b9, 9, 32, 32, head=_lib.HEAD_SOFTMAX9)
    assert (out - torch.softmax(F.conv2d(x, w9, b9, padding=1), 1)).abs().max() < 2e-3
    x64 = _bf(torch.randn(1, 64, 32, 32, generator=g))
    w2, b2 = _bf(torch.randn(2, 64, 3, 3, generator=g) * 0.05), torch.randn(2, generator=g) * 0.1
    out = _run_tc(_lib.CONV3, 1, [(x64, 0, False)], [w2], b2, 2, 32, 32, head=_lib.HEAD_TANH2)
    assert (out - torch.tanh(F.conv2d(x64, w2, b2, padding=1))).abs().max() < 2e-3
    # enhanceNet.inConv: fp32 L channel (weight channel 0) beside 64 bf16 feature channels
    gray = torch.rand(1, 1, 32, 32, generator=g) * 2 - 1
    w65 = torch.randn(64, 65, 3, 3, generator=g) / (65 * 9) ** 0.5
    w65[:, 1:] = _bf(w65[:, 1:])
    b = torch.randn(64, generator=g) * 0.1
    ref = F.relu(F.conv2d(torch.cat((gray, x64), 1), w65, b, padding=1))
    out = _run_tc(_lib.CONV3, 1, [(gray, 0, True), (x64, 0, False)], [w65[:, :1], w65[:, 1:]], b, 64, 32, 32, act=_lib.ACT_RELU)
    _check(out, ref)


@pytest.mark.parametrize("cout", [16, 32, 64])
def test_tc_narrow_layers_without_host_parameter_copies(cout):
    """The shared-memory parameter cache path (descriptor without the optional host copies) stays correct."""
    from disentangledcolorization_b200 import _lib
    g = torch.Generator().manual_seed(40 + cout)
    x = _bf(torch.randn(2, 32, 32, 32, generator=g))
    w = _bf(torch.randn(cout, 32, 3, 3, generator=g) / (32 * 9) ** 0.5)
    b = torch.randn(cout, generator=g) * 0.1
    ps, pb = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1
    ref = F.leaky_relu(F.conv2d(x, w, b, padding=1), 0.1) * ps.view(1, -1, 1, 1) + pb.view(1, -1, 1, 1)
    for hp in (False, True):
        out = _run_tc(_lib.CONV3, 1, [(x, 0, False)], [w], b, cout, 32, 32, act=_lib.ACT_LRELU, slope=0.1, post=(ps, pb),
                      host_params=hp)
        _check(out, ref)


def test_cin1_layers_stay_on_cuda_cores():
    from disentangledcolorization_b200 import _lib
    g = torch.Generator().manual_seed(14)
    gray = torch.rand(1, 1, 32, 32, generator=g) * 2 - 1
    w, b = torch.randn(64, 1, 3, 3, generator=g) / 3, torch.randn(64, generator=g) * 0.1
    out = _run_tc(_lib.CONV3, 1, [(gray, 0, True)], [w], b, 64, 32, 32, act=_lib.ACT_LRELU, slope=0.2, expect_tc=False)
    _check(out, F.leaky_relu(F.conv2d(gray, w, b, padding=1), 0.2))


# ---- narrow-layer mma.sync kernels (csrc/conv_narrow.cu): SpixelNet's decoder tail (deconv0, conv0_1, pred_mask0 + softmax);
# the 32-channel shapes below stay on the tcgen05 resident kernel and are kept as ragged-size cases for it
NARROW_SIZES = [(2, 32, 32), (1, 48, 80), (3, 20, 36), (1, 16, 100)]


@pytest.mark.parametrize("B,H,W", NARROW_SIZES, ids=str)
def test_narrow_conv0_1_two_16_channel_sources(B, H, W):
    from disentangledcolorization_b200 import _lib
    g = torch.Generator().manual_seed(B * 1000 + H + W)
    a, s = _bf(torch.randn(B, 16, H, W, generator=g)), _bf(torch.randn(B, 16, H, W, generator=g))
    w = _bf(torch.randn(16, 32, 3, 3, generator=g) / (32 * 9) ** 0.5)
    b = torch.randn(16, generator=g) * 0.1
    ref = F.leaky_relu(F.conv2d(torch.cat((a, s), 1), w, b, padding=1), 0.1)
    out = _run_tc(_lib.CONV3, 1, [(a, 0, False), (s, 0, False)], [w[:, :16], w[:, 16:]], b, 16, H, W, act=_lib.ACT_LRELU, slope=0.1)
    _check(out, ref)


@pytest.mark.parametrize("B,H,W", NARROW_SIZES, ids=str)
def test_narrow_pred_mask_softmax9(B, H, W):
    from disentangledcolorization_b200 import _lib
    g = torch.Generator().manual_seed(B * 1000 + H + W + 1)
    x = _bf(torch.randn(B, 16, H, W, generator=g))
    w9, b9 = _bf(torch.randn(9, 16, 3, 3, generator=g) * 0.2), torch.randn(9, generator=g)
    out = _run_tc(_lib.CONV3, 1, [(x, 0, False)], [w9], b9, 9, H, W, head=_lib.HEAD_SOFTMAX9)
    ref = torch.softmax(F.conv2d(x, w9, b9, padding=1), 1)
    assert (out - ref).abs().max() < 1e-5                  # bf16 operands are exact products: only fp32 summation order differs
    assert (out.sum(1) - 1).abs().max() < 1e-5


@pytest.mark.parametrize("B,H,W", NARROW_SIZES, ids=str)
@pytest.mark.parametrize("two", [False, True], ids=["conv1b", "conv1_1"])
def test_narrow_32_channel_layers(B, H, W, two):
    from disentangledcolorization_b200 import _lib
    g = torch.Generator().manual_seed(B * 1000 + H + W + 2 + int(two))
    cin = 64 if two else 32
    x = _bf(torch.randn(B, cin, H, W, generator=g))
    w = _bf(torch.randn(32, cin, 3, 3, generator=g) / (cin * 9) ** 0.5)
    b = torch.randn(32, generator=g) * 0.1
    ref = F.relu(F.conv2d(x, w, b, padding=1))
    srcs = [(x[:, :32], 0, False), (x[:, 32:], 0, False)] if two else [(x, 0, False)]
    ws = [w[:, :32], w[:, 32:]] if two else [w]
    out = _run_tc(_lib.CONV3, 1, srcs, ws, b, 32, H, W, act=_lib.ACT_RELU)
    _check(out, ref)


@pytest.mark.parametrize("B,H,W", [(2, 16, 16), (1, 12, 20), (1, 9, 7), (3, 8, 50)], ids=str)
def test_narrow_deconv0(B, H, W):
    from disentangledcolorization_b200 import _lib
    g = torch.Generator().manual_seed(B * 1000 + H + W + 3)
    x = _bf(torch.randn(B, 32, H, W, generator=g))
    w = _bf(torch.randn(32, 16, 4, 4, generator=g) / (32 * 4) ** 0.5)
    b = torch.randn(16, generator=g) * 0.1
    ref = F.leaky_relu(F.conv_transpose2d(x, w, b, stride=2, padding=1), 0.1)
    out = _run_tc(_lib.DECONV4, 1, [(x, 0, False)], [w], b, 16, 2 * H, 2 * W, act=_lib.ACT_LRELU, slope=0.1)
    _check(out, ref)


# ---- 64 -> 64 stride-1 layers: with DISCO_TS=1 these run on the tensor-memory-operand kernel (csrc/conv_ts.cu), otherwise on
# the resident-weight tcgen05 kernel; the expectations are the same
@pytest.mark.parametrize("B,H,W,res,post", [(1, 32, 128, False, False), (2, 40, 256, True, True), (1, 19, 200, False, True),
                                            (3, 8, 64, True, False), (1, 70, 136, False, False), (2, 1, 128, False, False)], ids=str)
def test_conv_64_to_64_full_resolution_shapes(B, H, W, res, post):
    from disentangledcolorization_b200 import _lib
    g = torch.Generator().manual_seed(B * 977 + H * 31 + W)
    x = _bf(torch.randn(B, 64, H, W, generator=g))
    w = _bf(torch.randn(64, 64, 3, 3, generator=g) / (64 * 9) ** 0.5)
    b = torch.randn(64, generator=g) * 0.1
    r = _bf(torch.randn(B, 64, H, W, generator=g)) if res else None
    ps, pb = torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.1
    ref = F.conv2d(x, w, b, padding=1)
    if res:
        ref = ref + r
    ref = F.leaky_relu(ref, 0.2)
    if post:
        ref = ref * ps.view(1, -1, 1, 1) + pb.view(1, -1, 1, 1)
    out = _run_tc(_lib.CONV3, 1, [(x, 0, False)], [w], b, 64, H, W, act=_lib.ACT_LRELU, slope=0.2, post=(ps, pb) if post else None, res=r)
    _check(out, ref)
