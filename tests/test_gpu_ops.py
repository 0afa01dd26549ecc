"""GPU: each C-ABI entry point against the CPU oracle / torch fp32 on seeded inputs (through ctypes)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
pytestmark = pytest.mark.gpu


def _dev():
    return torch.device("cuda:0")


def _handle():
    from disentangledcolorization_b200 import _lib
    return _lib.Handle.get(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _conv_call(dtype, kind, stride, srcs, weights, bias, cout, Ho, Wo, act=0, slope=0.0, post=None, res=None, head=0):
    """srcs: list of (tensor NHWC or gray NCHW(C=1), up2, is_f32); weights: list of torch conv weights."""
    from disentangledcolorization_b200 import _lib
    h = _handle()
    tdt = torch.float32 if dtype == _lib.F32 else torch.bfloat16
    B = srcs[0][0].shape[0]
    blocks, offs, off = [], [], 0
    for w in weights:
        blk = (w.permute(2, 3, 0, 1) if kind == _lib.DECONV4 else w.permute(2, 3, 1, 0)).reshape(-1)
        offs.append(off)
        off += blk.numel()
        blocks.append(blk)
    wdev = torch.cat(blocks).contiguous().cuda()
    d = _lib.ConvDesc()
    d.kind, d.stride, d.dtype, d.batch, d.Ho, d.Wo, d.Cout, d.n_src = kind, stride, dtype, B, Ho, Wo, cout, len(srcs)
    keep = [wdev]
    for i, (t, up2, is32) in enumerate(srcs):
        t = t.cuda().contiguous()
        if not is32:
            t = t.to(tdt)
        keep.append(t)
        d.src[i].ptr, d.src[i].H, d.src[i].W, d.src[i].C = t.data_ptr(), t.shape[1], t.shape[2], t.shape[3]
        d.src[i].up2, d.src[i].is_f32, d.src[i].w_off = up2, is32, offs[i]
    bias = bias.cuda().contiguous()
    keep.append(bias)
    d.weights, d.bias = wdev.data_ptr(), bias.data_ptr()
    if post is not None:
        ps, pb = post[0].cuda().contiguous(), post[1].cuda().contiguous()
        keep += [ps, pb]
        d.post_scale, d.post_shift = ps.data_ptr(), pb.data_ptr()
    if res is not None:
        r = res.cuda().contiguous().to(tdt)
        keep.append(r)
        d.residual = r.data_ptr()
    d.act, d.slope, d.head = act, slope, head
    out = (torch.empty(B, cout, Ho, Wo, device="cuda") if head else torch.empty(B, Ho, Wo, cout, device="cuda", dtype=tdt))
    d.out = out.data_ptr()
    _lib.check(h.lib.disco_conv(h.h, C.byref(d), _stream()), "disco_conv")
    torch.cuda.synchronize()
    return out.float().cpu()


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


CONV_CASES = [
    # cin, cout, H, W, stride, up2
    (1, 64, 32, 48, 1, 0), (16, 16, 32, 32, 1, 0), (64, 128, 32, 32, 2, 0), (128, 64, 16, 24, 1, 1),
    (65, 64, 24, 40, 1, 0), (512, 256, 8, 8, 1, 0), (16, 9, 32, 32, 1, 0), (32, 48, 20, 12, 1, 0),
]


@pytest.mark.parametrize("case", CONV_CASES, ids=str)
def test_conv3_fp32_matches_torch(case):
    from disentangledcolorization_b200 import _lib
    cin, cout, H, W, stride, up2 = case
    g = torch.Generator().manual_seed(hash(case) % 1000)
    x = torch.randn(2, cin, H, W, generator=g)
    w = torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
    b = torch.randn(cout, generator=g)
    xin = F.interpolate(x, scale_factor=2, mode="nearest") if up2 else x
    ref = F.leaky_relu(F.conv2d(xin, w, b, stride=stride, padding=1), 0.2)
    ps, pb = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g)
    ref = ref * ps.view(1, -1, 1, 1) + pb.view(1, -1, 1, 1)
    Ho, Wo = ref.shape[2:]
    out = _conv_call(_lib.F32, _lib.CONV3, stride, [(_nhwc(x), up2, 0)], [w], b, cout, Ho, Wo, act=_lib.ACT_LRELU,
                     slope=0.2, post=(ps, pb))
    assert (out - _nhwc(ref)).abs().max() < 2e-5


def test_conv3_two_sources_residual_relu_fp32():
    from disentangledcolorization_b200 import _lib
    g = torch.Generator().manual_seed(3)
    a, b2 = torch.randn(2, 32, 8, 8, generator=g), torch.randn(2, 16, 16, 16, generator=g)
    wa = torch.randn(24, 32, 3, 3, generator=g) * 0.05
    wb = torch.randn(24, 16, 3, 3, generator=g) * 0.05
    bias, res = torch.randn(24, generator=g), torch.randn(2, 24, 16, 16, generator=g)
    ref = F.relu(F.conv2d(F.interpolate(a, scale_factor=2, mode="nearest"), wa, None, padding=1) +
                 F.conv2d(b2, wb, bias, padding=1) + res)
    out = _conv_call(_lib.F32, _lib.CONV3, 1, [(_nhwc(a), 1, 0), (_nhwc(b2), 0, 0)], [wa, wb], bias, 24, 16, 16,
                     act=_lib.ACT_RELU, res=_nhwc(res))
    assert (out - _nhwc(ref)).abs().max() < 2e-5


def test_deconv4_fp32_matches_torch():
    from disentangledcolorization_b200 import _lib
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 32, 9, 7, generator=g)
    w = torch.randn(32, 16, 4, 4, generator=g) * 0.1
    b = torch.randn(16, generator=g)
    ref = F.leaky_relu(F.conv_transpose2d(x, w, b, stride=2, padding=1), 0.1)
    out = _conv_call(_lib.F32, _lib.DECONV4, 1, [(_nhwc(x), 0, 0)], [w], b, 16, 18, 14, act=_lib.ACT_LRELU, slope=0.1)
    assert (out - _nhwc(ref)).abs().max() < 2e-5


def test_heads_fp32():
    from disentangledcolorization_b200 import _lib
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 16, 16, 32, generator=g)
    w9, b9 = torch.randn(9, 16, 3, 3, generator=g) * 0.2, torch.randn(9, generator=g)
    ref = torch.softmax(F.conv2d(x, w9, b9, padding=1), 1)
    out = _conv_call(_lib.F32, _lib.CONV3, 1, [(_nhwc(x), 0, 0)], [w9], b9, 9, 16, 32, head=_lib.HEAD_SOFTMAX9)
    assert (out - ref).abs().max() < 1e-6
    w2, b2 = torch.randn(2, 16, 3, 3, generator=g) * 0.2, torch.randn(2, generator=g)
    ref = torch.tanh(F.conv2d(x, w2, b2, padding=1))
    out = _conv_call(_lib.F32, _lib.CONV3, 1, [(_nhwc(x), 0, 0)], [w2], b2, 2, 16, 32, head=_lib.HEAD_TANH2)
    assert (out - ref).abs().max() < 1e-6


@pytest.mark.parametrize("B,H,W,post", [(2, 32, 48, False), (1, 16, 16, True), (1, 48, 80, True), (3, 20, 70, False)], ids=str)
def test_conv_cin1_cout64_tensor_core_kernel(B, H, W, post):
    """Cin = 1 -> 64 channels on tensor cores (conv_c1_mma_kernel: L-channel taps and weights split hi + lo, bias through
    a constant-one column): the products are fp32-grade, so the only error against torch fp32 is the bf16 rounding of the
    stored output.  Ragged widths (not a multiple of the 64-pixel tile), post-activation affine, image borders."""
    import torch.nn.functional as F
    from disentangledcolorization_b200 import _lib
    g = torch.Generator().manual_seed(B * 100 + H)
    gray = torch.rand(B, 1, H, W, generator=g) * 2 - 1
    w = torch.randn(64, 1, 3, 3, generator=g) * 0.4
    b = torch.randn(64, generator=g) * 0.2
    ps, pb = (torch.rand(64, generator=g) + 0.5, torch.randn(64, generator=g) * 0.1) if post else (None, None)
    out = _conv_call(_lib.BF16, _lib.CONV3, 1, [(gray.permute(0, 2, 3, 1), 0, 1)], [w], b, 64, H, W, act=_lib.ACT_LRELU, slope=0.2,
                     post=(ps, pb) if post else None)
    ref = F.leaky_relu(F.conv2d(gray, w, b, padding=1), 0.2)
    if post:
        ref = ref * ps.view(1, -1, 1, 1) + pb.view(1, -1, 1, 1)
    ref = _nhwc(ref)
    err = (out - ref).abs()
    assert float((err / (ref.abs() + 1e-2)).max()) < 2.0 ** -8 + 1e-3       # one bf16 rounding of the result
    assert float((out - ref.to(torch.bfloat16).float()).abs().max()) <= float(ref.abs().max()) * 2.0 ** -7   # at most 1 ulp off


def test_conv_rejects_bad_descriptor():
    from disentangledcolorization_b200 import _lib
    h = _handle()
    d = _lib.ConvDesc()
    assert h.lib.disco_conv(h.h, C.byref(d), _stream()) == -1
    assert b"conv" in h.lib.disco_last_error()


def test_poolfeat_upfeat_match_oracle():
    import disco_oracle as O
    from disentangledcolorization_b200 import basic
    g = torch.Generator().manual_seed(6)
    B, H, W = 2, 48, 80
    feat = torch.randn(B, 66, H, W, generator=g)
    prob = torch.softmax(torch.randn(B, 9, H, W, generator=g) * 2, 1)
    ref_pool, ref_conf = O.poolfeat(feat, prob, 16)
    pool, conf = basic.poolfeat(feat.cuda(), prob.cuda(), 16, 16, True)
    assert (pool.cpu() - ref_pool).abs().max() < 1e-5
    assert (conf.cpu() - ref_conf).abs().max() < 1e-6
    sizes = basic.get_spixel_size(prob.cuda(), 16, 16)
    assert (sizes.cpu() - O.get_spixel_size(prob, 16)).abs().max() < 1e-6
    tok = torch.randn(B, 64, H // 16, W // 16, generator=g)
    up = basic.upfeat(tok.cuda(), prob.cuda(), 16, 16)
    assert (up.cpu() - O.upfeat(tok, prob, 16)).abs().max() < 1e-5
    # hard one-hot affinity: pool(up(x)) round trip reproduces tokens where every cell keeps its own pixels
    onehot = torch.zeros(B, 9, H, W)
    onehot[:, 4] = 1
    rt, _ = basic.poolfeat(basic.upfeat(tok.cuda(), onehot.cuda(), 16, 16), onehot.cuda(), 16, 16, True)
    assert (rt.cpu() - tok).abs().max() < 1e-5


@pytest.mark.parametrize("B,H,W", [(2, 48, 80), (1, 16, 16), (3, 64, 64)], ids=lambda v: str(v))
def test_poolfeat_bf16_tensor_core_kernel_matches_oracle(B, H, W):
    """disco_poolfeat on bf16 features (the tensor-core kernel: affinity split hi + lo, mma.sync) against the oracle's
    poolfeat / get_spixel_size on the same bf16-rounded features: the split product is fp32-grade, so the tolerance is
    the fp32 one.  Includes a hard one-hot affinity (ties in the hard mass) and borders."""
    import disco_oracle as O
    from disentangledcolorization_b200 import _lib
    h = _handle()
    g = torch.Generator().manual_seed(3 + H)
    feat = torch.relu(torch.randn(B, 64, H, W, generator=g)).to(torch.bfloat16)
    ab = torch.rand(B, 2, H, W, generator=g) - 0.5
    prob = torch.softmax(torch.randn(B, 9, H, W, generator=g) * 2, 1)
    prob[0, :, :8] = 0
    prob[0, 4, :8] = 1                                   # rows with a one-hot assignment
    hh, ww, S = H // 16, W // 16, (H // 16) * (W // 16)
    f32 = dict(dtype=torch.float32, device="cuda")
    feats = feat.permute(0, 2, 3, 1).contiguous().cuda()
    partial = torch.empty(B, hh, ww, 9, 68, **f32)
    tokens, spix = torch.empty(B, S, 64, **f32), torch.empty(B, 2, hh, ww, **f32)
    conf, sizes = torch.empty(B, S, **f32), torch.empty(B, S, **f32)
    p = lambda t: C.c_void_p(t.data_ptr())
    ab_d, prob_d = ab.cuda(), prob.cuda()                # keep the device copies alive across the asynchronous launch
    _lib.check(h.lib.disco_poolfeat(h.h, _lib.BF16, p(feats), p(ab_d), p(prob_d), B, H, W, 64, p(partial), p(tokens),
                                    p(spix), p(conf), p(sizes), _stream()), "disco_poolfeat")
    torch.cuda.synchronize()
    want, mass = O.poolfeat(torch.cat([feat.float(), ab], 1), prob, 16)
    got = tokens.view(B, hh, ww, 64).permute(0, 3, 1, 2).cpu()
    scale = max(1.0, float(want.abs().max()))
    assert float((got - want[:, :64]).abs().max()) < 2e-5 * scale
    assert float((spix.cpu() - want[:, 64:]).abs().max()) < 1e-5
    assert float((conf.view(B, 1, hh, ww).cpu() - mass).abs().max()) < 3e-6          # affinity split hi + lo: 2^-17 relative
    assert float((sizes.view(B, 1, hh, ww).cpu() - O.get_spixel_size(prob, 16)).abs().max()) < 1e-6


def test_encoder_stack_matches_oracle(synth_sd):
    import disco_oracle as O
    from disentangledcolorization_b200.engine import Engine
    eng = Engine(synth_sd, _dev(), precision="fp32", n_clusters=4)
    B, H, W = 3, 64, 96
    ws = eng._workspace(B, H, W)
    S = ws["S"]
    g = torch.Generator().manual_seed(7)
    x = torch.randn(B, S, 64, generator=g)
    pos = O.position_sine(H // 16, W // 16).flatten(1).t().unsqueeze(0).expand(B, -1, -1)
    assert (eng._pos[(H // 16, W // 16)].cpu() - pos[0]).abs().max() == 0
    ref = O.encoder_stack(synth_sd, "hintpath", x, pos)
    out = torch.empty(B * S, 64, device="cuda")
    eng._encoder_stack("hintpath", x.cuda().view(B * S, 64).contiguous(), out, ws, B, _stream())
    torch.cuda.synchronize()
    assert (out.cpu().view(B, S, 64) - ref).abs().max() < 2e-5


@pytest.mark.parametrize("B,H,W", [(3, 64, 96), (2, 256, 256), (1, 160, 160), (1, 512, 512), (5, 16, 16), (2, 176, 240)],
                         ids=lambda v: str(v))
def test_encoder_stack_fused_matches_oracle(synth_sd, B, H, W):
    """The one-launch tensor-core stack (split-bf16 mma.sync, csrc/encoder_stack.cu) against the fp32 oracle stack:
    S = 24 (one partly filled CTA), 256 (2-CTA cluster), 100 and 165 (ragged tails), 1024 (8-CTA cluster), 1 token."""
    import disco_oracle as O
    from disentangledcolorization_b200.engine import Engine
    eng = Engine(synth_sd, _dev(), precision="bf16", n_clusters=1)
    assert eng.fused_tokens
    ws = eng._workspace(B, H, W)
    S = ws["S"]
    g = torch.Generator().manual_seed(7 + S)
    x = torch.randn(B, S, 64, generator=g)
    pos = O.position_sine(H // 16, W // 16).flatten(1).t().unsqueeze(0).expand(B, -1, -1)
    for stack in ("wildpath", "hintpath"):
        ref = O.encoder_stack(synth_sd, stack, x, pos)
        out = torch.full((B * S, 64), float("nan"), device="cuda")
        before = eng.handle.launches()
        eng._encoder_stack(stack, x.cuda().view(B * S, 64).contiguous(), out, ws, B, _stream())
        torch.cuda.synchronize()
        assert eng.handle.launches() - before == 1
        err = (out.cpu().view(B, S, 64) - ref).abs().max()
        print(f"fused {stack} B={B} S={S}: max err {float(err):.2e}")
        assert err < 5e-4, float(err)          # P V on the f16 path: P carries 11 significant bits (exact variant: -DES_PV_F16=0, 5e-5)
        # a second run into the ping-pong scratch gives the same bits (no stale key/value reads)
        out2 = torch.empty_like(out)
        eng._encoder_stack(stack, x.cuda().view(B * S, 64).contiguous(), out2, ws, B, _stream())
        assert torch.equal(out, out2)


@pytest.mark.parametrize("B,H,W", [(2, 64, 96), (1, 48, 80), (3, 16, 16), (1, 256, 256), (2, 144, 112)], ids=lambda v: str(v))
def test_segnet_fused_head_matches_layer_by_layer(synth_sd, B, H, W):
    """conv0a -> conv0b -> conv1a in one launch (csrc/segnet_fused.cu, 16-channel intermediate in shared memory) against
    the same three layers launched one by one (tcgen05 / CUDA-core kernels) and against torch fp32: partial tiles, tiles
    on every image border, images smaller than a tile."""
    import torch.nn.functional as F
    from disentangledcolorization_b200 import synth, netspec
    from disentangledcolorization_b200.engine import Engine
    gray = torch.from_numpy(synth.make_gray(B, H, W, seed=17 + H)).cuda()
    outs = {}
    for fused in (True, False):
        eng = Engine(synth_sd, _dev(), precision="bf16", n_clusters=1)
        assert eng.seg_head is not None
        if not fused:
            eng.seg_head = None
        ws = eng._workspace(B, H, W)
        before = eng.handle.launches()
        eng._run_net("segnet", ws, B, gray, _stream())
        torch.cuda.synchronize()
        outs[fused] = (ws["bufs"]["sg.out1"].float().cpu(), ws["bufs"]["sg.1a"].float().cpu(), ws["bufs"]["affinity"].cpu(),
                       eng.handle.launches() - before)
    assert outs[False][3] - outs[True][3] == 2                       # three launches became one
    # torch fp32 on the folded weights
    ops = netspec.segnet_ops()[:3]
    x = gray.cpu()
    ref = []
    for op in ops:
        f = netspec.fold(synth_sd, op)
        x = F.leaky_relu(F.conv2d(x, f.weights[0], f.bias, stride=op.stride, padding=1), 0.1)
        ref.append(x.permute(0, 2, 3, 1))
    for i, name in ((0, "out1"), (1, "conv1a")):
        a, b, r = outs[True][i], outs[False][i], ref[i + 1]
        scale = float(r.abs().max())
        print(f"{name}: fused vs layer-by-layer max {float((a - b).abs().max()):.3e}, fused vs torch fp32 max "
              f"{float((a - r).abs().max()):.3e} (layer-by-layer: {float((b - r).abs().max()):.3e}), scale {scale:.2f}")
        assert float((a - r).abs().max()) < 2e-2 * scale            # two / three bf16-stored layers
        assert float((a - b).abs().max()) < 1e-2 * scale            # same arithmetic, different fp32 summation order
        assert float((a - b).abs().mean()) < 5e-4 * scale
    assert float((outs[True][2] - outs[False][2]).abs().max()) < 2e-2


def _kmeans_gpu(X, K, sizes, seed):
    from disentangledcolorization_b200 import _lib
    h = _handle()
    B, S, _ = X.shape
    np.random.seed(seed)
    torch.manual_seed(seed)
    init = torch.from_numpy(np.stack([np.random.choice(S, K, replace=False) for _ in range(B)]).astype(np.int32)).cuda()
    state = torch.get_rng_state()
    draws = torch.randint(S, (512,)).to(torch.int32).cuda()
    torch.set_rng_state(state)
    Xd, sd_ = X.cuda().contiguous(), sizes.cuda().contiguous()
    assign = torch.empty(B, S, dtype=torch.int32, device="cuda")
    hint = torch.empty(B, S, device="cuda")
    events = torch.zeros(B + 2, dtype=torch.int32, device="cuda")
    iters = torch.zeros(B, dtype=torch.int32, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    _lib.check(h.lib.disco_kmeans_anchor(h.h, p(Xd), p(init), p(draws), 512, p(sd_), B, S, K, 20, 1e-4, p(assign), p(hint),
                                         p(events), p(iters), _stream()), "kmeans")
    torch.cuda.synchronize()
    return assign.cpu(), hint.cpu(), events.cpu(), iters.cpu()


@pytest.mark.parametrize("dup", [False, True], ids=["generic", "empty-clusters"])
def test_kmeans_anchor_matches_oracle(dup):
    """Includes the empty-cluster re-seed path: duplicated points make clusters go empty, which consumes
    torch.randint draws in image-major order (clusterkit.py:178-184)."""
    import disco_oracle as O
    g = torch.Generator().manual_seed(8)
    B, h, w, K = 5, 6, 8, 6
    S = h * w
    X = torch.randn(B, S, 64, generator=g)
    if dup:
        X[:, : S // 2] = X[:, :1]            # half of the tokens identical -> duplicate initial centres
        X[3] = torch.randn(S, 64, generator=g)
    sizes = torch.rand(B, S, generator=g)
    np.random.seed(21)
    torch.manual_seed(21)
    ref_hint, ref_cluster = O.anchor_mask(X.transpose(1, 2).reshape(B, 64, h, w), K, sizes.view(B, 1, h, w))
    ref_next = int(torch.randint(1 << 30, (1,)))
    assign, hint, events, iters = _kmeans_gpu(X, K, sizes, 21)
    assert torch.equal(assign.long(), ref_cluster.flatten(2).argmax(1))
    assert torch.equal(hint.view(B, 1, h, w), ref_hint)
    if dup:
        assert int(events[B]) > 0, "test should exercise the empty-cluster path"
    # advancing the host generator by the reported number of draws reproduces the reference stream position
    torch.manual_seed(21)
    if int(events[B]):
        torch.randint(S, (int(events[B]),))
    assert int(torch.randint(1 << 30, (1,))) == ref_next
    assert int(events[B + 1]) == 0 and int(iters.max()) <= 20


def test_token_labels():
    import disco_oracle as O
    from disentangledcolorization_b200 import _lib, basic
    g = torch.Generator().manual_seed(9)
    B, h, w = 2, 3, 5
    logits = torch.randn(B, 313, h, w, generator=g)
    hd = _handle()
    labels = torch.empty(B * h * w, dtype=torch.int32, device="cuda")
    colors = torch.empty(B, 2, h, w, device="cuda")
    table = torch.from_numpy(__import__("disentangledcolorization_b200.cielab", fromlist=["Q_TO_AB"]).Q_TO_AB.copy()).cuda()
    p = lambda t: C.c_void_p(t.data_ptr())
    _lib.check(hd.lib.disco_token_labels(hd.h, 0, p(logits.cuda()), p(table), B, h * w, p(labels), p(colors), _stream()), "labels")
    torch.cuda.synchronize()
    ref_colors = O.sample_anchor_colors(torch.softmax(logits, 1), T=0)
    assert torch.equal(labels.cpu().view(B, h, w).long(), logits.argmax(1))
    assert (colors.cpu() - ref_colors).abs().max() < 1e-7
    ab = (torch.rand(B, 2, h, w, generator=g) - 0.5) * 0.8
    hard = basic.ColorLabel().encode_ab2ind_hard(ab.cuda())
    assert torch.equal(hard.cpu(), O.encode_ab2ind(ab).max(dim=1, keepdim=True)[1])


def _attention_ref(qkv, B, S):
    """softmax(q k^T) v per (image, head) in float64; q is already scaled (the QKV projection folds 1/sqrt(d_head))."""
    x = qkv.double().view(B, S, 3, 8, 8)
    q, k, v = x[:, :, 0].permute(0, 2, 1, 3), x[:, :, 1].permute(0, 2, 1, 3), x[:, :, 2].permute(0, 2, 1, 3)   # B,8,S,8
    att = torch.softmax(q @ k.transpose(-1, -2), dim=-1) @ v
    return att.permute(0, 2, 1, 3).reshape(B * S, 64)


@pytest.mark.parametrize("B,S", [(2, 24), (1, 100), (3, 256), (1, 1024), (1, 1030)], ids=str)
def test_attention_matches_float64_softmax(B, S):
    """disco_attention (two queries per thread, packed fp32 FMAs, fixed softmax reference m = 0 with exact fall-back)
    against float64 softmax attention: ragged token counts, both CTA shapes (S <= 256 and S > 256)."""
    from disentangledcolorization_b200 import _lib
    g = torch.Generator().manual_seed(1000 + S)
    qkv = torch.randn(B * S, 192, generator=g)
    qkv[:, :64] *= 1.5
    hd = _handle()
    out = torch.empty(B * S, 64, device="cuda")
    d = qkv.cuda()
    _lib.check(hd.lib.disco_attention(hd.h, C.c_void_p(d.data_ptr()), B, S, C.c_void_p(out.data_ptr()), _stream()), "attention")
    torch.cuda.synchronize()
    ref = _attention_ref(qkv, B, S)
    assert (out.cpu().double() - ref).abs().max() < 5e-6


def test_attention_falls_back_outside_the_exp2_range():
    """Scores beyond fp32 exp2's range (one head overflows: |s| ~ 500 log2 units; another underflows: all scores ~ -500)
    invalidate the m = 0 evaluation: the online-softmax routine must take over for exactly those queries."""
    from disentangledcolorization_b200 import _lib
    g = torch.Generator().manual_seed(77)
    B, S = 1, 64
    qkv = torch.randn(B * S, 192, generator=g)
    x = qkv.view(S, 3, 8, 8)
    x[:, 0, 1] *= 60.0             # head 1: scores of magnitude ~ +-500 (overflow AND underflow of exp2 without a reference)
    x[:, 1, 2] = -torch.abs(x[:, 1, 2])
    x[:, 0, 2] = 40.0 * torch.abs(x[:, 0, 2])   # head 2: every score strongly negative (sum underflows to 0)
    x[9, 1, 3, :] = 40.0 * x[2, 0, 3, :] / x[2, 0, 3, :].norm()   # head 3: one genuinely large score (sharp softmax)
    hd = _handle()
    out = torch.empty(B * S, 64, device="cuda")
    d = qkv.cuda()
    _lib.check(hd.lib.disco_attention(hd.h, C.c_void_p(d.data_ptr()), B, S, C.c_void_p(out.data_ptr()), _stream()), "attention")
    torch.cuda.synchronize()
    ref = _attention_ref(qkv, B, S)
    assert torch.isfinite(out).all()
    # scores of magnitude ~500 carry ~3e-5 of fp32 rounding themselves: this test is about the fall-back logic
    assert (out.cpu().double() - ref).abs().max() < 1e-4
