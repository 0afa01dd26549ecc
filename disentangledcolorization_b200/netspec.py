"""Declarative description of the three convolutional networks on the DISCO forward path.

One table drives three things:
  * the `state_dict` schema of the drop-in modules (461 keys, identical to the reference's
    `AnchorColorProb.state_dict()`);
  * weight folding at load time (spectral-norm sigma, eval-mode BatchNorm) into per-op
    (weight, bias, post-scale, post-shift) tensors;
  * the fused-op launch plan executed by the CUDA library (one `ConvOp` = one kernel launch).

Reference structure restated here (not copied): `SpixelNet` models/network.py:260-313,
`ColorProbNet` models/network.py:147-236, `HourGlass2` + blocks models/network.py:10-47,66-101,125-144.

Fusion rules (all exact in real arithmetic):
  * conv(no bias) -> BN -> LeakyReLU  (segnet `conv()`, network.py:240-246): BN folded into W, b.
  * [SN]conv -> act -> ... -> BN      (repnet / enhanceNet blocks): BN becomes a per-channel affine
    applied *after* the activation in the epilogue of the block's last conv (it cannot be pushed
    into the next conv because of zero padding).
  * nn.Upsample(x2 nearest) -> conv   (network.py:188,195,199) and interpolate -> cat -> conv
    (network.py:96-101): the source is tagged `up2`; kernels index the low-res tensor directly.
  * conv8up(f7) + conv3short8(f3) -> ReLU (network.py:228,190): two sources accumulated in one op.
  * torch.cat((a, b), 1) -> conv: two sources with a split of the weight's input channels.
  * spectral_norm in eval mode: W = weight_orig / (u . (W_mat v)) with the stored u, v.
"""
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

BN_EPS = 1e-5


@dataclass
class Src:
    buf: str                 # name of the activation buffer read
    wkey: str                # state_dict prefix of the conv whose weight this source uses
    cin: Tuple[int, int]     # [lo, hi) slice of that weight's input channels
    up2: bool = False        # source is nearest-upsampled x2 before the conv


@dataclass
class ConvOp:
    name: str
    srcs: List[Src]
    cout: int
    out: str
    kind: str = "conv3"      # 'conv3' (3x3, pad 1) | 'deconv4' (ConvTranspose2d 4x4, stride 2, pad 1)
    stride: int = 1
    sn: bool = False         # weight is spectral-normalised (weight_orig/u/v)
    bias: bool = True
    fold_bn: Optional[str] = None   # BN applied before the activation -> folded into W, b
    act: str = "none"        # 'none' | 'relu' | 'lrelu'
    slope: float = 0.0
    post_bn: Optional[str] = None   # BN applied after the activation -> epilogue affine
    res: Optional[str] = None       # buffer added before the activation
    head: Optional[str] = None      # 'softmax9' | 'tanh2' (fp32 NCHW outputs)
    scale: int = 1           # spatial size of `out` relative to the network input = 1/scale

    @property
    def cin(self):
        return sum(s.cin[1] - s.cin[0] for s in self.srcs)


def _seg(name, srcs, cout, out, scale, stride=1):
    cin_off = 0
    ss = []
    for buf, c in srcs:
        ss.append(Src(buf, f"segnet.net.{name}.0", (cin_off, cin_off + c)))
        cin_off += c
    return ConvOp(f"segnet.net.{name}", ss, cout, out, stride=stride, bias=False,
                  fold_bn=f"segnet.net.{name}.1", act="lrelu", slope=0.1, scale=scale)


def _segdc(name, src, cin, cout, out, scale):
    return ConvOp(f"segnet.net.{name}", [Src(src, f"segnet.net.{name}.0", (0, cin))], cout, out,
                  kind="deconv4", act="lrelu", slope=0.1, scale=scale)


def segnet_ops() -> List[ConvOp]:
    """SpixelNet (models/network.py:260-313): input 'gray' -> 'affinity' (softmax over 9)."""
    ops = [
        _seg("conv0a", [("gray", 1)], 16, "sg.0a", 1),
        _seg("conv0b", [("sg.0a", 16)], 16, "sg.out1", 1),
        _seg("conv1a", [("sg.out1", 16)], 32, "sg.1a", 2, stride=2),
        _seg("conv1b", [("sg.1a", 32)], 32, "sg.out2", 2),
        _seg("conv2a", [("sg.out2", 32)], 64, "sg.2a", 4, stride=2),
        _seg("conv2b", [("sg.2a", 64)], 64, "sg.out3", 4),
        _seg("conv3a", [("sg.out3", 64)], 128, "sg.3a", 8, stride=2),
        _seg("conv3b", [("sg.3a", 128)], 128, "sg.out4", 8),
        _seg("conv4a", [("sg.out4", 128)], 256, "sg.4a", 16, stride=2),
        _seg("conv4b", [("sg.4a", 256)], 256, "sg.out5", 16),
        _segdc("deconv3", "sg.out5", 256, 128, "sg.dc3", 8),
        _seg("conv3_1", [("sg.out4", 128), ("sg.dc3", 128)], 128, "sg.c31", 8),
        _segdc("deconv2", "sg.c31", 128, 64, "sg.dc2", 4),
        _seg("conv2_1", [("sg.out3", 64), ("sg.dc2", 64)], 64, "sg.c21", 4),
        _segdc("deconv1", "sg.c21", 64, 32, "sg.dc1", 2),
        _seg("conv1_1", [("sg.out2", 32), ("sg.dc1", 32)], 32, "sg.c11", 2),
        _segdc("deconv0", "sg.c11", 32, 16, "sg.dc0", 1),
        _seg("conv0_1", [("sg.out1", 16), ("sg.dc0", 16)], 16, "sg.c01", 1),
        ConvOp("segnet.net.pred_mask0", [Src("sg.c01", "segnet.net.pred_mask0", (0, 16))], 9, "affinity",
               head="softmax9", scale=1),
    ]
    return ops


def repnet_ops() -> List[ConvOp]:
    """ColorProbNet (models/network.py:147-236): input 'gray' -> 'pred_feats' (64 ch, >= 0)."""
    P = "repnet."
    ops: List[ConvOp] = []

    def sn(block, idx, src, cin, cout, out, scale, stride=1, post_bn=None):
        key = f"{P}{block}.{idx}"
        ops.append(ConvOp(key, [Src(src, key, (0, cin))], cout, out, stride=stride, sn=True, act="lrelu",
                          slope=0.2, post_bn=(f"{P}{block}.{post_bn}" if post_bn is not None else None),
                          scale=scale))

    sn("conv1_2", 0, "gray", 1, 64, "rp.1a", 1)
    sn("conv1_2", 2, "rp.1a", 64, 64, "rp.f1", 1, post_bn=4)
    sn("conv2_3", 0, "rp.f1", 64, 128, "rp.2a", 2, stride=2)
    sn("conv2_3", 2, "rp.2a", 128, 128, "rp.2b", 2)
    sn("conv2_3", 4, "rp.2b", 128, 128, "rp.f2", 2, post_bn=6)
    sn("conv3_3", 0, "rp.f2", 128, 256, "rp.3a", 4, stride=2)
    sn("conv3_3", 2, "rp.3a", 256, 256, "rp.3b", 4)
    sn("conv3_3", 4, "rp.3b", 256, 256, "rp.f3", 4, post_bn=6)
    sn("conv4_3", 0, "rp.f3", 256, 512, "rp.4a", 8, stride=2)
    sn("conv4_3", 2, "rp.4a", 512, 512, "rp.4b", 8)
    sn("conv4_3", 4, "rp.4b", 512, 512, "rp.f4", 8, post_bn=6)
    prev = "rp.f4"
    for blk in ("conv5_3", "conv6_3", "conv7_3"):
        n = blk[4]
        sn(blk, 0, prev, 512, 512, f"rp.{n}a", 8)
        sn(blk, 2, f"rp.{n}a", 512, 512, f"rp.{n}b", 8)
        sn(blk, 4, f"rp.{n}b", 512, 512, f"rp.f{n}", 8, post_bn=6)
        prev = f"rp.f{n}"
    # f8_up = conv8up(up2(f7)) + conv3short8(f3); conv8_3 starts with ReLU (network.py:188-190,228)
    ops.append(ConvOp(P + "conv8up+conv3short8",
                      [Src("rp.f7", P + "conv8up.1", (0, 512), up2=True),
                       Src("rp.f3", P + "conv3short8.0", (0, 256))],
                      256, "rp.8a", act="relu", scale=4))
    ops.append(ConvOp(P + "conv8_3.1", [Src("rp.8a", P + "conv8_3.1", (0, 256))], 256, "rp.8b", act="relu", scale=4))
    ops.append(ConvOp(P + "conv8_3.3", [Src("rp.8b", P + "conv8_3.3", (0, 256))], 256, "rp.f8", act="relu",
                      post_bn=P + "conv8_3.5", scale=4))
    ops.append(ConvOp(P + "conv9up.1", [Src("rp.f8", P + "conv9up.1", (0, 256), up2=True)], 128, "rp.9a", scale=2))
    ops.append(ConvOp(P + "conv9_2.0", [Src("rp.9a", P + "conv9_2.0", (0, 128))], 128, "rp.f9", act="relu",
                      post_bn=P + "conv9_2.2", scale=2))
    # conv10_2 starts with ReLU -> fused into conv10up's epilogue (network.py:199-201)
    ops.append(ConvOp(P + "conv10up.1", [Src("rp.f9", P + "conv10up.1", (0, 128), up2=True)], 64, "rp.10a",
                      act="relu", scale=1))
    ops.append(ConvOp(P + "conv10_2.1", [Src("rp.10a", P + "conv10_2.1", (0, 64))], 64, "pred_feats", act="relu",
                      scale=1))
    return ops


def enhancenet_ops(res_num=3) -> List[ConvOp]:
    """HourGlass2(65 -> 2) + tanh (models/network.py:125-144, models/model.py:196-197).

    Input is cat[gray, full_feats]: channel 0 of inConv's weight meets 'gray', channels 1..64 meet
    'full_feats' (the upfeat output).
    """
    P = "enhanceNet."
    ops: List[ConvOp] = []

    def c(key, src, cin, cout, out, scale, stride=1, act="relu", post_bn=None, res=None, sn_=False):
        ops.append(ConvOp(P + key, [Src(src, P + key, (0, cin))], cout, out, stride=stride, act=act,
                          post_bn=(P + post_bn if post_bn else None), res=res, sn=sn_, scale=scale))

    ops.append(ConvOp(P + "inConv.inConv.0",
                      [Src("gray", P + "inConv.inConv.0", (0, 1)), Src("full_feats", P + "inConv.inConv.0", (1, 65))],
                      64, "en.1a", act="relu", scale=1))
    c("inConv.conv.0", "en.1a", 64, 64, "en.f1", 1, post_bn="inConv.conv.2")
    c("down1.conv.0", "en.f1", 64, 128, "en.2a", 2, stride=2)
    c("down1.conv.2", "en.2a", 128, 128, "en.f2", 2, post_bn="down1.conv.4")
    c("down2.conv.0", "en.f2", 128, 256, "en.3a", 4, stride=2)
    c("down2.conv.2", "en.3a", 256, 256, "en.f3", 4, post_bn="down2.conv.4")
    x = "en.f3"
    for i in range(res_num):
        c(f"residual.{i}.conv.0", x, 256, 256, f"en.r{i}a", 4, act="none")
        c(f"residual.{i}.conv.1", f"en.r{i}a", 256, 256, f"en.r{i}b", 4, sn_=True)
        c(f"residual.{i}.conv.3", f"en.r{i}b", 256, 256, f"en.r{i}", 4, res=x)
        x = f"en.r{i}"
    for lvl, cin, cout, skip, scale, outbuf in ((2, 256, 128, "en.f2", 2, "en.u2"), (1, 128, 64, "en.f1", 1, "en.u1")):
        c(f"up{lvl}.conv1", x, cin, cout, f"en.u{lvl}a", scale * 2, act="none")
        ops.append(ConvOp(P + f"up{lvl}.combine",
                          [Src(f"en.u{lvl}a", P + f"up{lvl}.combine", (0, cout), up2=True),
                           Src(skip, P + f"up{lvl}.combine", (cout, 2 * cout))],
                          cout, f"en.u{lvl}b", act="relu", scale=scale))
        c(f"up{lvl}.conv2.0", f"en.u{lvl}b", cout, cout, f"en.u{lvl}c", scale)
        c(f"up{lvl}.conv2.2", f"en.u{lvl}c", cout, cout, outbuf, scale, post_bn=f"up{lvl}.conv2.4")
        x = outbuf
    ops.append(ConvOp(P + "outConv", [Src(x, P + "outConv", (0, 64))], 2, "pred_colors", head="tanh2", scale=1))
    return ops


# ----------------------------------------------------------------------------------------------
# state_dict schema
# ----------------------------------------------------------------------------------------------

def _conv_entries(op: ConvOp):
    """(key, shape, is_param) for every conv weight/bias touched by `op` (deduplicated by caller)."""
    out = []
    seen = set()
    for s in op.srcs:
        if s.wkey in seen:
            continue
        seen.add(s.wkey)
        cin_total = sum(t.cin[1] - t.cin[0] for t in op.srcs if t.wkey == s.wkey)
        if op.kind == "deconv4":
            wshape = (cin_total, op.cout, 4, 4)
        else:
            wshape = (op.cout, cin_total, 3, 3)
        if op.sn:
            out.append((s.wkey + ".bias", (op.cout,), True))
            out.append((s.wkey + ".weight_orig", wshape, True))
            out.append((s.wkey + ".weight_u", (op.cout,), False))
            out.append((s.wkey + ".weight_v", (cin_total * 9,), False))
        else:
            out.append((s.wkey + ".weight", wshape, True))
            if op.bias:
                out.append((s.wkey + ".bias", (op.cout,), True))
    for bn in (op.fold_bn, op.post_bn):
        if bn:
            out.append((bn + ".weight", (op.cout,), True))
            out.append((bn + ".bias", (op.cout,), True))
            out.append((bn + ".running_mean", (op.cout,), False))
            out.append((bn + ".running_var", (op.cout,), False))
            out.append((bn + ".num_batches_tracked", (), False))
    return out


D_MODEL, N_HEAD, D_FF, N_LAYERS, N_VOCAB = 64, 8, 256, 6, 313


def transformer_entries(stack: str):
    """`TransformerEncoder` of 6 `EncoderLayer`s (models/transformer2d.py:9-60)."""
    out = []
    for i in range(N_LAYERS):
        p = f"{stack}.layers.{i}."
        out += [
            (p + "self_attn.in_proj_weight", (3 * D_MODEL, D_MODEL), True),
            (p + "self_attn.in_proj_bias", (3 * D_MODEL,), True),
            (p + "self_attn.out_proj.weight", (D_MODEL, D_MODEL), True),
            (p + "self_attn.out_proj.bias", (D_MODEL,), True),
            (p + "linear1.weight", (D_FF, D_MODEL), True),
            (p + "linear1.bias", (D_FF,), True),
            (p + "linear2.weight", (D_MODEL, D_FF), True),
            (p + "linear2.bias", (D_MODEL,), True),
            (p + "norm1.weight", (D_MODEL,), True),
            (p + "norm1.bias", (D_MODEL,), True),
            (p + "norm2.weight", (D_MODEL,), True),
            (p + "norm2.bias", (D_MODEL,), True),
        ]
    return out


def schema(enhanced=True, nets=("segnet", "repnet", "enhanceNet", "tokens")):
    """Ordered list of (key, shape, is_param) matching `AnchorColorProb.state_dict()` of the reference."""
    entries = []
    groups = []
    if "segnet" in nets:
        groups.append(segnet_ops())
    if "repnet" in nets:
        groups.append(repnet_ops())
    if "enhanceNet" in nets and enhanced:
        groups.append(enhancenet_ops())
    seen = set()
    for ops in groups:
        for op in ops:
            for e in _conv_entries(op):
                if e[0] not in seen:
                    seen.add(e[0])
                    entries.append(e)
    if "tokens" in nets:
        entries += transformer_entries("wildpath")
        entries += transformer_entries("hintpath")
        entries += [("mid_word_prj.weight", (N_VOCAB, D_MODEL), True),
                    ("trg_word_emb.weight", (D_MODEL, D_MODEL + N_VOCAB + 1), True),
                    ("trg_word_prj.weight", (N_VOCAB, D_MODEL), True)]
    return entries


# ----------------------------------------------------------------------------------------------
# folding: state_dict -> per-op effective tensors (fp32, on the state_dict's device)
# ----------------------------------------------------------------------------------------------

def effective_weight(sd, wkey, sn):
    """Conv weight as the reference's eval-mode forward sees it.

    spectral_norm eval path: sigma = u . (W_mat v), W = weight_orig / sigma (torch
    nn/utils/spectral_norm.py `compute_weight` with do_power_iteration=False; call sites
    models/network.py:152-185, :36).
    """
    import torch
    if not sn:
        return sd[wkey + ".weight"].float()
    w = sd[wkey + ".weight_orig"].float()
    u = sd[wkey + ".weight_u"].float()
    v = sd[wkey + ".weight_v"].float()
    sigma = torch.dot(u, torch.mv(w.reshape(w.shape[0], -1), v))
    return w / sigma


def bn_affine(sd, key):
    """Eval-mode BatchNorm2d as y = a*x + b (per channel)."""
    import torch
    a = sd[key + ".weight"].float() / torch.sqrt(sd[key + ".running_var"].float() + BN_EPS)
    b = sd[key + ".bias"].float() - sd[key + ".running_mean"].float() * a
    return a, b


@dataclass
class FoldedOp:
    op: ConvOp
    weights: list            # per source: conv3 -> (cout, cin_s, 3, 3); deconv4 -> (cin_s, cout, 4, 4)
    bias: object             # (cout,)
    post_scale: object = None
    post_shift: object = None


def fold(sd, op: ConvOp) -> FoldedOp:
    import torch
    ws = []
    bias = None
    done_bias = set()
    for s in op.srcs:
        w = effective_weight(sd, s.wkey, op.sn)
        if op.kind == "deconv4":
            w = w[s.cin[0]:s.cin[1]]
        else:
            w = w[:, s.cin[0]:s.cin[1]]
        ws.append(w.contiguous())
        if op.bias and s.wkey not in done_bias:
            done_bias.add(s.wkey)
            b = sd[s.wkey + ".bias"].float()
            bias = b.clone() if bias is None else bias + b
    if bias is None:
        bias = torch.zeros(op.cout, dtype=torch.float32, device=ws[0].device)
    if op.fold_bn:
        a, b = bn_affine(sd, op.fold_bn)
        ws = [(w * a.view(-1, 1, 1, 1)).contiguous() for w in ws]
        bias = bias * a + b
    f = FoldedOp(op, ws, bias.contiguous())
    if op.post_bn:
        f.post_scale, f.post_shift = (t.contiguous() for t in bn_affine(sd, op.post_bn))
    return f
