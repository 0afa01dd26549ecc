"""Host-side executor of the DISCO forward on the CUDA library (one Engine per model instance).

Responsibilities (plumbing only -- all arithmetic on the path runs in libdisco_b200.so):
  * fold + pack the reference-schema state_dict into device tensors (netspec.fold);
  * own the activation workspace for each (batch, H, W) it has seen;
  * issue the launch plan of `AnchorColorProb.forward` (reference models/model.py:103-199,
    test_mode branch) on the caller's current CUDA stream;
  * reproduce the reference's host RNG consumption: `np.random.choice(S, K, replace=False)` once per
    image in batch order (clusterkit.py:107) and one `torch.randint(S, (1,))` per empty cluster
    (clusterkit.py:182).
"""
import collections
import ctypes as C
import math
import os

import numpy as np
import torch

from . import _lib, netspec
from .cielab import Q_TO_AB

_DT = {"fp32": (_lib.F32, torch.float32), "bf16": (_lib.BF16, torch.bfloat16)}
_ACT = {"none": _lib.ACT_NONE, "relu": _lib.ACT_RELU, "lrelu": _lib.ACT_LRELU}
_HEAD = {None: _lib.HEAD_NONE, "softmax9": _lib.HEAD_SOFTMAX9, "tanh2": _lib.HEAD_TANH2, "raw2": _lib.HEAD_RAW2}
N_DRAWS = 4096


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def position_table(h, w, device):
    """PositionEmbeddingSine(32, normalize=True) as an (S, 64) table (models/position_encoding.py:26-47).
    Input-independent: computed once per grid size on the host and cached on the device."""
    ones = torch.ones(h, w)
    y = ones.cumsum(0, dtype=torch.float32)
    x = ones.cumsum(1, dtype=torch.float32)
    y = y / (y[-1:, :] + 1e-6) * (2 * math.pi)
    x = x / (x[:, -1:] + 1e-6) * (2 * math.pi)
    d = torch.arange(32, dtype=torch.float32)
    d = 10000.0 ** (2 * torch.div(d, 2, rounding_mode="floor") / 32)
    px, py = x[:, :, None] / d, y[:, :, None] / d
    px = torch.stack((px[:, :, 0::2].sin(), px[:, :, 1::2].cos()), dim=3).flatten(2)
    py = torch.stack((py[:, :, 0::2].sin(), py[:, :, 1::2].cos()), dim=3).flatten(2)
    return torch.cat((py, px), dim=2).reshape(h * w, 64).contiguous().to(device)


class _PackedConv:
    """Device-resident folded parameters of one fused conv op."""

    def __init__(self, folded, device):
        op = folded.op
        blocks, offs, off = [], [], 0
        for w in folded.weights:
            if op.kind == "deconv4":
                blk = w.permute(2, 3, 0, 1).reshape(16, w.shape[0], w.shape[1])   # [tap][cin][cout]
            else:
                blk = w.permute(2, 3, 1, 0).reshape(9, w.shape[1], w.shape[0])
            offs.append(off)
            off += blk.numel()
            blocks.append(blk.reshape(-1))
        self.op = op
        self.w32_host = torch.cat(blocks).contiguous()
        self.w32 = self.w32_host.to(device)
        self.w16 = None        # tensor-core (bf16) packing, filled on first use
        self.w_off = offs
        self.bias = folded.bias.to(device)
        # host copies: narrow tensor-core layers take their epilogue parameters through the kernel-parameter block
        self.bias_host = folded.bias.detach().float().contiguous().cpu()
        self.post_scale_host = folded.post_scale.detach().float().contiguous().cpu() if folded.post_scale is not None else None
        self.post_shift_host = folded.post_shift.detach().float().contiguous().cpu() if folded.post_shift is not None else None
        self.post_scale = folded.post_scale.to(device) if folded.post_scale is not None else None
        self.post_shift = folded.post_shift.to(device) if folded.post_shift is not None else None


class _FusedOp:
    """Profile record of a launch that covers several ConvOps of the plan (same fields profile_convs reads)."""

    def __init__(self, name, ops):
        self.name, self.ops = name, ops
        self.cout, self.cin, self.stride, self.kind, self.srcs = ops[-1].cout, ops[0].cin, 1, "fused", ops[0].srcs


class Engine:
    def __init__(self, state_dict, device, precision="bf16", n_clusters=8, sp_size=16, enhanced=True, random_hint=False):
        if precision not in _DT:
            raise ValueError(f"precision must be one of {list(_DT)}")
        if sp_size != 16:
            raise _lib.DiscoError("only sp_size=16 is built (all BASELINE configs); got %d" % sp_size)
        device = torch.device(device)
        if device.type != "cuda":
            raise _lib.DiscoError("disentangledcolorization_b200 runs on a CUDA (B200) device only; "
                                  "there is no CPU path -- move the model with .cuda()")
        self.device = device
        self.handle = _lib.Handle.get(device.index if device.index is not None else torch.cuda.current_device())
        self.lib = self.handle.lib
        self.precision = precision
        self.dt_code, self.dt_torch = _DT[precision]
        self.n_clusters = n_clusters
        self.enhanced = enhanced
        self.random_hint = random_hint
        # activation workspaces, least recently used first; bounded (the CLI's --no_resize mode feeds a new
        # image size per file and every (batch, H, W) owns ~9 GB at batch 64)
        self._ws = collections.OrderedDict()
        self.max_workspaces = 4
        self.max_workspace_bytes = 64 << 30
        self._pos = {}
        self._pending_rng = None
        self._prof = None
        self.use_graph = False        # replay the whole forward from a CUDA graph (fixed shapes, sampled_T == 0)
        self.static_outputs = False   # graph mode: return the graph-owned output tensors (valid until the next forward)
        self.lazy_rng = False         # advance torch's CPU generator at the next forward / sync_rng() instead of syncing now
        self.batched_diverse = False  # extension: sampled_T > 0 on batches (3N variants, variant-major)
        self._graphs = {}
        self.load_state_dict(state_dict)

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, sd):
        sd = {k: v.detach().to("cpu") for k, v in sd.items()}
        dev = self.device
        groups = [("segnet", netspec.segnet_ops()), ("repnet", netspec.repnet_ops())]
        if self.enhanced:
            groups.append(("enhanceNet", netspec.enhancenet_ops()))
        self.convs = {}
        for name, ops in groups:
            packed = []
            for op in ops:
                for f in self._maybe_split(netspec.fold(sd, op)):
                    packed.append(_PackedConv(f, dev))
            self.convs[name] = packed
        # fused head of SpixelNet (conv0a -> conv0b -> conv1a in one launch, csrc/segnet_fused.cu), bf16 path
        self.seg_head = None
        seg = self.convs["segnet"]
        if (self.precision == "bf16" and os.environ.get("DISCO_SEG_FUSED", "1") != "0"
                and [pc.op.name.rsplit(".", 1)[-1] for pc in seg[:3]] == ["conv0a", "conv0b", "conv1a"]):
            def tc_pack(pc, cin, cout):          # fp32 [tap][cin][cout] -> bf16 [tap][cout][cin]
                w = pc.w32_host.view(9, cin, cout).permute(0, 2, 1).contiguous().to(torch.bfloat16)
                return w.view(torch.int16).to(dev)
            self.seg_head = dict(w0b=tc_pack(seg[1], 16, 16), w1a=tc_pack(seg[2], 16, 32), slope=float(seg[0].op.slope),
                                 op=_FusedOp("segnet.net.conv0a+conv0b+conv1a", [pc.op for pc in seg[:3]]))
        f = lambda k: sd[k].float().contiguous().to(dev)
        self.stacks = {}
        for stack in ("wildpath", "hintpath"):
            layers = []
            for i in range(netspec.N_LAYERS):
                p = f"{stack}.layers.{i}."
                layers.append(dict(in_w=f(p + "self_attn.in_proj_weight"), in_b=f(p + "self_attn.in_proj_bias"),
                                   out_w=f(p + "self_attn.out_proj.weight"), out_b=f(p + "self_attn.out_proj.bias"),
                                   l1_w=f(p + "linear1.weight"), l1_b=f(p + "linear1.bias"),
                                   l2_w=f(p + "linear2.weight"), l2_b=f(p + "linear2.bias"),
                                   n1_w=f(p + "norm1.weight"), n1_b=f(p + "norm1.bias"),
                                   n2_w=f(p + "norm2.weight"), n2_b=f(p + "norm2.bias")))
            self.stacks[stack] = layers
        # fused tensor-core stack kernel (bf16 path): weights split into bf16 hi + lo 64x64 chunks by the library
        self.fused_tokens = self.precision == "bf16" and os.environ.get("DISCO_TOKENS_FUSED", "1") != "0"
        self.stack_packed = {}
        if self.fused_tokens:
            for stack in ("wildpath", "hintpath"):
                arr = (_lib.EncoderLayerWeights * netspec.N_LAYERS)()
                keep = []
                for i in range(netspec.N_LAYERS):
                    p = f"{stack}.layers.{i}."
                    for field, key in (("in_w", "self_attn.in_proj_weight"), ("in_b", "self_attn.in_proj_bias"),
                                       ("out_w", "self_attn.out_proj.weight"), ("out_b", "self_attn.out_proj.bias"),
                                       ("l1_w", "linear1.weight"), ("l1_b", "linear1.bias"), ("l2_w", "linear2.weight"),
                                       ("l2_b", "linear2.bias"), ("n1_w", "norm1.weight"), ("n1_b", "norm1.bias"),
                                       ("n2_w", "norm2.weight"), ("n2_b", "norm2.bias")):
                        t = sd[p + key].float().contiguous()
                        keep.append(t)
                        setattr(arr[i], field, t.data_ptr())
                w = torch.empty(netspec.N_LAYERS * 12 * 2 * 4096, dtype=torch.int16)
                vec = torch.empty(netspec.N_LAYERS * 832, dtype=torch.float32)
                _lib.check(self.lib.disco_encoder_stack_pack(arr, netspec.N_LAYERS, C.c_void_p(w.data_ptr()),
                                                             C.c_void_p(vec.data_ptr())), "disco_encoder_stack_pack")
                self.stack_packed[stack] = (w.to(dev), vec.to(dev))
        self.mid_w = f("mid_word_prj.weight")
        self.trg_w = f("trg_word_prj.weight")
        emb = sd["trg_word_emb.weight"].float()
        self.emb_src = emb[:, :64].contiguous().to(dev)                 # (64, 64): acts on the token features
        self.emb_tab = emb[:, 64:].t().contiguous().to(dev)             # (314, 64): 313 label columns + mask column
        self.q_to_ab = torch.from_numpy(Q_TO_AB.copy()).to(dev)
        self._ws.clear()

    def _maybe_split(self, folded):
        """bf16 path only: an op that mixes a nearest-upsampled source with a full-resolution skip source and has
        Cout <= 64 (HourGlass2 up1.combine) is issued as two launches -- the 4-phase 2x2 convolution of the low-res
        source writes a bf16 partial sum, the plain 3x3 convolution of the skip source adds it as a residual.  In one
        launch the skip source has to be sampled on every second pixel (stride-2 TMA boxes, one 128-byte request per
        pixel), which is TMA-issue-bound for such narrow tiles: 1.07 ms vs 0.55 ms for the pair at batch 64."""
        op = folded.op
        if (self.precision != "bf16" or len(op.srcs) != 2 or not op.srcs[0].up2 or op.srcs[1].up2 or op.cout > 64
                or op.kind != "conv3" or op.res is not None or op.head is not None):
            return [folded]
        import copy
        op_a = copy.copy(op)
        op_a.name, op_a.srcs, op_a.out = op.name + "#up", [op.srcs[0]], op.out + "#partial"
        op_a.act, op_a.post_bn, op_a.fold_bn = "none", None, None
        fa = netspec.FoldedOp(op_a, [folded.weights[0]], torch.zeros_like(folded.bias))
        op_b = copy.copy(op)
        op_b.name, op_b.srcs, op_b.res = op.name + "#skip", [op.srcs[1]], op_a.out
        fb = netspec.FoldedOp(op_b, [folded.weights[1]], folded.bias, folded.post_scale, folded.post_shift)
        return [fa, fb]

    # ------------------------------------------------------------------ workspace / plan
    def _workspace(self, B, H, W):
        key = (B, H, W)
        ws = self._ws.get(key)
        if ws is not None:
            self._ws.move_to_end(key)
            return ws
        if H % 16 or W % 16 or H <= 0 or W <= 0:
            raise _lib.DiscoError(f"H and W must be positive multiples of 16 (got {H}x{W}); the reference has the "
                                  "same constraint (SpixelNet has four stride-2 stages)")
        dev = self.device
        h, w = H // 16, W // 16
        S = h * w
        if self.n_clusters > S:
            raise _lib.DiscoError(f"n_clusters={self.n_clusters} exceeds the {S} super-pixel tokens of a {H}x{W} image")
        bufs = {}
        plans = {}
        f32 = dict(dtype=torch.float32, device=dev)
        for net, packed in self.convs.items():
            descs = []
            for pc in packed:
                op = pc.op
                Ho, Wo = H // op.scale, W // op.scale
                if op.head == "softmax9":
                    out = torch.empty(B, 9, Ho, Wo, **f32)
                elif op.head in ("tanh2", "raw2"):
                    out = torch.empty(B, 2, Ho, Wo, **f32)
                else:
                    out = torch.empty(B, Ho, Wo, op.cout, dtype=self.dt_torch, device=dev)
                bufs[op.out] = out
                descs.append((pc, Ho, Wo))
            plans[net] = descs
        bufs["full_feats"] = torch.empty(B, H, W, 64, dtype=self.dt_torch, device=dev)
        M = B * S
        tok = dict(partial=torch.empty(B, h, w, 9, 68, **f32), tokens=torch.empty(B, S, 64, **f32),
                   spix_ab=torch.empty(B, 2, h, w, **f32), conf=torch.empty(B, S, **f32),
                   sizes=torch.empty(B, S, **f32), qkv=torch.empty(M, 192, **f32), att=torch.empty(M, 64, **f32),
                   x1=torch.empty(M, 64, **f32), hid=torch.empty(M, 256, **f32), xa=torch.empty(M, 64, **f32),
                   xb=torch.empty(M, 64, **f32), enc=torch.empty(M, 64, **f32), dec=torch.empty(M, 64, **f32),
                   hint_seq=torch.empty(M, 64, **f32), labels=torch.empty(M, dtype=torch.int32, device=dev),
                   assign=torch.empty(M, dtype=torch.int32, device=dev),
                   events=torch.zeros(B + 2, dtype=torch.int32, device=dev),
                   iters=torch.zeros(B, dtype=torch.int32, device=dev),
                   init_idx=torch.empty(B, self.n_clusters, dtype=torch.int32, device=dev),
                   draws=torch.zeros(N_DRAWS, dtype=torch.int32, device=dev),
                   # valid placeholders: graph warm-up / capture runs execute k-means before the first real draws
                   init_idx_host=torch.arange(self.n_clusters, dtype=torch.int32).repeat(B, 1).pin_memory(),
                   draws_host=torch.zeros(N_DRAWS, dtype=torch.int32).pin_memory(),
                   events_host=torch.zeros(B + 2, dtype=torch.int32).pin_memory())
        if (h, w) not in self._pos:
            self._pos[(h, w)] = position_table(h, w, dev)
        nbytes = sum(t.numel() * t.element_size() for t in list(bufs.values()) + list(tok.values()) if t.is_cuda)
        ws = dict(bufs=bufs, plans=plans, tok=tok, h=h, w=w, S=S, descs={}, nbytes=nbytes)
        self._ws[key] = ws
        self._evict(keep=key)
        return ws

    def _evict(self, keep):
        """Drops least-recently-used workspaces beyond `max_workspaces` / `max_workspace_bytes`, together with their CUDA
        graphs and the library's launch plans (which hold the dead buffers' addresses).  The two most recent workspaces
        always stay (--diverse uses the batch-1 and the batch-3 workspace in one forward)."""
        def over():
            return (len(self._ws) > self.max_workspaces
                    or sum(w["nbytes"] for w in self._ws.values()) > self.max_workspace_bytes)
        dropped = False
        while len(self._ws) > 2 and over():
            old = next(iter(self._ws))
            if old == keep:
                break
            del self._ws[old]
            self._graphs.pop(old, None)
            dropped = True
        if dropped:
            torch.cuda.synchronize(self.device)      # nothing in flight still reads the freed buffers
            _lib.check(self.lib.disco_conv_tc_cache_clear(self.handle.h), "disco_conv_tc_cache_clear")
            for w in self._ws.values():              # surviving descriptors are rebuilt lazily (their plans were dropped too)
                w["descs"] = {}

    def _conv_desc(self, ws, pc, Ho, Wo, B, bufs):
        op = pc.op
        d = _lib.ConvDesc()
        d.kind = _lib.DECONV4 if op.kind == "deconv4" else _lib.CONV3
        d.stride, d.dtype = op.stride, self.dt_code
        d.batch, d.Ho, d.Wo, d.Cout = B, Ho, Wo, op.cout
        d.n_src = len(op.srcs)
        for i, s in enumerate(op.srcs):
            t = bufs[s.buf]
            if s.buf == "gray":                     # (B,1,H,W) fp32 == NHWC with C = 1
                Hs, Ws, Cs, is32 = t.shape[2], t.shape[3], 1, 1
            else:
                Hs, Ws, Cs, is32 = t.shape[1], t.shape[2], t.shape[3], 0
            d.src[i].ptr = t.data_ptr()
            d.src[i].H, d.src[i].W, d.src[i].C = Hs, Ws, Cs
            d.src[i].up2, d.src[i].is_f32, d.src[i].w_off = int(s.up2), is32, pc.w_off[i]
        d.weights, d.bias = pc.w32.data_ptr(), pc.bias.data_ptr()
        d.post_scale = pc.post_scale.data_ptr() if pc.post_scale is not None else None
        d.post_shift = pc.post_shift.data_ptr() if pc.post_shift is not None else None
        d.bias_host = pc.bias_host.data_ptr()
        d.post_scale_host = pc.post_scale_host.data_ptr() if pc.post_scale_host is not None else None
        d.post_shift_host = pc.post_shift_host.data_ptr() if pc.post_shift_host is not None else None
        d.residual = bufs[op.res].data_ptr() if op.res else None
        d.act, d.slope, d.head = _ACT[op.act], op.slope, _HEAD[op.head]
        d.out = bufs[op.out].data_ptr()
        self.route(d, pc)
        return d

    def route(self, d, pc):
        """Points a bf16 descriptor at the tensor-core weight packing when the tcgen05 kernel takes it."""
        if d.dtype != _lib.BF16 or not self.lib.disco_conv_tc_supported(self.handle.h, C.byref(d)):
            return False
        if pc.w16 is None:
            n = int(self.lib.disco_conv_tc_weight_elems(C.byref(d)))
            w16 = torch.empty(n, dtype=torch.int16)
            _lib.check(self.lib.disco_conv_tc_pack_weights(C.byref(d), C.c_void_p(pc.w32_host.data_ptr()),
                                                           C.c_void_p(w16.data_ptr())), "disco_conv_tc_pack_weights")
            pc.w16 = w16.to(self.device)
        d.weights = pc.w16.data_ptr()
        for i in range(d.n_src):
            if d.src[i].is_f32:
                d.gray_weights = pc.w32.data_ptr() + 4 * int(d.src[i].w_off)
                d.gray_weights_host = pc.w32_host.data_ptr() + 4 * int(d.src[i].w_off)
        return True

    def _run_net(self, net, ws, B, gray, stream):
        bufs = ws["bufs"]
        bufs["gray"] = gray
        key = (net, gray.data_ptr())
        descs = ws["descs"].get(key)
        if descs is None:
            descs = [self._conv_desc(ws, pc, Ho, Wo, B, bufs) for pc, Ho, Wo in ws["plans"][net]]
            ws["descs"] = {k: v for k, v in ws["descs"].items() if k[0] != net}
            ws["descs"][key] = descs
        prof = self._prof
        skip = 0
        if net == "segnet" and self.seg_head is not None:
            sh, seg = self.seg_head, self.convs["segnet"]
            H, W = gray.shape[2], gray.shape[3]
            if prof is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            _lib.check(self.lib.disco_segnet_head(self.handle.h, _ptr(gray), _ptr(seg[0].w32), _ptr(seg[0].bias), _ptr(sh["w0b"]),
                                                  _ptr(seg[1].bias), _ptr(sh["w1a"]), _ptr(seg[2].bias), sh["slope"], B, H, W,
                                                  _ptr(bufs["sg.out1"]), _ptr(bufs["sg.1a"]), stream), "disco_segnet_head")
            if prof is not None:
                e1.record()
                prof.append((sh["op"], B, H, W, e0, e1))
            skip = 3
        for d, (pc, _, _) in list(zip(descs, ws["plans"][net]))[skip:]:
            if prof is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            _lib.check(self.lib.disco_conv(self.handle.h, C.byref(d), stream), "disco_conv")
            if prof is not None:
                e1.record()
                prof.append((pc.op, d.batch, d.Ho, d.Wo, e0, e1))

    # ------------------------------------------------------------------ token path
    def _linear(self, stream, X, W, Y, b=None, pos=None, pos_cols=0, S=0, col_scale=1.0, scale_cols=0, relu=False,
                residual=None, ln=None, hint=None, transpose_S=0):
        d = _lib.LinearDesc()
        d.X, d.W, d.b = X.data_ptr(), W.data_ptr(), (b.data_ptr() if b is not None else None)
        d.M, d.N, d.K = X.shape[0], W.shape[0], W.shape[1]
        d.pos, d.pos_cols, d.S = (pos.data_ptr() if pos is not None else None), pos_cols, S
        d.col_scale, d.scale_cols, d.relu = col_scale, scale_cols, int(relu)
        d.residual = residual.data_ptr() if residual is not None else None
        d.ln_gamma, d.ln_beta = (ln[0].data_ptr(), ln[1].data_ptr()) if ln is not None else (None, None)
        if hint is not None:
            d.hint_mask, d.labels, d.emb = hint[0].data_ptr(), hint[1].data_ptr(), hint[2].data_ptr()
        d.transpose_S, d.Y = transpose_S, Y.data_ptr()
        _lib.check(self.lib.disco_linear(self.handle.h, C.byref(d), stream), "disco_linear")

    def _encoder_stack(self, stack, x_in, out, ws, B, stream):
        """TransformerEncoder(use_dense_pos=True), 6 post-norm layers (models/transformer2d.py:17-28,52-60)."""
        tok, S = ws["tok"], ws["S"]
        pos = self._pos[(ws["h"], ws["w"])]
        if self.fused_tokens and S <= 1024:      # one launch per stack (csrc/encoder_stack.cu); larger images: per-layer path
            if "kv" not in tok:
                n = int(self.lib.disco_encoder_stack_scratch_elems(B, S))
                tok["kv"] = torch.zeros(n, dtype=torch.int16, device=self.device)
            w, vec = self.stack_packed[stack]
            _lib.check(self.lib.disco_encoder_stack(self.handle.h, _ptr(x_in), _ptr(pos), _ptr(w), _ptr(vec), netspec.N_LAYERS,
                                                    B, S, _ptr(tok["kv"]), _ptr(out), stream), "disco_encoder_stack")
            return
        x = x_in
        layers = self.stacks[stack]
        for i, L in enumerate(layers):
            self._linear(stream, x, L["in_w"], tok["qkv"], b=L["in_b"], pos=pos, pos_cols=128, S=S,
                         col_scale=8 ** -0.5, scale_cols=64)
            _lib.check(self.lib.disco_attention(self.handle.h, _ptr(tok["qkv"]), B, S, _ptr(tok["att"]), stream),
                       "disco_attention")
            y = out if i == len(layers) - 1 else (tok["xa"] if i % 2 == 0 else tok["xb"])
            _lib.check(self.lib.disco_encoder_tail(self.handle.h, _ptr(tok["att"]), _ptr(x), _ptr(y), x.shape[0],
                                                   _ptr(L["out_w"]), _ptr(L["out_b"]), _ptr(L["n1_w"]), _ptr(L["n1_b"]),
                                                   _ptr(L["l1_w"]), _ptr(L["l1_b"]), _ptr(L["l2_w"]), _ptr(L["l2_b"]),
                                                   _ptr(L["n2_w"]), _ptr(L["n2_b"]), stream), "disco_encoder_tail")
            x = y

    # ------------------------------------------------------------------ RNG protocol
    def sync_rng(self):
        """Public: bring torch's CPU generator to the reference's stream position (needed only with lazy_rng)."""
        self._resolve_rng()

    def _resolve_rng(self):
        """Advance torch's CPU generator by the number of empty-cluster draws the last forward consumed."""
        pend = self._pending_rng
        if pend is None:
            return
        self._pending_rng = None
        state, events_host, event, S = pend
        event.synchronize()
        if int(events_host[-1]) != 0:
            raise _lib.DiscoError("k-means consumed more than %d empty-cluster draws" % N_DRAWS)
        used = int(events_host[-2])
        if used:
            # the draws are taken from the generator as it stood when the forward was issued; if the caller has drawn from
            # torch's CPU generator since then (only possible with lazy_rng), rewinding it would silently replay their
            # numbers -- refuse instead
            if not torch.equal(torch.get_rng_state(), state):
                raise _lib.DiscoError("lazy_rng: torch's CPU generator was used between a forward that consumed empty-cluster "
                                      "draws and its sync_rng(); call engine.sync_rng() right after the forward")
            torch.randint(S, (used,))

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, gray, ab, sampled_T=0, hint_mask=None, sync_rng=True, init_idx=None):
        """Returns the reference 6-tuple (pal_logit, ref_logit, pred_colors, affinity_map, spix_colors, hint_mask)."""
        with torch.cuda.device(self.device):     # allocations, stream lookup and launches follow the model's device
            if (self.use_graph and self._prof is None and sampled_T == 0 and hint_mask is None and not self.random_hint
                    and gray.dim() == 4 and gray.is_cuda and ab.is_cuda):
                return self._forward_graph(gray, ab, sync_rng, init_idx)
            return self._forward_impl(gray, ab, sampled_T, hint_mask, sync_rng, init_idx)

    def _host_draws(self, tok, B, S, init_idx):
        """Host RNG consumption of one forward (see module docstring); fills the pinned staging buffers."""
        K = self.n_clusters
        if init_idx is not None:                 # sharded runs: rows drawn for the global batch (dist.py)
            tok["init_idx_host"].copy_(torch.as_tensor(np.asarray(init_idx, dtype=np.int32)).view(B, K))
        else:                                    # B x np.random.choice(S, K, replace=False), one native call
            tok["init_idx_host"].copy_(torch.from_numpy(_lib.choice_rows(S, K, B)))
        state = torch.get_rng_state()
        tok["draws_host"].copy_(torch.randint(S, (N_DRAWS,)).to(torch.int32))
        torch.set_rng_state(state)
        return state

    def _forward_graph(self, gray, ab, sync_rng, init_idx):
        """Whole forward as one CUDA-graph replay: removes ~140 launch gaps per step (0.6 ms of 17.9 at batch 64).
        Inputs are copied into static buffers, the host RNG draws go through the pinned staging buffers the graph
        copies from, outputs are cloned out of the graph's static buffers."""
        self._resolve_rng()
        B, _, H, W = gray.shape
        dev = self.device
        key = (B, H, W)
        ws = self._workspace(B, H, W)
        tok, S = ws["tok"], ws["S"]
        entry = self._graphs.get(key)
        if entry is None:
            sg = gray.to(device=dev, dtype=torch.float32).contiguous().clone()
            sa = ab.to(device=dev, dtype=torch.float32).contiguous().clone()
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):                      # warm-up: builds every plan / tensor map outside capture
                for _ in range(2):
                    self._forward_impl(sg, sa, 0, None, False, None, in_graph=True)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            before = self.handle.launches()
            g0, g1 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(g0):
                self._forward_impl(sg, sa, 0, None, False, None, in_graph=True, part=0)
            with torch.cuda.graph(g1, pool=g0.pool()):
                outs = self._forward_impl(sg, sa, 0, None, False, None, in_graph=True, part=1)
            entry = dict(g0=g0, g1=g1, gray=sg, ab=sa, outs=outs, launches=self.handle.launches() - before)
            self._graphs[key] = entry
        entry["gray"].copy_(gray, non_blocking=True)
        entry["ab"].copy_(ab, non_blocking=True)
        entry["g0"].replay()
        # the host RNG draws (B x np.random.choice, ~1-2 ms at batch 64) overlap with the first half on the GPU
        state = self._host_draws(tok, B, S, init_idx)
        entry["g1"].replay()
        self.lib.disco_add_launch_count(self.handle.h, entry["launches"])
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))
        self._pending_rng = (state, tok["events_host"], ev, S)
        if sync_rng and not self.lazy_rng:
            self._resolve_rng()
        if self.static_outputs:
            return entry["outs"]
        return tuple(o.clone() if o is not None else None for o in entry["outs"])

    def _forward_impl(self, gray, ab, sampled_T=0, hint_mask=None, sync_rng=True, init_idx=None, in_graph=False, part=None):
        """`part` (graph capture only): 0 = first half up to pal_logit, 1 = the rest (pal_logit taken from the workspace)."""
        if not in_graph:
            self._resolve_rng()
        if gray.dim() != 4 or gray.shape[1] != 1:
            raise _lib.DiscoError(f"input_grays must be (N,1,H,W), got {tuple(gray.shape)}")
        B, _, H, W = gray.shape
        if B == 0:
            raise _lib.DiscoError("empty batch")
        if sampled_T > 0 and B != 1 and not self.batched_diverse:
            raise _lib.DiscoError("sampled_T > 0 (--diverse) needs a batch of 1, as in the reference "
                                  "(the expand at models/model.py:155-159 fails otherwise); set model.batched_diverse = True "
                                  "for the batched extension")
        dev = self.device
        gray = gray.to(device=dev, dtype=torch.float32).contiguous()
        ab = ab.to(device=dev, dtype=torch.float32).contiguous()
        if tuple(ab.shape) != (B, 2, H, W):
            raise _lib.DiscoError(f"input_colors must be ({B},2,{H},{W}), got {tuple(ab.shape)}")
        ws = self._workspace(B, H, W)
        bufs, tok, S, h, w = ws["bufs"], ws["tok"], ws["S"], ws["h"], ws["w"]
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        lib, hd = self.lib, self.handle.h
        M = B * S

        # ---- first half: affinity, features, tokens, colour-probability branch (model.py:104-135)
        affinity = bufs["affinity"]
        tokens = tok["tokens"].view(M, 64)
        if part in (None, 0):
            self._run_net("segnet", ws, B, gray, stream)                                   # model.py:104
            self._run_net("repnet", ws, B, gray, stream)                                   # model.py:105
            _lib.check(lib.disco_poolfeat(hd, self.dt_code, _ptr(bufs["pred_feats"]), _ptr(ab), _ptr(affinity), B, H, W, 64,
                                          _ptr(tok["partial"]), _ptr(tok["tokens"]), _ptr(tok["spix_ab"]), _ptr(tok["conf"]),
                                          _ptr(tok["sizes"]), stream), "disco_poolfeat")      # model.py:114-121
            self._encoder_stack("wildpath", tokens, tok["enc"], ws, B, stream)              # model.py:133
            pal_logit = torch.empty(B, 313, h, w, dtype=torch.float32, device=dev)
            self._linear(stream, tok["enc"], self.mid_w, pal_logit, transpose_S=S)          # model.py:134-135
            if part == 0:
                ws["pal_logit_static"] = pal_logit
                return (pal_logit,)
        else:
            pal_logit = ws["pal_logit_static"]

        # ---- anchors (model.py:140-141)
        if hint_mask is not None:
            hint = hint_mask.to(device=dev, dtype=torch.float32).reshape(B, 1, h, w).contiguous()
        elif self.random_hint:                                                          # anchor_gen.py:102-106
            import random
            mask = np.zeros((B, S), np.float32)
            for n in range(B):                                                          # basic.py:42-47
                mask[n, random.sample(range(0, S), random.randint(self.n_clusters, self.n_clusters))] = 1
            hint = torch.from_numpy(mask.reshape(B, 1, h, w)).to(dev)
        else:
            K = self.n_clusters
            state = None if in_graph else self._host_draws(tok, B, S, init_idx)
            tok["init_idx"].copy_(tok["init_idx_host"], non_blocking=True)
            tok["draws"].copy_(tok["draws_host"], non_blocking=True)
            hint = torch.empty(B, 1, h, w, dtype=torch.float32, device=dev)
            _lib.check(lib.disco_kmeans_anchor(hd, _ptr(tok["enc"]), _ptr(tok["init_idx"]), _ptr(tok["draws"]), N_DRAWS,
                                               _ptr(tok["sizes"]), B, S, K, 20, 1e-4, _ptr(tok["assign"]), _ptr(hint),
                                               _ptr(tok["events"]), _ptr(tok["iters"]), stream), "disco_kmeans_anchor")
            tok["events_host"].copy_(tok["events"], non_blocking=True)
            if not in_graph:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(dev))
                self._pending_rng = (state, tok["events_host"], ev, S)

        # ---- anchor colours and token labels (model.py:142-168)
        if sampled_T > 0:                                                               # model.py:148-159, N = 3 (x B: extension)
            B2 = 3 * B
            ws2 = self._workspace(B2, H, W)
            tok2 = ws2["tok"]
            spix_colors = torch.empty(B2, 2, h, w, dtype=torch.float32, device=dev)
            _lib.check(lib.disco_token_sample3(hd, _ptr(pal_logit), _ptr(self.q_to_ab), B, S, _ptr(tok2["labels"]),
                                               _ptr(spix_colors), stream), "disco_token_sample3")
            tok2["tokens"].copy_(tok["tokens"].repeat(3, 1, 1))                         # variant-major: index v * B + n
            hint = hint.repeat(3, 1, 1, 1)
            gray2 = gray.repeat(3, 1, 1, 1)
            ws2["bufs"]["affinity"].copy_(affinity.repeat(3, 1, 1, 1))
        else:
            B2, ws2, tok2, gray2 = B, ws, tok, gray
            spix_colors = torch.empty(B, 2, h, w, dtype=torch.float32, device=dev)
            if sampled_T < 0:                                                           # model.py:145-147,166
                spix_colors.copy_(tok["spix_ab"])
                _lib.check(lib.disco_token_labels(hd, 1, _ptr(tok["spix_ab"]), _ptr(self.q_to_ab), B, S,
                                                  _ptr(tok["labels"]), None, stream), "disco_token_labels")
            else:                                                                       # model.py:161,166
                _lib.check(lib.disco_token_labels(hd, 0, _ptr(pal_logit), _ptr(self.q_to_ab), B, S,
                                                  _ptr(tok["labels"]), _ptr(spix_colors), stream), "disco_token_labels")

        # ---- second half: hint embedding, hint path, un-pooling, enhancement (model.py:175-197)
        M2 = B2 * S
        bufs2 = ws2["bufs"]
        affinity2 = bufs2["affinity"]
        self._linear(stream, tok2["tokens"].view(M2, 64), self.emb_src, tok2["hint_seq"],
                     hint=(hint, tok2["labels"], self.emb_tab))                         # model.py:175-185
        self._encoder_stack("hintpath", tok2["hint_seq"], tok2["dec"], ws2, B2, stream)  # model.py:186
        ref_logit = torch.empty(B2, 313, h, w, dtype=torch.float32, device=dev)
        self._linear(stream, tok2["dec"], self.trg_w, ref_logit, transpose_S=S)         # model.py:187-189
        pred = None
        if self.enhanced:
            _lib.check(lib.disco_upfeat(hd, self.dt_code, _ptr(tok2["dec"]), _ptr(affinity2), B2, H, W, 64,
                                        _ptr(bufs2["full_feats"]), stream), "disco_upfeat")   # model.py:194-195
            self._run_net("enhanceNet", ws2, B2, gray2, stream)                         # model.py:196-197
            pred = bufs2["pred_colors"] if in_graph else bufs2["pred_colors"].clone()
        if sync_rng and not in_graph and not self.lazy_rng:
            self._resolve_rng()
        # workspace buffers are reused by the next call: hand out copies (the graph path copies once, after the replay)
        return pal_logit, ref_logit, pred, (affinity2 if in_graph else affinity2.clone()), spix_colors, hint

    @staticmethod
    def algorithmic_flops(op, B, Ho, Wo):
        """2 x MACs of the reference's own formulation of this op (9 taps at the output resolution for
        nn.Upsample -> conv; 16 taps per input pixel for ConvTranspose2d)."""
        if op.kind == "fused":
            return sum(Engine.algorithmic_flops(o, B, Ho // o.scale, Wo // o.scale) for o in op.ops)
        if op.kind == "deconv4":
            return 2.0 * B * (Ho // 2) * (Wo // 2) * op.cin * op.cout * 16
        return 2.0 * B * Ho * Wo * op.cin * op.cout * 9

    @staticmethod
    def executed_flops(op, B, Ho, Wo):
        """2 x MACs the kernels actually issue: a nearest-upsampled source is convolved as four 2x2-tap parity phases
        on the low-resolution map (4/9 of the reference formulation's MACs); everything else equals algorithmic_flops."""
        if op.kind == "fused":
            return sum(Engine.executed_flops(o, B, Ho // o.scale, Wo // o.scale) for o in op.ops)
        if op.kind == "deconv4":
            return 2.0 * B * (Ho // 2) * (Wo // 2) * op.cin * op.cout * 16
        f = 0.0
        for s in op.srcs:
            c = s.cin[1] - s.cin[0]
            f += 2.0 * B * Ho * Wo * c * op.cout * (4 if s.up2 else 9)
        return f

    def profile_convs(self, gray, ab, steps=2):
        """Times every disco_conv launch of `steps` forwards with CUDA events on the launching stream."""
        self.forward(gray, ab)
        torch.cuda.synchronize()
        self._prof = []
        try:
            for _ in range(steps):
                self.forward(gray, ab)
            torch.cuda.synchronize()
            recs = self._prof
        finally:
            self._prof = None
        per_op = {}
        flops = ms = 0.0
        for op, B, Ho, Wo, e0, e1 in recs:
            t = e0.elapsed_time(e1)
            f = self.algorithmic_flops(op, B, Ho, Wo)
            flops += f
            ms += t
            a = per_op.setdefault(op.name, [0.0, 0.0, op.cout, 0.0, op])
            a[0] += f
            a[1] += t
            a[3] += self.executed_flops(op, B, Ho, Wo)
        top = sorted(per_op.items(), key=lambda kv: -kv[1][1])[:6]
        return {"flops": flops / steps, "ms": ms / steps,
                "top": [{"op": k, "ms": v[1] / steps, "tflops": v[0] / (v[1] / 1e3) / 1e12} for k, v in top],
                "per_op": {k: {"ms": v[1] / steps, "tflops": v[0] / (v[1] / 1e3) / 1e12, "flops": v[0] / steps, "cout": v[2],
                               "executed_flops": v[3] / steps, "cin": v[4].cin, "stride": v[4].stride, "n_src": len(v[4].srcs),
                               "up2": any(s.up2 for s in v[4].srcs), "kind": v[4].kind}
                           for k, v in per_op.items()}}

    def kmeans_iterations(self, B, H, W):
        return self._ws[(B, H, W)]["tok"]["iters"].cpu()

    def kmeans_draws(self, B, H, W):
        """Empty-cluster draws the last k-means of this shape consumed (total over the batch)."""
        torch.cuda.synchronize(self.device)
        return int(self._ws[(B, H, W)]["tok"]["events"][B].item())
