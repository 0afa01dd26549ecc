"""TEST INFRASTRUCTURE -- CPU oracle for the DISCO colorization forward (fp32, torch on CPU).

This file is a *checker*, not product code: only tests/, __graft_entry__.smoke() and the
`cpu_baseline` / `--impl reference` legs of bench.py may import it.  The product package
(disentangledcolorization_b200/) never imports it and fails loudly without its CUDA library.

It restates, as plain functions over a reference-schema `state_dict`, the algorithm of the
reference's eval-mode forward (each function cites the reference lines it follows).  The arithmetic
lives in a third-party dependency of the reference (torch; requirements.txt:12 `torch>=1.8.0`):
convolution, batch-norm, layer-norm, softmax etc. are called through `torch.nn.functional` here as
the reference does through `torch.nn`.  The oracle runs on the installed torch (2.11), the reference
pinned 1.8.0 -- "parity unpinned" with respect to the torch version.

Pinning: the reference ships no tests, golden vectors or checkpoints.  This oracle is pinned
against outputs of the reference itself, generated in the authoring container by
oracle/make_golden.py (which imports /root/reference through oracle/ref_harness.py) and committed
under tests/golden/.  tests/test_oracle_golden.py re-checks the oracle against those fixtures.

Structure is deliberately independent of disentangledcolorization_b200/netspec.py (no folding, no
fusion: conv -> activation -> BN exactly in the reference's order) so that it also checks the
product's weight folding.
"""
import contextlib
import math

import numpy as np
import torch
import torch.nn.functional as F

N_VOCAB = 313

# ----------------------------------------------------------------------------------------------
# bf16-storage emulation (tolerance derivation only; the parity oracle proper is the fp32 path)
# ----------------------------------------------------------------------------------------------
# With `emulate_bf16()` active, conv weights and every activation map a bf16 pipeline would keep in
# memory between two convolutions are rounded to bf16 (round-to-nearest-even) while all arithmetic
# stays fp32 -- "torch fp32 vs torch with bf16 storage" on the reference's own formulation.  The
# distance between the two is the error ANY bf16-storage implementation of this network carries;
# oracle/derive_bf16_tolerance.py measures it and tests/ gate the CUDA bf16 path at 1.5x that.
_EMU = False


@contextlib.contextmanager
def emulate_bf16(on=True):
    global _EMU
    prev, _EMU = _EMU, bool(on)
    try:
        yield
    finally:
        _EMU = prev


def _q(x):
    return x.to(torch.bfloat16).to(torch.float32) if _EMU else x


# ----------------------------------------------------------------------------------------------
# gamut table  (utils/cielab.py:38-64, models/basic.py:150-152)
# ----------------------------------------------------------------------------------------------
def q_to_ab():
    from disentangledcolorization_b200.cielab import Q_TO_AB  # data table only
    return torch.from_numpy(Q_TO_AB.copy())


# ----------------------------------------------------------------------------------------------
# conv building blocks
# ----------------------------------------------------------------------------------------------
def _w(sd, key):
    """Eval-mode weight of a (possibly spectral-normalised) conv.  torch spectral_norm in eval:
    sigma = u.(W_mat v); W = weight_orig / sigma (call sites models/network.py:152-185,36)."""
    if key + ".weight" in sd:
        return _q(sd[key + ".weight"])
    w = sd[key + ".weight_orig"]
    sigma = torch.dot(sd[key + ".weight_u"], torch.mv(w.reshape(w.shape[0], -1), sd[key + ".weight_v"]))
    return _q(w / sigma)


def _conv(sd, key, x, stride=1):
    return F.conv2d(x, _w(sd, key), sd.get(key + ".bias"), stride=stride, padding=1)


def _bn(sd, key, x):
    return F.batch_norm(x, sd[key + ".running_mean"], sd[key + ".running_var"], sd[key + ".weight"],
                        sd[key + ".bias"], False, 0.0, 1e-5)


def spixelnet(sd, gray, prefix="segnet.net."):
    """SpixelNet.forward, models/network.py:293-313; conv()/deconv() blocks :240-258."""
    def cbl(name, x, stride=1):
        y = F.conv2d(x, _q(sd[prefix + name + ".0.weight"]), None, stride=stride, padding=1)
        return _q(F.leaky_relu(_bn(sd, prefix + name + ".1", y), 0.1))

    def dcl(name, x):
        y = F.conv_transpose2d(x, _q(sd[prefix + name + ".0.weight"]), sd[prefix + name + ".0.bias"], stride=2, padding=1)
        return _q(F.leaky_relu(y, 0.1))

    o1 = cbl("conv0b", cbl("conv0a", gray))
    o2 = cbl("conv1b", cbl("conv1a", o1, 2))
    o3 = cbl("conv2b", cbl("conv2a", o2, 2))
    o4 = cbl("conv3b", cbl("conv3a", o3, 2))
    o5 = cbl("conv4b", cbl("conv4a", o4, 2))
    c3 = cbl("conv3_1", torch.cat((o4, dcl("deconv3", o5)), 1))
    c2 = cbl("conv2_1", torch.cat((o3, dcl("deconv2", c3)), 1))
    c1 = cbl("conv1_1", torch.cat((o2, dcl("deconv1", c2)), 1))
    c0 = cbl("conv0_1", torch.cat((o1, dcl("deconv0", c1)), 1))
    mask = F.conv2d(c0, _q(sd[prefix + "pred_mask0.weight"]), sd[prefix + "pred_mask0.bias"], padding=1)
    return torch.softmax(mask, dim=1)


def colorprobnet(sd, gray, prefix="repnet."):
    """ColorProbNet.forward, models/network.py:220-236 (layer lists :152-201)."""
    def sn_block(name, x, n_convs, first_stride=1):
        for i in range(n_convs):
            x = F.leaky_relu(_conv(sd, f"{prefix}{name}.{2 * i}", x, first_stride if i == 0 else 1), 0.2)
            if i + 1 < n_convs:
                x = _q(x)
        return _q(_bn(sd, f"{prefix}{name}.{2 * n_convs}", x))

    up = lambda t: F.interpolate(t, scale_factor=2, mode="nearest")
    f1 = sn_block("conv1_2", gray, 2)
    f2 = sn_block("conv2_3", f1, 3, 2)
    f3 = sn_block("conv3_3", f2, 3, 2)
    f4 = sn_block("conv4_3", f3, 3, 2)
    f5 = sn_block("conv5_3", f4, 3)
    f6 = sn_block("conv6_3", f5, 3)
    f7 = sn_block("conv7_3", f6, 3)
    f8u = _conv(sd, prefix + "conv8up.1", up(f7)) + _conv(sd, prefix + "conv3short8.0", f3)
    x = _q(F.relu(f8u))
    x = _q(F.relu(_conv(sd, prefix + "conv8_3.1", x)))
    x = F.relu(_conv(sd, prefix + "conv8_3.3", x))
    f8 = _q(_bn(sd, prefix + "conv8_3.5", x))
    f9u = _q(_conv(sd, prefix + "conv9up.1", up(f8)))
    f9 = _q(_bn(sd, prefix + "conv9_2.2", F.relu(_conv(sd, prefix + "conv9_2.0", f9u))))
    f10u = _conv(sd, prefix + "conv10up.1", up(f9))
    return _q(F.relu(_conv(sd, prefix + "conv10_2.1", _q(F.relu(f10u)))))


def hourglass2(sd, x, prefix="enhanceNet.", res_num=3):
    """HourGlass2.forward, models/network.py:136-144; blocks :10-47,66-101."""
    r = lambda k, t, s=1: F.relu(_conv(sd, prefix + k, t, s))
    f1 = _q(_bn(sd, prefix + "inConv.conv.2", r("inConv.conv.0", _q(r("inConv.inConv.0", x)))))
    f2 = _q(_bn(sd, prefix + "down1.conv.4", r("down1.conv.2", _q(r("down1.conv.0", f1, 2)))))
    f3 = _q(_bn(sd, prefix + "down2.conv.4", r("down2.conv.2", _q(r("down2.conv.0", f2, 2)))))
    y = f3
    for i in range(res_num):
        t = _q(_conv(sd, f"{prefix}residual.{i}.conv.0", y))
        t = _q(F.relu(_conv(sd, f"{prefix}residual.{i}.conv.1", t)))
        t = _conv(sd, f"{prefix}residual.{i}.conv.3", t)
        y = _q(F.relu(y + t))

    def upblock(name, t, skip):
        t = F.interpolate(_q(_conv(sd, f"{prefix}{name}.conv1", t)), scale_factor=2, mode="nearest")
        t = _q(F.relu(_conv(sd, f"{prefix}{name}.combine", torch.cat((t, skip), 1))))
        t = r(f"{name}.conv2.2", _q(r(f"{name}.conv2.0", t)))
        return _q(_bn(sd, f"{prefix}{name}.conv2.4", t))

    y = upblock("up2", y, f2)
    y = upblock("up1", y, f1)
    return _conv(sd, prefix + "outConv", y)


# ----------------------------------------------------------------------------------------------
# superpixel pooling / un-pooling  (models/basic.py:274-376)
# ----------------------------------------------------------------------------------------------
_DIRS = [(ky, kx) for ky in (-1, 0, 1) for kx in (-1, 0, 1)]  # affinity channel k <-> neighbour offset


def poolfeat(feat, prob, sp=16):
    """basic.poolfeat (models/basic.py:274-324) with need_entry_prob=True.

    A pixel in cell (i, j) sends prob_k * [feat, 1] to cell (i+ky, j+kx); per-cell masses are
    16x16 *averages* (avg_pool2d), out-of-grid targets are dropped; result = feat_mass /
    (prob_mass + 1e-8).  Returns (pooled, prob_mass)."""
    b, c, H, W = feat.shape
    h, w = H // sp, W // sp
    ext = torch.cat([feat, feat.new_ones(b, 1, H, W)], 1)
    acc = feat.new_zeros(b, c + 1, h, w)
    for k, (ky, kx) in enumerate(_DIRS):
        mass = F.avg_pool2d(ext * prob[:, k:k + 1], sp, sp)           # mass emitted by each cell in direction k
        # receiving cell (i, j) <- emitting cell (i-ky, j-kx)
        ys0, ys1 = max(0, -ky), h - max(0, ky)                           # valid emitter rows
        xs0, xs1 = max(0, -kx), w - max(0, kx)
        acc[:, :, ys0 + ky:ys1 + ky, xs0 + kx:xs1 + kx] += mass[:, :, ys0:ys1, xs0:xs1]
    return acc[:, :-1] / (acc[:, -1:] + 1e-8), acc[:, -1:]


def get_spixel_size(affinity, sp=16):
    """basic.get_spixel_size (models/basic.py:327-335): hard (argmax, ties counted) assignment mass."""
    hard = (affinity == affinity.max(dim=1, keepdim=True)[0]).to(affinity.dtype)
    _, mass = poolfeat(affinity.new_ones(affinity.shape[0], 1, *affinity.shape[2:]), hard, sp)
    return mass


def upfeat(tok, prob, sp=16):
    """basic.upfeat (models/basic.py:338-376): pixel = sum_k prob_k * token of neighbour cell k
    (zero outside the grid)."""
    b, c, h, w = tok.shape
    padded = F.pad(tok, (1, 1, 1, 1))
    out = None
    for k, (ky, kx) in enumerate(_DIRS):
        nb = padded[:, :, 1 + ky:1 + ky + h, 1 + kx:1 + kx + w]
        term = F.interpolate(nb, size=(h * sp, w * sp), mode="nearest") * prob[:, k:k + 1]
        out = term if out is None else out + term
    return out


# ----------------------------------------------------------------------------------------------
# tokens: position encoding, transformer, labels
# ----------------------------------------------------------------------------------------------
def position_sine(h, w, num_pos_feats=32, temperature=10000.0):
    """PositionEmbeddingSine(normalize=True).forward, models/position_encoding.py:26-47 -> (64,h,w)."""
    ones = torch.ones(h, w)
    y = ones.cumsum(0, dtype=torch.float32)
    x = ones.cumsum(1, dtype=torch.float32)
    y = y / (y[-1:, :] + 1e-6) * (2 * math.pi)
    x = x / (x[:, -1:] + 1e-6) * (2 * math.pi)
    d = torch.arange(num_pos_feats, dtype=torch.float32)
    d = temperature ** (2 * torch.div(d, 2, rounding_mode="floor") / num_pos_feats)
    px, py = x[:, :, None] / d, y[:, :, None] / d
    px = torch.stack((px[:, :, 0::2].sin(), px[:, :, 1::2].cos()), dim=3).flatten(2)
    py = torch.stack((py[:, :, 0::2].sin(), py[:, :, 1::2].cos()), dim=3).flatten(2)
    return torch.cat((py, px), dim=2).permute(2, 0, 1)


def encoder_stack(sd, stack, x, pos, n_layers=6, n_head=8):
    """TransformerEncoder(use_dense_pos=True) of post-norm EncoderLayers in eval mode
    (models/transformer2d.py:17-28,52-60).  x, pos: (B, S, 64) batch-first here."""
    B, S, D = x.shape
    dh = D // n_head
    for i in range(n_layers):
        p = f"{stack}.layers.{i}."
        Wi, bi = sd[p + "self_attn.in_proj_weight"], sd[p + "self_attn.in_proj_bias"]
        qk_in = x + pos
        q = F.linear(qk_in, Wi[:D], bi[:D]).view(B, S, n_head, dh).transpose(1, 2)
        k = F.linear(qk_in, Wi[D:2 * D], bi[D:2 * D]).view(B, S, n_head, dh).transpose(1, 2)
        v = F.linear(x, Wi[2 * D:], bi[2 * D:]).view(B, S, n_head, dh).transpose(1, 2)
        att = torch.softmax((q * (dh ** -0.5)) @ k.transpose(-1, -2), dim=-1)
        o = (att @ v).transpose(1, 2).reshape(B, S, D)
        o = F.linear(o, sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"])
        x = F.layer_norm(x + o, (D,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
        f = F.linear(F.relu(F.linear(x, sd[p + "linear1.weight"], sd[p + "linear1.bias"])),
                     sd[p + "linear2.weight"], sd[p + "linear2.bias"])
        x = F.layer_norm(x + f, (D,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
    return x


def encode_ab2ind(ab, neighbours=5, sigma=5.0):
    """ColorLabel.encode_ab2ind (models/basic.py:177-194): soft 5-NN Gaussian code over the 313 bins.
    ab: (N,2,h,w) normalised by 110 -> (N,313,h,w)."""
    table = q_to_ab().to(ab.device)
    n, _, h, w = ab.shape
    pts = (ab * 110.0).permute(1, 0, 2, 3).reshape(2, -1)                 # (2, m)
    d = torch.cdist(table, pts.t())                                        # (313, m)
    nn_idx = d.argsort(dim=0)[:neighbours]                                 # (5, m)
    wts = []
    for i in range(neighbours):
        mu = table[nn_idx[i]].t()
        wts.append(torch.exp(-((mu - pts) ** 2).sum(0) / (2 * sigma ** 2)) / (2 * math.pi * sigma))
    wts = torch.stack(wts)
    wts = wts / wts.sum(0, keepdim=True)
    q = ab.new_zeros(N_VOCAB, pts.shape[1])
    q[nn_idx, torch.arange(pts.shape[1], device=ab.device).repeat(neighbours, 1)] = wts
    return q.reshape(N_VOCAB, n, h, w).permute(1, 0, 2, 3)


def decode_ind2ab(logit, T=0):
    """ColorLabel.decode_ind2ab (models/basic.py:196-218) for integer T: T-th most probable bin."""
    table = q_to_ab().to(logit.device)
    idx = torch.sort(torch.softmax(logit, dim=1), dim=1, descending=True)[1][:, T]    # (N,h,w)
    return table[idx].permute(0, 3, 1, 2) / 110.0


# ----------------------------------------------------------------------------------------------
# anchors: k-means + pick + colours
# ----------------------------------------------------------------------------------------------
def kmeans_lloyd(X, K, iter_limit=20, tol=1e-4):
    """clusterkit.kmeans with euclidean distance (models/clusterkit.py:112-208, init :99-109,
    distance :253-269).  Consumes np.random (init) and the torch CPU generator (empty clusters)
    exactly as the reference does."""
    idx = np.random.choice(len(X), K, replace=False)
    centers = X[idx]
    it = 0
    while True:
        dist = ((X[:, None, :] - centers[None, :, :]) ** 2.0).sum(-1)
        assign = dist.argmin(1)
        prev = centers.clone()
        for k in range(K):
            members = X[assign == k]
            if members.shape[0] == 0:
                members = X[torch.randint(len(X), (1,))]
            centers[k] = members.mean(0)
        shift = torch.sqrt(((centers - prev) ** 2).sum(1)).sum()
        it += 1
        if shift ** 2 < tol or it >= iter_limit:
            break
    return assign


def anchor_mask(tokens, K, spixel_sizes):
    """AnchorAnalysis.__call__ in 'clustering' mode (models/anchor_gen.py:92-101) on top of
    batch_kmeans_pytorch (models/clusterkit.py:31-58).  tokens: (N,C,h,w)."""
    N, C, h, w = tokens.shape
    masks = []
    for n in range(N):
        assign = kmeans_lloyd(tokens[n].permute(1, 2, 0).reshape(-1, C).float(), K)
        masks.append(F.one_hot(assign, K).t().reshape(1, K, h, w).float())
    cluster = torch.cat(masks, 0)
    score = (cluster + spixel_sizes * 0.01).flatten(2)
    picks = score.argmax(-1)                                              # (N,K)
    hint = F.one_hot(picks, h * w).float().sum(1, keepdim=True).view(N, 1, h, w)
    return hint, cluster


def sample_anchor_colors(prob, T=0, topk=10):
    """AnchorAnalysis._sample_anchor_colors (models/anchor_gen.py:54-90).  prob: (N,313,h,w)."""
    table = q_to_ab().to(prob.device)
    order = torch.sort(prob, dim=1, descending=True)[1][:, :topk]          # (N,topk,h,w)
    cand = table[order] / 110.0                                           # (N,topk,h,w,2)
    if T == 0:
        pick = cand[:, 0]
    else:
        d0 = torch.norm(cand - cand[:, :1], p=2, dim=4, keepdim=True)
        far = torch.sort(d0, dim=1, descending=True)[1][:, :1].expand(-1, -1, -1, -1, 2)
        if T == 1:
            pick = torch.gather(cand, 1, far).squeeze(1)
        else:
            c1 = torch.gather(cand, 1, far)
            d1 = torch.norm(cand - c1, p=2, dim=4, keepdim=True)
            sel = torch.sort(d0 + d1, dim=1, descending=True)[1][:, [T - 2]].expand(-1, -1, -1, -1, 2)
            pick = torch.gather(cand, 1, sel).squeeze(1)
    return pick.permute(0, 3, 1, 2)


# ----------------------------------------------------------------------------------------------
# the whole forward  (models/model.py:103-199, test_mode branch)
# ----------------------------------------------------------------------------------------------
def forward(sd, gray, ab, n_clusters=8, sampled_T=0, sp=16, hint_mask=None, stages=None):
    """AnchorColorProb.forward(input_grays, input_colors, test_mode=True, sampled_T) in eval mode.

    `hint_mask`: inject anchors instead of running k-means (used to decouple anchor selection from
    numerics in parity tests).  `stages`: optional dict that receives intermediate tensors.
    Returns the reference's 6-tuple (pal_logit, ref_logit, pred_colors, affinity, spix_colors, hint_mask).
    """
    gray, ab = gray.float(), ab.float()
    affinity = spixelnet(sd, gray)                                         # :104
    feats = colorprobnet(sd, gray)                                         # :105
    pooled, _ = poolfeat(torch.cat([feats, ab], 1), affinity, sp)          # :114-115
    tokens, spix_colors = pooled[:, :64], pooled[:, 64:]                   # :116-117
    N, C, h, w = tokens.shape
    pos = position_sine(h, w).to(gray.device).unsqueeze(0).expand(N, -1, -1, -1)           # :118
    sizes = get_spixel_size(affinity, sp)                                  # :121
    src = tokens.flatten(2).transpose(1, 2)                                # (N,S,64)
    pos_seq = pos.flatten(2).transpose(1, 2)
    enc = encoder_stack(sd, "wildpath", src, pos_seq)                      # :133
    pal_logit = F.linear(enc, sd["mid_word_prj.weight"]).transpose(1, 2).reshape(N, N_VOCAB, h, w)  # :134-135
    if hint_mask is None:
        hint_mask, _ = anchor_mask(enc.transpose(1, 2).reshape(N, C, h, w), n_clusters, sizes)      # :140-141
    prob = torch.softmax(pal_logit, dim=1)                                 # :142
    if sampled_T < 0:
        sampled = spix_colors                                              # :147
    elif sampled_T > 0:                                                    # :148-159 (N must be 1)
        sampled = torch.cat([sample_anchor_colors(prob, T=t) for t in (0, 1, 2)], 0)
        N = 3 * N
        gray, hint_mask, affinity = (t.expand(N, -1, -1, -1) for t in (gray, hint_mask, affinity))
        src, pos_seq = src.expand(N, -1, -1), pos_seq.expand(N, -1, -1)
    else:
        sampled = sample_anchor_colors(prob, T=0)                          # :161
    labels = encode_ab2ind(sampled).max(dim=1, keepdim=True)[1]            # :166
    onehot = F.one_hot(labels.squeeze(1), N_VOCAB).float().flatten(1, 2)   # (N,S,313)   :183-184
    m = hint_mask.flatten(2).transpose(1, 2)                               # (N,S,1)
    hint_seq = F.linear(torch.cat([src, m * onehot, m], 2), sd["trg_word_emb.weight"])   # :185
    dec = encoder_stack(sd, "hintpath", hint_seq, pos_seq)                 # :186
    ref_logit = F.linear(dec, sd["trg_word_prj.weight"]).transpose(1, 2).reshape(N, N_VOCAB, h, w)  # :187-189
    full = _q(upfeat(dec.transpose(1, 2).reshape(N, 64, h, w), affinity, sp))  # :194-195
    pred = torch.tanh(hourglass2(sd, torch.cat((gray, full), 1)))          # :196-197
    if stages is not None:
        stages.update(feats=feats, tokens=tokens, sizes=sizes, enc=enc, labels=labels, dec=dec, full=full)
    return pal_logit, ref_logit, pred, affinity, sampled, hint_mask


# ----------------------------------------------------------------------------------------------
# training-side losses on the forward's outputs (config 5; models/loss.py:12-87, models/basic.py:120-134,153-175)
# ----------------------------------------------------------------------------------------------
def class_weights(lambda_=0.5):
    """ColorLabel.weights (models/basic.py:153-157) from the gamut prior (utils/gamut_probs.npy)."""
    from disentangledcolorization_b200.cielab import gamut_prior  # data table only
    prior = torch.from_numpy(gamut_prior().astype(np.float32))      # ABGamut.DTYPE = float32 (utils/cielab.py:8-11)
    uniform = torch.zeros_like(prior)
    uniform[prior > 0] = 1 / (prior > 0).sum().type_as(uniform)
    w = 1 / ((1 - lambda_) * prior + lambda_ * uniform)
    return w / torch.sum(prior * w)


class _Rebalance(torch.autograd.Function):
    """basic.RebalanceLoss (models/basic.py:120-134): identity forward, gradient multiplied by the weights."""

    @staticmethod
    def forward(ctx, x, w):
        ctx.save_for_backward(w)
        return x.clone()

    @staticmethod
    def backward(ctx, g):
        (w,) = ctx.saved_tensors
        return g * w, None


def anchor_color_prob_loss(pal_logit, ref_logit, target_label, class_weight):
    """AnchorColorProbLoss.__call__ with hint2regress=False, enhanced=False (models/loss.py:59-87): returns the dict."""
    N, C, H, W = target_label.shape
    labels = target_label.permute(0, 2, 3, 1).contiguous().view(N * H * W)
    out = {}
    for key, logit in (("palLoss", pal_logit), ("refLoss", ref_logit)):
        probs = _Rebalance.apply(logit, class_weight).permute(0, 2, 3, 1).contiguous().view(N * H * W, -1)
        out[key] = F.cross_entropy(probs, labels, ignore_index=-1)
    out["recLoss"] = torch.zeros_like(out["palLoss"])
    out["totalLoss"] = out["palLoss"] + out["refLoss"] + out["recLoss"]
    return out


def spixel_loss(prob, feat, k=16):
    """SPixelLoss.__call__ (models/loss.py:17-30)."""
    pooled, _ = poolfeat(feat, prob, k)
    recon = upfeat(pooled, prob, k)
    d = recon - feat
    f = torch.norm(d[:, :-2], p=2, dim=1).mean()
    p_ = torch.norm(d[:, -2:], p=2, dim=1).mean() / k
    return {"totalLoss": 10 * f + 0.003 * p_, "featLoss": f, "posLoss": p_}


# ---- perceptual term of AnchorColorProbLoss (config 5): models/basic.py:422-475, models/loss.py:45-49,138-223 ----------
def lab2rgb(lab_rs):
    """basic.lab2rgb: normalised Lab (N,3,H,W) -> RGB [0,1] (lab2xyz, models/basic.py:439-454; xyz2rgb, :409-422)."""
    L = lab_rs[:, 0] * 50.0 + 50.0
    a, b = lab_rs[:, 1] * 110.0, lab_rs[:, 2] * 110.0
    y = (L + 16.0) / 116.0
    x = a / 500.0 + y
    z = torch.clamp(y - b / 200.0, min=0.0)
    out = torch.stack((x, y, z), dim=1)
    mask = (out > 0.2068966).float()
    out = (out ** 3.0) * mask + (out - 16.0 / 116.0) / 7.787 * (1 - mask)
    out = out * torch.tensor((0.95047, 1.0, 1.08883), dtype=out.dtype, device=out.device)[None, :, None, None]
    r = 3.24048134 * out[:, 0] - 1.53715152 * out[:, 1] - 0.49853633 * out[:, 2]
    g = -0.96925495 * out[:, 0] + 1.87599 * out[:, 1] + 0.04155593 * out[:, 2]
    bl = 0.05564664 * out[:, 0] - 0.20404134 * out[:, 1] + 1.05731107 * out[:, 2]
    rgb = torch.clamp(torch.stack((r, g, bl), dim=1), min=0.0)
    mask = (rgb > 0.0031308).float()
    return (1.055 * (rgb ** (1.0 / 2.4)) - 0.055) * mask + 12.92 * rgb * (1 - mask)


VGG19_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512, "M"]
VGG_SLICES = {"liu": [2, 7, 12, 21, 30], "lei": [4, 9, 14, 23, 32]}                    # models/loss.py:160-172: slice ends
VGG_WEIGHTS = {"liu": [1.0 / 32, 1.0 / 16, 1.0 / 8, 1.0 / 4, 1.0], "lei": [1.0 / 2.6, 1.0 / 4.8, 1.0 / 3.7, 1.0 / 5.6, 10.0 / 1.5]}


def vgg19_loss(convs, x, y, feat_type="liu"):
    """VGG19Loss.forward (models/loss.py:205-223).  convs: [(weight, bias)] of torchvision vgg19.features in order;
    x ground truth, y prediction, RGB (N,3,H,W) in [0,1]."""
    mean = torch.tensor([0.485, 0.456, 0.406], dtype=x.dtype, device=x.device)[None, :, None, None]
    std = torch.tensor([0.229, 0.224, 0.225], dtype=x.dtype, device=x.device)[None, :, None, None]
    z = torch.cat(((x - mean) / std, (y - mean) / std), 0)
    N = x.shape[0]
    ends = VGG_SLICES.get(feat_type, [28])
    wts = VGG_WEIGHTS.get(feat_type, [1.0])
    loss, idx, ci, k = 0.0, 0, 0, 0
    for v in VGG19_CFG:
        if v == "M":
            z = F.max_pool2d(z, 2, 2)
            idx += 1
        else:
            w, b = convs[ci]
            ci += 1
            z = F.relu(F.conv2d(z, w, b, padding=1))
            idx += 2
        if k < len(ends) and idx == ends[k]:
            loss = loss + wts[k] * (z[:N] - z[N:]).abs().mean()
            k += 1
            if k == len(ends):
                break
    return loss


def perceptual_loss(convs, gray, colors_x, colors_y, feat_type="liu"):
    """AnchorColorProbLoss._perceptual_loss (models/loss.py:45-49)."""
    return vgg19_loss(convs, lab2rgb(torch.cat([gray, colors_x], 1)), lab2rgb(torch.cat([gray, colors_y], 1)), feat_type)


def laplace_gradient(pred_ab, target_ab):
    """AnchorColorProbLoss._laplace_gradient (models/loss.py:51-57)."""
    Cc = pred_ab.shape[1]
    kernel = torch.tensor([[1, 1, 1], [1, -8, 1], [1, 1, 1]], dtype=pred_ab.dtype, device=pred_ab.device).view(1, 1, 3, 3).repeat(Cc, 1, 1, 1)
    return F.l1_loss(F.conv2d(target_ab, kernel, groups=Cc), F.conv2d(pred_ab, kernel, groups=Cc))
