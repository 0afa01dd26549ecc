"""Drop-in `basic` module: the free functions and ColorLabel that callers of the forward path use
directly (reference main/colorizer/inference.py:114-115,119,129; models/basic.py:10-12,149-218,
274-376).  NCHW fp32 tensors in and out, CUDA only; arithmetic runs in libdisco_b200.so.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from .cielab import Q_TO_AB, class_weights


def tensor2array(tensors):
    """(N,C,H,W) tensor -> (N,H,W,C) numpy (models/basic.py:10-12).  Host I/O helper."""
    return np.transpose(tensors.detach().to("cpu").numpy(), (0, 2, 3, 1))


def _ctx(t):
    if t.device.type != "cuda":
        raise _lib.DiscoError("disentangledcolorization_b200.basic runs on a CUDA (B200) device only")
    handle = _lib.Handle.get(t.device.index if t.device.index is not None else torch.cuda.current_device())
    return handle, C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _pool(feat, prob, sp_h, sp_w):
    if sp_h != 16 or sp_w != 16:
        raise _lib.DiscoError("only 16x16 super-pixels are built")
    B, Cc, H, W = feat.shape
    if Cc > 66:
        raise _lib.DiscoError("poolfeat: at most 64 feature + 2 colour channels")
    handle, stream = _ctx(feat)
    dev = feat.device
    h, w = H // 16, W // 16
    f32 = dict(dtype=torch.float32, device=dev)
    # split into the (<=64 feature, 2 colour) layout of the fused kernel; pad features with zeros
    nf = min(Cc, 64)
    feats = torch.zeros(B, H, W, 64, **f32)
    feats[..., :nf] = feat[:, :nf].permute(0, 2, 3, 1)
    ab = None
    if Cc > 64:
        ab = torch.zeros(B, 2, H, W, **f32)
        ab[:, :Cc - 64] = feat[:, 64:]
    partial = torch.empty(B, h, w, 9, 68, **f32)
    tokens, spix = torch.empty(B, h * w, 64, **f32), torch.empty(B, 2, h, w, **f32)
    conf, sizes = torch.empty(B, h * w, **f32), torch.empty(B, h * w, **f32)
    _lib.check(handle.lib.disco_poolfeat(handle.h, _lib.F32, _p(feats), _p(ab), _p(prob.float().contiguous()), B, H, W,
                                         64, _p(partial), _p(tokens), _p(spix), _p(conf), _p(sizes), stream),
               "disco_poolfeat")
    pooled = tokens.view(B, h, w, 64).permute(0, 3, 1, 2)[:, :nf]
    if Cc > 64:
        pooled = torch.cat([pooled, spix[:, :Cc - 64]], 1)
    return pooled.contiguous(), conf.view(B, 1, h, w), sizes.view(B, 1, h, w)


def poolfeat(input, prob, sp_h=2, sp_w=2, need_entry_prob=False):
    """models/basic.py:274-324."""
    pooled, conf, _ = _pool(input.float(), prob, sp_h, sp_w)
    return (pooled, conf) if need_entry_prob else pooled


def get_spixel_size(affinity_map, sp_h=2, sp_w=2, elem_thres=25):
    """models/basic.py:327-335."""
    ones = torch.ones(affinity_map.shape[0], 1, *affinity_map.shape[2:], device=affinity_map.device)
    return _pool(ones, affinity_map, sp_h, sp_w)[2]


def upfeat(input, prob, up_h=2, up_w=2):
    """models/basic.py:338-376: (B,C,h,w) tokens + (B,9,H,W) affinity -> (B,C,H,W), C <= 64."""
    if up_h != 16 or up_w != 16:
        raise _lib.DiscoError("only 16x16 super-pixels are built")
    B, Cc, h, w = input.shape
    if Cc > 64:
        raise _lib.DiscoError("upfeat: at most 64 channels")
    handle, stream = _ctx(input)
    H, W = h * 16, w * 16
    tok = torch.zeros(B, h * w, 64, dtype=torch.float32, device=input.device)
    tok[..., :Cc] = input.float().flatten(2).transpose(1, 2)
    out = torch.empty(B, H, W, 64, dtype=torch.float32, device=input.device)
    _lib.check(handle.lib.disco_upfeat(handle.h, _lib.F32, _p(tok), _p(prob.float().contiguous()), B, H, W, 64, _p(out),
                                       stream), "disco_upfeat")
    return out[..., :Cc].permute(0, 3, 1, 2).contiguous()


class ColorLabel:
    """models/basic.py:149-218 (the parts on the inference path)."""

    def __init__(self, lambda_=0.5, device="cuda"):
        self.q_to_ab = torch.from_numpy(Q_TO_AB.copy()).to(device)
        self.weights = torch.from_numpy(class_weights(lambda_)).to(device)        # float32, like the reference (:153-157)

    def get_classweights(self, batch_gt_indx):
        """models/basic.py:173-175."""
        return self.weights.to(batch_gt_indx.device)[batch_gt_indx]

    def encode_ab2ind(self, batch_ab, neighbours=5, sigma=5.0):
        """models/basic.py:177-194: soft 313-way code (5 nearest bins, Gaussian weights); (N,2,h,w) -> (N,313,h,w)."""
        if neighbours != 5 or sigma != 5.0:
            raise _lib.DiscoError("encode_ab2ind is built for neighbours=5, sigma=5.0 (the only values the reference uses)")
        B, _, h, w = batch_ab.shape
        handle, stream = _ctx(batch_ab)
        q = torch.empty(B, 313, h, w, dtype=torch.float32, device=batch_ab.device)
        _lib.check(handle.lib.disco_encode_ab2ind(handle.h, _p(batch_ab.float().contiguous()), _p(self.q_to_ab.to(batch_ab.device)),
                                                  B, h * w, _p(q), stream), "disco_encode_ab2ind")
        return q

    def decode_ind2ab(self, batch_q, T=0.38):
        """T-th most probable bin -> ab/110 (integer T) or annealed mean (fractional T); host-side glue on
        (N,313,h,w) logits whose result the reference CLI discards (inference.py:114-115)."""
        q = torch.softmax(batch_q, dim=1)
        table = self.q_to_ab.to(q.device)
        if T % 1 == 0:
            idx = torch.sort(q, dim=1, descending=True)[1][:, int(T)]
            ab = table[idx].permute(0, 3, 1, 2)
        else:
            q = torch.exp(q / T)
            q = q / q.sum(dim=1, keepdim=True)
            ab = torch.einsum("nqhw,qc->nchw", q, table)
        return (ab / 110.0).type(batch_q.dtype)

    def encode_ab2ind_hard(self, batch_ab):
        """argmax of encode_ab2ind (models/basic.py:177-194 + models/model.py:120,166) == nearest bin."""
        B, _, h, w = batch_ab.shape
        handle, stream = _ctx(batch_ab)
        labels = torch.empty(B * h * w, dtype=torch.int32, device=batch_ab.device)
        _lib.check(handle.lib.disco_token_labels(handle.h, 1, _p(batch_ab.float().contiguous()),
                                                 _p(self.q_to_ab.to(batch_ab.device)), B, h * w, _p(labels), None,
                                                 stream), "disco_token_labels")
        return labels.view(B, 1, h, w).long()


def lab2rgb(lab_rs, l_mean=50, l_norm=50, ab_norm=110):
    """reference models/basic.py:468-475 (-> lab2xyz -> xyz2rgb, :422-466): normalised Lab (N,3,H,W) in [-1,1] -> RGB in
    [0,1], fp32 NCHW, one kernel (disco_lab2rgb_norm).  Only the default normalisation constants are built."""
    if (l_mean, l_norm, ab_norm) != (50, 50, 110):
        raise _lib.DiscoError("lab2rgb: only l_mean=50, l_norm=50, ab_norm=110 are built")
    if lab_rs.dim() != 4 or lab_rs.shape[1] != 3:
        raise _lib.DiscoError(f"lab2rgb: input must be (N,3,H,W), got {tuple(lab_rs.shape)}")
    handle, stream = _ctx(lab_rs)
    with torch.cuda.device(lab_rs.device):
        lab = lab_rs.detach().float()
        gray, ab = lab[:, 0:1].contiguous(), lab[:, 1:3].contiguous()
        N, _, H, W = lab.shape
        out = torch.empty(N, 3, H, W, dtype=torch.float32, device=lab.device)
        _lib.check(handle.lib.disco_lab2rgb_norm(handle.h, _p(gray), _p(ab), N, H, W, _p(out), None, _lib.F32, 3, None, None, stream),
                   "disco_lab2rgb_norm")
    return out


def init_spixel_grid(img_height, img_width, spixel_size=16):
    """reference models/basic.py:221-262: (9,H,W) float map of the ids of the 9 neighbour cells of every pixel's cell (edge
    replicated) and the (2,H,W) pixel coordinate map (x first).  Host-side numpy, as in the reference."""
    n_h, n_w = int(np.floor(img_height / spixel_size)), int(np.floor(img_width / spixel_size))
    sp_h, sp_w = int(img_height / (1.0 * n_h)), int(img_width / (1.0 * n_w))
    cells = np.pad(np.int32(np.arange(0, n_w * n_h).reshape((n_h, n_w))), ((1, 1), (1, 1)), mode="edge")
    shifts = [cells[dy:dy + n_h, dx:dx + n_w] for dy in range(3) for dx in range(3)]     # top-left ... bottom-right
    grid = np.repeat(np.repeat(np.stack(shifts, 0), sp_h, axis=1), sp_w, axis=2)
    yy, xx = np.meshgrid(np.arange(0, img_height, 1), np.arange(0, img_width, 1), indexing="ij")
    return torch.from_numpy(grid).float(), torch.from_numpy(np.stack([xx, yy], 0)).float()


def split_spixels(assign_map, sp_size=16):
    """`split_spixels` of the reference's SpixelSeg inference script (main/spixelseg/inference.py:67-75): winner-take-all
    super-pixel id map (N,1,H,W) int32 from the (N,9,H,W) assignment map, on the init_spixel_grid ids (disco_spixel_ids)."""
    if assign_map.dim() != 4 or assign_map.shape[1] != 9:
        raise _lib.DiscoError(f"split_spixels: assignment map must be (N,9,H,W), got {tuple(assign_map.shape)}")
    handle, stream = _ctx(assign_map)
    with torch.cuda.device(assign_map.device):
        prob = assign_map.detach().float().contiguous()
        N, _, H, W = prob.shape
        ids = torch.empty(N, 1, H, W, dtype=torch.int32, device=prob.device)
        _lib.check(handle.lib.disco_spixel_ids(handle.h, _p(prob), N, H, W, int(sp_size), _p(ids), stream), "disco_spixel_ids")
    return ids
