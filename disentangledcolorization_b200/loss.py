"""Drop-in `loss` module, first slice (BASELINE config 5): the loss terms the trainers
compute on the hot path's outputs, with the reference's class names, constructor arguments, `__call__(data, epoch_no)`
contract and result keys (reference models/loss.py:12-87).

Built: `AnchorColorProbLoss` token terms (palLoss, refLoss: cross-entropy with gradient re-balancing) as one CUDA kernel each
(forward value AND d loss / d logits; `loss.backward()` delivers the re-balanced gradient to `pal_prob` / `ref_prob` when
they require grad), `SPixelLoss` (forward value).  NOT built: the perceptual term (`enhanced=True`: VGG19 with downloaded
weights, which the reference feeds with a detached input so that it contributes no gradient -- SURVEY 3.3), `hint2regress`,
`with_grad`, and the backward of the conv / transformer kernels; they raise.
"""
import ctypes as C

import torch

from . import _lib


def _ctx(t):
    if t.device.type != "cuda":
        raise _lib.DiscoError("disentangledcolorization_b200.loss runs on a CUDA (B200) device only")
    handle = _lib.Handle.get(t.device.index if t.device.index is not None else torch.cuda.current_device())
    return handle, C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class _CERebalance(torch.autograd.Function):
    """loss = CrossEntropy(RebalanceLoss(logits, weights), labels): value and gradient from disco_ce_rebalance."""

    @staticmethod
    def forward(ctx, logits, labels, weights):
        N, V, h, w = logits.shape
        if V != 313:
            raise _lib.DiscoError("AnchorColorProbLoss: logits must have 313 classes")
        handle, stream = _ctx(logits)
        with torch.cuda.device(logits.device):
            lg = logits.detach().float().contiguous()
            lab = labels.reshape(-1).to(torch.int32).contiguous()
            tw = weights.reshape(-1).float().contiguous()
            if lab.numel() != N * h * w or tw.numel() != N * h * w:
                raise _lib.DiscoError("AnchorColorProbLoss: target_label / class_weight must be (N,1,h,w)")
            tok = torch.empty(N * h * w, dtype=torch.float32, device=lg.device)
            out = torch.empty(2, dtype=torch.float32, device=lg.device)
            grad = torch.empty_like(lg) if logits.requires_grad else None
            _lib.check(handle.lib.disco_ce_rebalance(handle.h, _p(lg), _p(lab), _p(tw), N, h * w, _p(tok), _p(out), _p(grad), stream),
                       "disco_ce_rebalance")
        ctx.grad = grad
        return out[0].clone()

    @staticmethod
    def backward(ctx, grad_out):
        return (ctx.grad * grad_out if ctx.grad is not None else None), None, None


class AnchorColorProbLoss:
    """reference models/loss.py:33-87."""

    def __init__(self, hint2regress=False, enhanced=False, with_grad=False, mpdist=False, gpu_no=0):
        if hint2regress or enhanced or with_grad:
            raise _lib.DiscoError("AnchorColorProbLoss: hint2regress / enhanced (VGG19 perceptual term) / with_grad are not built; "
                                  "the token-level cross-entropy terms are (SURVEY 8f N3, first slice)")
        self.mpdist, self.gpu_no = mpdist, gpu_no
        self.hint2regress, self.enhanced, self.with_grad = hint2regress, enhanced, with_grad

    def __call__(self, data, epoch_no):
        pal = _CERebalance.apply(data["pal_prob"], data["target_label"], data["class_weight"])       # loss.py:61-70
        ref = _CERebalance.apply(data["ref_prob"], data["target_label"], data["class_weight"])       # loss.py:75-77
        rec = torch.zeros_like(pal)                                                                    # loss.py:78
        return {"totalLoss": pal + ref + rec, "palLoss": pal, "refLoss": ref, "recLoss": rec}


class SPixelLoss:
    """reference models/loss.py:12-30 (forward value; only 16 x 16 super-pixels are built)."""

    def __init__(self, psize=8, mpdist=False, gpu_no=0):
        self.mpdist, self.gpu_no, self.sp_size = mpdist, gpu_no, psize

    def __call__(self, data, epoch_no):
        from . import basic
        k = self.sp_size
        prob, feat = data["pred_prob"], data["target_feat"]
        N, Cc, H, W = feat.shape
        pooled = basic.poolfeat(feat, prob, k, k)
        recon = basic.upfeat(pooled, prob, k, k)
        handle, stream = _ctx(feat)
        with torch.cuda.device(feat.device):
            tgt = feat.float().contiguous()
            nb = 4 * 148
            partial = torch.empty(nb, 2, dtype=torch.float32, device=feat.device)
            out = torch.empty(2, dtype=torch.float32, device=feat.device)
            _lib.check(handle.lib.disco_spixel_recon_loss(handle.h, _p(recon), _p(tgt), N, Cc, H, W, _p(partial), nb, _p(out), stream),
                       "disco_spixel_recon_loss")
        feat_loss, pos_loss = out[0].clone(), out[1] / k
        return {"totalLoss": 10 * feat_loss + 0.003 * pos_loss, "featLoss": feat_loss, "posLoss": pos_loss}
