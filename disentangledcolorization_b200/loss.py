"""Drop-in `loss` module, first slice (BASELINE config 5): the loss terms the trainers
compute on the hot path's outputs, with the reference's class names, constructor arguments, `__call__(data, epoch_no)`
contract and result keys (reference models/loss.py:12-87).

Built: `AnchorColorProbLoss` token terms (palLoss, refLoss: cross-entropy with gradient re-balancing) as one CUDA kernel each
(forward value AND d loss / d logits; `loss.backward()` delivers the re-balanced gradient to `pal_prob` / `ref_prob` when
they require grad), its perceptual term (`enhanced=True`: `VGG19Loss` = Lab -> RGB, normalisation, the VGG19 feature stack
through disco_conv, 2x2 max pooling and the five weighted L1 means, all on the device; a VALUE only -- the reference feeds
it so that no gradient reaches the model), the Laplacian term (`with_grad=True`: value and d loss / d pred_color),
`SPixelLoss` (forward value).  NOT built: `hint2regress` (raises) and the backward of the conv / transformer kernels.
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


class _LaplaceL1(torch.autograd.Function):
    """AnchorColorProbLoss._laplace_gradient (loss.py:51-57): value and d loss / d pred from disco_laplace_l1."""

    @staticmethod
    def forward(ctx, pred, target):
        if pred.dim() != 4 or tuple(pred.shape) != tuple(target.shape):
            raise _lib.DiscoError(f"_laplace_gradient: pred and target must both be (N,C,H,W), got {tuple(pred.shape)} and {tuple(target.shape)}")
        handle, stream = _ctx(pred)
        with torch.cuda.device(pred.device):
            p, t = pred.detach().float().contiguous(), target.detach().float().contiguous()
            N, Cc, H, W = p.shape
            need = pred.requires_grad
            sign = torch.empty(N * Cc * (H - 2) * (W - 2), dtype=torch.float32, device=p.device) if need else None
            grad = torch.empty_like(p) if need else None
            partial = torch.empty(1184, dtype=torch.float32, device=p.device)
            out = torch.empty(1, dtype=torch.float32, device=p.device)
            _lib.check(handle.lib.disco_laplace_l1(handle.h, _p(p), _p(t), N, Cc, H, W, _p(sign), _p(partial), partial.numel(), _p(out),
                                                   _p(grad), stream), "disco_laplace_l1")
        ctx.grad = grad
        return out[0].clone()

    @staticmethod
    def backward(ctx, grad_out):
        return (ctx.grad * grad_out if ctx.grad is not None else None), None


class AnchorColorProbLoss:
    """reference models/loss.py:33-87."""

    def __init__(self, hint2regress=False, enhanced=False, with_grad=False, mpdist=False, gpu_no=0, vgg_loss=None):
        """`vgg_loss` [extension]: a ready `VGG19Loss` (e.g. built from a local torchvision model); by default the
        constructor builds one exactly as the reference does (loss.py:42-43: downloads the pretrained VGG19)."""
        if hint2regress:
            raise _lib.DiscoError("AnchorColorProbLoss: hint2regress is not built (the reference's own training branch for it is "
                                  "broken, SURVEY 8b); the cross-entropy, perceptual and Laplacian terms are (SURVEY 8a a18)")
        self.mpdist, self.gpu_no = mpdist, gpu_no
        self.hint2regress, self.enhanced, self.with_grad = hint2regress, enhanced, with_grad
        if self.enhanced:
            self.VGGLoss = vgg_loss if vgg_loss is not None else VGG19Loss(gpu_no=gpu_no, is_ddp=mpdist)

    def _perceptual_loss(self, input_grays, input_colors, pred_colors):
        """reference loss.py:45-49: VGGLoss(lab2rgb(gray, input_colors), lab2rgb(gray, pred_colors)); Lab -> RGB and the VGG
        normalisation are one kernel per image set (no RGB tensor in HBM)."""
        return self.VGGLoss.from_lab(input_grays, input_colors, pred_colors)

    def _laplace_gradient(self, pred_AB, target_AB):
        """reference loss.py:51-57."""
        return _LaplaceL1.apply(pred_AB, target_AB)

    def __call__(self, data, epoch_no):
        pal = _CERebalance.apply(data["pal_prob"], data["target_label"], data["class_weight"])       # loss.py:61-70
        ref = _CERebalance.apply(data["ref_prob"], data["target_label"], data["class_weight"])       # loss.py:75-77
        rec = torch.zeros_like(pal)                                                                    # loss.py:78
        if self.enhanced:                                                                              # loss.py:79-81 (argument order as there)
            rec = 5.0 * self._perceptual_loss(data["input_gray"], data["pred_color"], data["input_color"])
            if self.with_grad:                                                                         # loss.py:82-84
                rec = rec + self._laplace_gradient(data["pred_color"], data["input_color"])
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


# torchvision vgg19 `features` (cfg E): index of every convolution and its (cin, cout); 'M' = MaxPool2d(2, 2)
_VGG19_CFG = [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512, "M"]


def _vgg19_layers():
    """[(features index, 'conv', cin, cout) | (index, 'relu') | (index, 'pool')] as torchvision builds them."""
    out, cin, i = [], 3, 0
    for v in _VGG19_CFG:
        if v == "M":
            out.append((i, "pool"))
            i += 1
        else:
            out.append((i, "conv", cin, v))
            out.append((i + 1, "relu"))
            cin = v
            i += 2
    return out


class VGG19Loss(torch.nn.Module):
    """reference models/loss.py:138-223: weighted L1 distances between VGG19 features of two RGB images.

    Same constructor arguments, `forward(x, y)` contract (x, y: (N,3,H,W) RGB in [0,1]) and `state_dict` keys
    (`slice<k>.<i>.weight|bias`, or `featureExactor.<i>...`) as the reference.  Extensions: `vgg_model` (a torchvision vgg19
    to take the weights from instead of downloading the pretrained one), `precision` ('bf16' tensor-core path | 'fp32').
    The result is a value: no autograd graph is built (the reference detaches x and, as AnchorColorProbLoss calls it, y does
    not depend on the model either)."""
    _SLICES = {"liu": [(0, 2), (2, 7), (7, 12), (12, 21), (21, 30)], "lei": [(0, 4), (4, 9), (9, 14), (14, 23), (23, 32)]}
    _WEIGHTS = {"liu": [1.0 / 32, 1.0 / 16, 1.0 / 8, 1.0 / 4, 1.0], "lei": [1.0 / 2.6, 1.0 / 4.8, 1.0 / 3.7, 1.0 / 5.6, 10.0 / 1.5]}
    CPAD = 16                                     # the RGB input is padded to 16 channels (one tensor-core K step)

    def __init__(self, feat_type="liu", gpu_no=0, is_ddp=False, requires_grad=False, vgg_model=None, precision="bf16"):
        super().__init__()
        from .network import ParamTree
        self.mean = [0.485, 0.456, 0.406]
        self.std = [0.229, 0.224, 0.225]
        self.feat_type, self.precision = feat_type, precision
        layers = _vgg19_layers()
        if feat_type in self._SLICES:
            self.weights = self._WEIGHTS[feat_type]
            groups = [(f"slice{k + 1}", a, b) for k, (a, b) in enumerate(self._SLICES[feat_type])]
        else:
            self.weights = [1.0]
            groups = [("featureExactor", 0, 28)]
        # execution plan: ('conv', key, cin, cout) | ('pool',) | ('tap', weight); a ReLU always follows a conv (fused)
        self._plan, entries = [], []
        for name, a, b in groups:
            for lay in layers:
                if not a <= lay[0] < b:
                    continue
                if lay[1] == "conv":
                    key = f"{name}.{lay[0] - a}"
                    entries += [(key + ".weight", (lay[3], lay[2], 3, 3), True), (key + ".bias", (lay[3],), True)]
                    self._plan.append(("conv", key, lay[2], lay[3]))
                elif lay[1] == "pool":
                    self._plan.append(("pool",))
            self._plan.append(("tap", self.weights[len([p for p in self._plan if p[0] == "tap"])]))
        tree = ParamTree(entries)
        for name, mod in tree.named_children():
            self.add_module(name, mod)
        if vgg_model is None:
            import torchvision
            vgg_model = torchvision.models.vgg19(pretrained=True)          # as the reference (loss.py:147): needs the download
        feats = list(vgg_model.features)
        with torch.no_grad():
            own = dict(self.named_parameters())
            for name, a, b in groups:
                for i in range(a, b):
                    if isinstance(feats[i], torch.nn.Conv2d):
                        own[f"{name}.{i - a}.weight"].copy_(feats[i].weight)
                        own[f"{name}.{i - a}.bias"].copy_(feats[i].bias)
        if not requires_grad:
            for prm in self.parameters():
                prm.requires_grad = False
        self.eval()
        self._packed = None
        self.register_load_state_dict_post_hook(lambda m, keys: setattr(m, "_packed", None))
        if torch.cuda.is_available():
            self.cuda(gpu_no)
        print("[*] VGG19Loss init!")

    def _apply(self, fn, *a, **k):
        self._packed = None
        return super()._apply(fn, *a, **k)

    def _convs(self, dev):
        from . import netspec
        from .engine import _PackedConv
        key = (str(dev), self.precision)
        if self._packed is None or self._packed[0] != key:
            sd = {k: v.detach().cpu() for k, v in self.state_dict().items()}
            packed = {}
            for step in self._plan:
                if step[0] != "conv":
                    continue
                _, wkey, cin, cout = step
                cin_p = max(cin, self.CPAD)
                op = netspec.ConvOp(wkey, [netspec.Src("x", wkey, (0, cin))], cout, "y", bias=True, act="relu")
                f = netspec.fold(sd, op)
                if cin_p != cin:                                            # zero weights for the padding channels of the RGB input
                    f.weights = [torch.nn.functional.pad(f.weights[0], (0, 0, 0, 0, 0, cin_p - cin)).contiguous()]
                packed[wkey] = _PackedConv(f, dev)
            self._packed = (key, packed)
        return self._packed[1]

    def _features_loss(self, z, handle, stream):
        """z: (2N, H, W, C) NHWC activations, ground truth first, prediction second -> scalar loss (device tensor)."""
        from .engine import _DT
        from .network import _ConvNet
        dev = z.device
        code, tdt = _DT[self.precision]
        convs = self._convs(dev)
        lib, hd = handle.lib, handle.h
        out = torch.zeros(1, dtype=torch.float32, device=dev)
        partial = torch.empty(1184, dtype=torch.float32, device=dev)
        n_tap = 0
        for step in self._plan:
            B2, H, W, Cc = z.shape
            if step[0] == "conv":
                pc = convs[step[1]]
                y = torch.empty(B2, H, W, step[3], dtype=tdt, device=dev)
                d = _lib.ConvDesc()
                d.kind, d.stride, d.dtype, d.batch, d.Ho, d.Wo, d.Cout, d.n_src = _lib.CONV3, 1, code, B2, H, W, step[3], 1
                d.src[0].ptr, d.src[0].H, d.src[0].W, d.src[0].C = z.data_ptr(), H, W, Cc
                d.src[0].up2, d.src[0].is_f32, d.src[0].w_off = 0, 0, 0
                d.weights, d.bias = pc.w32.data_ptr(), pc.bias.data_ptr()
                d.bias_host = pc.bias_host.data_ptr()
                d.act, d.slope, d.head, d.out = _lib.ACT_RELU, 0.0, _lib.HEAD_NONE, y.data_ptr()
                _ConvNet._route(handle, d, pc, dev)
                _lib.check(lib.disco_conv(hd, C.byref(d), stream), "disco_conv")
                z = y
            elif step[0] == "pool":
                if H % 2 or W % 2:
                    raise _lib.DiscoError(f"VGG19Loss: feature map {H}x{W} is not even at a pooling layer; use H, W multiples of 16")
                y = torch.empty(B2, H // 2, W // 2, Cc, dtype=tdt, device=dev)
                _lib.check(lib.disco_maxpool2(hd, code, _p(z), B2, H, W, Cc, _p(y), stream), "disco_maxpool2")
                z = y
            else:
                half = z.numel() // 2
                zy = z.view(-1)[half:]
                _lib.check(lib.disco_l1_mean(hd, code, _p(z), _p(zy), half, float(step[1]), 1 if n_tap else 0, _p(partial),
                                             partial.numel(), _p(out), stream), "disco_l1_mean")
                n_tap += 1
        return out[0].clone()

    def _mean_std(self):
        import numpy as np
        m, s = np.asarray(self.mean, np.float32), np.asarray(self.std, np.float32)
        return m, s

    def from_lab(self, grays, colors_x, colors_y):
        """VGGLoss(lab2rgb(cat(grays, colors_x)), lab2rgb(cat(grays, colors_y))) without materialising the RGB images."""
        from .engine import _DT
        handle, stream = _ctx(grays)
        code, tdt = _DT[self.precision]
        with torch.cuda.device(grays.device), torch.no_grad():
            g = grays.detach().float().contiguous()
            N, _, H, W = g.shape
            z = torch.empty(2 * N, H, W, self.CPAD, dtype=tdt, device=g.device)
            m, s = self._mean_std()
            for i, ab in enumerate((colors_x, colors_y)):
                ab = ab.detach().float().contiguous()
                if tuple(ab.shape) != (N, 2, H, W):
                    raise _lib.DiscoError(f"VGG19Loss.from_lab: colours must be ({N},2,{H},{W}), got {tuple(ab.shape)}")
                _lib.check(handle.lib.disco_lab2rgb_norm(handle.h, _p(g), _p(ab), N, H, W, None, _p(z[i * N:]), code, self.CPAD,
                                                         C.c_void_p(m.ctypes.data), C.c_void_p(s.ctypes.data), stream), "disco_lab2rgb_norm")
            return self._features_loss(z, handle, stream)

    def forward(self, x, y):
        """x: ground truth, y: prediction, (N,3,H,W) RGB in [0,1] (reference loss.py:205-223)."""
        from .engine import _DT
        handle, stream = _ctx(x)
        code, tdt = _DT[self.precision]
        with torch.cuda.device(x.device), torch.no_grad():
            N, Cc, H, W = x.shape
            if Cc != 3 or tuple(y.shape) != tuple(x.shape):
                raise _lib.DiscoError(f"VGG19Loss: x and y must both be (N,3,H,W), got {tuple(x.shape)} and {tuple(y.shape)}")
            z = torch.empty(2 * N, H, W, self.CPAD, dtype=tdt, device=x.device)
            m, s = self._mean_std()
            for i, img in enumerate((x, y)):
                img = img.detach().float().contiguous()
                _lib.check(handle.lib.disco_rgb_norm(handle.h, _p(img), N, H, W, _p(z[i * N:]), code, self.CPAD,
                                                     C.c_void_p(m.ctypes.data), C.c_void_p(s.ctypes.data), stream), "disco_rgb_norm")
            return self._features_loss(z, handle, stream)
