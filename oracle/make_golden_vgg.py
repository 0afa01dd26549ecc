#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- fixture of the perceptual loss term, generated with the UNMODIFIED reference
(models/loss.py VGG19Loss / AnchorColorProbLoss._perceptual_loss and models/basic.py lab2rgb imported from /root/reference
through oracle/ref_harness.py, CPU).

    python oracle/make_golden_vgg.py        # writes tests/golden/vgg_loss.npz

The reference downloads torchvision's pretrained VGG19; there is no network here, so `torchvision.models.vgg19` is replaced
by a wrapper that returns the RANDOM-INIT torchvision model built under torch.manual_seed(VGG_SEED) -- the reference code
that slices it, normalises the images and forms the weighted L1 sums runs unmodified.  Tests rebuild the same weights from
the same seed (the fixture stores a checksum of them), so only inputs and results are committed.  One more CPU shim for
the Laplacian term: `Tensor.get_device()` is -1 on CPU, which `torch.tensor(..., device=-1)` (loss.py:53) rejects.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_harness  # noqa: E402

VGG_SEED = 7


_TV_VGG19 = None


def seeded_vgg19():
    """torchvision's own (un-patched) vgg19 constructor, random init under VGG_SEED"""
    global _TV_VGG19
    if _TV_VGG19 is None:
        import torchvision
        _TV_VGG19 = torchvision.models.vgg19
    torch.manual_seed(VGG_SEED)
    return _TV_VGG19(weights=None).eval()


def main():
    ref_harness.load()
    import torchvision
    cwd = os.getcwd()
    os.chdir(os.path.join(ref_harness.REF, "main", "colorizer"))
    seeded_vgg19()                                   # binds torchvision's constructor before it is patched
    orig = torchvision.models.vgg19
    try:
        import basic as ref_basic
        import loss as ref_loss
        torchvision.models.vgg19 = lambda *a, **k: seeded_vgg19()
        g = torch.Generator().manual_seed(4321)
        N, H, W = 2, 64, 96
        gray = torch.rand(N, 1, H, W, generator=g) * 2 - 1
        ab_x = (torch.rand(N, 2, H, W, generator=g) - 0.5) * 0.9
        ab_y = (ab_x + 0.25 * torch.randn(N, 2, H, W, generator=g)).clamp(-1, 1)
        rgb_x = ref_basic.lab2rgb(torch.cat([gray, ab_x], 1))
        rgb_y = ref_basic.lab2rgb(torch.cat([gray, ab_y], 1))
        out = dict(gray=gray.numpy(), ab_x=ab_x.numpy(), ab_y=ab_y.numpy(), rgb_x=rgb_x.numpy(), vgg_seed=VGG_SEED)
        with torch.no_grad():
            for ft in ("liu", "lei", "conv4_4"):
                crit = ref_loss.VGG19Loss(feat_type=ft)
                out["loss_" + ft] = float(crit(rgb_x, rgb_y))
            full = ref_loss.AnchorColorProbLoss(hint2regress=False, enhanced=True, with_grad=False, mpdist=False, gpu_no=0)
            out["perceptual"] = float(full._perceptual_loss(gray, ab_x, ab_y))
        # Laplacian term (with_grad=True): value and the gradient autograd delivers to the prediction
        torch.Tensor.get_device = lambda self: self.device if self.device.type == "cpu" else self.device.index
        pred = ab_y.clone().requires_grad_(True)
        lap = full._laplace_gradient(pred, ab_x)
        lap.backward()
        out["laplace"], out["laplace_grad"] = float(lap), pred.grad.numpy()
        vgg = seeded_vgg19()
        out["weight_checksum"] = float(sum(p.double().abs().sum() for p in vgg.features.parameters()))
    finally:
        torchvision.models.vgg19 = orig
        os.chdir(cwd)
    path = os.path.join(ROOT, "tests", "golden", "vgg_loss.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
