#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- fixture of the training-side losses, generated with the UNMODIFIED reference
(models/loss.py, models/basic.py imported from /root/reference through oracle/ref_harness.py, CPU).

    python oracle/make_golden_loss.py        # writes tests/golden/loss_terms.npz

Inputs are drawn from a seeded generator; outputs: AnchorColorProbLoss (hint2regress=False, enhanced=False) loss terms
and the gradients autograd delivers to pal_prob / ref_prob, ColorLabel.encode_ab2ind / get_classweights, SPixelLoss terms.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_harness  # noqa: E402


def main():
    ref_harness.load()
    cwd = os.getcwd()
    os.chdir(os.path.join(ref_harness.REF, "main", "colorizer"))
    try:
        import basic as ref_basic
        import loss as ref_loss
        color_class = ref_basic.ColorLabel()
        g = torch.Generator().manual_seed(1234)
        N, h, w = 3, 4, 6
        pal = (torch.randn(N, 313, h, w, generator=g) * 2.0).requires_grad_(True)
        ref = (torch.randn(N, 313, h, w, generator=g) * 3.0).requires_grad_(True)
        spix = (torch.rand(N, 2, h, w, generator=g) * 1.2 - 0.6)
        q = color_class.encode_ab2ind(spix)
        labels = torch.max(q, dim=1, keepdim=True)[1]                          # train_colorizer.py:143
        cw = color_class.get_classweights(labels)                               # :144
        crit = ref_loss.AnchorColorProbLoss(hint2regress=False, enhanced=False, with_grad=False, mpdist=False, gpu_no=0)
        d = crit({"target_label": labels, "pal_prob": pal, "ref_prob": ref, "class_weight": cw.float(), "pred_color": None,
                  "input_gray": None, "input_color": None, "spix_color": spix}, 0)
        d["totalLoss"].backward()
        # SPixelLoss on a 32 x 48 map, 16 x 16 super-pixels, Lab + xy features
        prob = torch.softmax(torch.randn(2, 9, 32, 48, generator=g), dim=1)
        feat = torch.randn(2, 5, 32, 48, generator=g)
        sp = ref_loss.SPixelLoss(psize=16)({"pred_prob": prob, "target_feat": feat}, 0)
        out = dict(pal=pal.detach().numpy(), ref=ref.detach().numpy(), spix=spix.numpy(), soft=q.numpy(),
                   labels=labels.numpy().astype(np.int32), class_weight=cw.numpy(), weights_table=color_class.weights.numpy(),
                   palLoss=d["palLoss"].item(), refLoss=d["refLoss"].item(), totalLoss=d["totalLoss"].item(),
                   pal_grad=pal.grad.numpy(), ref_grad=ref.grad.numpy(), prob=prob.numpy(), feat=feat.numpy(),
                   sp_total=sp["totalLoss"].item(), sp_feat=sp["featLoss"].item(), sp_pos=sp["posLoss"].item())
    finally:
        os.chdir(cwd)
    path = os.path.join(ROOT, "tests", "golden", "loss_terms.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: (v.shape if hasattr(v, "shape") else v) for k, v in out.items()})


if __name__ == "__main__":
    main()
