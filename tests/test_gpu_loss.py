"""GPU: the training-side loss kernels (config 5 first slice) against the fixture generated with the unmodified reference
(tests/golden/loss_terms.npz) and against the oracle's autograd on a larger seeded case.  Through the drop-in `loss` /
`basic` modules, i.e. through the C ABI (disco_ce_rebalance, disco_encode_ab2ind, disco_spixel_recon_loss)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden

sys.path.insert(0, os.path.join(ROOT, "oracle"))
pytestmark = pytest.mark.gpu


def test_color_label_tables_and_soft_encoding_match_reference():
    from disentangledcolorization_b200 import basic
    g = load_golden("loss_terms")
    cl = basic.ColorLabel(device="cuda")
    assert cl.weights.dtype == torch.float32 and np.allclose(cl.weights.cpu().numpy(), g["weights_table"], rtol=1e-6, atol=0)
    soft = cl.encode_ab2ind(torch.from_numpy(g["spix"]).cuda())
    assert tuple(soft.shape) == g["soft"].shape and np.abs(soft.cpu().numpy() - g["soft"]).max() < 1e-5
    labels = torch.max(soft, dim=1, keepdim=True)[1]
    assert np.array_equal(labels.cpu().numpy(), g["labels"])
    assert np.array_equal(cl.encode_ab2ind_hard(torch.from_numpy(g["spix"]).cuda()).cpu().numpy(), g["labels"])
    assert np.allclose(cl.get_classweights(labels).cpu().numpy(), g["class_weight"])


def test_anchor_color_prob_loss_matches_reference_values_and_gradients():
    from disentangledcolorization_b200 import loss
    g = load_golden("loss_terms")
    pal = torch.from_numpy(g["pal"]).cuda().requires_grad_(True)
    ref = torch.from_numpy(g["ref"]).cuda().requires_grad_(True)
    data = {"target_label": torch.from_numpy(g["labels"]).cuda().long(), "pal_prob": pal, "ref_prob": ref,
            "class_weight": torch.from_numpy(g["class_weight"]).cuda().float()}
    d = loss.AnchorColorProbLoss(hint2regress=False, enhanced=False)(data, 0)
    assert set(d) == {"totalLoss", "palLoss", "refLoss", "recLoss"}
    assert abs(d["palLoss"].item() - float(g["palLoss"])) < 2e-5 and abs(d["refLoss"].item() - float(g["refLoss"])) < 2e-5
    assert abs(d["totalLoss"].item() - float(g["totalLoss"])) < 4e-5 and d["recLoss"].item() == 0.0
    d["totalLoss"].backward()
    assert np.abs(pal.grad.cpu().numpy() - g["pal_grad"]).max() < 1e-6
    assert np.abs(ref.grad.cpu().numpy() - g["ref_grad"]).max() < 1e-6
    with pytest.raises(Exception):
        loss.AnchorColorProbLoss(enhanced=True)              # VGG19 perceptual term: not built


def test_ce_rebalance_large_case_with_ignored_tokens_matches_oracle():
    """Config-5 shape per GPU (32 x 313 x 16 x 16) with a few ignored tokens (label -1)."""
    import disco_oracle as O
    from disentangledcolorization_b200 import loss
    gen = torch.Generator().manual_seed(9)
    N, h, w = 32, 16, 16
    pal = torch.randn(N, 313, h, w, generator=gen) * 4
    ref = torch.randn(N, 313, h, w, generator=gen)
    labels = torch.randint(0, 313, (N, 1, h, w), generator=gen)
    labels[0, 0, :2] = -1
    cw = O.class_weights()[labels.clamp(min=0)].float()
    a, b = pal.clone().requires_grad_(True), ref.clone().requires_grad_(True)
    want = O.anchor_color_prob_loss(a, b, labels, cw)
    want["totalLoss"].backward()
    pa, pb = pal.cuda().requires_grad_(True), ref.cuda().requires_grad_(True)
    got = loss.AnchorColorProbLoss()({"target_label": labels.cuda(), "pal_prob": pa, "ref_prob": pb, "class_weight": cw.cuda()}, 0)
    got["totalLoss"].backward()
    assert abs(got["palLoss"].item() - want["palLoss"].item()) < 1e-4 and abs(got["refLoss"].item() - want["refLoss"].item()) < 1e-4
    assert (pa.grad.cpu() - a.grad).abs().max() < 1e-7 + 1e-4 * a.grad.abs().max()
    assert (pb.grad.cpu() - b.grad).abs().max() < 1e-7 + 1e-4 * b.grad.abs().max()
    assert float(pa.grad[0, :, 0, :].abs().max()) == 0.0     # ignored tokens receive no gradient


def test_spixel_loss_matches_reference_fixture():
    from disentangledcolorization_b200 import loss
    g = load_golden("loss_terms")
    d = loss.SPixelLoss(psize=16)({"pred_prob": torch.from_numpy(g["prob"]).cuda(), "target_feat": torch.from_numpy(g["feat"]).cuda()}, 0)
    assert abs(d["featLoss"].item() - float(g["sp_feat"])) < 1e-5
    assert abs(d["posLoss"].item() - float(g["sp_pos"])) < 1e-6
    assert abs(d["totalLoss"].item() - float(g["sp_total"])) < 1e-4
