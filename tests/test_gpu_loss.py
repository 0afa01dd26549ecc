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
        loss.AnchorColorProbLoss(hint2regress=True)          # not built (broken in the reference's own training branch)


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


def _seeded_vgg():
    import make_golden_vgg
    return make_golden_vgg.seeded_vgg19()


def test_lab2rgb_matches_reference_fixture():
    """basic.lab2rgb (disco_lab2rgb_norm) == the reference's torch formula (models/basic.py:431-475) on the fixture."""
    from disentangledcolorization_b200 import basic
    g = load_golden("vgg_loss")
    lab = torch.cat([torch.from_numpy(g["gray"]), torch.from_numpy(g["ab_x"])], 1).cuda()
    rgb = basic.lab2rgb(lab)
    assert tuple(rgb.shape) == g["rgb_x"].shape and rgb.dtype == torch.float32
    # fp32 on both sides; the cube, the 1/2.4 power (CUDA powf vs the CPU's pow) and FMA contraction differ in the last bits
    assert np.abs(rgb.cpu().numpy() - g["rgb_x"]).max() < 2e-5


# bf16 activations through up to 13 convolutions: the fp32 path must reproduce the reference's value; the bf16 tolerance is
# 3 x the relative error of a torch run of the same stack with bf16 storage between layers, measured in the test itself
@pytest.mark.parametrize("feat_type", ["liu", "lei", "conv4_4"])
def test_vgg19_loss_matches_reference_fixture(feat_type):
    import disco_oracle as O
    from disentangledcolorization_b200 import basic, loss
    g = load_golden("vgg_loss")
    vgg = _seeded_vgg()
    gray, ab_x, ab_y = (torch.from_numpy(g[k]).cuda() for k in ("gray", "ab_x", "ab_y"))
    rgb_x, rgb_y = basic.lab2rgb(torch.cat([gray, ab_x], 1)), basic.lab2rgb(torch.cat([gray, ab_y], 1))
    want = float(g["loss_" + feat_type])
    crit = loss.VGG19Loss(feat_type=feat_type, vgg_model=vgg, precision="fp32")
    keys = set(crit.state_dict())
    assert ("slice1.0.weight" in keys) == (feat_type != "conv4_4") and all(k.split(".")[0] in
                                                                           ("slice1", "slice2", "slice3", "slice4", "slice5", "featureExactor") for k in keys)
    got = float(crit(rgb_x, rgb_y))
    assert abs(got - want) < 2e-5 * want, (got, want)
    # bf16: tolerance from a torch emulation of bf16 storage (weights and every activation rounded to bf16)
    convs = [(m.weight.detach().cuda(), m.bias.detach().cuda()) for m in vgg.features if isinstance(m, torch.nn.Conv2d)]
    q = lambda t: t.to(torch.bfloat16).float()

    def emulated():
        import torch.nn.functional as F
        mean = torch.tensor([0.485, 0.456, 0.406], device="cuda")[None, :, None, None]
        std = torch.tensor([0.229, 0.224, 0.225], device="cuda")[None, :, None, None]
        z = q(torch.cat(((rgb_x - mean) / std, (rgb_y - mean) / std), 0))
        ends, wts = O.VGG_SLICES.get(feat_type, [28]), O.VGG_WEIGHTS.get(feat_type, [1.0])
        total, idx, ci, k, N = 0.0, 0, 0, 0, rgb_x.shape[0]
        for v in O.VGG19_CFG:
            if v == "M":
                z, idx = F.max_pool2d(z, 2, 2), idx + 1
            else:
                w, b = convs[ci]
                ci += 1
                z, idx = q(F.relu(F.conv2d(z, q(w), b, padding=1))), idx + 2
            if k < len(ends) and idx == ends[k]:
                total = total + wts[k] * (z[:N] - z[N:]).abs().mean()
                k += 1
                if k == len(ends):
                    break
        return float(total)

    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    emu_err = abs(emulated() - want) / want
    got16 = float(loss.VGG19Loss(feat_type=feat_type, vgg_model=vgg, precision="bf16")(rgb_x, rgb_y))
    tol = max(3.0 * emu_err, 5e-3)
    print(f"\n[{feat_type}] fp32 {got:.6f} bf16 {got16:.6f} reference {want:.6f}; bf16-storage emulation error {emu_err:.2e}, gate {tol:.2e}")
    assert abs(got16 - want) < tol * want, (got16, want, tol)


def test_anchor_color_prob_loss_enhanced_adds_five_times_the_perceptual_term():
    """loss.py:79-81: recLoss = 5 * VGGLoss(lab2rgb(gray, pred_color), lab2rgb(gray, input_color)) (argument order as the
    reference passes them)."""
    from disentangledcolorization_b200 import loss
    g, gl = load_golden("vgg_loss"), load_golden("loss_terms")
    crit = loss.AnchorColorProbLoss(hint2regress=False, enhanced=True,
                                    vgg_loss=loss.VGG19Loss(vgg_model=_seeded_vgg(), precision="fp32"))
    data = {"target_label": torch.from_numpy(gl["labels"]).cuda().long(), "pal_prob": torch.from_numpy(gl["pal"]).cuda(),
            "ref_prob": torch.from_numpy(gl["ref"]).cuda(), "class_weight": torch.from_numpy(gl["class_weight"]).cuda().float(),
            "input_gray": torch.from_numpy(g["gray"]).cuda(), "input_color": torch.from_numpy(g["ab_x"]).cuda(),
            "pred_color": torch.from_numpy(g["ab_y"]).cuda()}
    d = crit(data, 0)
    want = 5.0 * float(g["perceptual"])
    assert abs(d["recLoss"].item() - want) < 1e-4 * want
    assert abs(d["totalLoss"].item() - (float(gl["totalLoss"]) + want)) < 1e-4


def test_laplacian_term_matches_reference_value_and_gradient():
    """with_grad=True (loss.py:51-57,82-84): value and d loss / d pred_color == the reference's autograd."""
    from disentangledcolorization_b200 import loss
    g, gl = load_golden("vgg_loss"), load_golden("loss_terms")
    crit = loss.AnchorColorProbLoss(hint2regress=False, enhanced=True, with_grad=True,
                                    vgg_loss=loss.VGG19Loss(vgg_model=_seeded_vgg(), precision="fp32"))
    pred = torch.from_numpy(g["ab_y"]).cuda().requires_grad_(True)
    lap = crit._laplace_gradient(pred, torch.from_numpy(g["ab_x"]).cuda())
    assert abs(lap.item() - float(g["laplace"])) < 2e-6
    lap.backward()
    assert np.abs(pred.grad.cpu().numpy() - g["laplace_grad"]).max() < 1e-9
    data = {"target_label": torch.from_numpy(gl["labels"]).cuda().long(), "pal_prob": torch.from_numpy(gl["pal"]).cuda(),
            "ref_prob": torch.from_numpy(gl["ref"]).cuda(), "class_weight": torch.from_numpy(gl["class_weight"]).cuda().float(),
            "input_gray": torch.from_numpy(g["gray"]).cuda(), "input_color": torch.from_numpy(g["ab_x"]).cuda(),
            "pred_color": torch.from_numpy(g["ab_y"]).cuda()}
    d = crit(data, 0)
    want = 5.0 * float(g["perceptual"]) + float(g["laplace"])
    assert abs(d["recLoss"].item() - want) < 1e-4 * want
    with pytest.raises(Exception):
        loss.AnchorColorProbLoss(hint2regress=True)          # broken in the reference's own training branch: not built


def test_vgg_side_ops_reject_bad_shapes():
    from disentangledcolorization_b200 import _lib, loss
    crit = loss.VGG19Loss(vgg_model=_seeded_vgg(), precision="bf16")
    with pytest.raises(_lib.DiscoError):
        crit(torch.rand(1, 3, 24, 24).cuda(), torch.rand(1, 3, 24, 32).cuda())
    with pytest.raises(_lib.DiscoError):
        crit(torch.rand(1, 3, 20, 20).cuda(), torch.rand(1, 3, 20, 20).cuda())      # 20 -> 10 -> 5: odd at the third pooling layer
