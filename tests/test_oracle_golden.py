"""CPU: the oracle (oracle/disco_oracle.py) against fixtures generated from the unmodified reference
(oracle/make_golden.py).  This is what pins the oracle; GPU parity tests then compare the CUDA path
with the oracle and with the same fixtures."""
import json
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, GOLDEN, golden_cases, load_golden, case_inputs

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import disco_oracle as O  # noqa: E402

TOL = 2e-4   # fp32 CPU conv kernels may differ across hosts (thread count / ISA); same host gives 0.0


@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["name"])
def test_oracle_matches_reference_fixture(case, synth_sd):
    torch.set_flush_denormal(True)
    g = load_golden(case["name"])
    gray, ab = (torch.from_numpy(t) for t in case_inputs(case))
    np.random.seed(case["seed"])
    torch.manual_seed(case["seed"])
    with torch.no_grad():
        pal, ref, pred, aff, spix, hint = O.forward(synth_sd, gray, ab, case["K"], case["T"])
    st = int(g["affinity_stride"])
    assert np.array_equal(hint.numpy(), g["hint_mask"]), "anchor sites differ from the reference"
    assert np.abs(aff.numpy()[:, :, ::st, ::st] - g["affinity"]).max() < TOL
    assert np.abs(pal.numpy() - g["pal_logit"]).max() < 50 * TOL      # logits are O(10)
    assert np.abs(ref.numpy() - g["ref_logit"]).max() < 50 * TOL
    assert np.abs(spix.numpy() - g["spix_colors"]).max() < 1e-6
    assert np.abs(pred.numpy() - g["pred_colors"]).max() < 1e-3       # north-star fp32 tolerance on ab
    # RNG streams were consumed exactly as the reference consumes them
    assert int(np.random.randint(1 << 30)) == int(g["np_next"])
    assert int(torch.randint(1 << 30, (1,))) == int(g["torch_next"])


def test_gamut_table_properties():
    from disentangledcolorization_b200.cielab import Q_TO_AB
    assert Q_TO_AB.shape == (313, 2) and Q_TO_AB.dtype == np.float32
    assert np.all(Q_TO_AB % 10 == 0)
    order = np.lexsort((Q_TO_AB[:, 1], Q_TO_AB[:, 0]))
    assert np.array_equal(order, np.arange(313))           # sorted by (a, b) like ab[mask]
    assert Q_TO_AB[:, 0].min() == -90 and Q_TO_AB[:, 0].max() == 100
    assert Q_TO_AB[:, 1].min() == -110 and Q_TO_AB[:, 1].max() == 100


def test_encode_argmax_is_nearest_bin():
    """argmax(encode_ab2ind(x)) == nearest gamut bin."""
    rng = np.random.Generator(np.random.PCG64(5))
    ab = torch.from_numpy((rng.random((3, 2, 5, 7), dtype=np.float32) - 0.5) * 0.8)  # well inside the gamut hull
    lab = O.encode_ab2ind(ab).max(dim=1)[1]
    table = O.q_to_ab()
    d = torch.cdist((ab * 110).permute(0, 2, 3, 1).reshape(-1, 2), table)
    assert torch.equal(lab.flatten(), d.argmin(1))


def test_state_dict_schema_matches_reference():
    from disentangledcolorization_b200 import netspec
    with open(os.path.join(GOLDEN, "state_dict_schema.json")) as f:
        ref = json.load(f)
    mine = {k: (list(s), p) for k, s, p in netspec.schema()}
    assert len(mine) == len(ref) == 461
    for k, s, p in ref:
        assert mine[k] == (s, p), k


def test_oracle_losses_match_reference_fixture():
    """The oracle's restatement of AnchorColorProbLoss / RebalanceLoss / class weights / SPixelLoss / encode_ab2ind against
    tests/golden/loss_terms.npz, generated with the unmodified reference by oracle/make_golden_loss.py."""
    import numpy as np
    import torch
    import disco_oracle as O
    g = load_golden("loss_terms")
    w = O.class_weights()
    assert np.allclose(w.numpy(), g["weights_table"], rtol=1e-6, atol=0)
    soft = O.encode_ab2ind(torch.from_numpy(g["spix"]))
    assert np.abs(soft.numpy() - g["soft"]).max() < 1e-6
    labels = soft.max(dim=1, keepdim=True)[1]
    assert np.array_equal(labels.numpy(), g["labels"])
    cw = w[labels]
    assert np.allclose(cw.numpy(), g["class_weight"])
    pal = torch.from_numpy(g["pal"]).requires_grad_(True)
    ref = torch.from_numpy(g["ref"]).requires_grad_(True)
    d = O.anchor_color_prob_loss(pal, ref, labels, cw.float())
    d["totalLoss"].backward()
    assert abs(d["palLoss"].item() - float(g["palLoss"])) < 1e-5 and abs(d["refLoss"].item() - float(g["refLoss"])) < 1e-5
    assert np.abs(pal.grad.numpy() - g["pal_grad"]).max() < 1e-6 and np.abs(ref.grad.numpy() - g["ref_grad"]).max() < 1e-6
    sp = O.spixel_loss(torch.from_numpy(g["prob"]), torch.from_numpy(g["feat"]), 16)
    assert abs(sp["totalLoss"].item() - float(g["sp_total"])) < 1e-4 and abs(sp["posLoss"].item() - float(g["sp_pos"])) < 1e-6


def _seeded_vgg_convs():
    """[(weight, bias)] of torchvision's random-init vgg19 under the fixture's seed (oracle/make_golden_vgg.py)."""
    import make_golden_vgg
    vgg = make_golden_vgg.seeded_vgg19()
    convs = [(m.weight.detach(), m.bias.detach()) for m in vgg.features if isinstance(m, torch.nn.Conv2d)]
    return vgg, convs


def test_oracle_perceptual_term_matches_reference_fixture():
    """lab2rgb and VGG19Loss ('liu', 'lei', conv4_4 variants) of the oracle == the unmodified reference's values
    (tests/golden/vgg_loss.npz, oracle/make_golden_vgg.py), with the VGG weights rebuilt from the fixture's seed."""
    import disco_oracle as O
    g = load_golden("vgg_loss")
    vgg, convs = _seeded_vgg_convs()
    checksum = float(sum(p.detach().double().abs().sum() for p in vgg.features.parameters()))
    assert abs(checksum - float(g["weight_checksum"])) < 1e-6 * checksum, "torchvision's seeded init no longer reproduces the fixture's weights"
    gray, ab_x, ab_y = (torch.from_numpy(g[k]) for k in ("gray", "ab_x", "ab_y"))
    rgb_x = O.lab2rgb(torch.cat([gray, ab_x], 1))
    assert np.abs(rgb_x.numpy() - g["rgb_x"]).max() < 1e-6
    rgb_y = O.lab2rgb(torch.cat([gray, ab_y], 1))
    with torch.no_grad():
        for ft in ("liu", "lei", "conv4_4"):
            v = float(O.vgg19_loss(convs, rgb_x, rgb_y, ft))
            assert abs(v - float(g["loss_" + ft])) < 1e-5 * abs(float(g["loss_" + ft])) + 1e-7, (ft, v, float(g["loss_" + ft]))
        assert abs(float(O.perceptual_loss(convs, gray, ab_x, ab_y)) - float(g["perceptual"])) < 1e-6
    pred = ab_y.clone().requires_grad_(True)
    lap = O.laplace_gradient(pred, ab_x)
    lap.backward()
    assert abs(float(lap) - float(g["laplace"])) < 1e-6 and np.abs(pred.grad.numpy() - g["laplace_grad"]).max() < 1e-9
