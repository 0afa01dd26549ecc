"""GPU: the drop-in super-pixel CLI (reference main/spixelseg/inference.py:39-108): checkpoint file -> SpixelSeg
forward -> winner-take-all super-pixel map, colour reconstruction from per-super-pixel means, gray image.

Expected outputs come from the oracle (fp32 torch CPU SpixelNet, poolfeat, upfeat), the reference's own `split_spixels`
arithmetic on the reference's id grid, and OpenCV / PIL exactly as utils/util.py saves them."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
pytestmark = pytest.mark.gpu


def _ref_split_spixels(assign_map, spixel_ids):
    """main/spixelseg/inference.py:67-75 (torch, CPU)."""
    N = assign_map.shape[0]
    spixel_id_map = spixel_ids.expand(N, -1, -1, -1)
    assig_max, _ = torch.max(assign_map, dim=1, keepdim=True)
    assignment_ = torch.where(assign_map == assig_max, torch.ones(assign_map.shape), torch.zeros(assign_map.shape))
    return torch.sum(spixel_id_map * assignment_, dim=1, keepdim=True).type(torch.int)


def test_split_spixels_and_grid_match_the_reference_arithmetic():
    from disentangledcolorization_b200 import basic
    g = torch.Generator().manual_seed(11)
    prob = torch.softmax(torch.randn(2, 9, 64, 96, generator=g) * 2, dim=1)
    prob[0, :, 5, 7] = 1.0 / 9                                    # a nine-way tie: the reference sums all nine ids
    ids_grid, _ = basic.init_spixel_grid(64, 96, 16)
    want = _ref_split_spixels(prob, ids_grid.unsqueeze(0))
    got = basic.split_spixels(prob.cuda(), 16).cpu()
    assert got.dtype == torch.int32 and torch.equal(got, want)


def test_spixel_cli_writes_the_reference_outputs(tmp_path, synth_sd):
    import cv2
    import disco_oracle as O
    from PIL import Image
    from disentangledcolorization_b200 import basic, spixel_inference
    data = tmp_path / "data"
    data.mkdir()
    rng = np.random.default_rng(4)
    names = []
    for i, (h, w) in enumerate([(64, 96), (128, 64)]):
        small = rng.random((h // 16 + 2, w // 16 + 2, 3)).astype(np.float32)
        img = np.clip(cv2.resize(small, (w, h), interpolation=cv2.INTER_CUBIC) * 255, 0, 255).astype(np.uint8)
        names.append(f"im{i}.png")
        cv2.imwrite(str(data / names[-1]), img)
    ck_dir = tmp_path / "run" / "checkpts"
    ck_dir.mkdir(parents=True)
    ckpt = ck_dir / "model_last.pth.tar"
    seg_sd = {k[len("segnet."):]: v for k, v in synth_sd.items() if k.startswith("segnet.")}
    torch.save({"state_dict": seg_sd}, str(ckpt))
    n = spixel_inference.main(["--data", str(data), "--checkpt", str(ckpt), "--name", "result", "--precision", "fp32"])
    out_dir = tmp_path / "run" / "result"
    assert n == 2 and sorted(os.listdir(out_dir)) == sorted(sum(([f, f.replace(".png", "-recon.png"), f.replace(".png", "-g.png")] for f in names), []))
    for f in names:
        gray, color, bgr = spixel_inference.fetch_data(str(data / f))
        with torch.no_grad():
            prob = O.spixelnet(synth_sd, gray)
            recon = O.upfeat(O.poolfeat(color, prob, 16)[0], prob, 16)
        H, W = gray.shape[2:]
        # gray image: exact
        want_g = (127.5 * (gray[0, 0].numpy() + 1.0)).astype(np.uint8)
        assert np.array_equal(np.asarray(Image.open(out_dir / f.replace(".png", "-g.png"))), want_g)
        # reconstruction: Lab -> RGB through OpenCV, within one 8-bit level
        lab = torch.cat((gray, recon), 1).permute(0, 2, 3, 1).numpy().copy()
        lab[..., 0] = lab[..., 0] * 50.0 + 50.0
        lab[..., 1:3] = lab[..., 1:3] * 110.0
        want_rgb = (cv2.cvtColor(lab[0], cv2.COLOR_LAB2RGB) * 255.0).astype(np.uint8)
        got_rgb = np.asarray(Image.open(out_dir / f.replace(".png", "-recon.png")))
        d = np.abs(got_rgb.astype(np.int32) - want_rgb.astype(np.int32))
        assert d.max() <= 1 and d.mean() < 0.05, (f, int(d.max()), float(d.mean()))
        # marked image: boundaries (white) exactly where the oracle's super-pixel map, run through the same boundary rule, has them
        ids_grid, _ = basic.init_spixel_grid(H, W, 16)
        want_ids = _ref_split_spixels(prob, ids_grid.unsqueeze(0))[0, 0].numpy()
        base = bgr[0].permute(1, 2, 0).numpy() * 0.5 + 0.5
        want_marked = (spixel_inference.mark_boundaries(base, want_ids.astype(int), color=(1, 1, 1)) * 255.0).astype(np.uint8)
        got_marked = np.asarray(Image.open(out_dir / f))
        assert got_marked.shape == want_marked.shape
        assert (got_marked != want_marked).any(axis=2).mean() < 2e-3, f       # an arg-max flip at a near tie moves a boundary pixel


def test_boundary_rule_marks_label_changes():
    """The restated skimage 'outer' rule on a hand-made label map: boundaries sit on the pixels next to a label change, never
    inside a constant region."""
    from disentangledcolorization_b200 import spixel_inference
    lab = np.zeros((8, 8), np.int64)
    lab[:, 4:] = 1
    lab[4:, :4] = 2
    b = spixel_inference.find_boundaries(lab)
    assert b.any() and not b[0:2, 0:2].any() and not b[6:, 6:].any() and not b[6:, 0:2].any()
    assert b[0, 3] or b[0, 4]
