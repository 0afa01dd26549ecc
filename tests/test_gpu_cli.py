"""GPU: the drop-in inference CLI end to end (reference main/colorizer/inference.py:56-139):
checkpoint file -> load_checkpoint -> sorted image files -> fetch_data -> forward -> Lab -> RGB PNG.

Expected images come from the oracle (fp32 torch CPU) fed with the same `fetch_data` tensors under the CLI's seeding
(np/torch seeded once with --seed, images in sorted order, inference.py:58-60,93) and converted with OpenCV exactly as
`util.save_normLabs_from_batch` does (utils/util.py:91-106).
"""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
pytestmark = pytest.mark.gpu


def _write_images(folder, sizes, seed=0):
    """Smooth synthetic colour JPEGs (random low-frequency fields), like natural images for the k-means step."""
    import cv2
    rng = np.random.default_rng(seed)
    names = []
    for i, (h, w) in enumerate(sizes):
        small = rng.random((h // 16 + 2, w // 16 + 2, 3)).astype(np.float32)
        img = cv2.resize(small, (w, h), interpolation=cv2.INTER_CUBIC)
        img = np.clip(img * 255.0, 0, 255).astype(np.uint8)
        name = f"img{i:02d}.jpg"
        cv2.imwrite(os.path.join(folder, name), img, [cv2.IMWRITE_JPEG_QUALITY, 95])
        names.append(name)
    return names


def _expected_pngs(sd, folder, names, no_resize, K, seed):
    """Oracle forward + the reference's Lab->RGB save path, one image at a time in sorted order (as the reference CLI)."""
    import cv2
    import disco_oracle as O
    from disentangledcolorization_b200 import inference
    np.random.seed(seed)
    torch.manual_seed(seed)
    out = {}
    for name in sorted(names):
        gray, ab, (H, W) = inference.fetch_data(os.path.join(folder, name), no_resize)
        with torch.no_grad():
            pred = O.forward(sd, gray, ab, K, 0)[2]
        lab = torch.cat((gray, pred), 1).permute(0, 2, 3, 1).numpy().copy()
        if no_resize:
            lab = lab[:, :H, :W, :]
        lab[..., 0] = lab[..., 0] * 50.0 + 50.0
        lab[..., 1:3] = lab[..., 1:3] * 110.0
        rgb = cv2.cvtColor(lab[0], cv2.COLOR_LAB2RGB)
        out[os.path.splitext(name)[0] + ".png"] = (rgb * 255.0).astype(np.uint8)
    return out


@pytest.mark.parametrize("no_resize", [False, True], ids=["resize256", "no_resize"])
def test_cli_writes_the_reference_pngs(tmp_path, synth_sd, no_resize):
    from PIL import Image
    from disentangledcolorization_b200 import inference
    data = tmp_path / "data"
    data.mkdir()
    sizes = [(96, 128), (70, 100), (128, 96)] if no_resize else [(96, 128), (120, 90)]
    names = _write_images(str(data), sizes)
    ckpt = tmp_path / "model_last.pth.tar"
    torch.save({"epoch": 0, "state_dict": synth_sd, "best_loss": 0.0}, str(ckpt))
    K = 4 if no_resize else 8
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        argv = ["--data", str(data), "--checkpt", str(ckpt), "--name", "out", "--n_clusters", str(K), "--precision", "fp32"]
        if no_resize:
            argv.append("--no_resize")
        inference.main(argv)
    finally:
        os.chdir(cwd)
    want = _expected_pngs(synth_sd, str(data), names, no_resize, K, seed=130)
    save_dir = tmp_path / f"out-anchor{K}"
    got_files = sorted(os.listdir(save_dir))
    assert got_files == sorted(want), got_files
    for f, exp in want.items():
        got = np.asarray(Image.open(save_dir / f))
        assert got.shape == exp.shape, (f, got.shape, exp.shape)
        diff = np.abs(got.astype(np.int32) - exp.astype(np.int32))
        # |d ab| < 1e-3 (x110 = 0.11 Lab units) moves an 8-bit channel by at most one level, and rarely
        assert diff.max() <= 1 and diff.mean() < 0.02, (f, int(diff.max()), float(diff.mean()))


def test_cli_batched_bf16_matches_one_at_a_time(tmp_path, synth_sd):
    """--batch N (extension) groups equal-size images into one forward; the PNGs equal the one-image-per-forward run
    (the forward is independent per image and the host RNG draws are consumed in the same order)."""
    from PIL import Image
    from disentangledcolorization_b200 import inference
    data = tmp_path / "data"
    data.mkdir()
    _write_images(str(data), [(80, 80)] * 5, seed=3)
    ckpt = tmp_path / "m.pth.tar"
    torch.save({"state_dict": synth_sd}, str(ckpt))
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        inference.main(["--data", str(data), "--checkpt", str(ckpt), "--name", "one"])
        inference.main(["--data", str(data), "--checkpt", str(ckpt), "--name", "four", "--batch", "4"])
    finally:
        os.chdir(cwd)
    a, b = tmp_path / "one-anchor8", tmp_path / "four-anchor8"
    files = sorted(os.listdir(a))
    assert len(files) == 5 and files == sorted(os.listdir(b))
    for f in files:
        x, y = np.asarray(Image.open(a / f)), np.asarray(Image.open(b / f))
        assert x.shape == (256, 256, 3) and np.array_equal(x, y), f
