"""CPU: the inference CLI keeps the reference's flag surface (main/colorizer/inference.py:144-167) and the
fetch_data contract (inference.py:23-42)."""
import os
import sys

import numpy as np

REF_FLAGS = ["--name", "--seed", "--psize", "--data", "--model", "--checkpt", "--n_enc", "--n_dec", "--d_model",
             "--dense_pos", "--spix_pos", "--learning_pos", "--hint2regress", "--n_clusters", "--random_hint",
             "--no_resize", "--diverse"]


def test_cli_flags_and_defaults_match_reference():
    from disentangledcolorization_b200 import inference
    p = inference.build_parser()
    flags = {s for a in p._actions for s in a.option_strings}
    for f in REF_FLAGS:
        assert f in flags, f
    d = p.parse_args([])
    assert (d.seed, d.psize, d.n_clusters, d.d_model, d.name) == (130, 16, 8, 64, "test")
    assert not d.no_resize and not d.diverse and not d.random_hint


def test_fetch_data_shapes_and_ranges(tmp_path):
    import cv2
    from disentangledcolorization_b200 import inference
    rng = np.random.default_rng(0)
    img = (rng.random((50, 70, 3)) * 255).astype(np.uint8)
    path = str(tmp_path / "a.png")
    cv2.imwrite(path, img)
    gray, ab, (H, W) = inference.fetch_data(path, org_size=False)
    assert tuple(gray.shape) == (1, 1, 256, 256) and tuple(ab.shape) == (1, 2, 256, 256) and (H, W) == (50, 70)
    assert -1.0001 <= float(gray.min()) and float(gray.max()) <= 1.0001
    gray, ab, _ = inference.fetch_data(path, org_size=True)
    assert gray.shape[2] % 16 == 0 and gray.shape[3] % 16 == 0 and gray.shape[2] >= 50 and gray.shape[3] >= 70


def test_bare_name_compat_modules():
    from conftest import ROOT
    sys.path.insert(0, os.path.join(ROOT, "disentangledcolorization_b200", "compat"))
    try:
        import importlib
        m = importlib.import_module("model")
        assert hasattr(m, "AnchorColorProb") and hasattr(m, "SpixelSeg")
        b = importlib.import_module("basic")
        assert hasattr(b, "upfeat") and hasattr(b, "ColorLabel") and hasattr(b, "tensor2array")
    finally:
        sys.path.pop(0)
        for k in ("model", "basic", "network"):
            sys.modules.pop(k, None)


def test_fast_reader_is_bit_identical_to_fetch_data(tmp_path):
    """The pipeline's reader (`_fetch_into`: table lookup instead of the float64 division, results written into batch
    buffer views) feeds the network exactly what `fetch_data` (reference inference.py:23-42) feeds it."""
    import cv2
    from disentangledcolorization_b200 import inference
    rng = np.random.default_rng(5)
    for i, (h, w) in enumerate([(50, 70), (256, 256), (300, 481), (17, 9)]):
        img = (rng.random((h, w, 3)) * 255).astype(np.uint8)
        path = str(tmp_path / f"im{i}.png")
        cv2.imwrite(path, img)
        gray, ab, hw = inference.fetch_data(path, org_size=False)
        g = np.full((1, 256, 256), np.nan, np.float32)
        a = np.full((2, 256, 256), np.nan, np.float32)
        hw2 = inference._fetch_into(path, g, a)
        assert hw2 == hw
        assert np.array_equal(g, gray[0].numpy()) and np.array_equal(a, ab[0].numpy())
