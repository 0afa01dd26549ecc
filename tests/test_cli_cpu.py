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


def _png_cases():
    import cv2
    rng = np.random.default_rng(5)
    small = rng.random((18, 18, 3)).astype(np.float32)
    smooth = np.clip(cv2.resize(small, (256, 256), interpolation=cv2.INTER_CUBIC) * 255, 0, 255).astype(np.uint8)
    noisy = np.clip(smooth + rng.normal(0, 3, smooth.shape), 0, 255).astype(np.uint8)
    skew = np.cumsum(np.minimum(rng.geometric(0.5, (120, 200, 3)), 255), axis=1, dtype=np.uint64).astype(np.uint8)
    big = rng.integers(0, 256, (64, 80, 3), dtype=np.uint8)
    return {"smooth": smooth, "noisy": noisy, "flat": np.full((64, 48, 3), 7, np.uint8), "zeros": np.zeros((5, 3, 3), np.uint8),
            "1x1": np.array([[[1, 2, 3]]], np.uint8), "random": rng.integers(0, 256, (100, 77, 3), dtype=np.uint8),
            "deep_tree": skew, "row_stride_view": big[:50, :60], "long_runs": np.repeat(big[:8, :3], 100, axis=1)}


def test_native_png_writer_stores_the_pixels_pil_would(tmp_path):
    """disco_host_png_write replaces Image.fromarray(rgb).save(path, 'PNG') (reference utils/util.py:106): PIL and OpenCV
    decode the file to exactly the array that was written, and the stream passes PIL's CRC / Adler verification."""
    import cv2
    from PIL import Image
    from disentangledcolorization_b200 import _lib
    for name, img in _png_cases().items():
        path = str(tmp_path / f"{name}.png")
        _lib.png_write(path, img)
        Image.open(path).verify()
        with Image.open(path) as im:
            assert im.mode == "RGB" and im.size == (img.shape[1], img.shape[0]), name
            assert np.array_equal(np.asarray(im), img), name
        assert np.array_equal(cv2.cvtColor(cv2.imread(path, cv2.IMREAD_COLOR), cv2.COLOR_BGR2RGB), img), name
        with open(path, "rb") as f:
            assert f.read() == _lib.png_encode(img), name
        ref = str(tmp_path / "ref.png")
        Image.fromarray(np.ascontiguousarray(img)).save(ref, "PNG")
        if img.size > 10000:                                  # not LZ77: larger than PIL's level 6, but never by much
            assert os.path.getsize(path) < 1.6 * os.path.getsize(ref) + 2048, (name, os.path.getsize(path), os.path.getsize(ref))


def test_native_png_writer_rejects_bad_input(tmp_path):
    import pytest
    from disentangledcolorization_b200 import _lib
    with pytest.raises(ValueError):
        _lib.png_encode(np.zeros((4, 4, 3), np.float32))
    with pytest.raises(ValueError):
        _lib.png_encode(np.zeros((4, 4, 4), np.uint8))
    with pytest.raises(ValueError):
        _lib.png_encode(np.zeros((4, 8, 3), np.uint8)[:, ::2])
    with pytest.raises(RuntimeError):
        _lib.png_write(str(tmp_path / "no_such_dir" / "a.png"), np.zeros((4, 4, 3), np.uint8))


def test_spixel_cli_flags_and_grid_match_reference():
    """main/spixelseg/inference.py:130-137 flag surface; basic.init_spixel_grid == the reference's (models/basic.py:221-262)
    when the reference is present (it is not on the GPU box)."""
    import torch
    from disentangledcolorization_b200 import basic, spixel_inference
    d = spixel_inference.build_parser().parse_args([])
    assert (d.name, d.psize, d.model) == ("result", 16, "SpixelSeg")
    ids, coords = basic.init_spixel_grid(64, 96, 16)
    assert tuple(ids.shape) == (9, 64, 96) and tuple(coords.shape) == (2, 64, 96)
    assert ids[4, 17, 33] == 1 * 6 + 2 and ids[0, 0, 0] == 0 and ids[8, 63, 95] == 3 * 6 + 5     # centre channel = own cell; edges replicate
    assert coords[0, 5, 9] == 9 and coords[1, 5, 9] == 5
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import ref_harness
    if ref_harness.available():
        ref_harness.load()
        import basic as ref_basic
        for H, W, sp in ((256, 256, 16), (64, 96, 16), (128, 64, 8)):
            a, b = ref_basic.init_spixel_grid(H, W, sp)
            c, e = basic.init_spixel_grid(H, W, sp)
            assert torch.equal(a, c) and torch.equal(b, e)
