"""Deterministic synthetic checkpoint in the reference's `state_dict` schema, plus synthetic inputs.

No checkpoint ships with the reference (checkpoints/disco_download.sh is a network download) and
its own random init overflows in eval mode, so parity tests, smoke() and
bench.py all use this generator: random weights drawn from a numpy PCG64 stream (bit-stable across
machines), spectral-norm vectors (u, v) obtained by power iteration so that sigma is the true
spectral norm as in a trained model, and BatchNorm running statistics set from per-layer scalars
calibrated once against the reference forward (synth_calib.json, produced by
oracle/calibrate_synth.py) so activations stay O(1) through all 71 convolutions.

The result loads into the *unmodified* reference model with `load_state_dict(strict=True)` --
the same path main/colorizer/inference.py takes through `load_checkpoint`
(main/utils_train.py:140-156) -- and into the drop-in model of this package.
"""
import json
import os

import numpy as np

from . import netspec

_CALIB = os.path.join(os.path.dirname(__file__), "synth_calib.json")


def _load_calib():
    if os.path.exists(_CALIB):
        with open(_CALIB) as f:
            return json.load(f)
    return {}


def _unit(x):
    return x / max(np.linalg.norm(x), 1e-12)


_CACHE = {}


def make_state_dict(seed=0, calib=None, as_torch=True):
    """Returns an OrderedDict key -> tensor (fp32; num_batches_tracked int64) with all 461 entries."""
    from collections import OrderedDict
    if calib is None and as_torch and seed in _CACHE:
        return OrderedDict((k, v.clone()) for k, v in _CACHE[seed].items())
    res = _make_state_dict(seed, calib, as_torch)
    if calib is None and as_torch:
        _CACHE[seed] = OrderedDict((k, v.clone()) for k, v in res.items())
    return res


def _make_state_dict(seed, calib, as_torch):
    from collections import OrderedDict
    rng = np.random.Generator(np.random.PCG64(seed))
    calib = _load_calib() if calib is None else calib
    sch = netspec.schema()
    shapes = {k: s for k, s, _ in sch}
    out = OrderedDict()

    def randn(shape, std=1.0):
        return (rng.standard_normal(shape, dtype=np.float32) * np.float32(std)).astype(np.float32)

    for key, shape, _ in sch:
        leaf = key.rsplit(".", 1)[1]
        base = key.rsplit(".", 1)[0]
        if key in out:
            continue
        if leaf in ("weight", "weight_orig") and len(shape) == 4:
            if "deconv" in key:                       # ConvTranspose2d (cin, cout, 4, 4): 4 taps reach each output
                fan_in = shape[0] * 4
            else:
                fan_in = shape[1] * 9
            gain = 2.0
            if key.endswith("pred_mask0.weight"):
                gain = 6.0
            if key.endswith("outConv.weight"):
                gain = 0.1
            w = randn(shape, np.sqrt(gain / fan_in))
            out[key] = w
            if leaf == "weight_orig":
                wm = w.reshape(shape[0], -1).astype(np.float64)
                v = _unit(rng.standard_normal(wm.shape[1]))
                for _ in range(4):
                    u = _unit(wm @ v)
                    v = _unit(wm.T @ u)
                out[base + ".weight_u"] = u.astype(np.float32)
                out[base + ".weight_v"] = v.astype(np.float32)
        elif leaf in ("weight_u", "weight_v"):
            continue                                  # written together with weight_orig
        elif leaf == "num_batches_tracked":
            out[key] = np.asarray(1, dtype=np.int64)
        elif leaf in ("running_mean", "running_var"):
            m, v = calib.get(base, (0.0, 1.0))
            if leaf == "running_mean":
                out[key] = (np.float32(m) + np.float32(0.1 * np.sqrt(v)) * randn(shape)).astype(np.float32)
            else:
                out[key] = (np.float32(v) * (1.0 + 0.2 * np.abs(randn(shape)))).astype(np.float32)
        elif len(shape) == 1:
            # conv bias / BN affine / LayerNorm affine / linear bias
            if ("norm" in key or base + ".running_mean" in shapes) and leaf == "weight":
                out[key] = (1.0 + 0.1 * randn(shape)).astype(np.float32)
            else:
                out[key] = randn(shape, 0.05)
        elif len(shape) == 2:
            if key in ("mid_word_prj.weight", "trg_word_prj.weight"):
                out[key] = randn(shape, 3.0 / np.sqrt(shape[1]))
            elif key == "trg_word_emb.weight":
                out[key] = randn(shape, 1.0 / np.sqrt(netspec.D_MODEL))
            else:
                out[key] = randn(shape, 1.0 / np.sqrt(shape[1]))
        else:
            raise AssertionError(key)
    ordered = OrderedDict((k, out[k]) for k, _, _ in sch)
    if as_torch:
        import torch
        return OrderedDict((k, torch.from_numpy(np.ascontiguousarray(v))) for k, v in ordered.items())
    return ordered


def make_gray(batch, height, width, seed=0, smooth=True):
    """Synthetic L-channel batch in [-1, 1] = (L-50)/50, (batch,1,H,W) float32.

    `smooth=True`: sum of bilinear-ish blobs at three scales plus fine noise -- natural-image-like
    statistics so that super-pixels and k-means clusters are not degenerate.  `smooth=False`:
    i.i.d. U(-1,1) as in BASELINE configs."""
    rng = np.random.Generator(np.random.PCG64(1000 + seed))
    if not smooth:
        return (rng.random((batch, 1, height, width), dtype=np.float32) * 2 - 1).astype(np.float32)
    img = np.zeros((batch, 1, height, width), np.float32)
    for cells, amp in ((4, 0.55), (16, 0.3), (64, 0.12)):
        gh, gw = max(2, height // (height // cells or 1)), max(2, width // (width // cells or 1))
        gh, gw = min(gh, height), min(gw, width)
        g = rng.standard_normal((batch, 1, gh, gw), dtype=np.float32)
        ry, rx = -(-height // gh), -(-width // gw)
        up = np.repeat(np.repeat(g, ry, axis=2), rx, axis=3)[:, :, :height, :width]
        # separable box blur to soften block edges (pure numpy, deterministic)
        k = max(1, ry // 2)
        if k > 1:
            c = np.cumsum(np.pad(up, ((0, 0), (0, 0), (k, k), (0, 0)), mode="edge"), axis=2)
            up = (c[:, :, 2 * k:, :] - c[:, :, :-2 * k, :])[:, :, :height] / (2 * k)
            c = np.cumsum(np.pad(up, ((0, 0), (0, 0), (0, 0), (k, k)), mode="edge"), axis=3)
            up = (c[:, :, :, 2 * k:] - c[:, :, :, :-2 * k])[:, :, :, :width] / (2 * k)
        img += amp * up.astype(np.float32)
    img += 0.03 * rng.standard_normal(img.shape, dtype=np.float32)
    return np.clip(img, -1.0, 1.0).astype(np.float32)
