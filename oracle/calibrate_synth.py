"""TEST INFRASTRUCTURE -- one-off calibration of disentangledcolorization_b200/synth_calib.json.

Runs the *reference* model (through oracle/ref_harness.py, authoring container only) on a synthetic
batch with forward pre-hooks on every BatchNorm2d: each hook measures the layer-average mean and
variance of the BN input, stores the two scalars, and rewrites that layer's running statistics the
way synth.make_state_dict will (scalar * per-channel perturbation) before the layer executes, so
later layers are calibrated on already-normalised activations.  Output: {bn_key: [mean, var]}.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import ref_harness  # noqa: E402
from disentangledcolorization_b200 import synth  # noqa: E402


def main():
    torch.set_num_threads(os.cpu_count())
    torch.set_flush_denormal(True)
    model = ref_harness.build_model(n_clusters=8)
    sd_unit = synth.make_state_dict(seed=0, calib={})
    model.load_state_dict(sd_unit, strict=True)
    model.eval()
    calib = {}
    names = {m: n for n, m in model.named_modules()}

    def hook(mod, inp):
        x = inp[0]
        m = float(x.mean(dim=(0, 2, 3)).mean())
        v = float(x.var(dim=(0, 2, 3), unbiased=False).mean())
        v = max(v, 1e-12)
        key = names[mod]
        calib[key] = [m, v]
        mod.running_mean.copy_(m + np.sqrt(v) * sd_unit[key + ".running_mean"])
        mod.running_var.copy_(v * sd_unit[key + ".running_var"])

    for mod in model.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.register_forward_pre_hook(hook)
    gray = torch.from_numpy(synth.make_gray(4, 256, 256, seed=12345))
    np.random.seed(0)
    torch.manual_seed(0)
    with torch.no_grad():
        out = model(gray, torch.zeros(4, 2, 256, 256), True, 0)
    for name, t in zip(("pal", "ref", "pred", "aff", "spix", "hint"), out):
        print(name, tuple(t.shape), float(t.abs().mean()), float(t.std()), bool(torch.isfinite(t).all()))
    path = os.path.join(os.path.dirname(HERE), "disentangledcolorization_b200", "synth_calib.json")
    with open(path, "w") as f:
        json.dump(calib, f, indent=0, sort_keys=True)
    print("wrote", path, len(calib))
    # check: regenerated checkpoint reproduces the hooked model's buffers
    sd = synth.make_state_dict(seed=0)
    ref_sd = model.state_dict()
    worst = max(float((sd[k].float() - ref_sd[k].float()).abs().max() / (ref_sd[k].float().abs().max() + 1e-12))
                for k in sd if k.endswith(("running_mean", "running_var")))
    print("regen rel err", worst)


if __name__ == "__main__":
    main()
