"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz from the UNMODIFIED reference.

Run in the authoring container only (needs /root/reference; imports it through
oracle/ref_harness.py):  python oracle/make_golden.py

For each case the synthetic checkpoint (disentangledcolorization_b200/synth.py, seed in the case)
is loaded into the reference `AnchorColorProb` with load_state_dict(strict=True), the model is put
in eval() and called exactly as main/colorizer/inference.py:108-109 does
(`model(gray, ab, True, sampled_T)`) after seeding numpy and torch like inference.py:58-60.
Inputs are not stored: tests regenerate them from the seeds via synth.make_gray.
Large maps are stored on a fixed stride (`*_stride` entries) to keep fixtures small.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_harness  # noqa: E402
from disentangledcolorization_b200 import synth  # noqa: E402

CASES = [
    # name, B, H, W, K, T, gray_seed, smooth, rng_seed, ab_mode
    dict(name="c1_256_k8", B=1, H=256, W=256, K=8, T=0, gray_seed=1, smooth=True, seed=130, ab="zero"),
    dict(name="ragged_64x96_b2", B=2, H=64, W=96, K=8, T=0, gray_seed=2, smooth=True, seed=130, ab="zero"),
    dict(name="noise_128_k16", B=1, H=128, W=128, K=16, T=0, gray_seed=3, smooth=False, seed=7, ab="zero"),
    dict(name="diverse_64_T2", B=1, H=64, W=64, K=4, T=2, gray_seed=4, smooth=True, seed=130, ab="zero"),
    dict(name="gt_anchor_64_Tm1", B=2, H=64, W=64, K=4, T=-1, gray_seed=5, smooth=True, seed=130, ab="rand"),
    dict(name="b3_128x64", B=3, H=128, W=64, K=8, T=0, gray_seed=6, smooth=True, seed=11, ab="zero"),
]


def make_ab(case):
    if case["ab"] == "zero":
        return np.zeros((case["B"], 2, case["H"], case["W"]), np.float32)
    rng = np.random.Generator(np.random.PCG64(77 + case["gray_seed"]))
    g = rng.random((case["B"], 2, case["H"] // 16, case["W"] // 16), dtype=np.float32) * 1.2 - 0.6
    return np.repeat(np.repeat(g, 16, 2), 16, 3).astype(np.float32)


def main():
    torch.set_num_threads(os.cpu_count())
    torch.set_flush_denormal(True)
    outdir = os.path.join(ROOT, "tests", "golden")
    sd = synth.make_state_dict(seed=0)
    models = {}
    index = []
    for case in CASES:
        K = case["K"]
        if K not in models:
            m = ref_harness.build_model(n_clusters=K)
            m.load_state_dict(sd, strict=True)
            m.eval()
            models[K] = m
        gray = torch.from_numpy(synth.make_gray(case["B"], case["H"], case["W"], seed=case["gray_seed"],
                                                smooth=case["smooth"]))
        ab = torch.from_numpy(make_ab(case))
        np.random.seed(case["seed"])
        torch.manual_seed(case["seed"])
        with torch.no_grad():
            pal, ref, pred, aff, spix, hint = models[K](gray, ab, True, case["T"])
        np_draw = int(np.random.randint(1 << 30))        # RNG stream positions after the call
        th_draw = int(torch.randint(1 << 30, (1,)))
        st = 4 if case["H"] * case["W"] >= 256 * 256 else 1
        arrays = dict(pal_logit=pal.numpy(), ref_logit=ref.numpy(), pred_colors=pred.numpy(),
                      affinity=aff.numpy()[:, :, ::st, ::st].copy(), affinity_stride=np.asarray(st),
                      spix_colors=spix.numpy(), hint_mask=hint.numpy(),
                      np_next=np.asarray(np_draw), torch_next=np.asarray(th_draw))
        np.savez_compressed(os.path.join(outdir, case["name"] + ".npz"), **arrays)
        index.append(case)
        print(case["name"], {k: v.shape for k, v in arrays.items()}, "hint sum", float(hint.sum()),
              "pred std", float(pred.std()))
    with open(os.path.join(outdir, "cases.json"), "w") as f:
        json.dump({"checkpoint_seed": 0, "cases": index}, f, indent=1)


if __name__ == "__main__":
    main()
