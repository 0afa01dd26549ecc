import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_cases():
    with open(os.path.join(GOLDEN, "cases.json")) as f:
        return json.load(f)["cases"]


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def case_inputs(case):
    """Regenerates the (gray, ab) inputs of a golden case from its seeds (see oracle/make_golden.py)."""
    from disentangledcolorization_b200 import synth
    gray = synth.make_gray(case["B"], case["H"], case["W"], seed=case["gray_seed"], smooth=case["smooth"])
    if case["ab"] == "zero":
        ab = np.zeros((case["B"], 2, case["H"], case["W"]), np.float32)
    else:
        rng = np.random.Generator(np.random.PCG64(77 + case["gray_seed"]))
        g = rng.random((case["B"], 2, case["H"] // 16, case["W"] // 16), dtype=np.float32) * 1.2 - 0.6
        ab = np.repeat(np.repeat(g, 16, 2), 16, 3).astype(np.float32)
    return gray, ab


@pytest.fixture(scope="session")
def synth_sd():
    from disentangledcolorization_b200 import synth
    return synth.make_state_dict(seed=0)
