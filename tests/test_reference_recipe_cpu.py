"""CPU, authoring container only (skipped where /root/reference is absent, e.g. on the GPU box): the SURVEY 8(c)
"random-init" recipe on the reference's OWN weights -- seed -> construct -> train() -> 3 no-grad forwards -> eval() -- and
the oracle run on that instance's state_dict.  Pins the oracle's handling of the spectral-norm u / v vectors and of the
BatchNorm running statistics as the reference itself produces them (the other fixtures use the synthetic checkpoint):
VERDICT r1, missing item 10.  SURVEY fact 4: without the warm-up forwards the reference's eval() output is NaN."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_harness  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_harness.available(), reason="needs the reference checkout (/root/reference)")


def test_oracle_matches_reference_on_the_references_own_warmed_up_weights():
    import disco_oracle as O
    torch.set_flush_denormal(True)
    np.random.seed(130)
    torch.manual_seed(130)
    m = ref_harness.build_model(n_clusters=4)
    g = torch.Generator().manual_seed(5)
    warm = torch.rand(2, 1, 64, 64, generator=g) * 2 - 1
    warm_ab = torch.rand(2, 2, 64, 64, generator=g) * 0.6 - 0.3
    m.train()
    with torch.no_grad():
        for _ in range(3):
            m(warm, warm_ab, True, 0)                   # refreshes u / v (power iteration) and the BN running statistics
    m.eval()
    gray = torch.rand(2, 1, 64, 96, generator=g) * 2 - 1
    ab = torch.zeros(2, 2, 64, 96)
    np.random.seed(7)
    torch.manual_seed(7)
    with torch.no_grad():
        want = m(gray, ab, True, 0)
    assert all(torch.isfinite(t).all() for t in want)
    sd = {k: v.detach().clone() for k, v in m.state_dict().items()}
    np.random.seed(7)
    torch.manual_seed(7)
    with torch.no_grad():
        got = O.forward(sd, gray, ab, 4, 0)
    assert torch.equal(got[5], want[5])                                   # anchors
    assert float((got[2] - want[2]).abs().max()) < 1e-5                   # pred_colors
    assert float((got[0] - want[0]).abs().max()) < 1e-3 * max(1.0, float(want[0].abs().max()))
    assert float((got[3] - want[3]).abs().max()) < 1e-6                   # affinity
    # and the drop-in's state_dict schema is the reference's (461 keys, same shapes)
    from disentangledcolorization_b200 import netspec
    schema = {k: tuple(s) for k, s, _ in netspec.schema()}
    ref_shapes = {k: tuple(v.shape) for k, v in sd.items()}
    assert schema == ref_shapes
