"""GPU: the whole drop-in forward (through the C ABI) against the CPU oracle and the committed
reference fixtures.

Tolerances
  fp32 path : |d ab| <= 1e-3 on pred_colors (north-star gate), identical hint_mask, RNG streams advanced
              exactly as the reference advances them.
  bf16 path : activations are stored in bf16 (8 mantissa bits) through ~50 stacked convolutions.  The tolerance is
              DERIVED, not asserted: oracle/derive_bf16_tolerance.py runs the oracle in fp32 and again with bf16-rounded
              conv weights / stored activations (torch CPU) on the fixture inputs, the benchmarked inputs and a 512x512
              image; tests/golden/bf16_tolerance.json records max 0.057 / mean 0.0079 for that torch-vs-torch baseline and
              the gate is 1.5 x that (0.0857 / 0.01178), anchors injected (anchor choice is a discrete function of
              fp32-sensitive k-means).  The CUDA path measures max 0.04-0.062, mean 0.007: inside the baseline.
"""
import json
import os
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, GOLDEN, golden_cases, load_golden, case_inputs

sys.path.insert(0, os.path.join(ROOT, "oracle"))
pytestmark = pytest.mark.gpu

FP32_AB_TOL = 1e-3
with open(os.path.join(GOLDEN, "bf16_tolerance.json")) as _f:
    _TOL = json.load(_f)
BF16_AB_MAX, BF16_AB_MEAN = _TOL["BF16_AB_MAX"], _TOL["BF16_AB_MEAN"]


def _model(sd, K, precision):
    from disentangledcolorization_b200 import model
    m = model.AnchorColorProb(inChannel=1, outChannel=313, sp_size=16, d_model=64, use_dense_pos=True,
                              n_clusters=K, enhanced=True, precision=precision)
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval()


CASES = [c for c in golden_cases() if c["T"] <= 0]   # T=2 (--diverse) has its own test below


@pytest.mark.parametrize("case", CASES, ids=lambda c: c["name"])
def test_forward_fp32_matches_reference_fixture(case, synth_sd):
    g = load_golden(case["name"])
    gray, ab = (torch.from_numpy(t) for t in case_inputs(case))
    m = _model(synth_sd, case["K"], "fp32")
    np.random.seed(case["seed"])
    torch.manual_seed(case["seed"])
    pal, ref, pred, aff, spix, hint = m(gray.cuda(), ab.cuda(), True, case["T"])
    torch.cuda.synchronize()
    st = int(g["affinity_stride"])
    assert np.array_equal(hint.cpu().numpy(), g["hint_mask"]), "anchor sites differ from the reference"
    assert np.abs(aff.cpu().numpy()[:, :, ::st, ::st] - g["affinity"]).max() < 1e-4
    assert np.abs(pal.cpu().numpy() - g["pal_logit"]).max() < 1e-2
    assert np.abs(ref.cpu().numpy() - g["ref_logit"]).max() < 1e-2
    assert np.abs(spix.cpu().numpy() - g["spix_colors"]).max() < 1e-6
    assert np.abs(pred.cpu().numpy() - g["pred_colors"]).max() < FP32_AB_TOL
    assert int(np.random.randint(1 << 30)) == int(g["np_next"])
    assert int(torch.randint(1 << 30, (1,))) == int(g["torch_next"])


def test_forward_fp32_matches_oracle_seeded(synth_sd):
    import disco_oracle as O
    from disentangledcolorization_b200 import synth
    gray = torch.from_numpy(synth.make_gray(2, 96, 128, seed=42))
    ab = torch.zeros(2, 2, 96, 128)
    np.random.seed(5)
    torch.manual_seed(5)
    with torch.no_grad():
        want = O.forward(synth_sd, gray, ab, 8, 0)
    m = _model(synth_sd, 8, "fp32")
    np.random.seed(5)
    torch.manual_seed(5)
    got = m(gray.cuda(), ab.cuda(), True, 0)
    assert torch.equal(got[5].cpu(), want[5])
    assert (got[2].cpu() - want[2]).abs().max() < FP32_AB_TOL
    assert (got[3].cpu() - want[3]).abs().max() < 1e-4
    # image independence: image 1 alone (same draws) == image 1 inside the batch
    np.random.seed(5)
    torch.manual_seed(5)
    np.random.choice(48, 8, replace=False)
    alone = m(gray[1:].cuda(), ab[1:].cuda(), True, 0)
    assert (alone[2] - got[2][1:]).abs().max() < 1e-5


@pytest.mark.parametrize("case", CASES[:3], ids=lambda c: c["name"])
def test_forward_bf16_within_stated_tolerance(case, synth_sd):
    g = load_golden(case["name"])
    gray, ab = (torch.from_numpy(t) for t in case_inputs(case))
    m = _model(synth_sd, case["K"], "bf16")
    out = m(gray.cuda(), ab.cuda(), True, case["T"], hint_mask=torch.from_numpy(g["hint_mask"]).cuda())
    torch.cuda.synchronize()
    diff = np.abs(out[2].cpu().numpy() - g["pred_colors"])
    print(f"{case['name']}: bf16 max|d ab|={diff.max():.4f} mean={diff.mean():.5f}")
    assert diff.max() < BF16_AB_MAX and diff.mean() < BF16_AB_MEAN
    # anchor agreement when bf16 runs its own k-means.  k-means is a discrete function of the tokens, so ANY bf16-level
    # perturbation (even a different fp32 summation order inside one conv kernel) can move an anchor: on the smooth
    # fixtures every anchor is kept (the bf16-emulating oracle keeps them too, bf16_tolerance.json site_agreement 1.0);
    # on the i.i.d.-noise fixture (16 clusters on 64 tokens) up to two anchors move (measured 0.9375 .. 1.0 across kernel
    # revisions).  Gate: >= 0.9 of the sites; when all anchors agree the colours obey the bf16 tolerance.  The batch-64
    # benchmark inputs are gated in test_gpu_bench_config.py.
    np.random.seed(case["seed"])
    torch.manual_seed(case["seed"])
    own = m(gray.cuda(), ab.cuda(), True, case["T"])
    agree = float((own[5].cpu().numpy() == g["hint_mask"]).mean())
    print(f"{case['name']}: bf16 anchor-site agreement {agree:.3f}")
    assert agree >= 0.9
    if agree == 1.0:
        d_own = np.abs(own[2].cpu().numpy() - g["pred_colors"])
        assert d_own.max() < BF16_AB_MAX and d_own.mean() < BF16_AB_MEAN


def test_forward_diverse_T2_matches_reference_fixture(synth_sd):
    """--diverse: sampled_T=2 expands one image into the T=0/1/2 anchor-colour variants (model.py:148-159)."""
    case = [c for c in golden_cases() if c["name"] == "diverse_64_T2"][0]
    g = load_golden(case["name"])
    gray, ab = (torch.from_numpy(t) for t in case_inputs(case))
    m = _model(synth_sd, case["K"], "fp32")
    np.random.seed(case["seed"])
    torch.manual_seed(case["seed"])
    pal, ref, pred, aff, spix, hint = m(gray.cuda(), ab.cuda(), True, 2)
    assert tuple(pred.shape) == (3, 2, 64, 64) and tuple(ref.shape) == (3, 313, 4, 4) and tuple(aff.shape) == (3, 9, 64, 64)
    assert np.array_equal(hint.cpu().numpy(), g["hint_mask"])
    assert np.abs(spix.cpu().numpy() - g["spix_colors"]).max() < 1e-6, "diverse anchor colours differ"
    assert np.abs(ref.cpu().numpy() - g["ref_logit"]).max() < 1e-2
    assert np.abs(pred.cpu().numpy() - g["pred_colors"]).max() < FP32_AB_TOL
    with pytest.raises(Exception):
        m(torch.zeros(2, 1, 64, 64).cuda(), torch.zeros(2, 2, 64, 64).cuda(), True, 2)   # N must be 1, as in the reference


def test_diverse_batched_extension_equals_single_image_runs(synth_sd):
    """--diverse beyond N = 1.  With model.batched_diverse a batch of N returns 3N variants (variant-major),
    each equal to what the N == 1 path (the only one the reference supports) returns for that image."""
    from disentangledcolorization_b200 import synth
    m = _model(synth_sd, 4, "fp32")
    gray = torch.from_numpy(synth.make_gray(3, 64, 64, seed=61)).cuda()
    ab = torch.zeros(3, 2, 64, 64).cuda()
    with pytest.raises(Exception):
        m(gray, ab, True, 2)                                     # reference behaviour without the flag
    np.random.seed(8)
    torch.manual_seed(8)
    singles = [m(gray[i:i + 1], ab[i:i + 1], True, 2) for i in range(3)]
    m.batched_diverse = True
    np.random.seed(8)
    torch.manual_seed(8)
    out = m(gray, ab, True, 2)
    assert tuple(out[2].shape) == (9, 2, 64, 64) and tuple(out[1].shape) == (9, 313, 4, 4) and tuple(out[5].shape) == (9, 1, 4, 4)
    for n in range(3):
        for v in range(3):
            for j in (1, 2, 4, 5):
                assert torch.equal(out[j][v * 3 + n], singles[n][j][v]), (n, v, j)


def test_random_hint_mode_matches_oracle(synth_sd):
    """--random_hint: anchors from python's random.sample (basic.py:42-47) instead of k-means."""
    import random
    import disco_oracle as O
    from disentangledcolorization_b200 import model, synth
    gray = torch.from_numpy(synth.make_gray(2, 64, 96, seed=21))
    ab = torch.zeros(2, 2, 64, 96)
    m = model.AnchorColorProb(n_clusters=5, enhanced=True, random_hint=True, precision="fp32")
    m.load_state_dict(synth_sd, strict=True)
    m = m.cuda().eval()
    random.seed(3)
    got = m(gray.cuda(), ab.cuda(), True, 0)
    random.seed(3)
    mask = np.zeros((2, 24), np.float32)
    for n in range(2):
        mask[n, random.sample(range(0, 24), random.randint(5, 5))] = 1
    with torch.no_grad():
        want = O.forward(synth_sd, gray, ab, 5, 0, hint_mask=torch.from_numpy(mask.reshape(2, 1, 4, 6)))
    assert np.array_equal(got[5].cpu().numpy(), mask.reshape(2, 1, 4, 6)) and float(got[5].sum()) == 10
    assert (got[2].cpu() - want[2]).abs().max() < FP32_AB_TOL


def test_config4_512_no_resize_k16(synth_sd):
    """BASELINE config 4 shape (512x512, n_clusters=16, S=1024 tokens) at batch 1: fp32 parity with the oracle and a
    bf16 run (exercises the S=1024 attention / global-memory k-means paths and 32x32 token grids)."""
    import disco_oracle as O
    from disentangledcolorization_b200 import synth
    gray = torch.from_numpy(synth.make_gray(1, 512, 512, seed=33))
    ab = torch.zeros(1, 2, 512, 512)
    np.random.seed(9)
    torch.manual_seed(9)
    with torch.no_grad():
        want = O.forward(synth_sd, gray, ab, 16, 0)
    m = _model(synth_sd, 16, "fp32")
    np.random.seed(9)
    torch.manual_seed(9)
    got = m(gray.cuda(), ab.cuda(), True, 0)
    assert torch.equal(got[5].cpu(), want[5])
    assert (got[2].cpu() - want[2]).abs().max() < FP32_AB_TOL
    mb = _model(synth_sd, 16, "bf16")
    outb = mb(gray.cuda(), ab.cuda(), True, 0, hint_mask=want[5].cuda())
    diff = (outb[2].cpu() - want[2]).abs()
    print(f"512x512 bf16: max|d ab|={float(diff.max()):.4f} mean={float(diff.mean()):.5f}")
    assert float(diff.max()) < BF16_AB_MAX and float(diff.mean()) < BF16_AB_MEAN


def test_cuda_graph_mode_matches_eager_and_keeps_rng_protocol(synth_sd):
    """use_cuda_graph replays the same kernels: identical outputs, identical host-RNG consumption, new inputs honoured."""
    from disentangledcolorization_b200 import synth
    m = _model(synth_sd, 8, "bf16")
    grays = [torch.from_numpy(synth.make_gray(2, 64, 96, seed=s)).cuda() for s in (50, 51, 52)]
    ab = torch.zeros(2, 2, 64, 96).cuda()
    np.random.seed(4)
    torch.manual_seed(4)
    eager = [m(g, ab, True, 0) for g in grays]
    np_next, th_next = int(np.random.randint(1 << 30)), int(torch.randint(1 << 30, (1,)))
    m.use_cuda_graph = True
    np.random.seed(4)
    torch.manual_seed(4)
    graphed = [m(g, ab, True, 0) for g in grays]
    assert int(np.random.randint(1 << 30)) == np_next and int(torch.randint(1 << 30, (1,))) == th_next
    for e, g in zip(eager, graphed):
        for a, b in zip(e, g):
            assert torch.equal(a, b)
    assert m.engine().handle.launches() > 0


def test_error_behaviour(synth_sd):
    from disentangledcolorization_b200 import _lib
    m = _model(synth_sd, 8, "fp32")
    with pytest.raises(_lib.DiscoError):      # 250x250: the reference fails too (skip-concat size mismatch)
        m(torch.zeros(1, 1, 250, 250).cuda(), torch.zeros(1, 2, 250, 250).cuda(), True, 0)
    with pytest.raises(_lib.DiscoError):      # more clusters than tokens (np.random.choice raises in the reference)
        m(torch.zeros(1, 1, 32, 32).cuda(), torch.zeros(1, 2, 32, 32).cuda(), True, 0)
    with pytest.raises(_lib.DiscoError):      # training branch not built
        m(torch.zeros(1, 1, 64, 64).cuda(), torch.zeros(1, 2, 64, 64).cuda(), False, 0)


def test_standalone_networks_match_oracle(synth_sd):
    import disco_oracle as O
    from disentangledcolorization_b200 import model, network, synth
    gray = torch.from_numpy(synth.make_gray(1, 64, 64, seed=9))
    seg = model.SpixelSeg()
    seg.load_state_dict({k[len("segnet."):]: v for k, v in synth_sd.items() if k.startswith("segnet.")})
    seg.net.precision = "fp32"
    got = seg.cuda().eval()(gray.cuda())
    with torch.no_grad():
        want = O.spixelnet(synth_sd, gray)
    assert (got.cpu() - want).abs().max() < 1e-5
    rep = network.ColorProbNet(inChannel=1, outChannel=64)
    rep.load_state_dict({k[len("repnet."):]: v for k, v in synth_sd.items() if k.startswith("repnet.")})
    rep.precision = "fp32"
    with torch.no_grad():
        want = O.colorprobnet(synth_sd, gray)
    got = rep.cuda().eval()(gray.cuda())
    assert (got.cpu() - want).abs().max() < 1e-4 * max(1.0, float(want.abs().max()))


def test_pipeline_double_buffered_matches_direct_forward(synth_sd):
    """ColorizePipeline (H2D / forward / D2H of neighbouring steps overlapped on three streams) returns, per step, exactly
    what a direct forward on the same inputs returns, with every step's own inputs (no stale slot reuse)."""
    from disentangledcolorization_b200 import synth
    from disentangledcolorization_b200.pipeline import ColorizePipeline
    m = _model(synth_sd, 8, "bf16")
    B, H, W = 2, 64, 96
    batches = []
    for i in range(5):
        g = torch.from_numpy(synth.make_gray(B, H, W, seed=50 + i)).pin_memory()
        a = torch.zeros(B, 2, H, W).pin_memory()
        batches.append((g, a))
    direct = []
    for g, a in batches:
        np.random.seed(130)
        torch.manual_seed(130)
        direct.append(m(g.cuda(), a.cuda(), True, 0)[2].cpu().clone())
    pipe = ColorizePipeline(m, B, H, W, depth=2)
    got = []

    def reseed():
        np.random.seed(130)
        torch.manual_seed(130)

    # results live in slot buffers that are reused `depth` steps later: consume them step by step
    for i in range(0, len(batches), 2):
        outs = pipe.run(batches[i:i + 2], before_step=reseed)
        got += [o.clone() for o in outs]
    assert len(got) == len(direct)
    for i, (x, y) in enumerate(zip(got, direct)):
        assert torch.equal(x, y), f"step {i}: pipelined result differs from the direct forward"
    assert pipe.h2d_bytes == B * 3 * H * W * 4 and pipe.d2h_bytes == B * 2 * H * W * 4


@pytest.mark.parametrize("H,W", [(80, 112), (48, 208), (16, 16)], ids=str)
def test_bf16_affinity_on_ragged_tile_sizes_matches_oracle(synth_sd, H, W):
    """SpixelNet in bf16 on sizes that leave partial 32 x 16 tiles in the mma.sync decoder-tail kernels (csrc/conv_narrow.cu)
    and partial 128-pixel tiles in the tcgen05 kernels: the affinity map stays within the bf16 storage error of the fp32
    oracle and sums to one.  Bounds as measured for the fused head (test_gpu_ops): max 8e-2, mean 2e-3."""
    import disco_oracle as O
    from disentangledcolorization_b200 import model, synth
    gray = torch.from_numpy(synth.make_gray(3, H, W, seed=H + W))
    seg = model.SpixelSeg()
    seg.load_state_dict({k[len("segnet."):]: v for k, v in synth_sd.items() if k.startswith("segnet.")})
    got = seg.cuda().eval()(gray.cuda()).cpu()                 # precision: bf16 (default)
    with torch.no_grad():
        want = O.spixelnet(synth_sd, gray)
    assert tuple(got.shape) == (3, 9, H, W) and torch.isfinite(got).all()
    d = (got - want).abs()
    assert float(d.max()) < 8e-2 and float(d.mean()) < 2e-3, (float(d.max()), float(d.mean()))
    assert float((got.sum(1) - 1).abs().max()) < 1e-5
