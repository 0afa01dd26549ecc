"""GPU: parity of the BENCHMARKED configuration (BASELINE config 2: batch 64, 256x256, bf16, n_clusters=8, the
tensor-core path running its OWN k-means), of the sharded multi-GPU layout (config 3) and of the
standalone modules that round 1 left untested ("Next round" item 1).

The oracle (fp32 torch CPU restatement of the reference) needs ~15 s for the 64 images on the GPU box's 16 cores.
Tolerances come from tests/golden/bf16_tolerance.json (oracle/derive_bf16_tolerance.py: fp32 oracle vs the same oracle
with bf16-rounded stored activations; gate = 1.5 x measured).
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, GOLDEN

sys.path.insert(0, os.path.join(ROOT, "oracle"))
pytestmark = pytest.mark.gpu

with open(os.path.join(GOLDEN, "bf16_tolerance.json")) as _f:
    TOL = json.load(_f)
BF16_AB_MAX, BF16_AB_MEAN = TOL["BF16_AB_MAX"], TOL["BF16_AB_MEAN"]
# Anchor agreement of the bf16 path running its own k-means on bf16 conv features, against the fp32 oracle.  k-means is a
# discrete function of the tokens: on the smooth benchmark images most clusterings sit near a decision boundary, and the
# torch-vs-torch baseline itself (bf16_tolerance.json, row bench_c2_first8) keeps all 8 anchors on only 2 of 8 images
# while 98.8 % of the 256 sites agree (~1.5 of 8 anchors move per image).  The floors below are that baseline / the B200
# measurement (printed by the test) minus a margin; moved anchors change colours by an anchor flip, which
# is why the |d ab| gate with identical anchors is the parity statement and this is the agreement statement.
# Measured on B200 (round 2, 64 benchmark images): 26/64 images with all 8 anchors identical, site agreement 0.9927,
# anchor overlap 0.883 (on average 7.1 of 8 anchors are the reference's).
MIN_IMAGES_SAME_ANCHORS = 0.25
MIN_SITE_AGREEMENT = 0.985
MIN_ANCHOR_OVERLAP = 0.80


def _model(sd, K, precision):
    from disentangledcolorization_b200 import model
    m = model.AnchorColorProb(n_clusters=K, enhanced=True, precision=precision)
    m.load_state_dict(sd, strict=True)
    return m.cuda().eval()


@pytest.fixture(scope="module")
def bench_case(synth_sd):
    """The exact inputs bench.py times (rank 0): make_gray(64, 256, 256, seed=100), ab = 0, seeds 130, and the fp32
    oracle's outputs for them."""
    import disco_oracle as O
    from disentangledcolorization_b200 import synth
    torch.set_flush_denormal(True)
    gray = torch.from_numpy(synth.make_gray(64, 256, 256, seed=100))
    ab = torch.zeros(64, 2, 256, 256)
    np.random.seed(130)
    torch.manual_seed(130)
    outs = []
    with torch.no_grad():
        for i in range(0, 64, 8):            # same RNG stream as one batch of 64: draws are consumed per image, in order
            outs.append(O.forward(synth_sd, gray[i:i + 8], ab[i:i + 8], 8, 0))
    want = tuple(torch.cat([o[j] for o in outs]) for j in range(6))
    return gray, ab, want


def test_bench_config_bf16_with_oracle_anchors(bench_case, synth_sd):
    gray, ab, want = bench_case
    m = _model(synth_sd, 8, "bf16")
    out = m(gray.cuda(), ab.cuda(), True, 0, hint_mask=want[5].cuda())
    torch.cuda.synchronize()
    assert all(torch.isfinite(t).all() for t in out[:5])
    d = (out[2].cpu() - want[2]).abs()
    print(f"B=64 256x256 bf16, oracle anchors: max|d ab|={float(d.max()):.4f} mean={float(d.mean()):.5f} "
          f"(gate {BF16_AB_MAX} / {BF16_AB_MEAN})")
    assert float(d.max()) < BF16_AB_MAX and float(d.mean()) < BF16_AB_MEAN
    # affinity = softmax over 9 after 19 bf16-stored layers: probabilities move by a few 1e-2 at most (measured 0.039)
    assert (out[3].cpu() - want[3]).abs().max() < 8e-2
    assert (out[3].cpu() - want[3]).abs().mean() < 2e-3


def test_bench_config_bf16_own_kmeans(bench_case, synth_sd):
    """The path bench.py times: bf16 features -> fp32 tokens -> own k-means -> own anchors."""
    gray, ab, want = bench_case
    m = _model(synth_sd, 8, "bf16")
    np.random.seed(130)
    torch.manual_seed(130)
    out = m(gray.cuda(), ab.cuda(), True, 0)
    torch.cuda.synchronize()
    assert all(torch.isfinite(t).all() for t in out[:5])
    hint, hint_ref = out[5].cpu(), want[5]
    assert float(hint.sum()) == float(hint_ref.sum()) == 64 * 8
    same = (hint == hint_ref).flatten(1).all(1)
    site = float((hint == hint_ref).float().mean())
    overlap = float(torch.minimum(hint, hint_ref).flatten(1).sum(1).mean()) / 8.0     # shared anchors / K, mean over images
    d = (out[2].cpu() - want[2]).abs()
    d_same = d[same]
    print(f"B=64 256x256 bf16, own k-means: {int(same.sum())}/64 images with identical anchors, site agreement {site:.4f}, "
          f"anchor overlap {overlap:.3f}; max|d ab| (same-anchor images)="
          f"{float(d_same.max()) if same.any() else float('nan'):.4f}, all images: max={float(d.max()):.4f} "
          f"mean={float(d.mean()):.5f}")
    assert float(same.float().mean()) >= MIN_IMAGES_SAME_ANCHORS
    assert site >= MIN_SITE_AGREEMENT and overlap >= MIN_ANCHOR_OVERLAP
    # images whose anchors agree obey the bf16 tolerance; the rest differ by anchor flips
    if same.any():
        assert float(d_same.max()) < BF16_AB_MAX and float(d_same.mean()) < BF16_AB_MEAN
    assert float(d.mean()) < 2 * BF16_AB_MEAN
    # host RNG protocol: the numpy stream advanced exactly as the reference's (64 x np.random.choice)
    np.random.seed(130)
    for _ in range(64):
        np.random.choice(256, 8, replace=False)
    expect = int(np.random.randint(1 << 30))
    np.random.seed(130)
    m(gray.cuda(), ab.cuda(), True, 0)
    assert int(np.random.randint(1 << 30)) == expect


def test_bench_config_fp32_exact(bench_case, synth_sd):
    """fp32 path on a slice of the benchmarked batch: the north-star gate |d ab| <= 1e-3 with identical anchors."""
    gray, ab, want = bench_case
    m = _model(synth_sd, 8, "fp32")
    np.random.seed(130)
    torch.manual_seed(130)
    out = m(gray[:8].cuda(), ab[:8].cuda(), True, 0)
    assert torch.equal(out[5].cpu(), want[5][:8])
    assert float((out[2].cpu() - want[2][:8]).abs().max()) < 1e-3


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_shards_equal_single_batch_on_one_gpu(synth_sd, precision):
    """Config 3's layout without the collective: rank r runs images [r*b, (r+1)*b) with the k-means init rows
    `sharded_init_draws` gives it; the concatenation is bit-identical to the single-process batch."""
    from disentangledcolorization_b200 import dist as ddist, synth
    B, H, W, K, world = 6, 64, 96, 8, 3
    S = (H // 16) * (W // 16)
    gray = torch.from_numpy(synth.make_gray(B, H, W, seed=77)).cuda()
    ab = torch.zeros(B, 2, H, W).cuda()
    m = _model(synth_sd, K, precision)
    np.random.seed(130)
    torch.manual_seed(130)
    whole = m(gray, ab, True, 0)
    parts = []
    for r in range(world):
        np.random.seed(130)
        torch.manual_seed(130)
        idx = ddist.sharded_init_draws(B, S, K, world, r)
        lo, hi = ddist.shard_bounds(B, world, r)
        parts.append(m(gray[lo:hi], ab[lo:hi], True, 0, init_idx=idx))
    for j in (2, 3, 5):
        cat = torch.cat([p[j] for p in parts])
        assert torch.equal(cat, whole[j]), f"output {j}: sharded != single batch"


_TWO_RANK = r"""
import os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, {root!r})
from disentangledcolorization_b200 import model, synth, dist as ddist
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
B, H, W, K = 8, 128, 128, 8
S = (H // 16) * (W // 16)
sd = synth.make_state_dict(seed=0)
m = model.AnchorColorProb(n_clusters=K, enhanced=True, precision="bf16")
m.load_state_dict(sd, strict=True)
m = m.cuda().eval()
gray = torch.from_numpy(synth.make_gray(B, H, W, seed=5)).cuda()
ab = torch.zeros(B, 2, H, W).cuda()
lo, hi = ddist.shard_bounds(B, world, rank)
np.random.seed(130); torch.manual_seed(130)
idx = ddist.sharded_init_draws(B, S, K, world, rank)
out = m(gray[lo:hi], ab[lo:hi], True, 0, init_idx=idx)
gathered = ddist.gather_outputs(out[2], B)
hint = ddist.gather_outputs(out[5], B)
ok = True
if rank == 0:
    np.random.seed(130); torch.manual_seed(130)
    whole = m(gray, ab, True, 0)
    ok = bool(torch.equal(gathered, whole[2]) and torch.equal(hint, whole[5]))
    print(json.dumps({{"ok": ok, "shape": list(gathered.shape)}}))
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run with gpurun --gpus 2)")
def test_two_rank_nccl_allgather_equals_single_gpu_batch(tmp_path):
    """Config 3 on hardware: 2 ranks over NCCL, shards + sharded_init_draws + gather_outputs == the single-GPU batch,
    bit-exact (pred_colors and hint_mask)."""
    script = tmp_path / "two_rank.py"
    script.write_text(_TWO_RANK.format(root=ROOT))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert '"ok": true' in r.stdout


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_model_on_second_device_with_other_current_device(synth_sd):
    """one process, two devices -- per-device kernel attributes / error flag, and launches that follow the
    model's device rather than the caller's current device."""
    from disentangledcolorization_b200 import model, synth
    gray = torch.from_numpy(synth.make_gray(1, 64, 64, seed=3))
    ab = torch.zeros(1, 2, 64, 64)
    outs = []
    for dev in (0, 1):
        m = model.AnchorColorProb(n_clusters=8, enhanced=True, precision="bf16")
        m.load_state_dict(synth_sd, strict=True)
        m = m.to(f"cuda:{dev}").eval()
        torch.cuda.set_device(0)                       # current device stays 0 for both models
        np.random.seed(1)
        torch.manual_seed(1)
        outs.append(m(gray.to(f"cuda:{dev}"), ab.to(f"cuda:{dev}"), True, 0))
        assert outs[-1][2].device.index == dev
    torch.cuda.synchronize(0)
    torch.cuda.synchronize(1)
    assert torch.equal(outs[0][2].cpu(), outs[1][2].cpu())


def test_hourglass2_standalone_returns_pre_tanh(synth_sd):
    """`HourGlass2.forward` standalone = the reference's pre-tanh map (models/network.py:136-144), unbounded values."""
    import torch.nn as nn
    import disco_oracle as O
    from disentangledcolorization_b200 import network
    g = torch.Generator().manual_seed(11)
    x = torch.cat([torch.rand(2, 1, 64, 96, generator=g) * 2 - 1, torch.rand(2, 64, 64, 96, generator=g) * 4.0], 1)
    hg = network.HourGlass2(inChannel=65, outChannel=2, resNum=3, normLayer=nn.BatchNorm2d)
    hg.load_state_dict({k[len("enhanceNet."):]: v for k, v in synth_sd.items() if k.startswith("enhanceNet.")})
    hg.precision = "fp32"
    got = hg.cuda().eval()(x.cuda())
    with torch.no_grad():
        want = O.hourglass2(synth_sd, x)
    assert tuple(got.shape) == (2, 2, 64, 96)
    scale = max(1.0, float(want.abs().max()))
    assert float((got.cpu() - want).abs().max()) < 1e-4 * scale
    # a saturating input: the pre-tanh map exceeds atanh's range, which the round-1 atanh(tanh(.)) detour clipped at 8.3
    big = x.clone()
    big[:, 1:] *= 50.0
    got_big = hg(big.cuda())
    with torch.no_grad():
        want_big = O.hourglass2(synth_sd, big)
    assert float((got_big.cpu() - want_big).abs().max()) < 1e-4 * max(1.0, float(want_big.abs().max()))
    hg.precision = "bf16"
    hg._invalidate()
    got16 = hg(x.cuda())
    assert float((got16.cpu() - want).abs().max()) < 0.1 * scale


def test_workspace_lru_bounds_memory(synth_sd):
    """the engine keeps at most `max_workspaces` activation workspaces (the CLI's --no_resize mode feeds a new
    size per image) and results stay correct after evictions."""
    from disentangledcolorization_b200 import synth
    m = _model(synth_sd, 4, "bf16")
    eng = m.engine(torch.device("cuda", torch.cuda.current_device()))
    eng.max_workspaces = 3
    first = None
    sizes = [(64, 64), (64, 80), (80, 64), (96, 64), (64, 96), (64, 64)]
    for (H, W) in sizes:
        gray = torch.from_numpy(synth.make_gray(1, H, W, seed=8)).cuda()
        np.random.seed(2)
        torch.manual_seed(2)
        out = m(gray, torch.zeros(1, 2, H, W).cuda(), True, 0)
        assert torch.isfinite(out[2]).all()
        if first is None:
            first = out[2].clone()
        assert len(eng._ws) <= 3
    assert torch.equal(out[2], first)          # (64, 64) again, after its workspace had been evicted and rebuilt
