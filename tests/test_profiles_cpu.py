"""CPU: the committed profile evidence parses with the repo's own summarisers and is self-consistent."""
import json
import os

from conftest import ROOT


def test_launch_list_covers_two_whole_steps(tmp_path):
    from disentangledcolorization_b200.tools import ncu_summary
    out = tmp_path / "summary.md"
    ncu_summary.launches(os.path.join(ROOT, "profiles", "r1b_launches_bench.csv"), str(out), "test")
    text = out.read_text()
    total = [l for l in text.splitlines() if l.startswith("| **total**")][0]
    n_launches = int(total.split("|")[2])
    line = json.load(open(os.path.join(ROOT, "profiles", "r1b_bench_line.json")))
    # bench.py counts the launches of the timed region itself: the ncu window holds exactly two steps of them
    assert n_launches == 2 * line["gpu_launches"] // line["steps"]
    assert "conv_tc_grp_kernel<256, 64, 1, 1>" in text and "attention_kernel" in text


def test_bench_line_has_the_contract_keys():
    line = json.load(open(os.path.join(ROOT, "profiles", "r1b_bench_line.json")))
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "cpu_baseline"):
        assert key in line, key
    assert line["e2e"]["h2d_bytes_per_step"] == 64 * 3 * 256 * 256 * 4 and line["e2e"]["d2h_bytes_per_step"] == 64 * 2 * 256 * 256 * 4
    r = line["roofline"]
    assert r["bound"] == "tensor" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert abs(line["value"] - 64 * 1e3 / line["ms_per_step"]) < 1e-6 * line["value"]


def test_perop_table_matches_the_reference_flop_count():
    d = json.load(open(os.path.join(ROOT, "profiles", "r1b_perop_bench.json")))
    conv_flops = sum(v["flops"] for v in d["per_op"].values())
    # 255.17 GFLOP per image in convolutions (segnet 5.79 + repnet 137.78 + enhanceNet 111.36 = 254.93 conv
    # + linear parts counted elsewhere); the per-op table must reproduce the conv part for a batch of 64
    assert abs(conv_flops / 64 / 1e9 - 254.93) < 0.5
    executed = sum(v["executed_flops"] for v in d["per_op"].values())
    assert executed < conv_flops          # up-sampled layers run as 2x2-tap parity phases


def test_round2_evidence_is_committed_and_consistent():
    """profiles/r2_*: the bench line carries the round-2 keys, the per-op table still reproduces the reference's conv FLOPs
    (the fused SpixelNet head counts its three layers), and the SASS tally shows tcgen05 + TMA + mma.sync."""
    line = json.load(open(os.path.join(ROOT, "profiles", "r2_bench_line.json")))
    for key in ("parity_check", "gpu_eager_baseline", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in line, key
    assert line["parity_check"]["ok"] and line["parity_check"]["vs_cpu_oracle_ok"]
    r = line["roofline"]
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and "whole_step" in r and "config4" in r
    assert line["gpu_eager_baseline"]["fp32_tf32conv"]["speedup_ours_over_it"] > 1
    d = json.load(open(os.path.join(ROOT, "profiles", "r2_perop_bench.json")))
    conv_flops = sum(v["flops"] for v in d["per_op"].values())
    assert abs(conv_flops / 64 / 1e9 - 254.93) < 0.5
    assert "segnet.net.conv0a+conv0b+conv1a" in d["per_op"]
    # the ncu launch list of the same bench command holds exactly two timed steps of launches
    from disentangledcolorization_b200.tools import ncu_summary
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "s.md")
        ncu_summary.launches(os.path.join(ROOT, "profiles", "r2_launches_bench.csv"), out, "test")
        text = open(out).read()
    total = [l for l in text.splitlines() if l.startswith("| **total**")][0]
    assert int(total.split("|")[2]) == 2 * line["gpu_launches"] // line["steps"]
    assert "encoder_stack_kernel" in text and "segnet_head_kernel" in text and "attention_kernel" not in text
    sass = open(os.path.join(ROOT, "profiles", "sass_opcodes.md")).read()
    total = [l for l in sass.splitlines() if l.startswith("| **all kernels**")][0].split("|")
    utchmma, ldtm, utmaldg, hmma = int(total[3]), int(total[5]), int(total[6]), int(total[8])
    assert utchmma > 500 and ldtm > 100 and utmaldg > 500 and hmma > 300
    assert "encoder_stack_kernel" in sass and "segnet_head_kernel" in sass and "poolfeat_partial_mma_kernel" in sass


def test_integration_doc_maps_every_exported_symbol():
    from disentangledcolorization_b200 import _lib
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    assert [s for s in _lib.EXPORTS if s not in text] == []
