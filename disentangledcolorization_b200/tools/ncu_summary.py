"""Summarises ncu exports for profiles/ (run here, on the CPU box, over files brought back in gpurun_out/).

    python -m disentangledcolorization_b200.tools.ncu_summary launches <launches.csv> <out.md> "<command>"
    python -m disentangledcolorization_b200.tools.ncu_summary full <tag>=<report.ncu-rep> ... --out-prefix profiles/r1b
"""
import csv
import json
import re
import subprocess
import sys


def _short(name):
    name = re.sub(r"<unnamed>::|\(anonymous namespace\)::|void ", "", name)
    name = re.sub(r"\((?:const )?(?:<unnamed>::)?\w*Params\)|\(.*\)$", "", name)
    name = re.sub(r"\(int\)|\(bool\)", "", name)
    return name.strip()


def launches(path, out, command):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = {}
    for r in rows[1:]:
        if r[im] != "gpu__time_duration.sum":
            continue
        a = agg.setdefault(_short(r[ik]), [0, 0.0])
        a[0] += 1
        a[1] += float(r[iv].replace(",", "")) / 1e6
    total = sum(v[1] for v in agg.values())
    n = sum(v[0] for v in agg.values())
    conv = sum(v[1] for k, v in agg.items() if k.startswith(("conv_", "narrow_", "segnet_head")))
    with open(out, "w") as f:
        f.write(f"# ncu launch list of `{command}`\n\n")
        f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | total ms | share |\n|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1]:.3f} | {100 * v[1] / total:.1f} % |\n")
        f.write(f"| **total** | {n} | {total:.3f} | |\n\n")
        f.write(f"Convolution kernels (tcgen05, mma.sync narrow layers, fused SpixelNet head, Cin = 1 layers): {100 * conv / total:.1f} % of the captured device time.\n")
    print(open(out).read())


KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__cluster_size", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum"]


def _to_bytes(val, unit):
    v = float(val.replace(",", ""))
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)


def full(pairs, prefix):
    lines = ["| capture | kernel | ncu duration | tensor pipe active (elapsed) | DRAM read+write | DRAM % of peak | L2 hit | regs | "
             "dyn smem | cluster |", "|---|---|---|---|---|---|---|---|---|---|"]
    traffic = {"note": "dram__bytes_read.sum + dram__bytes_write.sum per launch", "bytes": {}}
    for tag, rep in pairs:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(l for l in raw.splitlines() if l.startswith('"')))
        hdr, units, vals = rows[0], rows[1], rows[2]
        d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
        with open(f"{prefix}_{tag}_ncu_raw.csv", "w") as f:
            w = csv.writer(f)
            keep = [i for i, h in enumerate(hdr) if h in KEYS or h in ("Kernel Name", "Block Size", "Grid Size")]
            for r in rows[:3]:
                w.writerow([r[i] for i in keep])
        g = lambda k: d.get(k, ("", "nan"))
        dram = _to_bytes(g("dram__bytes_read.sum")[1], g("dram__bytes_read.sum")[0]) + \
            _to_bytes(g("dram__bytes_write.sum")[1], g("dram__bytes_write.sum")[0])
        traffic["bytes"][f"{prefix.split('/')[-1]}_{tag}"] = dram
        lines.append(f"| {tag} | `{_short(d['Kernel Name'][1])}` | {g('gpu__time_duration.sum')[1]} {g('gpu__time_duration.sum')[0]} | "
                     f"{float(g('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed')[1]):.1f} % | {dram / 1e6:.1f} MB | "
                     f"{float(g('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed')[1]):.1f} % | "
                     f"{float(g('lts__t_sector_hit_rate.pct')[1]):.1f} % | {g('launch__registers_per_thread')[1]} | "
                     f"{g('launch__shared_mem_per_block_dynamic')[1]} {g('launch__shared_mem_per_block_dynamic')[0]} | "
                     f"{g('launch__cluster_size')[1]} |")
    with open(f"{prefix}_traffic.json", "w") as f:
        json.dump(traffic, f, indent=1)
    print("\n".join(lines))
    return "\n".join(lines)


def kernels(rep, hbm_peak_gbs=None):
    """One markdown row per captured launch of a `--set full` report: duration, DRAM bytes, achieved GB/s (and fraction of
    the measured HBM peak), tensor / SM / L1 / issue utilisation -- the table for the HBM-bound kernels."""
    import os
    if hbm_peak_gbs is None:
        peaks = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..", "MEASURED_PEAKS.json")
        hbm_peak_gbs = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(l for l in raw.splitlines() if l.startswith('"')))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, k, default="nan"):
        return r[col[k]] if k in col and r[col[k]] != "" else default
    out = ["| kernel | grid x block | ncu duration | DRAM read + write | achieved | of measured HBM peak | tensor pipe | SM throughput | "
           "L1/shared throughput | issue slots | regs |", "|---|---|---|---|---|---|---|---|---|---|---|"]
    for r in rows[2:]:
        dur_us = float(val(r, "gpu__time_duration.sum").replace(",", ""))
        if units[col["gpu__time_duration.sum"]] == "ms":
            dur_us *= 1e3
        elif units[col["gpu__time_duration.sum"]] == "ns":
            dur_us /= 1e3
        dram = _to_bytes(val(r, "dram__bytes_read.sum"), units[col["dram__bytes_read.sum"]]) + \
            _to_bytes(val(r, "dram__bytes_write.sum"), units[col["dram__bytes_write.sum"]])
        gbs = dram / dur_us / 1e3
        f = lambda k: f"{float(val(r, k)):.1f} %" if val(r, k) != "nan" else "-"
        out.append(f"| `{_short(val(r, 'Kernel Name'))}` | {val(r, 'Grid Size')} x {val(r, 'Block Size')} | {dur_us:.1f} us | {dram / 1e6:.1f} MB | "
                   f"{gbs:.0f} GB/s | {100 * gbs / hbm_peak_gbs:.1f} % | {f('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed')} | "
                   f"{f('sm__throughput.avg.pct_of_peak_sustained_elapsed')} | {f('l1tex__throughput.avg.pct_of_peak_sustained_elapsed')} | "
                   f"{f('smsp__issue_active.avg.pct_of_peak_sustained_active')} | {val(r, 'launch__registers_per_thread')} |")
    print("\n".join(out))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    elif sys.argv[1] == "kernels":
        kernels(sys.argv[2])
    else:
        prefix = sys.argv[sys.argv.index("--out-prefix") + 1]
        pairs = [a.split("=", 1) for a in sys.argv[2:] if "=" in a and not a.startswith("--")]
        full(pairs, prefix)
