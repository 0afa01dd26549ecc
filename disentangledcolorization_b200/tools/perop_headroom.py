"""Per-op roofline gap from a `bench.py --dump-profile` table: measured launch time against
max(algorithmic HBM bytes / measured copy bandwidth, executed FLOPs / measured burst bf16 peak).

    python -m disentangledcolorization_b200.tools.perop_headroom profiles/r1b_perop_bench.json [out.md]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    d = json.load(open(sys.argv[1]))["per_op"]
    peaks = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
    pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pp):
        peaks.update(json.load(open(pp)))
    B = 64
    rows = []
    for k, v in d.items():
        cin, cout = v["cin"], v["cout"]
        if v["kind"] == "deconv4":
            hw_in = v["flops"] / (2 * B * cin * cout * 16)
            hw_out, inb = hw_in * 4, hw_in * cin * 2
        else:
            hw_out = v["flops"] / (2 * B * cin * cout * 9)
            if v["up2"] and v["n_src"] == 1:
                inb = hw_out / 4 * cin * 2
            elif v["n_src"] == 2:
                inb = hw_out * cin * 2 * 0.6          # one low-resolution + one full-resolution source (approximate)
            else:
                inb = hw_out * v["stride"] ** 2 * cin * 2
        if cin == 1:
            inb = hw_out * 4
        outb = hw_out * cout * (4 if cout in (2, 9) else 2)
        nbytes = B * (inb + outb)
        if "#skip" in k:
            nbytes += B * hw_out * cout * 2               # partial sum read back as residual
        t_hbm = nbytes / (peaks["hbm_gbs"] * 1e9) * 1e3
        t_tc = v["executed_flops"] / (peaks["bf16_tflops"] * 1e12) * 1e3
        rows.append((v["ms"] - max(t_hbm, t_tc), k, v["ms"], t_hbm, t_tc))
    total = sum(r[2] for r in rows)
    gap = sum(r[0] for r in rows)
    lines = [f"# Per-op roofline gap ({os.path.basename(sys.argv[1])}; batch 64, 256x256, CUDA-event launch times)", "",
             f"Bound per op = max(HBM bytes / {peaks['hbm_gbs']:.0f} GB/s, executed FLOPs / {peaks['bf16_tflops']:.0f} TFLOP/s) "
             f"(measured peaks).  Sum of launch times {total:.2f} ms, sum of bounds {total - gap:.2f} ms, gap {gap:.2f} ms.", "",
             "| op | measured ms | HBM bound ms | tensor bound ms | gap ms |", "|---|---|---|---|---|"]
    for r in sorted(rows, reverse=True)[:30]:
        lines.append(f"| {r[1]} | {r[2]:.3f} | {r[3]:.3f} | {r[4]:.3f} | {r[0]:.3f} |")
    text = "\n".join(lines) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
