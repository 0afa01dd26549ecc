"""Per-kernel SASS opcode tally of libdisco_b200.so (run here, on the CPU box: cuobjdump needs no GPU).

    python -m disentangledcolorization_b200.tools.sass_opcodes > profiles/sass_opcodes.md

What the mnemonics prove (B200_PROFILING.md): UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st (tensor memory),
UTMALDG = TMA tensor loads, UTCBAR = tcgen05.commit, HMMA = mma.sync (legacy warp-level tensor path), LDSM = ldmatrix,
LDGSTS = cp.async, FFMA2 = packed fp32 pairs.
"""
import collections
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(HERE, "libdisco_b200.so")
WATCH = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UBLKCP", "HMMA", "LDSM", "LDGSTS", "FFMA2", "FFMA",
         "MUFU.EX2", "SHFL", "BAR", "UCGABAR", "SYNCS"]


def short(name):
    name = re.sub(r"<unnamed>::|\(anonymous namespace\)::|void ", "", name)
    return re.sub(r"\(.*\)$", "", name).strip()


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    per = collections.OrderedDict()
    cur = None
    it = iter(names)
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = short(next(it))
            per[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur is not None:
            op = m.group(1)
            per[cur]["_total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    per[cur][w] += 1
            if op.startswith("UTCHMMA") and ".2CTA" in op:
                per[cur]["UTCHMMA.2CTA"] += 1
    cols = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "UTMALDG", "UTCBAR", "HMMA", "LDSM", "LDGSTS", "FFMA2", "FFMA", "MUFU.EX2", "SHFL"]
    print("# SASS opcode tally per kernel (`cuobjdump -sass disentangledcolorization_b200/libdisco_b200.so`)\n")
    print("UTCHMMA = tcgen05.mma (`.2CTA` = cta_group::2), LDTM = tcgen05.ld, UTMALDG = TMA tensor load, UTCBAR = tcgen05.commit, "
          "HMMA = mma.sync, LDSM = ldmatrix, LDGSTS = cp.async, FFMA2 = packed fp32 pair FMA.  Static instruction counts "
          "(loops execute them many times); regenerate with `python -m disentangledcolorization_b200.tools.sass_opcodes`.\n")
    print("| kernel | instrs | " + " | ".join(cols) + " |")
    print("|---|---|" + "---|" * len(cols))
    tot = collections.Counter()
    for k, c in per.items():
        if not any(c[w] for w in cols):
            continue
        print(f"| `{k}` | {c['_total']} | " + " | ".join(str(c[w]) if c[w] else "" for w in cols) + " |")
        tot.update(c)
    print(f"| **all kernels** | {sum(c['_total'] for c in per.values())} | " + " | ".join(str(tot[w]) for w in cols) + " |")
    print(f"\n{len(per)} kernels in the library; no UTMASTG (epilogues store straight from registers), no HGMMA/QGMMA (Hopper wgmma).")


if __name__ == "__main__":
    sys.exit(main())
