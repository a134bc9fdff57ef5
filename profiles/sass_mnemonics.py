#!/usr/bin/env python
"""SASS mnemonic counts per kernel of libafp_b200.so (cuobjdump -sass), the evidence table kept under profiles/.

    python profiles/sass_mnemonics.py > profiles/r02/sass_mnemonics.txt

Kernel families are summarised as one row per family (instantiations counted, mnemonic counts of the largest
instantiation) followed by the full per-instantiation table of the kernels the benchmark lines run on.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "audio_filtering_pipeline_b200", "libafp_b200.so")
COLS = ["UTCHMMA", "UTCBAR", "LDTM", "UTCATOMSWS", "UBLKCP", "UBLKPF", "SYNCS", "FFMA2", "FFMA", "DADD", "MUFU", "LDCU", "LDS", "STS",
        "LDG", "STG", "LDGSTS", "SHFL", "BAR"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, cur = collections.OrderedDict(), None
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            counts[cur][m.group(1).split(".")[0]] += 1
    names = demangle(list(counts))
    short = {k: re.sub(r"\(.*$", "", names[k]).replace("void ", "").replace("afp::", "").replace("(anonymous namespace)::", "")
             for k in counts}
    fam = collections.defaultdict(list)
    for k in counts:
        fam[re.sub(r"<.*$", "", short[k])].append(k)
    print("SASS mnemonic counts per kernel in libafp_b200.so (cuobjdump -sass, sm_100a).")
    print("UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc/dealloc, UBLKCP = cp.async.bulk")
    print("(TMA bulk copy), UBLKPF = cp.async.bulk.prefetch.L2, SYNCS = mbarrier ops, FFMA2 = fma.rn.f32x2, LDGSTS = cp.async.\n")
    hdr = f"{'kernel family':28s} {'inst.':>5s} " + " ".join(f"{c:>7s}" for c in COLS)
    print(hdr)
    for f, ks in sorted(fam.items()):
        big = max(ks, key=lambda k: sum(counts[k].values()))
        print(f"{f:28s} {len(ks):5d} " + " ".join(f"{counts[big][c]:7d}" for c in COLS))
    print("\nper instantiation (kernels behind the bench lines: tensor, single-row, AGC, spectrum):")
    print(f"{'kernel':60s} " + " ".join(f"{c:>7s}" for c in COLS))
    for k in counts:
        if re.match(r"(fir_tc2_kernel|agc_kernel|spectrum_db_kernel|fir_row_kernel<[124], (16|8|4), [01], [02], 0,)", short[k]):
            print(f"{short[k][:60]:60s} " + " ".join(f"{counts[k][c]:7d}" for c in COLS))


if __name__ == "__main__":
    sys.exit(main())
