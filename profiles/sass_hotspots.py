#!/usr/bin/env python
"""Per-opcode and per-instruction view of an ncu report's SASS page (read here, no GPU needed).

    python profiles/sass_hotspots.py gpurun_out/prof.ncu-rep [top_n]
"""
import collections
import csv
import io
import subprocess
import sys


def main(path, top=25):
    raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[1]
    i_src, i_smp, i_exec = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    per_op, per_op_s = collections.Counter(), collections.Counter()
    body = [r for r in rows[2:] if len(r) > i_exec and r[i_exec].isdigit()]
    tot_e, tot_s = sum(int(r[i_exec]) for r in body), sum(int(r[i_smp]) for r in body)
    for r in body:
        op = r[i_src].split()[0] if not r[i_src].lstrip().startswith("@") else r[i_src].split()[1]
        op = op.split(".")[0]
        per_op[op] += int(r[i_exec]); per_op_s[op] += int(r[i_smp])
    print(f"kernel: {rows[0][1]}\ninstructions executed {tot_e}, stall samples {tot_s}")
    print(f"{'opcode':12s} {'executed':>12s} {'share':>7s} {'samples':>9s} {'share':>7s}")
    for op, e in per_op.most_common(22):
        print(f"{op:12s} {e:12d} {100 * e / tot_e:6.1f}% {per_op_s[op]:9d} {100 * per_op_s[op] / max(tot_s, 1):6.1f}%")
    print(f"\ntop {top} instructions by stall samples (index, samples, executed, main stall reasons, SASS):")
    order = sorted(range(len(body)), key=lambda k: -int(body[k][i_smp]))[:top]
    for k in sorted(order):
        r = body[k]
        st = sorted(((int(r[i]), h[6:]) for i, h in stall_cols if r[i].isdigit() and int(r[i]) > 0), reverse=True)[:3]
        print(f"{k:5d} {int(r[i_smp]):7d} {int(r[i_exec]):9d}  {' '.join(f'{h}:{v}' for v, h in st):42s} {r[i_src].strip()[:90]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
