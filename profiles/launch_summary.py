#!/usr/bin/env python
"""Per-kernel share of an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv`).

    python profiles/launch_summary.py gpurun_out/launches.csv "command line" > profiles/rNN_launches_summary.txt
"""
import collections
import csv
import sys


def main(path, cmd=""):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        if r[im] != "gpu__time_duration.sum":
            continue
        name = r[ik].split("(")[0][:100]
        tot[name] += float(r[iv].replace(",", "")) / 1e3          # ns -> us
        cnt[name] += 1
    all_us = sum(tot.values())
    print(f"ncu --metrics gpu__time_duration.sum --clock-control none, command: {cmd}")
    print("(cold-cache, serialised per-launch times: the kernel's SHARE of the step is what must agree with bench.py)")
    print(f"{'kernel':100s} {'launches':>8s} {'total us':>10s} {'share':>7s} {'avg us':>8s}")
    for name, us in tot.most_common():
        print(f"{name:100s} {cnt[name]:8d} {us:10.1f} {100 * us / all_us:6.1f}% {us / cnt[name]:8.1f}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "")
