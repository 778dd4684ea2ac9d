#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list.

  python profiles/summarize_launches.py profiles/r01_launches_c2_final.csv
"""
import collections
import csv
import re
import sys


def main():
    rows = []
    with open(sys.argv[1]) as f:
        lines = [l for l in f if l.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = collections.OrderedDict()
    for r in rd:
        if len(r) <= iv:
            continue
        name = re.sub(r"\(.*", "", r[ik])
        v = float(r[iv].replace(",", ""))
        unit = r[iu]
        ms = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v if unit in ("ms", "msecond") else v * 1e3
        t = tot.setdefault(name, [0, 0.0])
        t[0] += 1
        t[1] += ms
    total = sum(t[1] for t in tot.values())
    for name, (n, ms) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"{name:62s} {n:5d} launches {ms:10.3f} ms {100 * ms / total:5.1f}%")
    print(f"\n{sum(t[0] for t in tot.values())} launches, {total:.1f} ms under ncu")


if __name__ == "__main__":
    main()
