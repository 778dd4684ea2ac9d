#!/usr/bin/env python
"""Top instructions of an `ncu --page source --csv` dump by warp-stall samples, with the dominant stall reason.
  python profiles/read_source_page.py gpurun_out/x_source.csv [top]"""
import csv
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    items, tot = [], 0.0
    for n, r in enumerate(rows[2:]):
        try:
            v = float(r[ci["# Samples"]])
        except Exception:
            continue
        tot += v
        items.append((v, n, r))
    by_stall = {s: sum(float(r[ci[s]] or 0) for _, _, r in items) for s in stalls}
    print(f"total samples {tot:.0f}; by reason: " + ", ".join(f"{k[6:]} {v / tot * 100:.1f}%" for k, v in sorted(by_stall.items(), key=lambda kv: -kv[1])[:8]))
    for v, n, r in sorted(items, key=lambda x: -x[0])[:top]:
        reason = max(stalls, key=lambda s: float(r[ci[s]] or 0))
        print(f"{v / tot * 100:5.1f}%  #{n:5d}  exec {r[ci['Instructions Executed']]:>9}  {reason[6:]:<14} {r[ci['Source']].strip()[:80]}")


if __name__ == "__main__":
    main()
