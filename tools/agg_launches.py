#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name + grid."""
import collections
import csv
import sys


def main(path, top=25):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        if d.get("Metric Unit") == "us":
            v *= 1e3
        elif d.get("Metric Unit") == "ms":
            v *= 1e6
        agg[d["Kernel Name"][:100] + " grid=" + d["Grid Size"]][0] += 1
        agg[d["Kernel Name"][:100] + " grid=" + d["Grid Size"]][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"total {tot / 1e3:.1f} us over {sum(v[0] for v in agg.values())} launches")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1])[:top]:
        print(f"{v[1] / 1e3:10.1f} us  n={v[0]:4d}  avg {v[1] / 1e3 / v[0]:8.1f} us  {100 * v[1] / tot:5.1f}%  {k}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25)
