#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: share of the listed time per kernel."""
import collections
import csv
import re
import sys


def main(path):
    rows = list(csv.reader(open(path, errors="replace")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hi]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot, n, unit = 0.0, 0, ""
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        unit = r[ui]
        name = re.sub(r"<.*", "", r[ki])[:90]
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
        n += 1
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1.0)
    print(f"# {n} launches, {tot * scale / 1e3:.2f} ms of kernel time listed (cold-cache, serialised under ncu)")
    print(f"# {'share':>6s} {'count':>6s} {'avg_us':>10s}  kernel")
    for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"  {100 * v / tot:5.1f}% {c:6d} {v * scale / c:10.1f}  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
