#!/usr/bin/env python
"""Split an `ncu --page source --csv` (SASS view) export of one kernel at its BAR.SYNC instructions and
report, per phase, the stall samples, executed warp instructions and the top opcodes."""
import collections
import csv
import sys


def main(path, which=0):
    rows = list(csv.reader(open(path, errors="replace")))
    kernels, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "body": []}
            kernels.append(cur)
        elif cur is not None and r and r[0].startswith("0x"):
            cur["body"].append(r)
    k = kernels[which]
    print("#", k["name"], f"({len(kernels)} kernels in file, showing #{which})")
    seg, acc, tot = 0, collections.OrderedDict(), 0
    for idx, r in enumerate(k["body"]):
        sass = r[1].strip()
        s, n = int(r[2] or 0), int(r[5] or 0)
        a = acc.setdefault(seg, [0, 0, idx, collections.Counter(), collections.Counter()])
        op = sass.split()[1] if sass.startswith("@") else sass.split()[0]
        a[0] += s
        a[1] += n
        a[3][op] += n
        a[4][op] += s
        tot += s
        if "BAR.SYNC" in sass:
            seg += 1
    for sg, a in acc.items():
        print(f"phase {sg}: sass[{a[2]}..] samples {a[0]} ({100 * a[0] / max(tot, 1):.1f}%) warp-inst {a[1]}")
        print("   by count  :", ", ".join(f"{o}={c}" for o, c in a[3].most_common(8)))
        print("   by samples:", ", ".join(f"{o}={c}" for o, c in a[4].most_common(8)))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
