"""Hottest SASS instructions of one kernel instance in an `ncu --page source --csv` export, with their main stall reasons.
Usage: ncu_hot_sass.py file.csv[.gz] [kernel-substring] [instance] [top]"""
import csv
import gzip
import sys


def main():
    path = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    inst = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
    f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
    rows = list(csv.reader(f))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and want in r[1]]
    s = starts[inst]
    nxt = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and i > s]
    e = nxt[0] if nxt else len(rows)
    hdr = rows[s + 1]
    body = rows[s + 2:e]
    ix = {h: i for i, h in enumerate(hdr)}
    stall_cols = [h for h in hdr if h.startswith("stall_") and "(Not Issued)" not in h]
    total = sum(int(r[ix["# Samples"]] or 0) for r in body)
    print(f"# {rows[s][1]}  instance {inst}: {len(body)} SASS lines, {total} samples")
    ranked = sorted(enumerate(body), key=lambda kv: -int(kv[1][ix["# Samples"]] or 0))[:top]
    for n, r in ranked:
        smp = int(r[ix["# Samples"]] or 0)
        st = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
        print(f"{n:5d} {100.0 * smp / max(total, 1):5.1f}%  {r[ix['Source']][:70]:70s} " + " ".join(f"{c}={v}" for v, c in st if v))


if __name__ == "__main__":
    main()
