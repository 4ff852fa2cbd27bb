"""Stall samples of one kernel instance in an `ncu --page source --csv` export, bucketed between synchronisation /
asynchronous-unit instructions (mbarrier waits, tcgen05.ld/st/mma, TMA, fences): which phase of a warp-specialised
kernel the warps spend their time in.  Usage: ncu_sync_split.py file.csv[.gz] [kernel-substring] [instance] [min-pct]"""
import csv
import gzip
import sys

MARKS = ("SYNCS.PHASECHK", "LDTM", "STTM", "UTMASTG", "UTMALDG", "UTCHMMA", "UTCBAR", "BAR.SYNC", "BAR.RED", "DEPBAR",
         "SYNCS.ARRIVE", "FENCE", "EXIT", "USETMAXREG")


def main():
    path = sys.argv[1]
    want = sys.argv[2] if len(sys.argv) > 2 else ""
    inst = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    min_pct = float(sys.argv[4]) if len(sys.argv) > 4 else 0.3
    f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
    rows = list(csv.reader(f))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and want in r[1]]
    s = starts[inst]
    nxt = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and i > s]
    e = nxt[0] if nxt else len(rows)
    hdr = rows[s + 1]
    body = rows[s + 2:e]
    ix = {h: i for i, h in enumerate(hdr)}
    src = [r[ix["Source"]] for r in body]
    smp = [int(r[ix["# Samples"]] or 0) for r in body]
    stall_cols = [h for h in hdr if h.startswith("stall_") and "(Not Issued)" not in h]
    total = sum(smp)
    print(f"# {rows[s][1]} instance {inst}: {len(body)} SASS lines, {total} samples")
    marks = [i for i, x in enumerate(src) if any(k in x for k in MARKS)]
    prev = 0
    for m in marks + [len(src)]:
        seg = sum(smp[prev:m])
        if seg > min_pct / 100.0 * total:
            st = {}
            for r in body[prev:m]:
                for c in stall_cols:
                    st[c[6:]] = st.get(c[6:], 0) + int(r[ix[c]] or 0)
            top = sorted(st.items(), key=lambda kv: -kv[1])[:3]
            nxt_src = src[m].strip()[:60] if m < len(src) else "(end)"
            print(f"[{prev:4d},{m:4d}) {100.0 * seg / total:5.1f}%  " + " ".join(f"{k}={v}" for k, v in top if v) + f"   -> {nxt_src}")
        prev = m


if __name__ == "__main__":
    main()
