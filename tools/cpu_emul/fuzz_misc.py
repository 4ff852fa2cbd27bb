#!/usr/bin/env python
"""Randomised host-emulation check of al_ola_gather (against a float64 numpy overlap-add, plus bitwise shardability through
the halo path) and of the register-blocked al_resample_poly (against scipy.signal.resample_poly).

    python tools/cpu_emul/fuzz_misc.py [--n 40] [--seed 0]
"""
import argparse
import ctypes
import importlib.util
import os
import sys

import numpy as np
from scipy.signal import firwin, resample_poly

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def load():
    spec = importlib.util.spec_from_file_location("build_emul", os.path.join(HERE, "build_emul.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    lib = ctypes.CDLL(mod.build())
    P, LL, I, F = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_float
    lib.emul_ola_gather.argtypes = [P, I, I, I, I, P, P, P, P, LL, LL, LL, P, I, F, F, P, LL, I]
    lib.emul_ola_gather.restype = None
    lib.emul_resample.argtypes = [P, LL, P, LL, I, LL, LL, I, I, P, I, I]
    lib.emul_resample.restype = I
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def ola_case(lib, rs):
    C = int(rs.randint(8, 3000))
    nck = int(rs.randint(1, 9))
    steps = rs.randint(1, C + C // 2 + 1, size=nck - 1) if nck > 1 else np.zeros(0, int)
    offs = np.concatenate(([int(rs.randint(0, 50))], steps)).cumsum().astype(np.int64)
    n = int(offs[-1] + rs.randint(1, C + 1))                       # the last chunk may be cut by n_total
    rows = int(rs.choice([1, 2, 3, 4]))
    n_tab = int(rs.choice([1, 2]))
    wtab = rs.uniform(0.1, 1.0, size=(n_tab, C)).astype(np.float32) if rs.randint(0, 4) else None
    tab_id = rs.randint(0, n_tab, size=nck).astype(np.int32) if (wtab is not None and n_tab > 1) else None
    mult = rs.randint(1, 4, size=nck).astype(np.int32) if rs.randint(0, 2) else None
    chunks = rs.standard_normal((nck, rows, C)).astype(np.float32)
    res = np.zeros((rows, n), np.float64)
    cnt = np.zeros(n, np.float64)
    for c in range(nck):
        L = min(C, n - int(offs[c]))
        w = (wtab[tab_id[c] if tab_id is not None else 0] if wtab is not None else np.ones(C, np.float32))[:L]
        for _ in range(int(mult[c]) if mult is not None else 1):
            res[:, offs[c]:offs[c] + L] += chunks[c][:, :L].astype(np.float64) * w
            cnt[offs[c]:offs[c] + L] += w
    ref = res / np.maximum(cnt, 1e-10)

    def run(ch, of, p0, p1, halo=None, raw=False, dc0=0, mu=None, tid=None):
        track = np.zeros((rows, n), np.float32)
        lib.emul_ola_gather(_p(ch), len(of), dc0, rows, C, _p(np.ascontiguousarray(of)), _p(mu), _p(wtab), _p(tid), n, p0, p1,
                            _p(halo), int(raw), 1e-10, 1.0, _p(track), n, 64)
        return track

    got = run(chunks, offs, 0, n, mu=mult, tid=tab_id)
    err = float(np.abs(got - ref).max() / max(1.0, np.abs(ref).max()))
    # two "ranks" split at a random cut: bitwise equal to the single pass
    cut = int(rs.randint(1, n)) if n > 1 else 0
    k = int((offs < cut).sum())
    ok = True
    if 0 < k < nck and cut > 0:
        sl = lambda a: None if a is None else np.ascontiguousarray(a[:k])
        left = run(np.ascontiguousarray(chunks[:k]), offs[:k], 0, cut, mu=sl(mult), tid=sl(tab_id))
        halo = np.ascontiguousarray(run(np.ascontiguousarray(chunks[:k]), offs[:k], cut, n, raw=True, mu=sl(mult), tid=sl(tab_id))[:, cut:])
        right = run(np.ascontiguousarray(chunks[k:]), offs, cut, n, halo=halo, dc0=k, mu=mult, tid=tab_id)
        ok = np.array_equal(np.concatenate([left[:, :cut], right[:, cut:]], axis=1), got)
    return err, ok, dict(C=C, nck=nck, n=n, rows=rows, wtab=wtab is not None, tab=tab_id is not None, mult=mult is not None, cut=cut)


def res_case(lib, rs):
    up, down = [(147, 160), (160, 147), (147, 160), (3, 4), (5, 4), (80, 147), (441, 480)][int(rs.randint(0, 7))]
    rows = int(rs.choice([1, 2, 3]))
    n_in = int(rs.randint(1, 20000))
    pad = int(rs.choice([0, 1]))
    half = 10 * max(up, down)
    taps = (firwin(2 * half + 1, 1.0 / max(up, down), window=("kaiser", 5.0)) * up).astype(np.float32)
    x = rs.uniform(-1, 1, size=(rows, n_in)).astype(np.float32)
    n_out = (n_in * up + down - 1) // down
    in_stride = (n_in + 3) // 4 * 4 + pad
    out_stride = (n_out + 3) // 4 * 4 + pad
    xin = np.full((rows, in_stride), np.nan, np.float32)
    xin[:, :n_in] = x
    out = np.full((rows, out_stride), -77.0, np.float32)
    rc = lib.emul_resample(_p(xin), in_stride, _p(out), out_stride, rows, n_in, n_out, up, down, _p(taps), taps.size, int(rs.choice([1, 3, 9])))
    desc = dict(up=up, down=down, rows=rows, n_in=n_in, pad=pad, rc=rc)
    if rc < 0:
        return 0.0, True, desc                                     # ratio outside the register-blocked plan: generic kernel
    ref = resample_poly(x, up, down, axis=-1)                      # scipy designs the same Kaiser(5) filter
    ref = ref[:, :n_out]
    err = float(np.abs(out[:, :n_out] - ref).max())
    return err, bool((out[:, n_out:] == -77.0).all()), desc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=40)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    lib = load()
    rs = np.random.RandomState(args.seed)
    bad = 0
    for i in range(args.n):
        err, ok, d = (ola_case if i % 2 == 0 else res_case)(lib, rs)
        fail = err > 3e-6 or not ok
        bad += fail
        print(f"{i:3d} {'ola' if i % 2 == 0 else 'res'} err {err:.2e} exact {ok} {d}{'   <-- FAIL' if fail else ''}", flush=True)
    print("failures", bad)
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
