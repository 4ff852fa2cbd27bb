#!/usr/bin/env python
"""Randomised host-emulation check of al_stft (generic and packed kernels) against torch.stft on the materialised chunk.

    python tools/cpu_emul/fuzz_stft.py [--n 40] [--seed 0]

Draws (n_fft, hop, frames, chunk placement incl. negative first offsets and tails past the track, centre pad, layout,
cropped / zeroed bins, persistent-grid size).  Test infrastructure."""
import argparse
import ctypes
import importlib.util
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
_spec = importlib.util.spec_from_file_location("tke", os.path.join(ROOT, "tests", "test_kernel_emulation.py"))
_tke = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_tke)
_bind_fft, _p, _plan_tables = _tke._bind_fft, _tke._p, _tke._plan_tables


def load():
    spec = importlib.util.spec_from_file_location("build_emul", os.path.join(HERE, "build_emul.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    lib = ctypes.CDLL(mod.build())
    _bind_fft(lib)
    P, LL, I = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int
    lib.emul_stft_pk.argtypes = [P, LL, LL, LL, LL, I, I, I, I, I, P, P, P, P, I, I, I, I, I]
    lib.emul_stft_pk.restype = I
    return lib


def one(lib, rs, packed):
    if packed:
        n_fft, hop = 2048, int(rs.choice([441, 512, 300, 1024, 777]))
        layout = int(rs.choice([0, 3]))
    else:
        n_fft = int(rs.choice([2048, 4096, 6144]))
        hop = int(rs.choice([n_fft // 4, n_fft // 6, 441, 1024, n_fft // 2, 1000]))
        layout = int(rs.choice([0, 1, 2, 3]))
    F = n_fft // 2 + 1
    T = int(rs.randint(1, 24))
    center = int(rs.choice([n_fft // 2, n_fft // 2, 0, 1536 if n_fft >= 4096 else 700]))
    # frames must stay within a single reflection of the chunk
    need = (T - 1) * hop - center + n_fft - 1
    chunk_len = int(max(center + 1, (need + 2) // 2 + 1) + rs.randint(0, 2000))
    n_chunks = int(rs.choice([1, 2, 3]))
    step = int(rs.randint(1, chunk_len + 1))
    off0 = int(rs.choice([0, 0, -rs.randint(1, n_fft), rs.randint(0, 1000)]))
    n = int(max(1, off0 + (n_chunks - 1) * step + chunk_len - rs.randint(0, chunk_len // 2 + 1)))
    crop = int(rs.choice([F, F - 1, F // 2, 1000])) if not packed else int(rs.choice([F, 1000, 1024]))
    crop = min(crop, F)
    low = int(rs.choice([0, 0, 3]))
    x = np.zeros((2, n + 8), np.float32)
    x[:, :n] = rs.uniform(-1, 1, size=(2, n))
    x[:, n:] = np.nan
    wa, _, tw, ctw, _ = _plan_tables(n_fft, hop)
    desc = dict(packed=packed, n_fft=n_fft, hop=hop, T=T, center=center, chunk_len=chunk_len, n_chunks=n_chunks, step=step, off0=off0,
                n=n, crop=crop, low=low, layout=layout)
    if layout == 2:
        spec = np.full((n_chunks, 4, crop, T), np.nan, np.float32)
    else:
        spec = np.full(n_chunks * 2 * T * crop * 2, np.nan, np.float32)
    if packed:
        half = np.zeros((544, 2), np.float32)
        k = np.arange(513)
        half[:513, 0], half[:513, 1] = 0.5 * np.cos(-2 * np.pi * k / n_fft), 0.5 * np.sin(-2 * np.pi * k / n_fft)
        grid = int(rs.choice([1, 2, 7]))
        aligned = int(x.ctypes.data % 16 == 0 and x.shape[1] % 4 == 0)
        desc.update(grid=grid, aligned=aligned)
        rc = lib.emul_stft_pk(_p(x), n, x.shape[1], off0, step, n_chunks, chunk_len, center, hop, T, _p(wa), _p(tw), _p(half),
                              _p(spec), layout, crop, low, aligned, grid)
        assert rc >= 1, (rc, desc)
    else:
        rc = lib.emul_stft(n_fft, hop, _p(x), n, x.shape[1], 2, off0, step, n_chunks, chunk_len, center, T, _p(wa), _p(tw), _p(ctw),
                           _p(spec), layout, crop, low)
        assert rc == 0, (rc, desc)
    if not np.isfinite(spec).all():
        return float("inf"), desc
    err = 0.0
    win = torch.hann_window(n_fft)
    for c in range(n_chunks):
        a = off0 + c * step
        idx = np.arange(a, a + chunk_len)
        ok = (idx >= 0) & (idx < n)
        chunk = np.zeros((2, chunk_len), np.float32)
        chunk[:, ok] = x[:, idx[ok]]
        # the kernel reflects `center` samples about the chunk ends (torch.stft center=True does n_fft/2): pad by hand
        j = np.arange(-center, (T - 1) * hop - center + n_fft)
        j = np.where(j < 0, -j, j)
        j = np.where(j >= chunk_len, 2 * (chunk_len - 1) - j, j)
        padded = torch.tensor(chunk[:, j])
        ref = torch.stft(padded, n_fft, hop, window=win, center=False, return_complex=True)        # [2, F, T]
        ref = torch.view_as_real(ref).numpy()[:, :crop].copy()
        ref[:, :low] = 0
        if layout == 0:
            got = spec.reshape(n_chunks, 2, T, crop, 2)[c].transpose(0, 2, 1, 3)
        elif layout == 1:
            got = spec.reshape(n_chunks, 2, crop, T, 2)[c]
        elif layout == 2:
            got = spec[c].reshape(2, 2, crop, T).transpose(0, 2, 3, 1)
        else:
            got = spec.reshape(n_chunks, T, crop, 2, 2)[c].transpose(2, 1, 0, 3)
        err = max(err, float(np.abs(got - ref).max()) / max(1e-3, float(np.abs(ref).max())))
    return err, desc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=40)
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args()
    lib = load()
    rs = np.random.RandomState(args.seed)
    worst = 0.0
    for i in range(args.n):
        err, desc = one(lib, rs, bool(i % 2))
        worst = max(worst, err)
        flag = "" if err <= 5e-6 else "   <-- FAIL"
        print(f"{i:3d} err {err:.2e} {desc}{flag}", flush=True)
    print("worst", worst)
    sys.exit(0 if worst <= 5e-6 else 1)


if __name__ == "__main__":
    main()
