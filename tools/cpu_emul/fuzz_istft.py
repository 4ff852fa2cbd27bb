#!/usr/bin/env python
"""Randomised host-emulation check of al_istft (generic and packed kernels) against torch.istft.

    python tools/cpu_emul/fuzz_istft.py [--n 40] [--seed 0]

Draws (n_fft, hop, frames, frame_pad, cropped bins, out_start / out_len windows, stems, mask, weight, the packed kernel
(istft_pk2 / istft_pk4 / istft_pk5), warps per CTA, SM count for the tiling) and compares every emitted sample with a torch.istft of the same (padded) spectrum.  Test
infrastructure (like tests/test_kernel_emulation.py, which holds the fixed cases)."""
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
    lib.emul_istft_pk.argtypes = [P, P, I, I, I, I, P, P, P, P, I, I, P, P, LL, LL, LL, LL, LL, I, I, I, I, I]
    lib.emul_istft_pk.restype = I
    lib.emul_istft_pk4.argtypes = [P, P, I, I, I, I, P, P, P, P, I, I, P, P, LL, LL, LL, LL, LL, I, I, I]
    lib.emul_istft_pk4.restype = I
    return lib


def one(lib, rs, packed):
    if packed:
        n_fft = 2048
        hop = int(rs.choice([441, 512, 256, 1024, 300, 777]))
        kernel = int(rs.choice([2, 4, 5]))        # 2: istft_pk2_kernel (register pipeline), 4 / 5: the streaming kernels
        if kernel == 2 and (n_fft + hop - 1) // hop > 5:
            kernel = 4                            # the register-form overlap-add of istft_pk2 needs <= 5 frames per position
    else:
        n_fft = int(rs.choice([2048, 4096, 6144]))
        hop = int(rs.choice([n_fft // 4, n_fft // 6, 441, 1024, n_fft // 2, 1000]))
    F = n_fft // 2 + 1
    T = int(rs.randint(1, 40))
    frame_pad = 0 if packed else int(rs.choice([0, 0, 1, 2]))
    Tt = T + 2 * frame_pad
    ola_len = (Tt - 1) * hop + n_fft
    Fo = F if packed else int(rs.choice([F, F - 1, F // 2 + 3]))
    out_start = int(rs.randint(0, min(ola_len - 1, n_fft)))
    out_len = int(rs.randint(1, ola_len - out_start + 1))
    stems = int(rs.choice([1, 1, 2]))
    use_mask = bool(rs.randint(0, 2))
    use_w = bool(rs.randint(0, 2))
    cplx = lambda *s: (rs.standard_normal(s) + 1j * rs.standard_normal(s)).astype(np.complex64)
    _, ws, tw, ctw, env = _plan_tables(n_fft, hop, Tt)
    weight = rs.uniform(0.5, 1.5, out_len).astype(np.float32) if use_w else None
    dst = np.full((stems, 2, out_len), np.nan, np.float32)
    win = torch.hann_window(n_fft)
    desc = dict(packed=packed, n_fft=n_fft, hop=hop, T=T, frame_pad=frame_pad, Fo=Fo, out_start=out_start, out_len=out_len,
                stems=stems, mask=use_mask, weight=use_w)
    if packed:
        spec = cplx(1, T, F, 2)
        mask = cplx(1, stems, T, F, 2) if use_mask else None
        k = np.arange(1024)
        ctw_full = np.ascontiguousarray(np.stack((np.cos(-2 * np.pi * k / n_fft), np.sin(-2 * np.pi * k / n_fft)), -1).astype(np.float32))
        warps, n_sm, fast = int(rs.choice([4, 8])), int(rs.choice([1, 2, 5])), int(rs.randint(0, 2))
        desc.update(warps=warps, n_sm=n_sm, fast=fast, kernel=kernel)
        if kernel == 2:
            rc = lib.emul_istft_pk(_p(spec), _p(mask), T, stems, 0, hop, _p(ws), _p(tw), _p(ctw_full), _p(env), out_start, out_len,
                                   _p(weight), _p(dst), out_len, stems * 2 * out_len, 0, 0, out_len, 1, warps, n_sm, 0, fast)
        else:
            rc = lib.emul_istft_pk4(_p(spec), _p(mask), T, stems, 0, hop, _p(ws), _p(tw), _p(ctw_full), _p(env), out_start, out_len,
                                    _p(weight), _p(dst), out_len, stems * 2 * out_len, 0, 0, out_len, 1, kernel, n_sm)
        full = [spec[0] * (mask[0, s] if use_mask else 1.0) for s in range(stems)]          # [t, f, ch]
        full = [y.transpose(2, 1, 0) for y in full]                                          # [ch, f, t]
    else:
        layout = int(rs.choice([0, 1])) if use_mask else int(rs.choice([0, 1, 2]))
        desc.update(layout=layout)
        S = cplx(2, Fo, T)                                                                   # [ch, f, t]
        M = cplx(stems, 2, F, T) if use_mask else None
        if layout == 0:
            spec = np.ascontiguousarray(S.transpose(0, 2, 1))
            mask = None if M is None else np.ascontiguousarray(M.transpose(0, 1, 3, 2))
        elif layout == 1:
            spec, mask = np.ascontiguousarray(S), M
        else:
            spec, mask = np.ascontiguousarray(np.stack((S.real, S.imag), axis=1)).astype(np.float32), None
        rc = lib.emul_istft(n_fft, hop, _p(spec), _p(mask), layout, Fo, T, frame_pad, 1, stems, 2, 0, 0, _p(ws), _p(tw), _p(ctw),
                            _p(env), out_start, out_len, _p(weight), _p(dst), out_len, stems * 2 * out_len, 0, 0, out_len)
        full = []
        for s in range(stems):
            y = np.zeros((2, F, T), np.complex64)
            y[:, :Fo] = S
            if use_mask:
                y = y * M[s]
                y[:, Fo:] = 0
            full.append(y)
    assert rc >= 1, (rc, desc)
    err = 0.0
    for s in range(stems):
        y = np.zeros((2, F, Tt), np.complex64)
        y[:, :, frame_pad:frame_pad + T] = full[s]
        # torch.istft(center=True) trims n_fft/2 at both ends: rebuild the untrimmed overlap-add to compare any window
        frames = torch.fft.irfft(torch.tensor(y), n=n_fft, dim=1) * win[None, :, None]      # [ch, n_fft, Tt]
        ola = torch.zeros((2, ola_len), dtype=torch.float64)
        envd = torch.zeros(ola_len, dtype=torch.float64)
        for t in range(Tt):
            ola[:, t * hop:t * hop + n_fft] += frames[:, :, t].double()
            envd[t * hop:t * hop + n_fft] += win.double() ** 2
        ref = torch.where(envd > 1e-11, ola / envd.clamp(min=1e-30), torch.zeros_like(ola))
        ref = ref[:, out_start:out_start + out_len].float().numpy()
        if use_w:
            ref = ref * weight
        if not np.isfinite(dst[s]).all():
            return float("inf"), desc
        # near the untrimmed ends sum(w^2) -> 0 and 1/envelope amplifies fp32 rounding (these samples are normally cut by
        # the centre trim): the tolerance grows with 1/envelope there
        cond = np.maximum(1.0, 1e-3 / np.maximum(envd[out_start:out_start + out_len].numpy(), 1e-30))
        good = envd[out_start:out_start + out_len].numpy() > 1e-11
        scale = max(1.0, float(np.abs(ref[:, good]).max())) if good.any() else 1.0
        err = max(err, float((np.abs(dst[s] - ref) / cond).max()) / scale)
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
        packed = bool(i % 2)
        err, desc = one(lib, rs, packed)
        worst = max(worst, err)
        flag = "" if err <= 5e-5 else "   <-- FAIL"
        print(f"{i:3d} err {err:.2e} {desc}{flag}", flush=True)
    print("worst", worst)
    sys.exit(0 if worst <= 5e-5 else 1)


if __name__ == "__main__":
    main()
