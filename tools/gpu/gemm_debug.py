"""Bring-up diagnostics of the tcgen05 GEMM: error structure of the simplest cases (so a wrong descriptor /
swizzle shows up as a pattern, not just as "mismatch"), then timings of the mask-network shapes against torch
(cuBLAS) on the same tensors.  Each case runs in this process; run under `timeout`."""
import json
import math
import sys
import time

import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from audiolab_b200 import netops  # noqa: E402


def describe(out, ref, name):
    out = out.float()
    err = (out - ref).abs()
    tol = 2.0 ** -7 * ref.abs() + 1e-4 * ref.abs().max()
    bad = err > tol
    nan = torch.isnan(out)
    info = {"case": name, "shape": list(out.shape), "max_err": float(err[~nan].max()) if (~nan).any() else None,
            "ref_max": float(ref.abs().max()), "bad": int(bad.sum()), "nan": int(nan.sum()), "n": out.numel()}
    if bad.any() or nan.any():
        b = bad | nan
        rows = b.any(dim=1).nonzero().flatten()
        cols = b.any(dim=0).nonzero().flatten()
        info["bad_rows"] = [int(rows.min()), int(rows.max()), int(rows.numel())]
        info["bad_cols"] = [int(cols.min()), int(cols.max()), int(cols.numel())]
        info["bad_by_row_mod8"] = [int(b[i::8].sum()) for i in range(8)]
        info["bad_by_col_div8_mod8"] = [int(b[:, [c for c in range(out.shape[1]) if (c // 8) % 8 == i]].sum()) for i in range(8)]
        info["bad_by_col_div64"] = [int(b[:, i * 64:(i + 1) * 64].sum()) for i in range((out.shape[1] + 63) // 64)][:32]
        info["bad_by_row_div32"] = [int(b[i * 32:(i + 1) * 32].sum()) for i in range((out.shape[0] + 31) // 32)][:16]
        i, j = [int(v) for v in b.nonzero()[0]]
        info["first_bad"] = [i, j, float(out[i, j]), float(ref[i, j])]
        # is the output a permutation of the reference columns? (swizzle mix-up): best matching ref column for out[:, j]
        r0 = ref[: min(128, ref.shape[0])]
        o0 = out[: min(128, out.shape[0])]
        if not nan[: o0.shape[0]].any():
            d = torch.cdist(o0.t()[None], r0.t()[None])[0]
            match = d.argmin(dim=1)
            info["col_match_first64"] = [int(v) for v in match[:64]]
    print(json.dumps(info), flush=True)
    return not (bad.any() or nan.any())


def timed(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    torch.manual_seed(0)
    dev = "cuda"
    ok = True
    # 1. smallest case: one tile, one k-block
    for (m, n, k) in [(128, 256, 64), (128, 256, 128), (128, 256, 512), (256, 512, 512), (1000, 1536, 512)]:
        a = torch.randn(m, k, device=dev).bfloat16()
        w = (torch.randn(n, k, device=dev) * k ** -0.5).bfloat16()
        out = torch.full((m, n), float("nan"), device=dev, dtype=torch.bfloat16)
        netops.gemm_bf16(a, w, out)
        torch.cuda.synchronize()
        ok &= describe(out, a.float() @ w.float().t(), f"plain {m}x{n}x{k}")
        if not ok:
            break
    if ok:
        m, n, k = 1000, 512, 512
        a = torch.randn(m, k, device=dev).bfloat16()
        w = (torch.randn(n, k, device=dev) * k ** -0.5).bfloat16()
        x0 = torch.randn(m, n, device=dev)
        x32 = x0.clone()
        xb = torch.full((m, n), float("nan"), device=dev, dtype=torch.bfloat16)
        ss = torch.zeros(m, n // netops.resid_slab(n), device=dev)
        netops.gemm_bf16_residual(a, w, x32, xb, ss)
        torch.cuda.synchronize()
        ref = x0 + a.float() @ w.float().t()
        ok &= describe(x32, ref, "residual x32")
        ok &= describe(xb, ref, "residual xb")
    if not ok or "--no-time" in sys.argv:
        return 0 if ok else 1
    # 2. timings at the bench shapes (M = 27 chunks x 801 frames x 62 bands)
    M = 27 * 801 * 62
    res = []
    a512 = torch.randn(M, 512, device=dev).bfloat16()
    for (n, k, kind) in [(1552, 512, "qkv+gates"), (2048, 512, "ff1+gelu"), (512, 2048, "ff2+res"), (512, 512, "out+res")]:
        a = a512 if k == 512 else torch.randn(M, k, device=dev).bfloat16()
        w = (torch.randn(n, k, device=dev) * k ** -0.5).bfloat16()
        flops = 2.0 * M * n * k
        if kind.endswith("res"):
            x32 = torch.randn(M, n, device=dev)
            xb = torch.empty(M, n, device=dev, dtype=torch.bfloat16)
            ss = torch.empty(M, n // netops.resid_slab(n), device=dev)
            t = timed(lambda: netops.gemm_bf16_residual(a, w, x32, xb, ss))
            xres = torch.randn(M, n, device=dev).bfloat16()
            t_ref = timed(lambda: xres.addmm_(a, w.t()))
            bytes_ = M * (k * 2 + n * 4 * 2 + n * 2)
            del x32, xb, xres
        else:
            out = torch.empty(M, n, device=dev, dtype=torch.bfloat16)
            bias = torch.randn(n, device=dev)
            ssin = torch.rand(M, 2, device=dev) + 0.5
            act = "gelu" if "gelu" in kind else None
            t = timed(lambda: netops.gemm_bf16(a, w, out, bias=bias, row_ss=ssin, ss_scale=math.sqrt(512.0), act=act))
            t_ref = timed(lambda: torch.nn.functional.linear(a, w))
            t_plain = timed(lambda: netops.gemm_bf16(a, w, out))
            bytes_ = M * (k * 2 + n * 2)
            del out
        r = {"kind": kind, "M": M, "N": n, "K": k, "ms": round(t, 4), "tflops": round(flops / t / 1e9, 1),
             "gbs": round(bytes_ / t / 1e6, 1), "torch_ms": round(t_ref, 4), "torch_tflops": round(flops / t_ref / 1e9, 1)}
        if not kind.endswith("res"):
            r["plain_ms"] = round(t_plain, 4)
        print(json.dumps(r), flush=True)
        res.append(r)
        torch.cuda.empty_cache()
    return 0


if __name__ == "__main__":
    sys.exit(main())
