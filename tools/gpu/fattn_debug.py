"""Bring-up diagnostics of the tcgen05 time-axis attention (csrc/al_fattn.cu): error structure of small cases, then the
timing of the bench shape against F.scaled_dot_product_attention (cuDNN) + the separate gate pass it replaces."""
import json
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from audiolab_b200 import netops  # noqa: E402


def ref(q, k, v, B, T, I, H, gates=None):
    shp = (B, T, I * H, 64)
    o = F.scaled_dot_product_attention(q.view(shp).transpose(1, 2).float(), k.view(shp).transpose(1, 2).float(),
                                       v.view(shp).transpose(1, 2).float()).transpose(1, 2).reshape(q.shape)
    if gates is not None:
        o = (o.view(-1, H, 64) * torch.sigmoid(gates.float())[:, :, None]).reshape(q.shape)
    return o


def describe(name, got, want, B, T, I, H):
    err = (got - want).abs()
    tol = (2 ** -6 if "bfloat16" in name else 2 ** -9) * float(want.abs().max())
    bad = (err > tol) | ~torch.isfinite(got)
    info = {"case": name, "B": B, "T": T, "I": I, "H": H, "max_err": float(err[torch.isfinite(err)].max()),
            "ref_max": float(want.abs().max()), "bad": int(bad.sum()), "nan": int((~torch.isfinite(got)).sum()), "n": got.numel()}
    if bad.any():
        b4 = bad.view(B, T, I, H, 64)
        info["bad_by_t_div32"] = [int(b4[:, i * 32:(i + 1) * 32].sum()) for i in range((T + 31) // 32)][:32]
        info["bad_by_d_div8"] = [int(b4[..., i * 8:(i + 1) * 8].sum()) for i in range(8)]
        info["bad_by_h"] = [int(b4[:, :, :, h].sum()) for h in range(H)]
        info["bad_by_i"] = [int(b4[:, :, i].sum()) for i in range(I)][:16]
        idx = b4.nonzero()[0].tolist()
        info["first_bad"] = idx + [float(got.view(B, T, I, H, 64)[tuple(idx)]), float(want.view(B, T, I, H, 64)[tuple(idx)])]
        # does the output row equal the reference of ANOTHER row / a column permutation?
        g0 = got.view(B, T, I, H, 64)[0, :, 0, 0].float()
        w0 = want.view(B, T, I, H, 64)[0, :, 0, 0].float()
        if torch.isfinite(g0).all():
            d = torch.cdist(g0[None], w0[None])[0]
            info["row_match_first16"] = [int(x) for x in d.argmin(dim=1)[:16]]
            dc = torch.cdist(g0.t()[None], w0.t()[None])[0]
            info["col_match"] = [int(x) for x in dc.argmin(dim=1)]
    print(json.dumps(info), flush=True)
    return not bad.any()


def timed(fn, iters=10, warm=3):
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
    for dtype in (torch.float16, torch.bfloat16):
        for (B, T, I, H) in [(1, 128, 1, 1), (1, 64, 1, 1), (1, 256, 1, 1), (1, 129, 1, 1), (1, 801, 1, 1), (2, 801, 3, 2)]:
            q, k, v = (torch.randn(B * T * I, H * 64, device=dev).to(dtype) for _ in range(3))
            gates = torch.randn(B * T * I, H, device=dev).to(dtype)
            got = netops.time_attention(q, k, v, B, T, I, H, 64).float()
            torch.cuda.synchronize()
            ok &= describe(f"{str(dtype)[6:]} plain", got, ref(q, k, v, B, T, I, H), B, T, I, H)
            got = netops.time_attention(q, k, v, B, T, I, H, 64, gates=gates).float()
            torch.cuda.synchronize()
            ok &= describe(f"{str(dtype)[6:]} gated", got, ref(q, k, v, B, T, I, H, gates), B, T, I, H)
            if not ok:
                return 1
    if "--no-time" in sys.argv:
        return 0
    B, T, I, H = 27, 801, 62, 8
    q, k, v = (torch.randn(B * T * I, H * 64, device=dev).half() for _ in range(3))
    gates = torch.randn(B * T * I, 16, device=dev).half()[:, :H]
    t_ours = timed(lambda: netops.time_attention(q, k, v, B, T, I, H, 64, gates=gates))
    shp = (B, T, I * H, 64)

    def lib():
        o = F.scaled_dot_product_attention(q.view(shp).transpose(1, 2), k.view(shp).transpose(1, 2), v.view(shp).transpose(1, 2))
        o = o.transpose(1, 2)
        if not o.is_contiguous():
            o = o.contiguous()
        netops.gate_sigmoid_(o.view(-1, H * 64), gates, H, 64)
    t_lib = timed(lib)
    t_sdpa = timed(lambda: F.scaled_dot_product_attention(q.view(shp).transpose(1, 2), k.view(shp).transpose(1, 2),
                                                          v.view(shp).transpose(1, 2)))
    flops = 4.0 * B * I * H * T * T * 64
    print(json.dumps({"kind": "time attention 27x801x62x8", "ms": round(t_ours, 4), "tflops": round(flops / t_ours / 1e9, 1),
                      "cudnn_sdpa_plus_gate_ms": round(t_lib, 4), "cudnn_sdpa_ms": round(t_sdpa, 4),
                      "exp_per_s": round(B * I * H * T * T / t_ours / 1e6, 1)}), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
