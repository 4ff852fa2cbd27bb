"""Launch each mask-network GEMM shape of the bench step twice (for `ncu -k regex:gemm_bf16_kernel`)."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from audiolab_b200 import netops  # noqa: E402

M = 27 * 801 * 62
dev = "cuda"
a512 = torch.randn(M, 512, device=dev).bfloat16()
ssin = torch.rand(M, 4, device=dev) + 0.5
cs = torch.stack((torch.rand(801, 32, device=dev), torch.rand(801, 32, device=dev)), dim=-1).contiguous()
for (n, k, kind) in [(1552, 512, "qkv"), (2048, 512, "ff1"), (512, 2048, "ff2"), (512, 512, "out")]:
    a = a512 if k == 512 else torch.randn(M, k, device=dev).bfloat16()
    w = (torch.randn(n, k, device=dev) * k ** -0.5).bfloat16()
    for _ in range(2):
        if kind in ("ff2", "out"):
            x32 = torch.randn(M, n, device=dev)
            xb = torch.empty(M, n, device=dev, dtype=torch.bfloat16)
            ss = torch.empty(M, 4, device=dev)
            netops.gemm_bf16_residual(a, w, x32, xb, ss, bias=torch.randn(n, device=dev) if kind == "ff2" else None)
        elif kind == "qkv":
            q, kk, v = (torch.empty(M, 512, device=dev, dtype=torch.bfloat16) for _ in range(3))
            g = torch.empty(M, 16, device=dev, dtype=torch.bfloat16)
            netops.gemm_bf16(a, w, [q, kk, v, g], bias=torch.zeros(n, device=dev), row_ss=ssin, ss_scale=math.sqrt(512.0),
                             cos_sin=cs, pos_div=62, pos_mod=801, rot_cols=1024, out_split=512)
        else:
            out = torch.empty(M, n, device=dev, dtype=torch.bfloat16)
            netops.gemm_bf16(a, w, out, bias=torch.randn(n, device=dev), row_ss=ssin, ss_scale=math.sqrt(512.0), act="gelu")
        torch.cuda.synchronize()
print("done")
