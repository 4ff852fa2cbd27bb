"""Cost of each epilogue ingredient of al_gemm_bf16 on the FF Linear-1 shape (M = 27 x 801 x 62, N = 2048, K = 512)."""
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from audiolab_b200 import netops  # noqa: E402


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


M, N, K = 27 * 801 * 62, 2048, 512
dev = "cuda"
a = torch.randn(M, K, device=dev).half()
w = (torch.randn(N, K, device=dev) * K ** -0.5).half()
out = torch.empty(M, N, device=dev, dtype=torch.float16)
bias = torch.randn(N, device=dev)
ss = torch.rand(M, 2, device=dev) + 0.5
res = {}
for name, kw in [("plain", {}), ("bias", {"bias": bias}), ("rowscale", {"row_ss": ss, "ss_scale": math.sqrt(512.0)}),
                 ("gelu", {"act": "gelu"}), ("tanh", {"act": "tanh"}), ("bias+gelu", {"bias": bias, "act": "gelu"}),
                 ("all", {"bias": bias, "row_ss": ss, "ss_scale": math.sqrt(512.0), "act": "gelu"})]:
    res[name] = round(timed(lambda: netops.gemm_bf16(a, w, out, **kw)), 4)
print(json.dumps({"shape": [M, N, K], "pairs": os.environ.get("AL_GEMM_PAIRS", "1"), "ms": res}))
