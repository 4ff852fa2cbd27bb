"""Timing of the band-axis attention kernel at the bench shape (27 x 801 sequences of 62 bands, 8 heads, fp16, gated)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from audiolab_b200 import netops  # noqa: E402

n_seq, F, H = 27 * 801, 62, 8
q, k, v = (torch.randn(n_seq * F, H * 64, device="cuda").half() for _ in range(3))
gates = torch.randn(n_seq * F, 16, device="cuda").half()[:, :H]
for _ in range(3):
    netops.band_attention(q, k, v, n_seq, F, H, 64, gates=gates)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    netops.band_attention(q, k, v, n_seq, F, H, 64, gates=gates)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
gb = (3 * q.numel() * 2 + q.numel() * 2 + gates.numel() * 2) / 1e9
print(f'{{"kind": "band attention 21627x62x8", "ms": {ms:.4f}, "gbs": {gb / ms * 1e3:.1f}}}')
