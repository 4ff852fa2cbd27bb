"""TFC-TDF U-Net stand-in for the released MDX-Net ``.onnx`` models.

Contract (from /root/reference/handlers/patch_separate.py:45-63): ``model_run(spek)`` with
``spek [B, 4, dim_f, dim_t]`` fp32 returns ``spec_pred`` of the same shape.  The reference feeds
it through onnxruntime with a host round trip per chunk (``spek.cpu().numpy()``); here it is a
PyTorch module that stays on the device between al_stft and al_istft.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn


class TFC(nn.Module):
    def __init__(self, c: int, layers: int = 3):
        super().__init__()
        self.convs = nn.ModuleList([
            nn.Sequential(nn.Conv2d(c, c, 3, padding=1), nn.BatchNorm2d(c), nn.ReLU()) for _ in range(layers)
        ])

    def forward(self, x):
        for conv in self.convs:
            x = conv(x)
        return x


class TDF(nn.Module):
    """Bottlenecked Linear over the frequency axis (input [B, C, F, T])."""

    def __init__(self, c: int, f: int, bn: int = 8):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(f, max(f // bn, 1)), nn.ReLU(), nn.Linear(max(f // bn, 1), f), nn.ReLU())

    def forward(self, x):
        return self.net(x.transpose(-1, -2)).transpose(-1, -2)


class TfcTdfNet(nn.Module):
    def __init__(self, dim_f: int, channels: int = 32, depth: int = 2, growth: int = 16):
        super().__init__()
        self.first = nn.Sequential(nn.Conv2d(4, channels, 1), nn.ReLU())
        c, f = channels, dim_f
        self.enc, self.down, self.up, self.dec = nn.ModuleList(), nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        for _ in range(depth):
            self.enc.append(nn.ModuleList([TFC(c), TDF(c, f)]))
            self.down.append(nn.Sequential(nn.Conv2d(c, c + growth, 2, stride=2), nn.ReLU()))
            c, f = c + growth, f // 2
        self.mid = nn.ModuleList([TFC(c), TDF(c, f)])
        for _ in range(depth):
            self.up.append(nn.Sequential(nn.ConvTranspose2d(c, c - growth, 2, stride=2), nn.ReLU()))
            c, f = c - growth, f * 2
            self.dec.append(nn.ModuleList([TFC(c), TDF(c, f)]))
        self.final = nn.Conv2d(c, 4, 1)

    def forward(self, x):
        h = self.first(x)
        skips = []
        for (tfc, tdf), down in zip(self.enc, self.down):
            h = tfc(h)
            h = h + tdf(h)
            skips.append(h)
            h = down(h)
        tfc, tdf = self.mid
        h = tfc(h)
        h = h + tdf(h)
        for up, (tfc, tdf) in zip(self.up, self.dec):
            h = up(h) * skips.pop()
            h = tfc(h)
            h = h + tdf(h)
        return self.final(h)
