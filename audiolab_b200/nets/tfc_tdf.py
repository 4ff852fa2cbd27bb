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


# ---- full-size ConvTDFNet (KUIELab MDX-Net as released in the UVR .onnx models) --------------------
class _TFC(nn.Module):
    def __init__(self, c: int, l: int, k: int):
        super().__init__()
        self.H = nn.ModuleList([nn.Sequential(nn.Conv2d(c, c, k, 1, k // 2), nn.BatchNorm2d(c), nn.ReLU()) for _ in range(l)])

    def forward(self, x):
        for h in self.H:
            x = h(x)
        return x


class _TFC_TDF(nn.Module):
    def __init__(self, c: int, l: int, f: int, k: int, bn: int, bias: bool = True):
        super().__init__()
        self.use_tdf = bn is not None
        self.tfc = _TFC(c, l, k)
        if self.use_tdf:
            if bn == 0:
                self.tdf = nn.Sequential(nn.Linear(f, f, bias=bias), nn.BatchNorm2d(c), nn.ReLU())
            else:
                self.tdf = nn.Sequential(nn.Linear(f, f // bn, bias=bias), nn.BatchNorm2d(c), nn.ReLU(),
                                         nn.Linear(f // bn, f, bias=bias), nn.BatchNorm2d(c), nn.ReLU())

    def forward(self, x):
        x = self.tfc(x)
        return x + self.tdf(x) if self.use_tdf else x


class ConvTdfNet(nn.Module):
    """``Conv_TDF_net_trim`` without its STFT (that is al_stft / al_istft): [B, 4, dim_f, dim_t] -> same shape.

    Module / parameter names follow KUIELab's mdx-net (``first_conv``, ``encoding_blocks.i.tfc.H.j``, ``ds``,
    ``bottleneck_block``, ``us``, ``decoding_blocks``, ``final_conv``), so a state dict converted from a released
    checkpoint loads with ``strict=True``.  Released vocal models: L = 11 (5 down / 5 up), l = 3, g = 48, bn = 8, k = 3,
    dim_f 3072, dim_t 256, n_fft 6144 (the in-tree twin's defaults, mdxnet.py:241-253)."""

    def __init__(self, dim_f: int = 3072, L: int = 11, l: int = 3, g: int = 48, bn: int = 8, k: int = 3, dim_c: int = 4,
                 bias: bool = True):
        super().__init__()
        self.n = L // 2
        if dim_f % (2 ** self.n) != 0:
            raise ValueError(f"dim_f {dim_f} must be divisible by 2^{self.n}")
        self.first_conv = nn.Sequential(nn.Conv2d(dim_c, g, (1, 1)), nn.BatchNorm2d(g), nn.ReLU())
        f, c = dim_f, g
        self.encoding_blocks, self.ds = nn.ModuleList(), nn.ModuleList()
        for _ in range(self.n):
            self.encoding_blocks.append(_TFC_TDF(c, l, f, k, bn, bias=bias))
            self.ds.append(nn.Sequential(nn.Conv2d(c, c + g, kernel_size=(2, 2), stride=(2, 2)), nn.BatchNorm2d(c + g), nn.ReLU()))
            f, c = f // 2, c + g
        self.bottleneck_block = _TFC_TDF(c, l, f, k, bn, bias=bias)
        self.decoding_blocks, self.us = nn.ModuleList(), nn.ModuleList()
        for _ in range(self.n):
            self.us.append(nn.Sequential(nn.ConvTranspose2d(c, c - g, kernel_size=(2, 2), stride=(2, 2)), nn.BatchNorm2d(c - g),
                                         nn.ReLU()))
            f, c = f * 2, c - g
            self.decoding_blocks.append(_TFC_TDF(c, l, f, k, bn, bias=bias))
        self.final_conv = nn.Sequential(nn.Conv2d(c, dim_c, (1, 1)))

    def forward(self, x):
        x = self.first_conv(x)
        x = x.transpose(-1, -2)                 # [B, C, T, F]: TDF's Linear runs over the frequency axis
        skips = []
        for enc, ds in zip(self.encoding_blocks, self.ds):
            x = enc(x)
            skips.append(x)
            x = ds(x)
        x = self.bottleneck_block(x)
        for us, dec in zip(self.us, self.decoding_blocks):
            x = us(x)
            x = x * skips.pop()
            x = dec(x)
        x = x.transpose(-1, -2)
        return self.final_conv(x)

    @staticmethod
    def from_state_dict(sd: dict, dim_f: int) -> "ConvTdfNet":
        """Hyper-parameters read off a converted checkpoint's tensor shapes."""
        g = sd["first_conv.0.weight"].shape[0]
        n = len({key.split(".")[1] for key in sd if key.startswith("encoding_blocks.")})
        l = len({key.split(".")[4] for key in sd if key.startswith("encoding_blocks.0.tfc.H.")})
        k = sd["encoding_blocks.0.tfc.H.0.0.weight"].shape[-1]
        bn = None
        if "encoding_blocks.0.tdf.0.weight" in sd:
            out_f, in_f = sd["encoding_blocks.0.tdf.0.weight"].shape
            bn = 0 if out_f == in_f else in_f // out_f
        net = ConvTdfNet(dim_f=dim_f, L=2 * n + 1, l=l, g=g, bn=bn, k=k, bias="encoding_blocks.0.tdf.0.bias" in sd)
        net.load_state_dict(sd, strict=True)
        return net
