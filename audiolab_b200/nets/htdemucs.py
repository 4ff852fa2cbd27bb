"""HTDemucs network core (hybrid spectrogram / waveform U-Net with a cross-domain transformer).

Restated from facebookresearch/demucs v4 ``HTDemucs`` (the reference reaches it through
``audio_separator`` -> ``demucs.apply.apply_model``, /root/reference/modules/separator/
stem_separator.py:466-503).  ``_spec`` / ``_magnitude`` / ``_mask`` / ``_ispec`` and the branch
standardisation are NOT here: they are al_stft / al_istft and ``HTDemucsDemixer``.  ``forward``
takes the standardised CaC spectrogram and waveform and returns both branches' outputs:

    core(mag [B, 4, 2048, T], xt [B, 2, L]) -> (x_spec [B, S, 4, 2048, T], x_time [B, S, 2, L])

Defaults: channels 48, growth 2, depth 4, kernel 8, stride 4, DConv depth 2 / compress 8, 5 cross-
transformer layers, 8 heads, hidden scale 4 (A.3).
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn


class LayerScale(nn.Module):
    def __init__(self, channels: int, init: float = 0.0, channel_last: bool = False):
        super().__init__()
        self.channel_last = channel_last
        self.scale = nn.Parameter(torch.full((channels,), float(init)))

    def forward(self, x):
        return self.scale * x if self.channel_last else self.scale[:, None] * x


class DConv(nn.Module):
    """Residual branch of dilated 1-D convs: x + LayerScale(GLU(Conv1x1(GELU(Conv_dilated(x)))))."""

    def __init__(self, channels: int, compress: float = 8, depth: int = 2, init: float = 1e-3, kernel: int = 3):
        super().__init__()
        hidden = int(channels / compress)
        self.layers = nn.ModuleList()
        for d in range(depth):
            dilation = 2 ** d
            self.layers.append(nn.Sequential(
                nn.Conv1d(channels, hidden, kernel, dilation=dilation, padding=dilation * (kernel // 2)),
                nn.GroupNorm(1, hidden), nn.GELU(),
                nn.Conv1d(hidden, 2 * channels, 1), nn.GroupNorm(1, 2 * channels), nn.GLU(1),
                LayerScale(channels, init)))

    def forward(self, x):
        for layer in self.layers:
            x = x + layer(x)
        return x


class HEncLayer(nn.Module):
    def __init__(self, chin, chout, kernel_size=8, stride=4, freq=True, context=0):
        super().__init__()
        self.freq, self.stride = freq, stride
        pad = kernel_size // 4
        if freq:
            self.conv = nn.Conv2d(chin, chout, (kernel_size, 1), (stride, 1), (pad, 0))
            self.rewrite = nn.Conv2d(chout, 2 * chout, 1 + 2 * context, 1, context)
        else:
            self.conv = nn.Conv1d(chin, chout, kernel_size, stride, pad)
            self.rewrite = nn.Conv1d(chout, 2 * chout, 1 + 2 * context, 1, context)
        self.dconv = DConv(chout)

    def forward(self, x, inject=None):
        if not self.freq:
            le = x.shape[-1]
            if le % self.stride:
                x = F.pad(x, (0, self.stride - (le % self.stride)))
        y = self.conv(x)
        if inject is not None:
            y = y + inject
        y = F.gelu(y)
        if self.freq:
            B, C, Fr, T = y.shape
            y = self.dconv(y.permute(0, 2, 1, 3).reshape(-1, C, T)).view(B, Fr, C, T).permute(0, 2, 1, 3)
        else:
            y = self.dconv(y)
        return F.glu(self.rewrite(y), dim=1)


class HDecLayer(nn.Module):
    def __init__(self, chin, chout, last=False, kernel_size=8, stride=4, freq=True, context=1):
        super().__init__()
        self.freq, self.last, self.pad, self.chin = freq, last, kernel_size // 4, chin
        if freq:
            self.conv_tr = nn.ConvTranspose2d(chin, chout, (kernel_size, 1), (stride, 1))
            self.rewrite = nn.Conv2d(chin, 2 * chin, 1 + 2 * context, 1, context)
        else:
            self.conv_tr = nn.ConvTranspose1d(chin, chout, kernel_size, stride)
            self.rewrite = nn.Conv1d(chin, 2 * chin, 1 + 2 * context, 1, context)

    def forward(self, x, skip, length):
        if self.freq and x.dim() == 3:
            B, C, T = x.shape
            x = x.view(B, self.chin, -1, T)
        x = x + skip
        y = F.glu(self.rewrite(x), dim=1)
        z = self.conv_tr(y)
        if self.freq:
            z = z[..., self.pad: -self.pad, :]
        else:
            z = z[..., self.pad: self.pad + length]
        return z if self.last else F.gelu(z)


class ScaledEmbedding(nn.Module):
    def __init__(self, num_embeddings: int, embedding_dim: int, scale: float = 10.0, smooth: bool = True):
        super().__init__()
        self.embedding = nn.Embedding(num_embeddings, embedding_dim)
        if smooth:
            w = torch.cumsum(self.embedding.weight.data, dim=0)
            self.embedding.weight.data[:] = w / torch.arange(1, num_embeddings + 1).sqrt()[:, None]
        self.embedding.weight.data /= scale
        self.scale = scale

    def forward(self, x):
        return self.embedding(x) * self.scale


def sin_embedding_1d(length: int, dim: int, device, max_period: float = 10000.0):
    pos = torch.arange(length, device=device).view(-1, 1, 1).float()
    half = dim // 2
    adim = torch.arange(half, device=device).view(1, 1, -1).float()
    phase = pos / (max_period ** (adim / (half - 1)))
    return torch.cat([torch.cos(phase), torch.sin(phase)], dim=-1)          # [T, 1, dim]


def sin_embedding_2d(d_model: int, height: int, width: int, device, max_period: float = 10000.0):
    pe = torch.zeros(d_model, height, width, device=device)
    d = d_model // 2
    div = torch.exp(torch.arange(0.0, d, 2, device=device) * -(math.log(max_period) / d))
    pos_w = torch.arange(0.0, width, device=device).unsqueeze(1)
    pos_h = torch.arange(0.0, height, device=device).unsqueeze(1)
    pe[0:d:2] = torch.sin(pos_w * div).t().unsqueeze(1).repeat(1, height, 1)
    pe[1:d:2] = torch.cos(pos_w * div).t().unsqueeze(1).repeat(1, height, 1)
    pe[d::2] = torch.sin(pos_h * div).t().unsqueeze(2).repeat(1, 1, width)
    pe[d + 1::2] = torch.cos(pos_h * div).t().unsqueeze(2).repeat(1, 1, width)
    return pe[None]                                                           # [1, C, Fr, T]


class _TLayer(nn.Module):
    """norm-first transformer layer with LayerScale; cross=True attends to the other branch."""

    def __init__(self, dim: int, heads: int, hidden: int, cross: bool):
        super().__init__()
        self.cross = cross
        self.attn = nn.MultiheadAttention(dim, heads, batch_first=True)
        self.norm1, self.norm2 = nn.LayerNorm(dim), nn.LayerNorm(dim)
        self.norm3 = nn.LayerNorm(dim) if cross else None
        self.linear1, self.linear2 = nn.Linear(dim, hidden), nn.Linear(hidden, dim)
        self.gamma_1, self.gamma_2 = LayerScale(dim, 1e-4, True), LayerScale(dim, 1e-4, True)
        self.norm_out = nn.GroupNorm(1, dim)

    def forward(self, q, k=None):
        if self.cross:
            kn = self.norm2(k)
            x = q + self.gamma_1(self.attn(self.norm1(q), kn, kn, need_weights=False)[0])
            x = x + self.gamma_2(self.linear2(F.gelu(self.linear1(self.norm3(x)))))
        else:
            qn = self.norm1(q)
            x = q + self.gamma_1(self.attn(qn, qn, qn, need_weights=False)[0])
            x = x + self.gamma_2(self.linear2(F.gelu(self.linear1(self.norm2(x)))))
        return self.norm_out(x.transpose(1, 2)).transpose(1, 2)


class CrossTransformerEncoder(nn.Module):
    def __init__(self, dim: int, heads: int = 8, hidden_scale: float = 4.0, num_layers: int = 5):
        super().__init__()
        hidden = int(dim * hidden_scale)
        self.norm_in, self.norm_in_t = nn.LayerNorm(dim), nn.LayerNorm(dim)
        self.layers = nn.ModuleList([_TLayer(dim, heads, hidden, cross=bool(i % 2)) for i in range(num_layers)])
        self.layers_t = nn.ModuleList([_TLayer(dim, heads, hidden, cross=bool(i % 2)) for i in range(num_layers)])

    def forward(self, x, xt):
        B, C, Fr, T1 = x.shape
        pos2d = sin_embedding_2d(C, Fr, T1, x.device).permute(0, 3, 2, 1).reshape(1, T1 * Fr, C)
        x = self.norm_in(x.permute(0, 3, 2, 1).reshape(B, T1 * Fr, C)) + pos2d.to(x.dtype)
        T2 = xt.shape[-1]
        pos = sin_embedding_1d(T2, C, x.device).permute(1, 0, 2)
        xt = self.norm_in_t(xt.permute(0, 2, 1)) + pos.to(xt.dtype)
        for layer, layer_t in zip(self.layers, self.layers_t):
            if layer.cross:
                old_x = x
                x = layer(x, xt)
                xt = layer_t(xt, old_x)
            else:
                x = layer(x)
                xt = layer_t(xt)
        x = x.reshape(B, T1, Fr, C).permute(0, 3, 2, 1)
        return x, xt.permute(0, 2, 1)


class HTDemucsCore(nn.Module):
    def __init__(self, num_sources: int = 4, audio_channels: int = 2, channels: int = 48, growth: int = 2,
                 depth: int = 4, nfft: int = 4096, kernel_size: int = 8, stride: int = 4, freq_emb: float = 0.2,
                 t_layers: int = 5, t_heads: int = 8, t_hidden_scale: float = 4.0, bottom_channels: int = 0):
        super().__init__()
        self.num_sources, self.audio_channels, self.depth = num_sources, audio_channels, depth
        self.encoder, self.decoder = nn.ModuleList(), nn.ModuleList()
        self.tencoder, self.tdecoder = nn.ModuleList(), nn.ModuleList()
        chin, chin_z = audio_channels, audio_channels * 2
        chout = chout_z = channels
        freqs = nfft // 2
        for index in range(depth):
            self.encoder.append(HEncLayer(chin_z, chout_z, kernel_size, stride, freq=True))
            self.tencoder.append(HEncLayer(chin, chout, kernel_size, stride, freq=False))
            if index == 0:
                chin, chin_z = audio_channels * num_sources, audio_channels * 2 * num_sources
            self.decoder.insert(0, HDecLayer(chout_z, chin_z, last=index == 0, kernel_size=kernel_size,
                                             stride=stride, freq=True))
            self.tdecoder.insert(0, HDecLayer(chout, chin, last=index == 0, kernel_size=kernel_size,
                                              stride=stride, freq=False))
            chin, chin_z = chout, chout_z
            chout, chout_z = int(growth * chout), int(growth * chout_z)
            freqs //= stride
            if index == 0:
                self.freq_emb = ScaledEmbedding(freqs, chin_z)
                self.freq_emb_scale = freq_emb
        tdim = channels * growth ** (depth - 1)
        self.bottom_channels = bottom_channels
        if bottom_channels:
            self.channel_upsampler = nn.Conv1d(tdim, bottom_channels, 1)
            self.channel_downsampler = nn.Conv1d(bottom_channels, tdim, 1)
            self.channel_upsampler_t = nn.Conv1d(tdim, bottom_channels, 1)
            self.channel_downsampler_t = nn.Conv1d(bottom_channels, tdim, 1)
            tdim = bottom_channels
        self.crosstransformer = CrossTransformerEncoder(tdim, t_heads, t_hidden_scale, t_layers)

    def forward(self, x, xt):
        B, _, Fq, T = x.shape
        length = xt.shape[-1]
        saved, saved_t, lengths, lengths_t = [], [], [], []
        for idx, (enc, tenc) in enumerate(zip(self.encoder, self.tencoder)):
            lengths.append(x.shape[-1])
            lengths_t.append(xt.shape[-1])
            xt = tenc(xt)
            saved_t.append(xt)
            x = enc(x)
            if idx == 0:
                frs = torch.arange(x.shape[-2], device=x.device)
                emb = self.freq_emb(frs).t()[None, :, :, None].expand_as(x)
                x = x + self.freq_emb_scale * emb.to(x.dtype)
            saved.append(x)
        if self.bottom_channels:
            b, c, f, t = x.shape
            x = self.channel_upsampler(x.reshape(b, c, f * t)).view(b, -1, f, t)
            xt = self.channel_upsampler_t(xt)
        x, xt = self.crosstransformer(x, xt)
        if self.bottom_channels:
            b, c, f, t = x.shape
            x = self.channel_downsampler(x.reshape(b, c, f * t)).view(b, -1, f, t)
            xt = self.channel_downsampler_t(xt)
        for dec, tdec in zip(self.decoder, self.tdecoder):
            x = dec(x, saved.pop(-1), lengths.pop(-1))
            xt = tdec(xt, saved_t.pop(-1), lengths_t.pop(-1))
        S = self.num_sources
        return x.reshape(B, S, -1, Fq, T), xt.reshape(B, S, -1, length)
