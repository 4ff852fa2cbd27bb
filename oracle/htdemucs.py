"""HTDemucs spectral path + segment loop, CPU oracle.  Test infrastructure only.

PARITY UNPINNED by any reference test: the code this restates lives in
``audio-separator`` -> ``demucs`` (facebookresearch/demucs v4; pinned ``demucs>=4.0.1`` in
/root/reference/requirements.txt; call sites /root/reference/modules/separator/
stem_separator.py:466-503), not in /root/reference.  Restated from SURVEY.md Appendix A.3.

The hybrid network is any ``core(mag[B,4,F,T], xt[B,2,L]) -> (x_spec[B,S,4,F,T], x_time[B,S,2,L])``;
the standardisation of both branches, the CaC ("complex as channels") packing, the
iSTFT and the sum of the two branches are done here, as in ``HTDemucs.forward``.
"""
from __future__ import annotations

import math
import random
from dataclasses import dataclass
from typing import Callable, Optional, Tuple

import torch
import torch.nn.functional as F


@dataclass
class HTDemucsConfig:
    nfft: int = 4096
    samplerate: int = 44100
    segment_num: int = 39          # segment = 39/5 s  -> 343 980 samples
    segment_den: int = 5
    overlap: float = 0.25
    shifts: int = 1
    num_sources: int = 4
    transition_power: float = 1.0

    @property
    def hop(self) -> int:
        return self.nfft // 4

    @property
    def segment_samples(self) -> int:
        return int(self.samplerate * self.segment_num / self.segment_den)


def pad1d(x: torch.Tensor, paddings: Tuple[int, int], mode: str = "constant") -> torch.Tensor:
    """demucs.hdemucs.pad1d: reflect padding that tolerates short inputs."""
    length = x.shape[-1]
    left, right = paddings
    if mode == "reflect":
        max_pad = max(left, right)
        if length <= max_pad:
            extra = max_pad - length + 1
            extra_right = min(right, extra)
            extra_left = extra - extra_right
            paddings = (left - extra_left, right - extra_right)
            x = F.pad(x, (extra_left, extra_right))
    return F.pad(x, paddings, mode)


def spec(x: torch.Tensor, cfg: HTDemucsConfig) -> torch.Tensor:
    """HTDemucs._spec: [B, C, L] -> complex [B, C, nfft/2, ceil(L/hop)]."""
    hl, nfft = cfg.hop, cfg.nfft
    le = int(math.ceil(x.shape[-1] / hl))
    pad = hl // 2 * 3
    x = pad1d(x, (pad, pad + le * hl - x.shape[-1]), mode="reflect")
    *other, length = x.shape
    z = torch.stft(x.reshape(-1, length), nfft, hl, window=torch.hann_window(nfft).to(x),
                   win_length=nfft, normalized=True, center=True, return_complex=True, pad_mode="reflect")
    z = z.view(*other, z.shape[-2], z.shape[-1])[..., :-1, :]
    assert z.shape[-1] == le + 4
    return z[..., 2: 2 + le]


def ispec(z: torch.Tensor, length: int, cfg: HTDemucsConfig) -> torch.Tensor:
    """HTDemucs._ispec: complex [..., nfft/2, T] -> [..., length]."""
    hl = cfg.hop
    z = F.pad(z, (0, 0, 0, 1))
    z = F.pad(z, (2, 2))
    pad = hl // 2 * 3
    le = hl * int(math.ceil(length / hl)) + 2 * pad
    *other, freqs, frames = z.shape
    n_fft = 2 * freqs - 2
    x = torch.istft(z.reshape(-1, freqs, frames), n_fft, hl, window=torch.hann_window(n_fft).to(z.real),
                    win_length=n_fft, normalized=True, length=le, center=True)
    x = x.view(*other, x.shape[-1])
    return x[..., pad: pad + length]


def magnitude_cac(z: torch.Tensor) -> torch.Tensor:
    B, C, Fr, T = z.shape
    return torch.view_as_real(z).permute(0, 1, 4, 2, 3).reshape(B, C * 2, Fr, T)


def mask_cac(m: torch.Tensor) -> torch.Tensor:
    B, S, C, Fr, T = m.shape
    out = m.view(B, S, -1, 2, Fr, T).permute(0, 1, 2, 4, 5, 3)
    return torch.view_as_complex(out.contiguous())


def hybrid_forward(mix: torch.Tensor, core: Callable, cfg: HTDemucsConfig) -> torch.Tensor:
    """HTDemucs.forward around the network core.  mix [B, 2, L] -> [B, S, 2, L]."""
    length = mix.shape[-1]
    training_length = cfg.segment_samples
    length_pre_pad = None
    if mix.shape[-1] < training_length:
        length_pre_pad = mix.shape[-1]
        mix = F.pad(mix, (0, training_length - length_pre_pad))
    z = spec(mix, cfg)
    mag = magnitude_cac(z)
    mean = mag.mean(dim=(1, 2, 3), keepdim=True)
    std = mag.std(dim=(1, 2, 3), keepdim=True)
    x = (mag - mean) / (1e-5 + std)
    meant = mix.mean(dim=(1, 2), keepdim=True)
    stdt = mix.std(dim=(1, 2), keepdim=True)
    xt = (mix - meant) / (1e-5 + stdt)
    x, xt = core(x, xt)
    x = x * std[:, None] + mean[:, None]
    zout = mask_cac(x)
    x = ispec(zout, mix.shape[-1], cfg)
    xt = xt * stdt[:, None] + meant[:, None]
    x = xt + x
    if length_pre_pad:
        x = x[..., :length_pre_pad]
    return x


def _padded(tensor: torch.Tensor, offset: int, length: int, target: int) -> torch.Tensor:
    """demucs.apply.TensorChunk.padded: centre the chunk in ``target`` using real context."""
    delta = target - length
    total = tensor.shape[-1]
    start = offset - delta // 2
    end = start + target
    cs, ce = max(0, start), min(total, end)
    return F.pad(tensor[..., cs:ce], (cs - start, end - ce))


def _center_trim(t: torch.Tensor, length: int) -> torch.Tensor:
    delta = t.shape[-1] - length
    if delta:
        t = t[..., delta // 2: -(delta - delta // 2)]
    return t


def triangle_weight(segment: int, power: float = 1.0) -> torch.Tensor:
    w = torch.cat([torch.arange(1, segment // 2 + 1), torch.arange(segment - segment // 2, 0, -1)]).float()
    return (w / w.max()) ** power


def segment_offsets(length: int, cfg: HTDemucsConfig):
    stride = int((1 - cfg.overlap) * cfg.segment_samples)
    return list(range(0, length, stride))


def apply_split(tensor: torch.Tensor, core: Callable, cfg: HTDemucsConfig,
                base: int = 0, length: Optional[int] = None) -> torch.Tensor:
    """demucs.apply.apply_model(split=True) for one model on ``tensor[..., base:base+length]``.

    Like ``TensorChunk``, padding of a segment to the training length takes real context
    from ``tensor`` outside the ``[base, base+length)`` view when it exists.
    Returns [B, S, 2, length].
    """
    B, C, total = tensor.shape
    if length is None:
        length = total - base
    segment = cfg.segment_samples
    out = torch.zeros(B, cfg.num_sources, C, length)
    sum_weight = torch.zeros(length)
    weight = triangle_weight(segment, cfg.transition_power)
    for offset in segment_offsets(length, cfg):
        clen = min(length - offset, segment)
        padded = _padded(tensor, base + offset, clen, segment)
        with torch.no_grad():
            chunk_out = _center_trim(hybrid_forward(padded, core, cfg), clen)
        out[..., offset: offset + segment] += weight[:clen] * chunk_out
        sum_weight[offset: offset + segment] += weight[:clen]
    assert sum_weight.min() > 0
    return out / sum_weight


def shift_offsets(cfg: HTDemucsConfig, seed: int = 0):
    rng = random.Random(seed)
    max_shift = int(0.5 * cfg.samplerate)
    return [rng.randint(0, max_shift) for _ in range(cfg.shifts)]


def apply_model(mix: torch.Tensor, core: Callable, cfg: HTDemucsConfig, seed: int = 0) -> torch.Tensor:
    """shifts (random time-shift averaging) around ``apply_split``."""
    if not cfg.shifts:
        return apply_split(mix, core, cfg)
    length = mix.shape[-1]
    max_shift = int(0.5 * cfg.samplerate)
    padded_mix = _padded(mix, 0, length, length + 2 * max_shift)
    out = 0
    for offset in shift_offsets(cfg, seed):
        shifted_out = apply_split(padded_mix, core, cfg, base=offset, length=length + max_shift - offset)
        out = out + shifted_out[..., max_shift - offset:]
    return out / cfg.shifts


def demix_demucs(mix: torch.Tensor, core: Callable, cfg: HTDemucsConfig, seed: int = 0) -> torch.Tensor:
    """DemucsSeparator.demix_demucs: normalise by ref, apply_model, de-normalise.  mix [2, L] -> [S, 2, L]."""
    ref = mix.mean(0)
    mean, std = ref.mean(), ref.std()
    x = (mix - mean) / std
    sources = apply_model(x[None], core, cfg, seed=seed)[0]
    return sources * std + mean
