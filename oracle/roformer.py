"""BS-RoFormer / Mel-Band RoFormer demix, CPU oracle.  Test infrastructure only.

PARITY UNPINNED by any reference test: the code this restates lives in the un-vendored
third-party package ``audio-separator[gpu]>=0.32.0`` (/root/reference/setup.sh:96; call
sites /root/reference/modules/separator/stem_separator.py:102-107,281,394), which vendors
the lucidrains / ZFTurbo ``bs_roformer`` and ``mel_band_roformer`` modules.  The
algorithm is restated from SURVEY.md Appendix A.2; parameter names and ``state_dict``
keys follow upstream so that released checkpoints would load.  The spectral ground
truth is ``torch.stft`` / ``torch.istft`` (BASELINE.json north_star), which is what
``forward`` calls.

Layout notes (upstream einops strings, kept as comments because the CUDA path fuses them):
  stft -> view_as_real                      [b*s, f, t, c]
  'b s f t c -> b (f s) t c'                stereo interleaved inside frequency (row = f*2+s)
  'b f t c -> b t (f c)'                    band-split input
  mask 'b n t (f c) -> b n f t c'           complex mask per (f s) row
  'b n (f s) t -> (b n s) f t'              istft rows
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

# 24x2, 12x4, 8x12, 8x24, 8x48, 128, 129 -> 62 bands, sum 1025
DEFAULT_FREQS_PER_BANDS: Tuple[int, ...] = (
    (2,) * 24 + (4,) * 12 + (12,) * 8 + (24,) * 8 + (48,) * 8 + (128, 129)
)


@dataclass
class RoformerConfig:
    kind: str = "bs"                   # "bs" | "mel"
    dim: int = 512
    depth: int = 12
    stereo: bool = True
    num_stems: int = 1
    time_transformer_depth: int = 1
    freq_transformer_depth: int = 1
    freqs_per_bands: Tuple[int, ...] = DEFAULT_FREQS_PER_BANDS
    num_bands: int = 60                # mel only
    sample_rate: int = 44100           # mel only
    dim_head: int = 64
    heads: int = 8
    ff_mult: int = 4
    stft_n_fft: int = 2048
    stft_hop_length: int = 441
    stft_win_length: int = 2048
    stft_normalized: bool = False
    mask_estimator_depth: int = 2
    mlp_expansion_factor: int = 4
    # chunk loop (A.2 / BASELINE cfg2)
    chunk_size: int = 352800
    num_overlap: int = 4

    @property
    def audio_channels(self) -> int:
        return 2 if self.stereo else 1

    @property
    def step(self) -> int:
        return self.chunk_size // self.num_overlap


# --------------------------------------------------------------------------------------
# building blocks (upstream names -> state_dict compatible)
# --------------------------------------------------------------------------------------
class RMSNorm(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.scale = dim ** 0.5
        self.gamma = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        return F.normalize(x, dim=-1) * self.scale * self.gamma


class RotaryEmbedding(nn.Module):
    """rotary_embedding_torch.RotaryEmbedding(dim=dim_head): theta 10000, interleaved pairs."""

    def __init__(self, dim: int, theta: float = 10000.0):
        super().__init__()
        freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
        self.freqs = nn.Parameter(freqs, requires_grad=False)

    def rotate_queries_or_keys(self, t: torch.Tensor) -> torch.Tensor:
        n = t.shape[-2]
        pos = torch.arange(n, device=t.device, dtype=self.freqs.dtype)
        ang = torch.einsum("n,f->nf", pos, self.freqs)
        ang = torch.repeat_interleave(ang, 2, dim=-1)            # '... n -> ... (n r)', r=2
        cos, sin = ang.cos().to(t.dtype), ang.sin().to(t.dtype)
        x = t.reshape(*t.shape[:-1], -1, 2)
        x1, x2 = x.unbind(-1)
        rot = torch.stack((-x2, x1), dim=-1).reshape(t.shape)     # rotate_half
        return t * cos + rot * sin


class FeedForward(nn.Module):
    def __init__(self, dim: int, mult: int = 4):
        super().__init__()
        inner = int(dim * mult)
        self.net = nn.Sequential(RMSNorm(dim), nn.Linear(dim, inner), nn.GELU(), nn.Dropout(0.0),
                                 nn.Linear(inner, dim), nn.Dropout(0.0))

    def forward(self, x):
        return self.net(x)


class Attention(nn.Module):
    def __init__(self, dim: int, heads: int, dim_head: int, rotary_embed: Optional[RotaryEmbedding]):
        super().__init__()
        self.heads = heads
        inner = heads * dim_head
        self.rotary_embed = rotary_embed
        self.norm = RMSNorm(dim)
        self.to_qkv = nn.Linear(dim, inner * 3, bias=False)
        self.to_gates = nn.Linear(dim, heads)
        self.to_out = nn.Sequential(nn.Linear(inner, dim, bias=False), nn.Dropout(0.0))

    def forward(self, x):
        b, n, _ = x.shape
        x = self.norm(x)
        qkv = self.to_qkv(x).reshape(b, n, 3, self.heads, -1).permute(2, 0, 3, 1, 4)  # qkv b h n d
        q, k, v = qkv[0], qkv[1], qkv[2]
        if self.rotary_embed is not None:
            q = self.rotary_embed.rotate_queries_or_keys(q)
            k = self.rotary_embed.rotate_queries_or_keys(k)
        out = F.scaled_dot_product_attention(q, k, v)
        gates = self.to_gates(x)                                                       # b n h
        out = out * gates.permute(0, 2, 1).unsqueeze(-1).sigmoid()
        out = out.permute(0, 2, 1, 3).reshape(b, n, -1)
        return self.to_out(out)


class Transformer(nn.Module):
    def __init__(self, dim, depth, heads, dim_head, ff_mult, rotary_embed, norm_output):
        super().__init__()
        self.layers = nn.ModuleList([
            nn.ModuleList([Attention(dim, heads, dim_head, rotary_embed), FeedForward(dim, ff_mult)])
            for _ in range(depth)
        ])
        self.norm = RMSNorm(dim) if norm_output else nn.Identity()

    def forward(self, x):
        for attn, ff in self.layers:
            x = attn(x) + x
            x = ff(x) + x
        return self.norm(x)


class BandSplit(nn.Module):
    def __init__(self, dim: int, dim_inputs: Tuple[int, ...]):
        super().__init__()
        self.dim_inputs = tuple(dim_inputs)
        self.to_features = nn.ModuleList([nn.Sequential(RMSNorm(d), nn.Linear(d, dim)) for d in dim_inputs])

    def forward(self, x):
        parts = x.split(self.dim_inputs, dim=-1)
        return torch.stack([f(p) for p, f in zip(parts, self.to_features)], dim=-2)


def _mlp(dim_in, dim_out, dim_hidden, depth):
    dims = (dim_in,) + (dim_hidden,) * (depth - 1) + (dim_out,)
    net: List[nn.Module] = []
    for i, (a, b) in enumerate(zip(dims[:-1], dims[1:])):
        net.append(nn.Linear(a, b))
        if i != len(dims) - 2:
            net.append(nn.Tanh())
    return nn.Sequential(*net)


class MaskEstimator(nn.Module):
    def __init__(self, dim, dim_inputs, depth, mlp_expansion_factor=4):
        super().__init__()
        self.dim_inputs = tuple(dim_inputs)
        hidden = dim * mlp_expansion_factor
        self.to_freqs = nn.ModuleList([
            nn.Sequential(_mlp(dim, d * 2, hidden, depth), nn.GLU(dim=-1)) for d in dim_inputs
        ])

    def forward(self, x):
        bands = x.unbind(dim=-2)
        return torch.cat([mlp(b) for b, mlp in zip(bands, self.to_freqs)], dim=-1)


# --------------------------------------------------------------------------------------
# mel filter bank (librosa.filters.mel, Slaney scale + norm; librosa is absent here)
# --------------------------------------------------------------------------------------
def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def mel_filter_bank(sr: int, n_fft: int, n_mels: int) -> np.ndarray:
    fftfreqs = np.linspace(0, sr / 2.0, 1 + n_fft // 2)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(0.0), _hz_to_mel(sr / 2.0), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float64)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2: n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, None]
    return weights.astype(np.float32)


def mel_band_masks(sr: int, n_fft: int, n_mels: int) -> np.ndarray:
    """Boolean [n_mels, n_fft/2+1] band membership used by Mel-Band RoFormer (A.2)."""
    fb = mel_filter_bank(sr, n_fft, n_mels)
    fb[0, 0] = 1.0
    fb[-1, -1] = 1.0
    member = fb > 0
    assert member.any(axis=0).all(), "all frequencies need to be covered by a band"
    return member


# --------------------------------------------------------------------------------------
# the models
# --------------------------------------------------------------------------------------
class _RoformerBase(nn.Module):
    cfg: RoformerConfig

    def _stft(self, raw_audio):
        c = self.cfg
        b, s, t = raw_audio.shape
        window = torch.hann_window(c.stft_win_length, device=raw_audio.device)
        X = torch.stft(raw_audio.reshape(b * s, t), n_fft=c.stft_n_fft, hop_length=c.stft_hop_length,
                       win_length=c.stft_win_length, normalized=c.stft_normalized, window=window,
                       return_complex=True)
        X = torch.view_as_real(X).reshape(b, s, X.shape[-2], X.shape[-1], 2)
        # 'b s f t c -> b (f s) t c'
        return X.permute(0, 2, 1, 3, 4).reshape(b, -1, X.shape[-2], 2), window

    def _istft(self, Y, window, length):
        c = self.cfg
        b, n, fs, t = Y.shape
        s = c.audio_channels
        # 'b n (f s) t -> (b n s) f t'
        Y = Y.reshape(b, n, fs // s, s, t).permute(0, 1, 3, 2, 4).reshape(b * n * s, fs // s, t)
        y = torch.istft(Y, n_fft=c.stft_n_fft, hop_length=c.stft_hop_length, win_length=c.stft_win_length,
                        normalized=c.stft_normalized, window=window, return_complex=False, length=length)
        y = y.reshape(b, n, s, -1)
        return y[:, 0] if n == 1 else y

    def _axial(self, x):
        for time_transformer, freq_transformer in self.layers:
            b, t, f, d = x.shape
            x = x.permute(0, 2, 1, 3).reshape(b * f, t, d)
            x = time_transformer(x)
            x = x.reshape(b, f, t, d).permute(0, 2, 1, 3).reshape(b * t, f, d)
            x = freq_transformer(x)
            x = x.reshape(b, t, f, d)
        return x


class BSRoformer(_RoformerBase):
    def __init__(self, cfg: RoformerConfig):
        super().__init__()
        self.cfg = cfg
        c = cfg
        time_rot = RotaryEmbedding(c.dim_head)
        freq_rot = RotaryEmbedding(c.dim_head)
        self.layers = nn.ModuleList([
            nn.ModuleList([
                Transformer(c.dim, c.time_transformer_depth, c.heads, c.dim_head, c.ff_mult, time_rot, False),
                Transformer(c.dim, c.freq_transformer_depth, c.heads, c.dim_head, c.ff_mult, freq_rot, False),
            ]) for _ in range(c.depth)
        ])
        self.final_norm = RMSNorm(c.dim)
        assert sum(c.freqs_per_bands) == c.stft_n_fft // 2 + 1
        dims = tuple(2 * f * c.audio_channels for f in c.freqs_per_bands)
        self.band_split = BandSplit(c.dim, dims)
        self.mask_estimators = nn.ModuleList([
            MaskEstimator(c.dim, dims, c.mask_estimator_depth, c.mlp_expansion_factor) for _ in range(c.num_stems)
        ])

    def mask_from_spec(self, stft_repr):
        """stft_repr [b, (f s), t, 2] -> complex mask [b, n, (f s), t]."""
        b, fs, t, _ = stft_repr.shape
        x = stft_repr.permute(0, 2, 1, 3).reshape(b, t, fs * 2)      # 'b f t c -> b t (f c)'
        x = self.band_split(x)
        x = self._axial(x)
        x = self.final_norm(x)
        mask = torch.stack([fn(x) for fn in self.mask_estimators], dim=1)   # b n t (f c)
        mask = mask.reshape(b, len(self.mask_estimators), t, fs, 2).permute(0, 1, 3, 2, 4)
        return torch.view_as_complex(mask.contiguous())

    def forward(self, raw_audio):
        if raw_audio.ndim == 2:
            raw_audio = raw_audio[:, None]
        stft_repr, window = self._stft(raw_audio)
        mask = self.mask_from_spec(stft_repr)
        Y = torch.view_as_complex(stft_repr.contiguous())[:, None] * mask
        return self._istft(Y, window, raw_audio.shape[-1])


class MelBandRoformer(_RoformerBase):
    def __init__(self, cfg: RoformerConfig):
        super().__init__()
        self.cfg = cfg
        c = cfg
        time_rot = RotaryEmbedding(c.dim_head)
        freq_rot = RotaryEmbedding(c.dim_head)
        self.layers = nn.ModuleList([
            nn.ModuleList([
                Transformer(c.dim, c.time_transformer_depth, c.heads, c.dim_head, c.ff_mult, time_rot, True),
                Transformer(c.dim, c.freq_transformer_depth, c.heads, c.dim_head, c.ff_mult, freq_rot, True),
            ]) for _ in range(c.depth)
        ])
        member = torch.from_numpy(mel_band_masks(c.sample_rate, c.stft_n_fft, c.num_bands))
        freqs = member.shape[1]
        rep = torch.arange(freqs)[None].expand(c.num_bands, freqs)
        freq_indices = rep[member]
        if c.stereo:
            freq_indices = (freq_indices[:, None] * 2 + torch.arange(2)[None]).reshape(-1)
        self.register_buffer("freq_indices", freq_indices, persistent=False)
        self.register_buffer("freqs_per_band", member, persistent=False)
        self.register_buffer("num_freqs_per_band", member.sum(dim=1), persistent=False)
        self.register_buffer("num_bands_per_freq", member.sum(dim=0), persistent=False)
        dims = tuple(2 * f * c.audio_channels for f in self.num_freqs_per_band.tolist())
        self.band_split = BandSplit(c.dim, dims)
        self.mask_estimators = nn.ModuleList([
            MaskEstimator(c.dim, dims, c.mask_estimator_depth, c.mlp_expansion_factor) for _ in range(c.num_stems)
        ])

    def mask_from_spec(self, stft_repr):
        """stft_repr [b, (f s), t, 2] -> averaged complex mask [b, n, (f s), t]."""
        c = self.cfg
        b, fs, t, _ = stft_repr.shape
        x = stft_repr[:, self.freq_indices]                              # gather band rows
        x = x.permute(0, 2, 1, 3).reshape(b, t, -1)
        x = self.band_split(x)
        x = self._axial(x)
        n = len(self.mask_estimators)
        masks = torch.stack([fn(x) for fn in self.mask_estimators], dim=1)   # b n t (f c)
        masks = masks.reshape(b, n, t, -1, 2).permute(0, 1, 3, 2, 4)
        masks = torch.view_as_complex(masks.contiguous())                     # b n F' t
        idx = self.freq_indices[None, None, :, None].expand(b, n, -1, t)
        summed = torch.zeros(b, n, fs, t, dtype=masks.dtype).scatter_add_(2, idx, masks)
        denom = torch.repeat_interleave(self.num_bands_per_freq, c.audio_channels)[:, None]
        return summed / denom.clamp(min=1e-8)

    def forward(self, raw_audio):
        if raw_audio.ndim == 2:
            raw_audio = raw_audio[:, None]
        stft_repr, window = self._stft(raw_audio)
        mask = self.mask_from_spec(stft_repr)
        Y = torch.view_as_complex(stft_repr.contiguous())[:, None] * mask
        return self._istft(Y, window, raw_audio.shape[-1])


def build_roformer(cfg: RoformerConfig, seed: int = 4321) -> nn.Module:
    torch.manual_seed(seed)
    model = BSRoformer(cfg) if cfg.kind == "bs" else MelBandRoformer(cfg)
    return model.eval()


# --------------------------------------------------------------------------------------
# chunk loop (A.2, tail-aligned last chunks, Hamming(chunk) weights)
# --------------------------------------------------------------------------------------
def hamming_sym(n: int) -> np.ndarray:
    """scipy.signal.windows.hamming(n) (symmetric)."""
    if n == 1:
        return np.ones(1)
    k = np.arange(n, dtype=np.float64)
    return 0.54 - 0.46 * np.cos(2.0 * np.pi * k / (n - 1))


def chunk_schedule(n: int, chunk: int, step: int) -> List[Tuple[int, int]]:
    """[(read/write offset, multiplicity)] of the A.2 loop; tail-aligned chunks are merged.

    ``for i in range(0, n, step)``: offset ``i`` while ``i + chunk <= n``; every later ``i``
    re-runs the tail chunk at offset ``n - chunk``.
    """
    sched: List[Tuple[int, int]] = []
    tail = 0
    for i in range(0, n, step):
        if i + chunk > n:
            tail += 1
        else:
            sched.append((i, 1))
    if tail:
        if sched and sched[-1][0] == n - chunk:
            sched[-1] = (n - chunk, sched[-1][1] + tail)
        else:
            sched.append((n - chunk, tail))
    return sched


def demix_roformer(mix: torch.Tensor, model: nn.Module, cfg: RoformerConfig) -> torch.Tensor:
    """mix [s, n] -> stems [num_stems, s, n].  SURVEY.md A.2.

    Tracks shorter than one chunk (undefined upstream) are zero-padded to one chunk,
    run once and cropped -- the same rule the CUDA path implements.
    """
    C, step = cfg.chunk_size, cfg.step
    n = mix.shape[-1]
    stems = cfg.num_stems
    with torch.no_grad():
        if n < C:
            part = F.pad(mix, (0, C - n))
            x = model(part[None])[0]
            x = x.reshape(stems, mix.shape[0], C)
            return x[..., :n].clone()
        window = torch.tensor(hamming_sym(C), dtype=torch.float32)
        result = torch.zeros((stems,) + tuple(mix.shape), dtype=torch.float32)
        counter = torch.zeros_like(result)
        for i in range(0, n, step):
            off = i
            if i + C > n:
                off = n - C
            part = mix[:, off: off + C]
            x = model(part[None])[0].reshape(stems, mix.shape[0], C)
            result[..., off: off + C] += x * window
            counter[..., off: off + C] += window
        return result / counter.clamp(min=1e-10)
