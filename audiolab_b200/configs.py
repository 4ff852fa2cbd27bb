"""Architecture / chunking parameters of the separator families on the hot path.

Defaults follow BASELINE.json's configs and SURVEY.md Appendix A (reference call sites:
/root/reference/modules/separator/stem_separator.py:109-121 model list;
/root/reference/modules/rvc/infer/modules/uvr5/mdxnet.py:241-253 MDX parameters).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Tuple

# 24x2, 12x4, 8x12, 8x24, 8x48, 128, 129 -> 62 bands, sum 1025
DEFAULT_FREQS_PER_BANDS: Tuple[int, ...] = (
    (2,) * 24 + (4,) * 12 + (12,) * 8 + (24,) * 8 + (48,) * 8 + (128, 129)
)


@dataclass
class MdxConfig:
    n_fft: int = 6144
    hop: int = 1024
    dim_f: int = 3072
    dim_t_log2: int = 8
    compensate: float = 1.0
    overlap: float = 0.25
    denoise: bool = False
    zero_low_bins: int = 0

    @property
    def dim_t(self) -> int:
        return 2 ** self.dim_t_log2

    @property
    def n_bins(self) -> int:
        return self.n_fft // 2 + 1

    @property
    def chunk_size(self) -> int:
        return self.hop * (self.dim_t - 1)

    @property
    def trim(self) -> int:
        return self.n_fft // 2

    @property
    def gen_size(self) -> int:
        return self.chunk_size - 2 * self.trim


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
    num_bands: int = 60
    sample_rate: int = 44100
    dim_head: int = 64
    heads: int = 8
    ff_mult: int = 4
    stft_n_fft: int = 2048
    stft_hop_length: int = 441
    stft_win_length: int = 2048
    stft_normalized: bool = False
    mask_estimator_depth: int = 2
    mlp_expansion_factor: int = 4
    chunk_size: int = 352800
    num_overlap: int = 4

    @property
    def audio_channels(self) -> int:
        return 2 if self.stereo else 1

    @property
    def step(self) -> int:
        return self.chunk_size // self.num_overlap


@dataclass
class HTDemucsConfig:
    nfft: int = 4096
    samplerate: int = 44100
    segment_num: int = 39
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
