"""Seeded synthetic inputs.  Oracle / test infrastructure only.

Stereo fp32 ``[2, N]`` = 0.25 * (sum of 8 sinusoids with log-uniform 50 Hz - 16 kHz
frequencies and random phases, different per channel) + 0.05 * uniform(-1, 1) noise,
peak-normalised to 0.9.  ``numpy.random.RandomState`` (MT19937) keeps the stream
stable across numpy versions so golden fixtures can store only the seed.
"""
from __future__ import annotations

import numpy as np

SR = 44100


def synth_mix(n_samples: int, seed: int = 1234, sr: int = SR, channels: int = 2) -> np.ndarray:
    rs = np.random.RandomState(seed)
    t = np.arange(n_samples, dtype=np.float64) / sr
    out = np.zeros((channels, n_samples), dtype=np.float64)
    for c in range(channels):
        freqs = np.exp(rs.uniform(np.log(50.0), np.log(16000.0), size=8))
        phases = rs.uniform(0.0, 2.0 * np.pi, size=8)
        for f, p in zip(freqs, phases):
            out[c] += np.sin(2.0 * np.pi * f * t + p)
        out[c] *= 0.25
        out[c] += 0.05 * rs.uniform(-1.0, 1.0, size=n_samples)
    peak = np.abs(out).max()
    if peak > 0:
        out *= 0.9 / peak
    return out.astype(np.float32)


def synth_noise(shape, seed: int = 0, scale: float = 1.0) -> np.ndarray:
    rs = np.random.RandomState(seed)
    return (scale * rs.standard_normal(size=shape)).astype(np.float32)
