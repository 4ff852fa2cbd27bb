"""TEST INFRASTRUCTURE -- numpy restatement of the reference's post-blend de-bleed (not a product path).

Follows /root/reference/modules/separator/stem_separator.py line by line:
  residual_subtract    :173-239  (+-12 ms cross-correlation probe over <= 1 s, LS gain clipped to [0, 1.25])
  blend_tracks         :241-262
  debleed_instrumental :414-456  (blend the residual into the instrumental only if it lowers |cos| with the vocals)
  pcm16_roundtrip      :57-75    (sf.write PCM_16 then a float load)
Pinned by construction: these ARE the reference's numpy statements, only un-methodised; tests/test_orchestrator.py
compares the device implementation (audiolab_b200/orchestrator.py) against them on seeded inputs.
"""
from __future__ import annotations

from typing import List

import numpy as np


def residual_subtract(base: np.ndarray, component: np.ndarray, sr: int, max_shift_ms: float = 12.0) -> np.ndarray:
    if base.ndim == 1:
        base = np.stack([base, base], axis=0)
    if component.ndim == 1:
        component = np.stack([component, component], axis=0)
    channels = base.shape[0]
    max_shift = max(0, int((max_shift_ms / 1000.0) * float(sr)))
    n = min(base.shape[-1], component.shape[-1])
    residual = np.copy(base)

    def shift_signal(x, lag):
        if lag == 0:
            return x
        if lag > 0:
            return np.concatenate([np.zeros(lag, dtype=x.dtype), x[:-lag]])
        lag = -lag
        return np.concatenate([x[lag:], np.zeros(lag, dtype=x.dtype)])

    for ch in range(channels):
        ref = base[ch, :n]
        sig = component[ch, :n]
        if max_shift > 0 and ref.size > 0 and sig.size > 0:
            probe_len = min(n, 44100)
            corr = np.correlate(ref[:probe_len], sig[:probe_len], mode="full")
            center = len(corr) // 2
            window = corr[center - max_shift:center + max_shift + 1]
            best_rel = int(np.argmax(window)) - max_shift
        else:
            best_rel = 0
        sig_aligned = shift_signal(sig, best_rel)
        denom = float(np.dot(sig_aligned, sig_aligned)) + 1e-8
        alpha = float(np.dot(ref, sig_aligned)) / denom
        alpha = float(np.clip(alpha, 0.0, 1.25))
        residual[ch, :n] = ref - alpha * sig_aligned
    if not np.isfinite(residual).all():
        residual = np.nan_to_num(residual, nan=0.0, posinf=0.0, neginf=0.0)
    return residual


def blend_tracks(tracks: List[np.ndarray], weights: List[float]) -> np.ndarray:
    max_length = max(t.shape[-1] for t in tracks)
    combined = np.zeros((tracks[0].shape[0], max_length), dtype=np.float32)
    total_weight = max(sum(weights), 1e-6)
    for idx, t in enumerate(tracks):
        weight = weights[idx] if idx < len(weights) else 1.0
        combined[:, :t.shape[-1]] += t * float(weight)
    combined = combined / total_weight
    peak = np.max(np.abs(combined))
    if peak > 0:
        combined /= peak
    return combined


def debleed_instrumental(mix: np.ndarray, vocals: np.ndarray, instrumental: np.ndarray, sr: int, blend: float) -> np.ndarray:
    def cosine_abs(a, b):
        a_flat, b_flat = a.reshape(-1), b.reshape(-1)
        denom = (np.linalg.norm(a_flat) * np.linalg.norm(b_flat)) + 1e-8
        return float(abs(np.dot(a_flat, b_flat)) / denom)

    out = instrumental
    resid = residual_subtract(mix, vocals, sr)
    min_len = min(resid.shape[-1], instrumental.shape[-1])
    resid_m = resid[:, :min_len]
    inst = instrumental[:, :min_len]
    sim_inst = cosine_abs(inst, vocals[:, :min_len])
    sim_resid = cosine_abs(resid_m, vocals[:, :min_len])
    if sim_resid + 1e-6 < sim_inst - 0.01:
        blend = 0.0 if blend < 0 else (1.0 if blend > 1.0 else blend)
        inst_refined = (1.0 - blend) * inst + blend * resid_m
        peak = float(np.max(np.abs(inst_refined)))
        if peak > 0.99:
            inst_refined = inst_refined * (0.99 / peak)
        out = inst_refined
    if float(np.max(np.abs(out))) < 1e-6:
        peak = float(np.max(np.abs(resid)))
        out = resid / peak if peak > 1.0 else resid
    return out


def pcm16_roundtrip(x: np.ndarray) -> np.ndarray:
    """libsndfile float -> PCM_16 (lrintf(x * 32768), clipped) and back (/ 32768)."""
    return (np.clip(np.rint(x.astype(np.float32) * np.float32(32768.0)), -32768, 32767) / np.float32(32768.0)).astype(np.float32)
