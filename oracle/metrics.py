"""Parity metrics.  Oracle / test infrastructure only."""
from __future__ import annotations

import numpy as np


def _np(x):
    if hasattr(x, "detach"):
        x = x.detach().to("cpu")
        if x.dtype.is_complex:
            return x.numpy().astype(np.complex128)
        return x.double().numpy()
    return np.asarray(x)


def max_abs_err(a, b) -> float:
    a, b = _np(a), _np(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)))


def rel_err(a, b) -> float:
    """max|a-b| / max|b| -- the scale-free bound used for spectra."""
    a, b = _np(a), _np(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    den = float(np.max(np.abs(b)))
    return float(np.max(np.abs(a - b))) / max(den, 1e-30)


def si_sdr_db(est, ref) -> float:
    """Scale-invariant SDR of ``est`` against ``ref`` in dB (flattened)."""
    est = _np(est).reshape(-1).astype(np.float64)
    ref = _np(ref).reshape(-1).astype(np.float64)
    alpha = float(np.dot(est, ref)) / max(float(np.dot(ref, ref)), 1e-30)
    target = alpha * ref
    noise = est - target
    num = float(np.dot(target, target))
    den = float(np.dot(noise, noise))
    if den == 0.0:
        return float("inf")
    return 10.0 * np.log10(max(num, 1e-30) / den)
