"""Polyphase resampler, CPU oracle.  Test infrastructure only.

The reference resamples on load with ``librosa.load(sr=44100)``
(/root/reference/modules/separator/stem_separator.py:865) and names ``polyphase`` as the
resampler of the VR band splits (/root/reference/modules/rvc/infer/lib/uvr5_pack/lib_v5/
model_param_init.py:22); librosa's ``res_type="polyphase"`` is
``scipy.signal.resample_poly``.  scipy is present, so the oracle calls it directly.
PARITY UNPINNED by any reference test (the reference has none).
"""
from __future__ import annotations

import numpy as np
from scipy import signal


def resample_poly_ref(x: np.ndarray, up: int = 147, down: int = 160) -> np.ndarray:
    return signal.resample_poly(np.asarray(x, dtype=np.float32), up, down, axis=-1).astype(np.float32)


def design_taps(up: int, down: int) -> np.ndarray:
    """The FIR ``resample_poly`` designs internally: firwin(2*10*max+1, 1/max, kaiser 5.0) * up.

    Returned in float64; scipy casts it to the input dtype (float32) before ``upfirdn``.
    """
    g = np.gcd(up, down)
    up, down = up // g, down // g
    max_rate = max(up, down)
    f_c = 1.0 / max_rate
    half_len = 10 * max_rate
    h = signal.firwin(2 * half_len + 1, f_c, window=("kaiser", 5.0))
    return h * up
