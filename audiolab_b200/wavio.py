"""WAV file I/O at the drop-in boundary (the reference moves audio between L2 and L1 as WAV files,
/root/reference/modules/separator/stem_separator.py:57-75, 278-282, 625-677).  soundfile / librosa are
not available here; scipy.io.wavfile covers PCM_16 / PCM_32 / FLOAT."""
from __future__ import annotations

import numpy as np
from scipy.io import wavfile


def read_wav(path: str):
    """-> (audio float32 [channels, n], sample_rate)."""
    sr, data = wavfile.read(path)
    if data.ndim == 1:
        data = data[:, None]
    if data.dtype == np.int16:
        x = data.astype(np.float32) / 32768.0
    elif data.dtype == np.int32:
        x = data.astype(np.float32) / 2147483648.0
    elif data.dtype == np.uint8:
        x = (data.astype(np.float32) - 128.0) / 128.0
    else:
        x = data.astype(np.float32)
    return np.ascontiguousarray(x.T), int(sr)


def write_wav(path: str, audio: np.ndarray, sr: int, subtype: str = "FLOAT") -> None:
    """audio [channels, n] float -> WAV.  subtype FLOAT (stem_separator.py:669) or PCM_16 (:74)."""
    x = np.asarray(audio, dtype=np.float32)
    if x.ndim == 1:
        x = x[None]
    x = np.ascontiguousarray(x.T)
    if subtype == "PCM_16":
        x = np.clip(np.round(x * 32768.0), -32768, 32767).astype(np.int16)
    elif subtype != "FLOAT":
        raise ValueError(f"unsupported subtype {subtype}")
    wavfile.write(path, int(sr), x)
