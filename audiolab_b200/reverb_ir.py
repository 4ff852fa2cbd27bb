"""Reverb impulse-response extraction on the device (reference:
/root/reference/handlers/reverb.py:113-172 extract_reverb, called from stem_separator.py:823-829 with the dry / wet outputs
of the de-reverb pass when ``store_reverb_ir`` is set).

Pre-delay by FFT cross-correlation, the impulse response by Wiener deconvolution ``conj(H) Y / (|H|^2 + eps)``, both as ONE
long 1-D real FFT over the whole stem (fp64, like numpy computes the reference's), the RT60 fit of the reference
(``scipy.optimize.curve_fit`` of ``a exp(-b t) + c`` to the envelope in dB) on the host, and the reference's JSON parameter
file.  The stems arrive as the tensors the transform chain already holds ([channels, n], any device): no file round trip.
The long FFTs are ``torch.fft`` (cuFFT on the device) -- a library call, reported as such; the hot-path kernels of this
repository are fixed-size STFT frames."""
from __future__ import annotations

import json
from typing import Dict

import numpy as np
import torch


def fft_xcorr(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """reverb.py:56-67: circular cross-correlation at the next power of two, first len(a) + len(b) - 1 lags."""
    n = a.numel() + b.numel() - 1
    n_fft = 1 << (n - 1).bit_length()
    fa = torch.fft.rfft(a, n=n_fft)
    fb = torch.fft.rfft(b, n=n_fft)
    return torch.fft.irfft(fa * fb.conj(), n=n_fft)[:n]


def wiener_deconvolution(signal: torch.Tensor, kernel: torch.Tensor, epsilon: float = 1e-6) -> torch.Tensor:
    """reverb.py:95-106."""
    h = torch.fft.rfft(kernel, n=signal.numel())
    y = torch.fft.rfft(signal)
    return torch.fft.irfft(h.conj() * y / (h.abs() ** 2 + epsilon))


def estimate_rt60(wet: torch.Tensor, sr: int, curve_fit_maxfev: int = 5000) -> float:
    """reverb.py:70-92: exponential-decay fit to the envelope in dB (host: scipy's Levenberg-Marquardt)."""
    from scipy.optimize import curve_fit
    x = wet                                    # in the stems' own precision (fp32), like numpy does on the loaded samples
    env = (torch.sqrt((x * x).sum(dim=0)) if x.dim() == 2 else x.abs()) + 1e-10
    env_db = (20.0 * torch.log10(env)).cpu().numpy().astype(np.float64)
    time = np.linspace(0, len(env_db) / sr, len(env_db))

    def exp_decay(t, a, b, c):
        return a * np.exp(-b * t) + c

    popt, _ = curve_fit(exp_decay, time, env_db, maxfev=curve_fit_maxfev)
    decay = 3.0 / popt[1] if popt[1] != 0 else 0.5
    return max(float(decay), 0.01)


@torch.no_grad()
def extract_reverb_params(dry: torch.Tensor, wet: torch.Tensor, sr: int, wiener_epsilon: float = 1e-6,
                          curve_fit_maxfev: int = 5000) -> Dict:
    """[channels, n] (or [n]) dry / wet stems -> the reference's parameter dict (reverb.py:118-167)."""
    dry64, wet64 = dry.double(), wet.double()
    dry_mono = dry64.mean(dim=0) if dry64.dim() == 2 else dry64
    wet_mono = wet64.mean(dim=0) if wet64.dim() == 2 else wet64
    corr = fft_xcorr(wet_mono, dry_mono)
    best_shift = max(int(torch.argmax(corr)) - (dry_mono.numel() - 1), 0)
    decay_time = estimate_rt60(wet, sr, curve_fit_maxfev=curve_fit_maxfev)
    ir = wiener_deconvolution(wet_mono, dry_mono, epsilon=wiener_epsilon)[: int(sr * 2)]
    early = int(0.05 * sr)
    early_energy = float((ir[:early] ** 2).sum())
    total_energy = float((ir ** 2).sum()) + 1e-10
    fft_ir = torch.fft.rfft(ir).abs()
    freqs = torch.fft.rfftfreq(ir.numel(), d=1.0 / sr, dtype=torch.float64, device=ir.device)
    return {
        "sample_rate": int(sr),
        "pre_delay": float(best_shift / sr),
        "decay_time": float(decay_time),
        "early_reflection_ratio": early_energy / total_energy,
        "late_reverb_ratio": (total_energy - early_energy) / total_energy,
        "diffusion": float(ir.abs().var(unbiased=False)),
        "spectral_centroid": float((freqs * fft_ir).sum() / (fft_ir.sum() + 1e-10)),
        "impulse_response": ir.cpu().tolist(),
    }


def extract_reverb(dry: torch.Tensor, wet: torch.Tensor, sr: int, param_output_path: str, **kw) -> str:
    """The reference's file product: ``json.dump(params, indent=2)`` (reverb.py:38-41, :169-171)."""
    params = extract_reverb_params(dry, wet, sr, **kw)
    with open(param_output_path, "w") as f:
        json.dump(params, f, indent=2)
    return param_output_path
