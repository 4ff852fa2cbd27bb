"""Oracle (test infrastructure, never imported by the product): numpy restatement of the reference's reverb impulse-response
extraction (reference: /root/reference/handlers/reverb.py:52-53 to_mono, :56-67 fft_xcorr, :70-92 estimate_rt60, :95-106
wiener_deconvolution, :113-172 extract_reverb; called from stem_separator.py:823-829 on the (No Reverb, Reverb) outputs of the
de-reverb pass when store_reverb_ir is set).  Signals are [n] or [n, channels] float32 arrays, like soundfile / pydub give."""
import numpy as np
from scipy.optimize import curve_fit


def to_mono(signal):
    return np.mean(signal, axis=1) if signal.ndim == 2 else signal


def fft_xcorr(a, b):
    n = len(a) + len(b) - 1
    n_fft = 1 << (n - 1).bit_length()
    fa = np.fft.rfft(a, n=n_fft)
    fb = np.fft.rfft(b, n=n_fft)
    return np.fft.irfft(fa * np.conjugate(fb), n=n_fft)[:n]


def estimate_rt60(signal, sr, curve_fit_maxfev=5000):
    eps = 1e-10
    env = (np.sqrt(np.sum(signal ** 2, axis=1)) if signal.ndim == 2 else np.abs(signal)) + eps
    env_db = 20.0 * np.log10(env)
    time = np.linspace(0, len(env) / sr, len(env))

    def exp_decay(x, a, b, c):
        return a * np.exp(-b * x) + c

    popt, _ = curve_fit(exp_decay, time, env_db, maxfev=curve_fit_maxfev)
    decay_time = 3.0 / popt[1] if popt[1] != 0 else 0.5
    return max(decay_time, 0.01)


def wiener_deconvolution(signal, filter_kernel, epsilon=1e-6):
    h = np.fft.rfft(filter_kernel, len(signal))
    y = np.fft.rfft(signal)
    return np.fft.irfft((np.conjugate(h) * y) / (np.abs(h) ** 2 + epsilon))


def extract_params(dry_signal, wet_signal, sr, wiener_epsilon=1e-6, curve_fit_maxfev=5000):
    dry_mono, wet_mono = to_mono(dry_signal), to_mono(wet_signal)
    corr = fft_xcorr(wet_mono, dry_mono)
    best_shift = max(int(np.argmax(corr)) - (len(dry_mono) - 1), 0)
    decay_time = estimate_rt60(wet_signal, sr, curve_fit_maxfev=curve_fit_maxfev)
    ir = wiener_deconvolution(wet_mono, dry_mono, epsilon=wiener_epsilon)[: int(sr * 2)]
    early = int(0.05 * sr)
    early_energy = np.sum(np.square(ir[:early]))
    total_energy = np.sum(np.square(ir)) + 1e-10
    fft_ir = np.abs(np.fft.rfft(ir))
    freqs = np.fft.rfftfreq(len(ir), d=1.0 / sr)
    return {
        "sample_rate": sr,
        "pre_delay": float(best_shift / sr),
        "decay_time": float(decay_time),
        "early_reflection_ratio": float(early_energy / total_energy),
        "late_reverb_ratio": float((total_energy - early_energy) / total_energy),
        "diffusion": float(np.var(np.abs(ir))),
        "spectral_centroid": float(np.sum(freqs * fft_ir) / (np.sum(fft_ir) + 1e-10)),
        "impulse_response": ir.tolist(),
    }
