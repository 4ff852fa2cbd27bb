#!/usr/bin/env python
"""Realtime factor of the other BASELINE.json configurations on ONE B200, through the public in-memory API
(`Separator.separate_tensor`, host mix in, stems left on the device) -- parity-test shapes, not the contract line
(bench.py measures configs[1]).  Seeded random-init networks, synthetic audio, CUDA events, one warm-up pass.

    python tools/config_bench.py [--budget-s 30]

cfg1  UVR-MDX-NET spectral path: 30 s stereo clip, n_fft 6144 / hop 1024, 256-frame chunks (windowed-OLA form, overlap 0.25)
cfg3  HTDemucs 4-stem hybrid (n_fft 4096 / hop 1024 + waveform branch), shifts 2, 10-minute track (2 minutes if a
      10-minute pass would not fit the time budget)
cfg4  MDX-Net on a 3-minute 48 kHz song incl. the 48k -> 44.1k polyphase resample (one GPU's share of the 512-song batch)
cfg5  Mel-Band RoFormer on a 5-minute track (the single-GPU slice of the 60-minute chunk-range case)
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def synth(n, sr, seed):
    """10 s of seeded synthetic stereo (sinusoids + noise, peak 0.9), tiled to n samples (content does not matter here)."""
    base_n = min(n, 10 * sr)
    base = _synth(base_n, sr, seed)
    reps = (n + base_n - 1) // base_n
    return base.repeat(1, reps)[:, :n].contiguous().pin_memory()


def _synth(n, sr, seed):
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(n, dtype=torch.float64) / sr
    out = torch.zeros((2, n), dtype=torch.float64)
    for c in range(2):
        f = torch.exp(torch.empty(8).uniform_(3.9, 9.6, generator=g)).double()
        ph = torch.empty(8).uniform_(0, 6.283, generator=g).double()
        out[c] = 0.25 * torch.sin(2 * torch.pi * f[:, None] * t[None] + ph[:, None]).sum(0)
        out[c] += 0.05 * (torch.rand(n, generator=g, dtype=torch.float64) * 2 - 1)
    out *= 0.9 / out.abs().max()
    return out.float()


def timed(sep, mix, sr, budget_s):
    from audiolab_b200 import _lib
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    sep.separate_tensor(mix, sr)                      # warm-up (cuDNN / cuBLAS heuristics, plan tables)
    torch.cuda.synchronize()
    warm = time.perf_counter() - t0
    reps = max(1, min(3, int(budget_s / max(warm, 1e-3))))
    n0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = sep.separate_tensor(mix, sr)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    return ms, reps, (_lib.launch_count() - n0) // reps, {k: tuple(v.shape) for k, v in out.items()}, warm


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--budget-s", type=float, default=30.0)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    from audiolab_b200.separator import Separator
    cases = [
        ("cfg1_mdx_30s", "UVR-MDX-NET-Inst_HQ_3.onnx", 44100, [30], dict(mdx_params={"batch_size": 8})),
        ("cfg3_htdemucs", "htdemucs.yaml", 44100, [120, 600], dict()),
        ("cfg4_mdx_48k_song", "UVR-MDX-NET-Inst_HQ_3.onnx", 48000, [180], dict(mdx_params={"batch_size": 16})),
        ("cfg5_mel_roformer", "model_mel_band_roformer_ep_3005_sdr_11.4360.ckpt", 44100, [300],
         dict(mdxc_params={"batch_size": 27, "overlap": 4})),
    ]
    for name, model, sr, seconds, kw in cases:
        if args.only and args.only not in name:
            continue
        try:
            sep = Separator(log_level=40, allow_random_init=True, use_autocast=True, device="cuda:0", **kw)
            sep.load_model(model)
            last = None
            for secs in seconds:
                if last is not None and last * (secs / seconds[0]) > args.budget_s * 1000:
                    print(json.dumps({"config": name, "audio_s": secs, "skipped": "would exceed the time budget"}), flush=True)
                    continue
                mix = synth(secs * sr, sr, seed=1234 + secs)
                ms, reps, launches, shapes, warm = timed(sep, mix, sr, args.budget_s)
                last = ms
                print(json.dumps({"config": name, "model": model, "input_sr": sr, "audio_s": secs, "ms": round(ms, 2),
                                  "realtime_factor": round(secs / (ms / 1e3), 1), "reps": reps, "warmup_s": round(warm, 2),
                                  "al_kernel_launches_per_pass": launches, "stems": shapes,
                                  "timed": "H2D of the mix + resample (if any) + demix; stems stay on the device"}),
                      flush=True)
                del mix
            del sep
            torch.cuda.empty_cache()
        except Exception as e:                                  # keep the other configurations running
            print(json.dumps({"config": name, "error": f"{type(e).__name__}: {e}"[:400]}), flush=True)


if __name__ == "__main__":
    main()
