#!/usr/bin/env python
"""bench.py -- realtime factor of the chunked STFT -> mask -> iSTFT + OLA demix (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--mode auto|tracks|chunk-range]
                    [--configs all|none|cfg1,cfg3,cfg4,cfg5]

Workload of the contract line (BASELINE.json configs[1]): BS-RoFormer vocals/instrumental, dim 512 / depth 12,
n_fft 2048, hop 441, stereo, 8 s chunks with overlap 4 (step 2 s), bf16 random-init weights, 60 s of 44.1 kHz
synthetic audio per GPU.  One "step" = one full demix of that audio.

  value : audio-seconds / second, whole job, mix already resident in HBM, CUDA events, max over ranks
  e2e   : the same through the public API (Separator.separate_tensor) from a pinned HOST buffer, with
          the host->device copy of the mix and the device->host copy of both stems inside the timed
          region; `e2e_process_audio` = the file-based plugin call Separate.process_audio (WAV in, WAV stems out)
  roofline     : the dominant spectral kernel (al_istft, fused mask multiply + iFFT + OLA), algorithmic
                 bytes / CUDA-event duration of its launches inside the timed steps, vs the measured
                 HBM peak in MEASURED_PEAKS.json
  cpu_baseline : the oracle (CPU restatement of the reference) on the host cores: one 8 s chunk evaluation,
                 1 warm-up + best of 3, all host threads; plus the spectral-only part on 1 thread
  configs      : the other BASELINE.json configurations (cfg1 MDX-Net 30 s, cfg3 HTDemucs 10 min, cfg4 MDX-Net +
                 48k->44.1k resample on 64 songs per GPU, cfg5 Mel-RoFormer 60 min split by chunk range), each with its
                 realtime factor and the roofline fraction of its spectral kernels
N > 1 (torchrun): the default mode is chunk-range -- ONE 60*N s track split by chunk range over the ranks with the
OLA halo exchange over NCCL (weak scaling: 60 s of audio per GPU, the only mode with a data-path collective); the
track-per-GPU figure (no collective) is reported beside it as `tracks_mode`.
--impl reference: the oracle on the host cores (rank 0 only), same metric / unit / config.
"""
from __future__ import annotations

import argparse
import dataclasses
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SR = 44100
TRACK_SECONDS = 60
PARITY_DB = {"fp16": 63.61, "bf16": 44.69}     # SI-SDR vs the fp32 oracle at the real size (profiles/r02i_parity_fullsize_fp16.jsonl)
METRIC = "realtime factor (audio-s/s) of chunked STFT->mask->iSTFT+OLA demix"
UNIT = "audio-s/s"


def workload_config(n_gpus: int, mode: str) -> dict:
    if n_gpus > 1 and mode == "chunk-range":
        sharding = (f"chunk-range: one {TRACK_SECONDS * n_gpus} s track split by chunk range over {n_gpus} ranks, "
                    "OLA halo exchange (isend/irecv over NCCL) overlapped with the interior chunks")
    elif n_gpus > 1:
        sharding = "track-per-GPU (no data-path collective)"
    else:
        sharding = "none"
    return {
        "workload": "BS-RoFormer vocals/instrumental (dim 512, depth 12, 62 bands), n_fft=2048 hop=441 stereo, "
                    "8 s chunks overlap=4, 16-bit random-init weights (seed 4321; fp16 operands by default, bf16 selectable -- "
                    "see precision_modes), "
                    f"{TRACK_SECONDS} s of 44.1 kHz synthetic audio per GPU (BASELINE.json configs[1])",
        "n_fft": 2048, "hop": 441, "chunk_samples": 352800, "overlap": 4, "track_seconds": TRACK_SECONDS,
        "sharding": sharding,
        "l2_policy": "inputs larger than L2: every step streams 27 chunks x (13 MB spectrum + 13 MB mask) plus "
                     "~GBs of network activations through the 126 MB L2 between two uses of any buffer",
    }


def peaks() -> dict:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "source": "fallback (B200_PROFILING.md)"}


# ------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int = 0):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except (FileNotFoundError, OSError):
            self.proc = None
            return
        self.thread = threading.Thread(target=self._pump, daemon=True)
        self.thread.start()

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# analytic work counters
# ------------------------------------------------------------------------------------------------
def net_flops_per_chunk(cfg) -> float:
    """Dense-layer + attention FLOPs of one BS-RoFormer chunk evaluation (2 x MACs)."""
    T = 1 + cfg.chunk_size // cfg.stft_hop_length
    bands = list(cfg.freqs_per_bands)
    nb, d, inner = len(bands), cfg.dim, cfg.heads * cfg.dim_head
    ch = cfg.audio_channels
    tokens = T * nb
    per_layer = d * inner * 3 + d * cfg.heads + inner * d + 2 * d * d * cfg.ff_mult
    n_time = cfg.depth * cfg.time_transformer_depth
    n_freq = cfg.depth * cfg.freq_transformer_depth
    flops = 2.0 * tokens * per_layer * (n_time + n_freq)
    flops += n_time * nb * (4.0 * T * T * inner)             # QK^T and PV over time
    flops += n_freq * T * (4.0 * nb * nb * inner)            # over bands
    for bw in bands:
        din = 2 * bw * ch
        flops += 2.0 * T * din * d                            # band split
        hidden = d * cfg.mlp_expansion_factor
        flops += cfg.num_stems * 2.0 * T * (d * hidden + hidden * 2 * din)   # mask estimator (depth 2)
    return flops


def k2_algorithmic_bytes(cfg, n_chunks: int) -> float:
    """al_istft per launch: spectrum + mask read (complex64), chunk wave written."""
    T = 1 + cfg.chunk_size // cfg.stft_hop_length
    F = cfg.stft_n_fft // 2 + 1
    ch, st = cfg.audio_channels, cfg.num_stems
    return n_chunks * (ch * T * F * 8.0 + st * ch * T * F * 8.0 + st * ch * cfg.chunk_size * 4.0)


def k1_algorithmic_bytes(cfg, n_chunks: int) -> float:
    T = 1 + cfg.chunk_size // cfg.stft_hop_length
    F = cfg.stft_n_fft // 2 + 1
    return n_chunks * cfg.audio_channels * (cfg.chunk_size * 4.0 + T * F * 8.0)


# ------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle on the host cores
# ------------------------------------------------------------------------------------------------
def oracle_chunk_seconds(n_evals: int, warm: int) -> dict:
    """Time `n_evals` chunk evaluations (STFT -> net -> mask (.) STFT -> iSTFT -> Hamming weight) of the
    oracle BS-RoFormer (fp32, CPU, all host threads) after `warm` untimed ones.  The reference loop makes one such
    evaluation per `step` = chunk/4 = 2 s of audio, whatever the track length, so RTF = 2 s / t_eval."""
    import torch

    from oracle import roformer as oro
    from oracle.synth import synth_mix
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = oro.RoformerConfig()
    model = oro.build_roformer(cfg, seed=4321)
    mix = torch.tensor(synth_mix(cfg.chunk_size, seed=1236))
    window = torch.tensor(oro.hamming_sym(cfg.chunk_size), dtype=torch.float32)
    times = []
    with torch.no_grad():
        for i in range(warm + n_evals):
            t0 = time.perf_counter()
            x = model(mix[None])[0]
            _ = x * window
            dt = time.perf_counter() - t0
            if i >= warm:
                times.append(dt)
    return {"seconds": times, "cores": torch.get_num_threads(), "step_audio_s": cfg.step / SR}


def oracle_spectral_seconds(threads: int, reps: int = 3) -> float:
    """The spectral part alone (torch.stft -> 0.5 * mask -> torch.istft -> Hamming weight on one 8 s chunk) on
    `threads` host threads: best of `reps` after one warm-up (SURVEY.md 8d asks for a 1-thread figure)."""
    import torch

    from oracle import roformer as oro
    from oracle.synth import synth_mix
    cfg = oro.RoformerConfig()
    mix = torch.tensor(synth_mix(cfg.chunk_size, seed=1236))
    win = torch.hann_window(cfg.stft_n_fft)
    weight = torch.tensor(oro.hamming_sym(cfg.chunk_size), dtype=torch.float32)
    prev = torch.get_num_threads()
    torch.set_num_threads(threads)
    best = float("inf")
    try:
        with torch.no_grad():
            for i in range(reps + 1):
                t0 = time.perf_counter()
                spec = torch.stft(mix, cfg.stft_n_fft, cfg.stft_hop_length, window=win, return_complex=True)
                y = torch.istft(spec * 0.5, cfg.stft_n_fft, cfg.stft_hop_length, window=win, length=cfg.chunk_size)
                _ = y * weight
                dt = time.perf_counter() - t0
                if i > 0:
                    best = min(best, dt)
    finally:
        torch.set_num_threads(prev)
    return best


def cpu_baseline_block() -> dict:
    res = oracle_chunk_seconds(3, 1)
    best = min(res["seconds"])
    one = oracle_spectral_seconds(1)
    return {"value": res["step_audio_s"] / best, "unit": UNIT, "cores": res["cores"], "kind": "port",
            "sample": "one 8 s chunk evaluation of the oracle (fp32, all host threads) = the work the reference loop does "
                      "per 2 s of audio; 1 warm-up, best of 3",
            "seconds": [round(t, 3) for t in res["seconds"]],
            "spectral_only_1_thread": {"value": res["step_audio_s"] / one, "unit": UNIT, "cores": 1,
                                       "sample": "torch.stft -> mask -> torch.istft -> Hamming weight of one 8 s chunk, no "
                                                 "network, 1 thread, best of 3"}}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    budget_s = float(os.environ.get("AUDIOLAB_REF_BUDGET_S", "240"))
    probe = oracle_chunk_seconds(1, 0)                      # doubles as the first warm-up evaluation
    t1 = probe["seconds"][0]
    warm = max(0, min(args.warmup - 1, int(budget_s * 0.25 / t1)))
    steps = max(1, min(args.steps, int(budget_s * 0.75 / t1)))
    res = oracle_chunk_seconds(steps, warm)
    t = statistics.mean(res["seconds"])
    rtf = res["step_audio_s"] / t
    sample = ("one chunk evaluation per step (8 s chunk = the unit the reference loop runs once per 2 s of audio); "
              f"{steps} of the requested {args.steps} steps timed within a {budget_s:.0f} s budget")
    line = {
        "impl": "reference", "metric": METRIC, "value": rtf, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warm + 1, "ms_per_step": 1e3 * t, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus, ("chunk-range" if args.gpus > 1 else "tracks") if args.mode == "auto" else args.mode),
        "cpu_baseline": {"value": rtf, "unit": UNIT, "cores": res["cores"], "kind": "port", "sample": sample},
        "e2e": {"value": rtf, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def synth_mix(n_samples: int, seed: int):
    """Seeded synthetic stereo mix: 8 log-spaced sinusoids per channel + uniform noise,
    peak 0.9.  Kept local so that this arm imports nothing from oracle/."""
    import numpy as np
    rs = np.random.RandomState(seed)
    t = np.arange(n_samples, dtype=np.float64) / SR
    out = np.zeros((2, n_samples), dtype=np.float64)
    for c in range(2):
        freqs = np.exp(rs.uniform(np.log(50.0), np.log(16000.0), size=8))
        phases = rs.uniform(0.0, 2.0 * np.pi, size=8)
        for f, p in zip(freqs, phases):
            out[c] += np.sin(2.0 * np.pi * f * t + p)
        out[c] *= 0.25
        out[c] += 0.05 * rs.uniform(-1.0, 1.0, size=n_samples)
    out *= 0.9 / np.abs(out).max()
    return out.astype(np.float32)


class TimedPlan:
    """Wraps StftPlan to bracket every al_stft / al_istft launch with CUDA events on the launch stream."""

    def __init__(self, plan, torch):
        self._plan, self._torch = plan, torch
        self.events = {"stft": [], "istft": []}
        self.enabled = False

    def __getattr__(self, name):
        return getattr(self._plan, name)

    def _timed(self, key, fn, *a, **kw):
        if not self.enabled:
            return fn(*a, **kw)
        e0, e1 = self._torch.cuda.Event(enable_timing=True), self._torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*a, **kw)
        e1.record()
        self.events[key].append((e0, e1, kw.get("n_chunks", 1)))
        return out

    def stft(self, *a, **kw):
        return self._timed("stft", self._plan.stft, *a, **kw)

    def istft(self, *a, **kw):
        return self._timed("istft", self._plan.istft, *a, **kw)

    def reset(self):
        self.events = {"stft": [], "istft": []}


def kernel_frac(plan: "TimedPlan", key: str, bytes_per_chunk: float, hbm_gbs: float):
    """achieved GB/s (algorithmic bytes / CUDA-event time) of the bracketed launches of one kernel and its fraction of
    the measured HBM peak."""
    ev = plan.events[key]
    if not ev:
        return None
    ms = sum(a.elapsed_time(b) for a, b, _ in ev)
    byt = sum(bytes_per_chunk * nc for _, _, nc in ev)
    gbs = byt / (ms / 1e3) / 1e9
    return {"launches": len(ev), "achieved_gbs": round(gbs, 1), "frac": round(gbs / hbm_gbs, 4), "total_ms": round(ms, 3)}


# ------------------------------------------------------------------------------------------------
# the other BASELINE.json configurations (each: realtime factor + roofline fraction of its spectral kernels)
# ------------------------------------------------------------------------------------------------
def _tile_synth(n: int, seed: int):
    """10 s of seeded synthetic stereo tiled to n samples (content does not matter for timing), pinned."""
    import torch
    base_n = min(n, 10 * SR)
    base = torch.from_numpy(synth_mix(base_n, seed=seed))
    reps = (n + base_n - 1) // base_n
    return base.repeat(1, reps)[:, :n].contiguous().pin_memory()


def _timed_passes(torch, fn, reps: int) -> float:
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def config_cfg1(torch, dev, pk) -> dict:
    """cfg1: UVR-MDX-NET spectral path, 30 s stereo clip, n_fft 6144 / hop 1024, 256-frame chunks, full-size (L = 11)
    TFC-TDF net with seeded random weights; trim-concat form (6 chunks) as in the in-tree twin mdxnet.py:143-197."""
    from audiolab_b200.separator import Separator
    sep = Separator(log_level=40, allow_random_init=True, use_autocast=True, device=str(dev), mdx_params={"batch_size": 6})
    inst = sep.load_model("UVR-MDX-NET-Voc_FT.onnx")
    d = inst.demixer
    d.plan = TimedPlan(d.plan, torch)
    n = 30 * SR
    mix = torch.from_numpy(synth_mix(n, seed=1235)).to(dev)
    run = lambda: d.demix_trim_concat(mix)
    run()
    d.plan.enabled = True
    ms = _timed_passes(torch, run, 3)
    c = d.cfg
    k1 = kernel_frac(d.plan, "stft", 2 * c.chunk_size * 4.0 + 4 * c.dim_f * c.dim_t * 4.0, pk["hbm_gbs"])
    k2 = kernel_frac(d.plan, "istft", 4 * c.dim_f * c.dim_t * 4.0 + 2 * c.gen_size * 4.0, pk["hbm_gbs"])
    return {"workload": "MDX-Net (TFC-TDF L=11, g=48, random init), 30 s stereo, n_fft 6144 hop 1024, 256-frame chunks, "
                        "trim-concat demix (6 chunks)", "audio_s": 30, "ms": round(ms, 3), "value": round(30e3 / ms, 1),
            "unit": UNIT, "al_stft": k1, "al_istft": k2}


def config_cfg3(torch, dev, pk) -> dict:
    """cfg3: HTDemucs 4-stem hybrid (STFT n_fft 4096 / hop 1024 + waveform branch), 10-minute track, shifts 1."""
    from audiolab_b200.separator import Separator
    sep = Separator(log_level=40, allow_random_init=True, use_autocast=True, device=str(dev), demucs_params={"shifts": 1})
    inst = sep.load_model("htdemucs.yaml")
    d = inst.demixer
    d.plan = TimedPlan(d.plan, torch)
    seconds = 600
    mix = _tile_synth(seconds * SR, 1237).to(dev)
    run = lambda: d.demix(mix)
    d.demix(mix[:, : 30 * SR].contiguous())          # warm-up on a short track (cuDNN heuristics, plan tables)
    d.plan.enabled = True
    ms = _timed_passes(torch, run, 1)
    c = d.cfg
    seg, le = c.segment_samples, -(-c.segment_samples // c.hop)
    k1 = kernel_frac(d.plan, "stft", 2 * seg * 4.0 + 2 * 2048 * le * 8.0, pk["hbm_gbs"])
    k2 = kernel_frac(d.plan, "istft", c.num_sources * (2 * 2048 * le * 8.0 + 2 * seg * 4.0), pk["hbm_gbs"])
    return {"workload": "HTDemucs 4-stem hybrid (random init), n_fft 4096 hop 1024 + waveform branch, 10-minute track, "
                        "segments of 7.8 s, overlap 0.25, shifts 1", "audio_s": seconds, "ms": round(ms, 2),
            "value": round(seconds * 1e3 / ms, 1), "unit": UNIT, "al_stft": k1, "al_istft": k2}


def config_cfg4(torch, dist, dev, pk, rank: int, world: int, max_over_ranks) -> dict:
    """cfg4: batch of 3-minute 48 kHz songs sharded track-per-GPU (64 songs per GPU; 512 on 8 GPUs): host -> device copy,
    48k -> 44.1k polyphase resample (K3), MDX-Net windowed-OLA demix.  No data-path collective."""
    from audiolab_b200.separator import Separator
    from audiolab_b200.sharding import assign_tracks
    songs_per_gpu = int(os.environ.get("AUDIOLAB_CFG4_SONGS_PER_GPU", "64"))
    n_songs = songs_per_gpu * world
    n_in = 180 * 48000
    mine = assign_tracks([n_in] * n_songs, world)[rank]
    sep = Separator(log_level=40, allow_random_init=True, use_autocast=True, device=str(dev), mdx_params={"batch_size": 16})
    sep.load_model("UVR-MDX-NET-Voc_FT.onnx")
    hosts = [_tile_synth(n_in, 1300 + i) for i in range(2)]       # two pinned songs, alternated (content is irrelevant)
    ev = []

    def one_song(i, timed):
        x = hosts[i & 1].to(dev, non_blocking=True)
        if timed:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        y = sep.prepare_mix(x, 48000)                              # K3: 48k -> 44.1k
        if timed:
            e1.record()
            ev.append((e0, e1))
        return sep.model_instance.run(y)

    one_song(0, False)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for j, _ in enumerate(mine):
        one_song(j, True)
    t1.record()
    torch.cuda.synchronize()
    ms = max_over_ranks(t0.elapsed_time(t1))
    res_ms = sum(a.elapsed_time(b) for a, b in ev) / max(len(ev), 1)
    n_out = 180 * 44100
    res_gbs = 2 * (n_in + n_out) * 4.0 / (res_ms / 1e3) / 1e9
    audio_s = 180.0 * n_songs
    return {"workload": f"{n_songs} x 3-minute 48 kHz songs, track-per-GPU ({songs_per_gpu} per GPU, assign_tracks), "
                        "48k->44.1k polyphase resample + MDX-Net (TFC-TDF L=11, random init) windowed-OLA demix",
            "audio_s": audio_s, "ms": round(ms, 1), "value": round(audio_s * 1e3 / ms, 1), "unit": UNIT,
            "al_resample": {"avg_ms": round(res_ms, 3), "achieved_gbs": round(res_gbs, 1), "frac": round(res_gbs / pk["hbm_gbs"], 4)}}


def config_cfg5(torch, dist, dev, pk, rank: int, world: int, max_over_ranks) -> dict:
    """cfg5: one 60-minute track, Mel-RoFormer, split by chunk range over the ranks with the OLA halo exchange (strong
    scaling: the track is the same at every N)."""
    from audiolab_b200.separator import Separator
    from audiolab_b200.sharding import ShardedRoformerDemixer
    minutes = int(os.environ.get("AUDIOLAB_CFG5_MINUTES", "60"))
    sep = Separator(log_level=40, allow_random_init=True, use_autocast=True, device=str(dev),
                    mdxc_params={"batch_size": 27, "overlap": 4})
    inst = sep.load_model("vocals_mel_band_roformer.ckpt")
    d = inst.demixer
    n = minutes * 60 * SR
    mix = _tile_synth(n, 1238).to(dev)
    sharded = ShardedRoformerDemixer(d, rank, world) if world > 1 else None
    run = (lambda: sharded.demix_span(mix)[0]) if sharded is not None else (lambda: d.demix(mix))
    warm = mix[:, : 40 * SR * max(world, 1)].contiguous()          # short warm-up of the same code path
    (ShardedRoformerDemixer(d, rank, world).demix_span(warm) if world > 1 else d.demix(warm))
    if world > 1:
        dist.barrier()
    ms = max_over_ranks(_timed_passes(torch, run, 1))
    c = d.cfg
    return {"workload": f"Mel-Band RoFormer (dim {c.dim}, depth {c.depth}, {c.num_bands} bands, random init), one "
                        f"{minutes}-minute track, 8 s chunks overlap 4, "
                        + (f"chunk ranges over {world} GPUs + NCCL halo exchange" if world > 1 else "single GPU"),
            "audio_s": minutes * 60, "ms": round(ms, 1), "value": round(minutes * 60e3 / ms, 1), "unit": UNIT,
            "scaling": "strong"}


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    from audiolab_b200 import _lib, netops
    from audiolab_b200 import spectral as sp
    from audiolab_b200.configs import RoformerConfig
    from audiolab_b200.demix import _dev_i32, _dev_i64, roformer_schedule
    from audiolab_b200.separator import Separator
    from audiolab_b200.sharding import ShardedRoformerDemixer, plan_chunk_ranges

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N > 1")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep NCCL's version banner off stdout (one JSON line)
        dist.init_process_group("nccl", device_id=dev)
    mode = args.mode
    if mode == "auto":
        mode = "chunk-range" if world > 1 else "tracks"
    pk = peaks()

    sep = Separator(log_level=40, allow_random_init=True, use_autocast=True, device=str(dev),
                    mdxc_params={"batch_size": args.batch, "overlap": 4})
    inst = sep.load_model("model_bs_roformer_ep_368_sdr_12.9628.ckpt")
    demixer = inst.demixer
    cfg: RoformerConfig = demixer.cfg
    demixer.plan = TimedPlan(demixer.plan, torch)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    track_seconds = TRACK_SECONDS
    if args.profile_mode and args.track_seconds:
        track_seconds = args.track_seconds     # shorter launch list under ncu; never a bench value

    def measure(step_mode: str, steps: int, n_warm: int, timed_plan: bool):
        """Device-resident timing of `steps` demixes in `step_mode` ('tracks' or 'chunk-range')."""
        chunk_range = step_mode == "chunk-range" and world > 1
        n = track_seconds * SR * (world if chunk_range else 1)
        seed = 1236 + (0 if chunk_range else rank)
        mix_host = torch.from_numpy(synth_mix(n, seed=seed)).pin_memory()
        mix_dev = mix_host.to(dev)
        sharded = ShardedRoformerDemixer(demixer, rank, world) if chunk_range else None
        stats = {}

        def step_device():
            if sharded is not None:
                return sharded.demix_span(mix_dev, stats)[0]
            return demixer.demix(mix_dev)

        for _ in range(n_warm):
            step_device()
        barrier()
        if timed_plan:
            demixer.plan.reset()
            demixer.plan.enabled = True
            netops.gemm_timer = []
        stats["time_wait"] = True
        launches0 = _lib.launch_count()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if args.profile_mode:
            torch.cuda.cudart().cudaProfilerStart()          # ncu --profile-from-start off lists the timed steps only
        e0.record()
        for _ in range(steps):
            span = step_device()
        e1.record()
        barrier()
        if args.profile_mode:
            torch.cuda.cudart().cudaProfilerStop()
        launches = _lib.launch_count() - launches0
        demixer.plan.enabled = False
        gemm_events, netops.gemm_timer = netops.gemm_timer, None
        dev_ms = max_over_ranks(e0.elapsed_time(e1))
        out = {"dev_ms": dev_ms, "launches": launches, "n": n, "mix_host": mix_host, "mix_dev": mix_dev,
               "sharded": sharded, "span": span, "stats": stats, "gemm_events": gemm_events or []}
        return out

    n_warm = args.warmup if args.profile_mode else max(args.warmup, 3)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    main = measure(mode, args.steps, n_warm, True)
    clocks = sampler.stop() if rank == 0 else None
    dev_ms, launches = main["dev_ms"], main["launches"]
    chunk_range = main["sharded"] is not None
    audio_s_total = track_seconds * world            # both modes process TRACK_SECONDS per GPU in aggregate
    value = audio_s_total * args.steps / (dev_ms / 1e3)

    def kernel_stats(key, bytes_fn):
        ev = demixer.plan.events[key]
        if not ev:
            return None
        ms = [a.elapsed_time(b) for a, b, _ in ev]
        byt = [bytes_fn(cfg, nc) for _, _, nc in ev]
        gbs = sum(byt) / (sum(ms) / 1e3) / 1e9
        return {"launches": len(ev), "avg_ms": sum(ms) / len(ms), "bytes_per_launch": sum(byt) / len(byt),
                "achieved_gbs": gbs, "share_of_step": sum(ms) / dev_ms}

    k2 = kernel_stats("istft", k2_algorithmic_bytes)
    k1 = kernel_stats("stft", k1_algorithmic_bytes)

    # the tensor-core GEMM launches of the timed steps, grouped by epilogue family
    gemm = {}
    for kind, flops, nbytes, a, b in main["gemm_events"]:
        fam = "residual" if kind == "residual" else ("glu" if kind == "glu" else "bf16")
        g_ = gemm.setdefault(fam, {"launches": 0, "ms": 0.0, "flops": 0.0, "bytes": 0.0})
        g_["launches"] += 1
        g_["ms"] += a.elapsed_time(b)
        g_["flops"] += flops
        g_["bytes"] += nbytes
    for g_ in gemm.values():
        g_["tflops"] = g_["flops"] / (g_["ms"] / 1e3) / 1e12
        g_["gbs"] = g_["bytes"] / (g_["ms"] / 1e3) / 1e9
        g_["share_of_step"] = g_["ms"] / dev_ms

    # ---- chunk-range: halo accounting + self-check against a single-GPU overlap-add of the same span ------------
    shard_info = None
    if chunk_range:
        st = main["stats"]
        waits = [a.elapsed_time(b) for a, b in st.get("wait_events", [])]
        n = main["n"]
        offs, mult = roformer_schedule(n, cfg.chunk_size, cfg.step)
        plan = plan_chunk_ranges(offs, cfg.chunk_size, n, world)
        me = plan[rank]
        # reference for this rank's span without any exchange: every chunk that touches the span, evaluated here
        c_ref0 = me.c0
        while c_ref0 > 0 and offs[c_ref0 - 1] + cfg.chunk_size > me.p0:
            c_ref0 -= 1
        rows = cfg.num_stems * cfg.audio_channels
        waves = demixer.chunk_waves(main["mix_dev"], offs[c_ref0:me.c1])
        ref_span = torch.empty((rows, me.p1 - me.p0), dtype=torch.float32, device=dev)
        from audiolab_b200.sharding import _ShiftedOut
        sp.ola_gather(waves, _dev_i64(offs, dev)[:me.c1], n, mult=_dev_i32(mult, dev)[:me.c1], wtab=demixer.window(dev),
                      p0=me.p0, p1=me.p1, eps=1e-10, data_chunk0=c_ref0, out=_ShiftedOut(ref_span, me.p0))
        diff = float((ref_span - main["span"]).abs().max())
        bitwise = bool(torch.equal(ref_span, main["span"]))
        agg = torch.tensor([diff, 0.0 if bitwise else 1.0, float(st.get("halo_bytes_out", 0)),
                            (sum(waits) / len(waits)) if waits else 0.0], dtype=torch.float64, device=dev)
        mx = agg.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = agg.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        shard_info = {"halo_bytes_per_boundary": int(mx[2].item()), "halo_bytes_total_per_step": int(sm[2].item()),
                      "halo_wait_ms_max_over_ranks": round(float(mx[3].item()), 4),
                      "tail_chunks_first": st.get("tail_chunks"),
                      # why the driver's weak-scaling efficiency against the N = 1 line (ONE 60 s track) cannot reach 1: a track of
                      # N x 60 s has overlap - 1 more chunks per interior boundary than N separate 60 s tracks (whose first and last
                      # 6 s are covered by fewer chunks); every chunk is still evaluated exactly once, nothing is recomputed
                      "chunks_in_this_track": len(offs),
                      "chunks_in_n_separate_tracks": world * len(roformer_schedule(track_seconds * SR, cfg.chunk_size, cfg.step)[0]),
                      "ideal_weak_efficiency_vs_n1": round(world * len(roformer_schedule(track_seconds * SR, cfg.chunk_size, cfg.step)[0])
                                                           / len(offs), 4),
                      "selfcheck_vs_single_gpu_ola": {"max_abs_diff": float(mx[0].item()), "bitwise_all_ranks": mx[1].item() == 0.0},
                      "note": "halo_wait_ms = time the compute stream blocks on the exchange after the interior chunks "
                              "(the isend/irecv were posted before them)"}
        del waves, ref_span

    # ---- the other mode, for context (N > 1 only) ---------------------------------------------------------------
    tracks_mode = None
    if world > 1 and chunk_range and not args.profile_mode:
        tm = measure("tracks", max(2, min(args.steps, 3)), 1, False)
        tracks_mode = {"value": audio_s_total * max(2, min(args.steps, 3)) / (tm["dev_ms"] / 1e3), "unit": UNIT,
                       "sharding": "track-per-GPU (no data-path collective)", "scaling": "weak"}
        del tm

    # ---- timed: end to end through the public API, host buffers -----------------------------------
    e2e_val, h2d, d2h, e2e_pa = None, 0, 0, None
    if not chunk_range and not args.profile_mode:
        n = main["n"]
        mix_host = main["mix_host"]
        out_host = {k: torch.empty((2, n), dtype=torch.float32).pin_memory() for k in ("Vocals", "Instrumental")}

        def step_e2e():
            stems = sep.separate_tensor(mix_host)            # H2D of the mix inside
            for k, v in stems.items():
                out_host[k].copy_(v, non_blocking=True)      # D2H of both stems
            return stems

        for _ in range(2):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        torch.cuda.synchronize()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        e2e_val = audio_s_total * args.steps / e2e_s
        h2d = mix_host.numel() * 4 * world          # whole job: every rank copies its own track in and both stems out
        d2h = 2 * 2 * n * 4 * world
        if world == 1:
            e2e_pa = process_audio_e2e(torch, mix_host, dev)
    elif chunk_range and not args.profile_mode:
        # chunk-range end to end: every rank receives the whole track from its pinned host buffer and returns its span
        n = main["n"]
        mix_host = main["mix_host"]
        sharded = main["sharded"]
        me_span = main["span"]
        span_host = torch.empty(me_span.shape, dtype=torch.float32).pin_memory()

        def step_e2e():
            x = mix_host.to(dev, non_blocking=True)
            sp_ = sharded.demix_span(x)[0]
            span_host.copy_(sp_, non_blocking=True)

        step_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_e2e()
        torch.cuda.synchronize()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        e2e_val = audio_s_total * args.steps / e2e_s
        h2d = mix_host.numel() * 4 * world
        d2h = 2 * n * 4 * cfg.num_stems

    # ---- the other 16-bit operand format, for context (same kernels) ----------------------------------------------
    net_dtype = demixer.net.compute_dtype
    net_dtype_name = "fp16" if net_dtype == torch.float16 else "bf16"
    other_mode = None
    if world == 1 and not args.profile_mode:
        other = "bf16" if net_dtype_name == "fp16" else "fp16"
        sep2 = Separator(log_level=40, allow_random_init=True, use_autocast=True, device=str(dev),
                         mdxc_params={"batch_size": args.batch, "overlap": 4, "compute_dtype": other})
        d2 = sep2.load_model("model_bs_roformer_ep_368_sdr_12.9628.ckpt").demixer
        mix2 = torch.from_numpy(synth_mix(track_seconds * SR, seed=1236)).to(dev)
        d2.demix(mix2)
        ms2 = _timed_passes(torch, lambda: d2.demix(mix2), 2)
        other_mode = {"name": other, "value": track_seconds * 1e3 / ms2}
        del sep2, d2, mix2
        torch.cuda.empty_cache()

    # ---- the other BASELINE configurations -----------------------------------------------------------------------
    want = [] if args.profile_mode else parse_configs(args.configs, world)
    for k in ("mix_host", "mix_dev", "span", "sharded"):
        main.pop(k, None)
    sep.model_instance = None
    torch.cuda.empty_cache()
    cfg_block = {}
    for name in want:
        try:
            if name == "cfg1" and rank == 0:
                cfg_block[name] = config_cfg1(torch, dev, pk)
            elif name == "cfg3" and rank == 0:
                cfg_block[name] = config_cfg3(torch, dev, pk)
            elif name == "cfg4":
                cfg_block[name] = config_cfg4(torch, dist, dev, pk, rank, world, max_over_ranks)
            elif name == "cfg5":
                cfg_block[name] = config_cfg5(torch, dist, dev, pk, rank, world, max_over_ranks)
        except torch.cuda.OutOfMemoryError as e:                      # report, do not lose the contract line
            cfg_block[name] = {"error": f"out of memory: {e}"[:200]}
        torch.cuda.empty_cache()
        if world > 1 and name in ("cfg1", "cfg3"):
            dist.barrier()

    if world > 1:
        dist.barrier()
    if rank == 0:
        offs, mult = roformer_schedule(track_seconds * SR, cfg.chunk_size, cfg.step)
        flops_step = net_flops_per_chunk(cfg) * len(offs) * world
        net_ms = dev_ms / args.steps * (1.0 - (k1["share_of_step"] if k1 else 0) - (k2["share_of_step"] if k2 else 0))
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))     # ncu --set full capture of one al_istft launch inside a bench step
            traffic = tj.get("al_istft_bytes_per_launch")
            if traffic and k2 and tj.get("chunks_per_launch"):
                ev = demixer.plan.events["istft"]
                traffic = traffic / tj["chunks_per_launch"] * (sum(nc for _, _, nc in ev) / len(ev))
        import audiolab_b200.nets.roformer as rof
        tc = rof._TC_GEMM and inst.demixer.net._tc_supported()
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": n_warm, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": f"{net_dtype_name} operands in the mask net (tcgen05 kind::f16, fp32 accumulate, fp32 residual stream), "
                     "f32 STFT/iSTFT/OLA",
            "data": "synthetic", "config": workload_config(world, mode),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roofline_block(gemm, k2, pk, traffic),
            "roofline_spectral": {
                "kernel": "al_istft (istft_pk4_kernel<mask>, stereo-packed, producer / consumer / overlap-add warps): complex mask (.) "
                          "spec + C2R iFFT + window + OLA + /env",
                "bound": "hbm", "achieved": k2["achieved_gbs"] if k2 else None, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": (k2["achieved_gbs"] / pk["hbm_gbs"]) if k2 else None, "traffic": traffic,
                "peak_source": pk["source"], "bytes_per_launch": k2["bytes_per_launch"] if k2 else None,
                "avg_launch_ms": k2["avg_ms"] if k2 else None, "share_of_step": k2["share_of_step"] if k2 else None,
            },
            "kernels": {"al_stft": k1, "al_istft": k2, "al_gemm_bf16": gemm},
            "mask_net": {"flops_per_step": flops_step, "tflops": flops_step / (net_ms / 1e3) / 1e12 / world,
                         "peak_tflops": pk["bf16_tflops"],
                         "frac_of_bf16_peak": flops_step / (net_ms / 1e3) / 1e12 / world / pk["bf16_tflops"],
                         "note": ("linear layers: al_gemm_bf16 (hand-written tcgen05 / TMA / TMEM GEMM, RMSNorm + rotary + GELU "
                                  "+ bias + fp32 residual fused into the epilogues); attention: "
                                  + ("al_band_attention (band axis, mma.sync) + " if rof._BAND_ATTN_TC else "cuDNN SDPA (band axis) + ")
                                  + ("al_time_attention (time axis, tcgen05 flash attention with the gate in its epilogue; no "
                                     "library call left in the step)" if rof._TIME_ATTN_TC else "cuDNN SDPA (time axis)")
                                  + ("; band split / mask estimator: grouped al_gemm_bf16 (GLU epilogue, fp32 mask)"
                                     if rof._GROUPED and inst.demixer.net._grouped_supported() else
                                     "; band split / mask estimator under bf16 autocast")) if tc else
                                 "dense layers via cuBLAS / cuDNN SDPA (AUDIOLAB_B200_TC_GEMM=0 comparison path)"},
        }
        if other_mode is not None:
            line["precision_modes"] = {
                net_dtype_name: {"value": value, "si_sdr_db_vs_fp32_oracle_full_size": PARITY_DB.get(net_dtype_name)},
                other_mode["name"]: {"value": other_mode["value"],
                                     "si_sdr_db_vs_fp32_oracle_full_size": PARITY_DB.get(other_mode["name"])},
                "note": "same kernels, 16-bit operand format selected by Separator(mdxc_params={'compute_dtype': ...}); SI-SDR "
                        "figures from profiles/r02i_parity_fullsize_fp16.jsonl (asserted by tests/test_demix_gpu.py::"
                        "test_roformer_full_size_parity_of_the_16bit_paths); BASELINE.json asks for >= 60 dB",
            }
        if shard_info is not None:
            line["sharding"] = shard_info
        if tracks_mode is not None:
            line["tracks_mode"] = tracks_mode
        if e2e_pa is not None:
            line["e2e_process_audio"] = e2e_pa
        if cfg_block:
            line["configs"] = cfg_block
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline_block()
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def roofline_block(gemm: dict, k2, pk: dict, k2_traffic):
    """`roofline` of the contract line = the DOMINANT kernel of the timed step.  With the tcgen05 path that is
    gemm_bf16_kernel<256, 4, EPI_BF16> (to_qkv+to_gates, FeedForward Linear 1 + GELU, mask-estimator Linear 1): tensor
    bound, algorithmic FLOPs (2 M N K of every launch) / CUDA-event time of those launches inside the timed steps, against
    the SUSTAINED bf16 peak of MEASURED_PEAKS.json (the kernel runs inside a long power-capped step).  `traffic` = DRAM
    bytes of one to_qkv launch from the `ncu --set full` capture profiles/r02h_ncu_full_gemm_shapes.txt (algorithmic:
    1.37 GB A + 4.16 GB outputs).  Without GEMM launches (fallback path) the spectral kernel is reported as before."""
    g_ = gemm.get("bf16")
    if not g_:
        return {"kernel": "al_istft", "bound": "hbm", "achieved": k2["achieved_gbs"] if k2 else None, "peak": pk["hbm_gbs"],
                "unit": "GB/s", "frac": (k2["achieved_gbs"] / pk["hbm_gbs"]) if k2 else None, "traffic": k2_traffic,
                "peak_source": pk["source"]}
    return {"kernel": "al_gemm_bf16: gemm_bf16_kernel<BN 256, 4 stages, EPI_BF16> (tcgen05.mma kind::f16 + TMA + TMEM; to_qkv+"
                      "to_gates with rowscale / rotary, FF Linear 1 with bias + GELU, mask-estimator Linear 1 with tanh)",
            "bound": "tensor", "achieved": g_["tflops"], "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
            "frac": g_["tflops"] / pk["bf16_tflops"], "traffic": 5.53e9, "traffic_of": "one to_qkv+to_gates launch (ncu)",
            "peak_source": pk["source"] + ", sustained figure", "launches": g_["launches"],
            "avg_launch_ms": g_["ms"] / g_["launches"], "share_of_step": g_["share_of_step"],
            "hbm_bound_sibling": None if "residual" not in gemm else {
                "kernel": "gemm_bf16_kernel<256, 3 stages, EPI_RES> (to_out / FF Linear 2 + fp32 residual stream)",
                "bound": "hbm", "achieved": gemm["residual"]["gbs"], "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": gemm["residual"]["gbs"] / pk["hbm_gbs"], "share_of_step": gemm["residual"]["share_of_step"]}}


def parse_configs(spec: str, world: int):
    if spec == "none":
        return []
    if spec == "all":
        return ["cfg1", "cfg3", "cfg4", "cfg5"] if world == 1 else ["cfg4", "cfg5"]
    return [c.strip() for c in spec.split(",") if c.strip() in ("cfg1", "cfg3", "cfg4", "cfg5")]


def process_audio_e2e(torch, mix_host, dev) -> dict:
    """The reference-facing plugin call: Separate.process_audio on a WAV file (ensemble_strength 1 would load the Mel model;
    here the contract's BS-RoFormer is injected as the only ensemble member), stems written as FLOAT WAV files.  Files live in
    /dev/shm when available; file I/O is inside the timed region."""
    import shutil
    import tempfile

    import audiolab_b200.orchestrator as orch
    from audiolab_b200 import project_files
    from audiolab_b200.wavio import write_wav
    from audiolab_b200.wrappers import Separate
    root = tempfile.mkdtemp(prefix="al_bench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        old_out, old_ens, old_kw = project_files.output_path, orch.ENSEMBLE, Separate.engine_kwargs
        project_files.output_path = os.path.join(root, "outputs")
        orch.ENSEMBLE = [("model_bs_roformer_ep_368_sdr_12.9628.ckpt", 8.4, 16.0)]
        Separate.engine_kwargs = dict(allow_random_init=True, model_file_dir=os.path.join(root, "models"))
        w = Separate()
        times = []
        for i in range(3):
            src = os.path.join(root, f"track{i}.wav")
            write_wav(src, mix_host.numpy(), SR, "FLOAT")
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            w.process_audio([project_files.ProjectFiles(src)], ensemble_strength=1)
            torch.cuda.synchronize()
            times.append(time.perf_counter() - t0)
        best = min(times[1:])
        return {"value": mix_host.shape[1] / SR / best, "unit": UNIT, "seconds": [round(t, 3) for t in times],
                "what": "Separate.process_audio(WAV file) -> stem WAV files (read, H2D, demix, blend, de-bleed, D2H, write); "
                        "every call builds the Separator and loads the model, as the reference's predict_with_model does; "
                        "best of calls 2 and 3"}
    finally:
        project_files.output_path, orch.ENSEMBLE, Separate.engine_kwargs = old_out, old_ens, old_kw
        shutil.rmtree(root, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="auto", choices=["auto", "tracks", "chunk-range"],
                    help="N > 1: chunk-range (default; halo exchange over NCCL) or tracks (one track per GPU)")
    ap.add_argument("--configs", default="all", help="other BASELINE configs to time after the contract workload: "
                                                     "all | none | comma list of cfg1,cfg3,cfg4,cfg5")
    ap.add_argument("--batch", type=int, default=27, help="chunks per mask-net call (27 = the whole 60 s track)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--track-seconds", type=int, default=0,
                    help="with --profile-mode only: length of the synthetic track (bounds the ncu launch list)")
    ap.add_argument("--profile-mode", action="store_true",
                    help="for runs under ncu: honour --warmup exactly and skip the e2e leg (never a bench value)")
    args = ap.parse_args()
    # The contract is ONE JSON line on stdout.  Libraries below us write to file descriptor 1 on their own (NCCL prints its
    # version banner there when NCCL_DEBUG is set in the box's environment): point fd 1 at stderr for the whole run and keep
    # the real stdout for the final line only.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
    real_stdout.flush()


if __name__ == "__main__":
    main()
