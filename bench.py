#!/usr/bin/env python
"""bench.py -- realtime factor of the chunked STFT -> mask -> iSTFT + OLA demix (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--mode tracks|chunk-range]

Workload at N=1 (BASELINE.json configs[1]): BS-RoFormer vocals/instrumental, dim 512 / depth 12,
n_fft 2048, hop 441, stereo, 8 s chunks with overlap 4 (step 2 s), bf16 random-init weights, one
60 s 44.1 kHz synthetic track per GPU.  One "step" = one full demix of that track.

  value : audio-seconds / second, whole job, mix already resident in HBM, CUDA events, max over ranks
  e2e   : the same through the public API (Separator.separate_tensor) from a pinned HOST buffer, with
          the host->device copy of the mix and the device->host copy of both stems inside the timed
          region
  roofline     : the dominant spectral kernel (al_istft, fused mask multiply + iFFT + OLA), algorithmic
                 bytes / CUDA-event duration of its launches inside the timed steps, vs the measured
                 HBM peak in MEASURED_PEAKS.json
  cpu_baseline : the oracle (CPU restatement of the reference) on the host cores, one chunk evaluation
N > 1 (torchrun): --mode tracks (default, weak scaling: one track per GPU, no data-path collective) or
--mode chunk-range (one 60*N s track split by chunk range, OLA halo exchange over NCCL).
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
METRIC = "realtime factor (audio-s/s) of chunked STFT->mask->iSTFT+OLA demix"
UNIT = "audio-s/s"


def workload_config(n_gpus: int, mode: str) -> dict:
    return {
        "workload": "BS-RoFormer vocals/instrumental (dim 512, depth 12, 62 bands), n_fft=2048 hop=441 stereo, "
                    "8 s chunks overlap=4, bf16 random-init weights (seed 4321), "
                    f"{TRACK_SECONDS} s 44.1 kHz synthetic track per GPU (BASELINE.json configs[1])",
        "n_fft": 2048, "hop": 441, "chunk_samples": 352800, "overlap": 4, "track_seconds": TRACK_SECONDS,
        "sharding": ("track-per-GPU" if mode == "tracks" else "chunk-range + NCCL halo exchange") if n_gpus > 1 else "none",
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
    oracle BS-RoFormer (fp32, CPU, all host threads).  The reference loop makes one such evaluation per
    `step` = chunk/4 = 2 s of audio, whatever the track length, so RTF = 2 s / t_eval."""
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
        "dtype": "f32", "data": "synthetic", "config": workload_config(args.gpus, args.mode),
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


def run_ours(args) -> None:
    import torch
    import torch.distributed as dist

    from audiolab_b200 import _lib
    from audiolab_b200 import spectral as sp
    from audiolab_b200.configs import RoformerConfig
    from audiolab_b200.demix import _dev_i32, _dev_i64, roformer_schedule
    from audiolab_b200.separator import Separator
    from audiolab_b200.sharding import ShardedRoformerDemixer

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

    sep = Separator(log_level=40, allow_random_init=True, use_autocast=True, device=str(dev),
                    mdxc_params={"batch_size": args.batch, "overlap": 4})
    inst = sep.load_model("model_bs_roformer_ep_368_sdr_12.9628.ckpt")
    demixer = inst.demixer
    cfg: RoformerConfig = demixer.cfg
    demixer.plan = TimedPlan(demixer.plan, torch)

    chunk_range = args.mode == "chunk-range" and world > 1
    track_seconds = TRACK_SECONDS
    if args.profile_mode and args.track_seconds:
        track_seconds = args.track_seconds     # shorter launch list under ncu; never a bench value
    n = track_seconds * SR * (world if chunk_range else 1)
    seed = 1236 + (0 if chunk_range else rank)
    mix_host = torch.from_numpy(synth_mix(n, seed=seed)).pin_memory()
    mix_dev = mix_host.to(dev)
    sharded = ShardedRoformerDemixer(demixer, rank, world) if chunk_range else None

    def step_device():
        if sharded is not None:
            return sharded.demix_span(mix_dev)[0]
        return demixer.demix(mix_dev)

    out_host = {k: torch.empty((2, n), dtype=torch.float32).pin_memory() for k in ("Vocals", "Instrumental")}

    def step_e2e():
        stems = sep.separate_tensor(mix_host)            # H2D of the mix inside
        for k, v in stems.items():
            out_host[k].copy_(v, non_blocking=True)      # D2H of both stems
        return stems

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

    # ---- warm-up ------------------------------------------------------------------------------
    n_warm = args.warmup if args.profile_mode else max(args.warmup, 3)
    for _ in range(n_warm):
        step_device()
    barrier()

    # ---- timed: device-resident ------------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    demixer.plan.enabled = True
    launches0 = _lib.launch_count()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    if args.profile_mode:
        torch.cuda.cudart().cudaProfilerStart()          # ncu --profile-from-start off lists the timed steps only
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    if args.profile_mode:
        torch.cuda.cudart().cudaProfilerStop()
    launches = _lib.launch_count() - launches0
    demixer.plan.enabled = False
    dev_ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None

    audio_s_total = track_seconds * world            # both modes process TRACK_SECONDS per GPU in aggregate
    value = audio_s_total * args.steps / (dev_ms / 1e3)

    # kernel timings from the bracketed launches
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

    # ---- timed: end to end through the public API, host buffers -----------------------------------
    if sharded is None and not args.profile_mode:
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
    else:
        e2e_val, h2d, d2h = None, 0, 0

    if world > 1:
        dist.barrier()
    if rank == 0:
        pk = peaks()
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
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": n_warm, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if chunk_range else "weak", "vs_baseline": None, "dtype": "bf16 mask net, f32 STFT/iSTFT/OLA",
            "data": "synthetic", "config": workload_config(world, args.mode),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {
                "kernel": "al_istft (istft_pk2_kernel<mask>, stereo-packed): complex mask (.) spec + C2R iFFT + window "
                          "+ OLA + /env",
                "bound": "hbm", "achieved": k2["achieved_gbs"] if k2 else None, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": (k2["achieved_gbs"] / pk["hbm_gbs"]) if k2 else None, "traffic": traffic,
                "peak_source": pk["source"], "bytes_per_launch": k2["bytes_per_launch"] if k2 else None,
                "avg_launch_ms": k2["avg_ms"] if k2 else None, "share_of_step": k2["share_of_step"] if k2 else None,
            },
            "kernels": {"al_stft": k1, "al_istft": k2},
            "mask_net": {"flops_per_step": flops_step, "tflops": flops_step / (net_ms / 1e3) / 1e12 / world,
                         "peak_tflops": pk["bf16_tflops"],
                         "frac_of_bf16_peak": flops_step / (net_ms / 1e3) / 1e12 / world / pk["bf16_tflops"],
                         "note": "dense layers via cuBLAS / cuDNN SDPA (library calls, bf16); RMSNorm / rotary / gating are "
                                 "fused al_netops kernels; band split and mask estimator run under bf16 autocast"},
        }
        if not args.no_cpu_baseline and world == 1:
            res = oracle_chunk_seconds(1, 0)
            rtf = res["step_audio_s"] / res["seconds"][0]
            line["cpu_baseline"] = {"value": rtf, "unit": UNIT, "cores": res["cores"], "kind": "port",
                                    "sample": "one 8 s chunk evaluation of the oracle (fp32, all host threads) = "
                                              "the work the reference loop does per 2 s of audio"}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mode", default="tracks", choices=["tracks", "chunk-range"])
    ap.add_argument("--batch", type=int, default=27, help="chunks per mask-net call (27 = the whole 60 s track)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--track-seconds", type=int, default=0,
                    help="with --profile-mode only: length of the synthetic track (bounds the ncu launch list)")
    ap.add_argument("--profile-mode", action="store_true",
                    help="for runs under ncu: honour --warmup exactly and skip the e2e leg (never a bench value)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
