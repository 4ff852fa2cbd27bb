#!/usr/bin/env python
"""Isolated roofline measurement of the spectral kernels (K1 al_stft, K2 al_istft, K2b al_ola_gather,
K3 al_resample_poly) at the BASELINE shapes.  Working sets are sized above the 126 MB L2; every kernel
is timed with CUDA events on the launch stream after warm-up.  Prints one JSON line per case:
algorithmic bytes, median launch time, achieved GB/s and fraction of the measured HBM peak.

    python tools/kernel_bench.py [--iters 20] [--only stft|istft|ola|resample] [--once]

--once runs each case exactly once after one warm-up (for ncu captures).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from audiolab_b200 import spectral as sp  # noqa: E402


def hbm_peak() -> float:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"])
    return 6650.0


REPS = 8   # launches per event pair: a lone ~50 us kernel would otherwise be timed together with the ~20 us the
           # host needs to enqueue it after the start event (the GPU sits idle in between)


def time_it(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    reps = 1 if iters == 1 else REPS
    ms = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn()                      # keeps the queue non-empty when the start event is reached
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1) / reps)
    return ms


def report(name, nbytes, ms, peak, extra=None):
    med = statistics.median(ms)
    gbs = nbytes / (med / 1e3) / 1e9
    line = {"case": name, "bytes": nbytes, "ms_median": round(med, 4), "ms_min": round(min(ms), 4),
            "gbs": round(gbs, 1), "frac_of_measured_hbm": round(gbs / peak, 3)}
    if extra:
        line.update(extra)
    print(json.dumps(line), flush=True)


CASES = {
    # name: (n_fft, hop, chunk_len, n_frames, layout, n_bins_out, n_chunks, center_pad, frame_pad, stems, mask)
    "roformer_2048_441": dict(n_fft=2048, hop=441, chunk=352800, T=801, layout=sp.FRAME_INTERLEAVED, Fo=1025,
                              nch=16, stems=1, mask=True, normalized=False),
    "mdx_6144_1024": dict(n_fft=6144, hop=1024, chunk=261120, T=256, layout=sp.CAC, Fo=3072, nch=16, stems=1,
                          mask=False, normalized=False),
    "htdemucs_4096_1024": dict(n_fft=4096, hop=1024, chunk=343980, T=336, layout=sp.CAC, Fo=2048, nch=16,
                               stems=4, mask=False, normalized=True),
    "binmajor_2048_441": dict(n_fft=2048, hop=441, chunk=352800, T=801, layout=sp.BIN_MAJOR, Fo=1025, nch=16,
                              stems=1, mask=True, normalized=False),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--only", default="")
    ap.add_argument("--cases", default="")
    ap.add_argument("--once", action="store_true")
    ap.add_argument("--gelu", action="store_true", help="also time the bf16 GELU kernel (RoFormer FeedForward hidden of 5 chunks)")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    peak = hbm_peak()
    iters, warm = (1, 1) if args.once else (args.iters, 3)
    g = torch.Generator(device="cpu").manual_seed(7)
    want = set(args.cases.split(",")) if args.cases else None

    for name, c in CASES.items():
        if want and name not in want:
            continue
        plan = sp.StftPlan(c["n_fft"], c["hop"], normalized=c["normalized"])
        nch, chunk, T, Fo = c["nch"], c["chunk"], c["T"], c["Fo"]
        n = nch * chunk
        track = (torch.rand((2, n), generator=g) - 0.5).to(dev)
        ht = name.startswith("htdemucs")
        kw = dict(chunk_len=chunk, n_chunks=nch, off0=0, off_step=chunk, n_frames=T, layout=c["layout"], n_bins_out=Fo)
        if ht:
            kw["center_pad"] = 1536
        if args.only in ("", "stft"):
            out = plan.stft(track, **kw)
            ms = time_it(lambda: plan.stft(track, out=out, **kw), iters, warm)
            b = nch * 2 * (chunk * 4 + T * Fo * 8)
            report(f"al_stft/{name}", b, ms, peak)
        if args.only in ("", "istft"):
            spec = plan.stft(track, **kw)
            stems = c["stems"]
            mask = None
            if c["mask"]:
                shp = list(spec.shape)
                if c["layout"] == sp.FRAME_INTERLEAVED:
                    shp = [nch, stems] + shp[1:]
                else:
                    shp = [nch * stems * 2] + shp[1:]
                mask = torch.view_as_complex((torch.rand(shp + [2], generator=g) - 0.5).to(dev))
            if ht:
                spec_in = spec[:, None].expand(nch, stems, 4, Fo, T).contiguous()
                ikw = dict(n_chunks=nch, channels=2, stems=stems, layout=c["layout"], spec_has_stems=True,
                           frame_pad=2, out_start=c["n_fft"] // 2 + 1536, out_len=chunk)
                b = nch * stems * 2 * (T * Fo * 8 + chunk * 4)
            else:
                spec_in = spec
                ikw = dict(n_chunks=nch, channels=2, stems=stems, layout=c["layout"], mask=mask, out_len=chunk)
                b = nch * 2 * (T * Fo * 8 * (2 if mask is not None else 1) + chunk * 4)
            dst = plan.istft(spec_in, **ikw)
            dkw = dict(dst=dst, dst_ch_stride=chunk, dst_chunk_stride=stems * 2 * chunk, dst_limit=chunk)
            ms = time_it(lambda: plan.istft(spec_in, **ikw, **dkw), iters, warm)
            report(f"al_istft/{name}", b, ms, peak)
            del spec_in, spec, mask, dst
        del track
        torch.cuda.empty_cache()

    if args.only in ("", "ola"):
        # RoFormer cfg: 60 chunks of 352800 every 88200, 2 rows, Hamming table
        from audiolab_b200.demix import hamming_sym
        nchunks, C, step, rows = 64, 352800, 88200, 2
        n_total = (nchunks - 1) * step + C
        waves = (torch.rand((nchunks, rows, C), generator=g) - 0.5).to(dev)
        offs = torch.arange(nchunks, dtype=torch.int64, device=dev) * step
        wtab = torch.from_numpy(hamming_sym(C)[None]).to(dev)
        out = torch.empty((rows, n_total), dtype=torch.float32, device=dev)
        ms = time_it(lambda: sp.ola_gather(waves, offs, n_total, wtab=wtab, out=out), iters, warm)
        b = waves.numel() * 4 + out.numel() * 4
        report("al_ola_gather/roformer", b, ms, peak)
        del waves, out

    if args.only in ("", "resample"):
        rows, n_in = 8, 8_640_000
        x = (torch.rand((rows, n_in), generator=g) - 0.5).to(dev)
        taps = torch.from_numpy(sp.resample_taps(147, 160)).to(dev)
        y = sp.resample_poly(x, 147, 160, taps=taps)
        ms = time_it(lambda: sp.resample_poly(x, 147, 160, taps=taps), iters, warm)
        b = x.numel() * 4 + y.numel() * 4
        report("al_resample_poly/48k_to_44k1", b, ms, peak)

    if args.gelu or args.only == "gelu":
        from audiolab_b200 import netops
        # in place: refill from h0 before every launch (fresh activations, as in the network) and subtract the refill
        h0 = torch.randn((5 * 801 * 62, 2048), device=dev, dtype=torch.bfloat16)
        h = torch.empty_like(h0)
        import statistics as st
        ms_copy = st.median(time_it(lambda: h.copy_(h0), iters, warm))
        ms_both = time_it(lambda: netops.gelu_(h.copy_(h0)), iters, warm)
        ms = [max(m - ms_copy, 1e-6) for m in ms_both]
        report("al_gelu_bf16/roformer_ff_hidden", h.numel() * 4, ms, peak, {"refill_ms_subtracted": round(ms_copy, 4)})


if __name__ == "__main__":
    main()
