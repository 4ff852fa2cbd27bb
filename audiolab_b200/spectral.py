"""Thin host wrappers over the C ABI: torch tensors in, torch tensors out, kernels in between.

Every function enqueues on torch's current CUDA stream and never synchronises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Union

import numpy as np
import torch

from . import _lib

FRAME_MAJOR, BIN_MAJOR, CAC, FRAME_INTERLEAVED = 0, 1, 2, 3


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _need_cuda(*ts: Optional[torch.Tensor]) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("audiolab_b200 kernels need CUDA tensors (there is no CPU fallback)")


class StftPlan:
    """(n_fft, hop, window, normalized) -> device tables.  Mirrors torch.stft/istft arguments."""

    def __init__(self, n_fft: int, hop: int, window: Optional[Union[np.ndarray, torch.Tensor]] = None,
                 normalized: bool = False):
        self.n_fft, self.hop, self.normalized = int(n_fft), int(hop), bool(normalized)
        self.n_bins = self.n_fft // 2 + 1
        handle = C.c_void_p()
        wptr = None
        if window is not None:
            w = np.ascontiguousarray(torch.as_tensor(window).detach().cpu().numpy(), dtype=np.float32)
            if w.shape != (self.n_fft,):
                raise ValueError("window must have n_fft samples")
            self._w = w
            wptr = w.ctypes.data_as(C.c_void_p)
        _lib.check(_lib.lib().al_plan_create(self.n_fft, self.hop, wptr, int(self.normalized), C.byref(handle)),
                   "al_plan_create")
        self._h = handle

    def __del__(self):
        h = getattr(self, "_h", None)
        if h is not None and h.value:
            try:
                _lib.lib().al_plan_destroy(h)
            except Exception:
                pass
            self._h = None

    # ---------------------------------------------------------------------------------------
    def spec_shape(self, n_chunks: int, channels: int, n_frames: int, layout: int, n_bins_out: int):
        if layout == FRAME_MAJOR:
            return (n_chunks * channels, n_frames, n_bins_out), torch.complex64
        if layout == BIN_MAJOR:
            return (n_chunks * channels, n_bins_out, n_frames), torch.complex64
        if layout == FRAME_INTERLEAVED:
            return (n_chunks, n_frames, n_bins_out, channels), torch.complex64
        return (n_chunks, channels * 2, n_bins_out, n_frames), torch.float32

    def stft(self, track: torch.Tensor, *, chunk_len: int, n_chunks: int = 1,
             offsets: Union[None, torch.Tensor] = None, off0: int = 0, off_step: int = 0,
             n_valid: Optional[int] = None, center_pad: Optional[int] = None,
             n_frames: Optional[int] = None, layout: int = BIN_MAJOR, n_bins_out: Optional[int] = None,
             zero_low_bins: int = 0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """track [channels, n] fp32 -> spectrogram of `n_chunks` chunks (see al_stft)."""
        _need_cuda(track, offsets, out)
        if track.dtype != torch.float32 or track.dim() != 2 or track.stride(1) != 1:
            raise ValueError("track must be fp32 [channels, n] with unit inner stride")
        channels, n = track.shape
        n_valid = n if n_valid is None else int(n_valid)
        center_pad = self.n_fft // 2 if center_pad is None else int(center_pad)
        n_frames = 1 + chunk_len // self.hop if n_frames is None else int(n_frames)
        n_bins_out = self.n_bins if n_bins_out is None else int(n_bins_out)
        shape, dtype = self.spec_shape(n_chunks, channels, n_frames, layout, n_bins_out)
        if out is None:
            out = torch.empty(shape, dtype=dtype, device=track.device)
        elif tuple(out.shape) != shape or out.dtype != dtype or not out.is_contiguous():
            raise ValueError(f"out must be contiguous {dtype} {shape}")
        if offsets is not None and (offsets.dtype != torch.int64 or offsets.numel() != n_chunks):
            raise ValueError("offsets must be int64 [n_chunks]")
        _lib.check(_lib.lib().al_stft(self._h, track.data_ptr(), n_valid, track.stride(0), channels,
                                      _ptr(offsets), int(off0), int(off_step), int(n_chunks), int(chunk_len),
                                      center_pad, n_frames, out.data_ptr(), int(layout), n_bins_out,
                                      int(zero_low_bins), _stream()), "al_stft")
        return out

    def istft(self, spec: torch.Tensor, *, n_chunks: int, channels: int, stems: int = 1,
              mask: Optional[torch.Tensor] = None, layout: int = BIN_MAJOR, spec_has_stems: bool = False,
              frame_pad: int = 0, zero_low_bins: int = 0, out_start: Optional[int] = None,
              out_len: Optional[int] = None, weight: Optional[torch.Tensor] = None,
              dst: Optional[torch.Tensor] = None, dst_ch_stride: Optional[int] = None,
              dst_chunk_stride: Optional[int] = None, dst_offsets: Optional[torch.Tensor] = None,
              dst_off0: int = 0, dst_off_step: int = 0, dst_limit: Optional[int] = None) -> torch.Tensor:
        """Inverse of `stft` with fused mask multiply / OLA / normalisation (see al_istft).

        Default placement: dense chunk waves [n_chunks, stems, channels, out_len].
        """
        _need_cuda(spec, mask, weight, dst, dst_offsets)
        if not spec.is_contiguous() or (mask is not None and not mask.is_contiguous()):
            raise ValueError("spec / mask must be contiguous")
        if layout == CAC:
            if spec.dtype != torch.float32:
                raise ValueError("CaC spectrogram must be fp32")
            n_bins_in, n_frames_in = spec.shape[-2], spec.shape[-1]
        else:
            if spec.dtype != torch.complex64:
                raise ValueError("complex layouts need complex64")
            if layout == FRAME_MAJOR:
                n_frames_in, n_bins_in = spec.shape[-2], spec.shape[-1]
            elif layout == FRAME_INTERLEAVED:
                n_frames_in, n_bins_in = spec.shape[-3], spec.shape[-2]
            else:
                n_bins_in, n_frames_in = spec.shape[-2], spec.shape[-1]
        rows_spec = n_chunks * channels * (stems if spec_has_stems else 1)
        per_row = n_bins_in * n_frames_in * (2 if layout == CAC else 1)
        if spec.numel() != rows_spec * per_row:
            raise ValueError(f"spec has {spec.numel()} elements, expected {rows_spec}x{n_bins_in}x{n_frames_in}")
        if mask is not None:
            if mask.dtype != torch.complex64 or mask.numel() != n_chunks * stems * channels * self.n_bins * n_frames_in:
                raise ValueError("mask must be complex64 [n_chunks*stems*channels, (T,F)|(F,T)] over all n_fft/2+1 bins")
        T = n_frames_in + 2 * frame_pad
        out_start = self.n_fft // 2 if out_start is None else int(out_start)
        out_len = (T - 1) * self.hop if out_len is None else int(out_len)
        if weight is not None and (weight.dtype != torch.float32 or weight.numel() < out_len):
            raise ValueError("weight must be fp32 with >= out_len samples")
        if dst is None:
            dst = torch.empty((n_chunks, stems, channels, out_len), dtype=torch.float32, device=spec.device)
            dst_ch_stride, dst_chunk_stride, dst_limit = out_len, stems * channels * out_len, out_len
            dst_offsets, dst_off0, dst_off_step = None, 0, 0
        else:
            if dst.dtype != torch.float32:
                raise ValueError("dst must be fp32")
            if dst_ch_stride is None or dst_chunk_stride is None or dst_limit is None:
                raise ValueError("explicit dst needs dst_ch_stride, dst_chunk_stride, dst_limit")
        _lib.check(_lib.lib().al_istft(self._h, spec.data_ptr(), _ptr(mask), int(layout), int(n_bins_in),
                                       int(n_frames_in), int(frame_pad), int(n_chunks), int(stems), int(channels),
                                       int(bool(spec_has_stems)), int(zero_low_bins), out_start, out_len,
                                       _ptr(weight), dst.data_ptr(), int(dst_ch_stride), int(dst_chunk_stride),
                                       _ptr(dst_offsets), int(dst_off0), int(dst_off_step), int(dst_limit),
                                       _stream()), "al_istft")
        return dst


def ola_gather(chunks: torch.Tensor, offsets: torch.Tensor, n_total: int, *,
               mult: Optional[torch.Tensor] = None, wtab: Optional[torch.Tensor] = None,
               tab_id: Optional[torch.Tensor] = None, p0: int = 0, p1: Optional[int] = None,
               halo_in: Optional[torch.Tensor] = None, raw_out: bool = False, eps: float = 1e-10,
               data_chunk0: int = 0,
               scale: float = 1.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """chunks [n_chunks - data_chunk0, rows, chunk_len] -> track [rows, n_total] (positions [p0, p1) written)."""
    shift, windowed = 0, False
    if out is not None and hasattr(out, "buf"):          # sharding._ShiftedOut: [rows, p1-p0] window of the track
        shift, out, windowed = int(out.shift), out.buf, True
    _need_cuda(chunks, offsets, mult, wtab, tab_id, halo_in, out)
    if chunks.dtype != torch.float32 or chunks.dim() != 3 or not chunks.is_contiguous():
        raise ValueError("chunks must be contiguous fp32 [n_chunks, rows, chunk_len]")
    n_data, rows, chunk_len = chunks.shape
    n_chunks = n_data + int(data_chunk0)
    p1 = n_total if p1 is None else int(p1)
    if offsets.dtype != torch.int64 or offsets.numel() != n_chunks:
        raise ValueError("offsets must be int64 [n_chunks]")
    if mult is not None and (mult.dtype != torch.int32 or mult.numel() != n_chunks):
        raise ValueError("mult must be int32 [n_chunks]")
    if tab_id is not None and (tab_id.dtype != torch.int32 or tab_id.numel() != n_chunks):
        raise ValueError("tab_id must be int32 [n_chunks]")
    if wtab is not None and (wtab.dtype != torch.float32 or wtab.shape[-1] != chunk_len or not wtab.is_contiguous()):
        raise ValueError("wtab must be contiguous fp32 [n_tab, chunk_len]")
    if halo_in is not None and (halo_in.dtype != torch.float32 or halo_in.numel() != rows * (p1 - p0)
                                or not halo_in.is_contiguous()):
        raise ValueError("halo_in must be contiguous fp32 [rows, p1-p0]")
    if windowed and out.shape[1] < p1 - shift:
        raise ValueError("shifted out buffer too small")
    if out is None:
        out = torch.zeros((rows, n_total), dtype=torch.float32, device=chunks.device)
    elif out.dtype != torch.float32 or out.dim() != 2 or out.shape[0] != rows or out.stride(1) != 1:
        raise ValueError("out must be fp32 [rows, >= p1]")
    _lib.check(_lib.lib().al_ola_gather(chunks.data_ptr(), n_chunks, int(data_chunk0), rows, chunk_len, offsets.data_ptr(),
                                        _ptr(mult), _ptr(wtab), _ptr(tab_id), int(n_total), int(p0), p1,
                                        _ptr(halo_in), int(bool(raw_out)), float(eps), float(scale),
                                        out.data_ptr() - 4 * shift, out.stride(0), _stream()), "al_ola_gather")
    return out


_TAPS_CACHE = {}


def resample_taps(up: int, down: int) -> np.ndarray:
    """Kaiser(5.0)-windowed sinc of scipy.signal.resample_poly, in float32 like scipy uses it.

    firwin(2*10*max(up,down)+1, 1/max(up,down), window=('kaiser', 5.0)) * up
    """
    g = int(np.gcd(up, down))
    up, down = up // g, down // g
    max_rate = max(up, down)
    half = 10 * max_rate
    n = np.arange(-half, half + 1, dtype=np.float64)
    fc = 1.0 / max_rate
    h = fc * np.sinc(fc * n) * np.kaiser(2 * half + 1, 5.0)
    h /= h.sum()                       # firwin scales the pass band (DC) to unity gain
    return (h * up).astype(np.float32)


def resample_poly(x: torch.Tensor, up: int = 147, down: int = 160,
                  taps: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x [rows, n_in] fp32 -> [rows, ceil(n_in*up/down)] (scipy.signal.resample_poly semantics)."""
    _need_cuda(x, taps)
    if x.dtype != torch.float32 or x.dim() != 2 or x.stride(1) != 1:
        raise ValueError("x must be fp32 [rows, n] with unit inner stride")
    g = int(np.gcd(up, down))
    up, down = up // g, down // g
    if taps is None:
        key = (up, down, str(x.device))
        taps = _TAPS_CACHE.get(key)
        if taps is None:                                   # one firwin design + upload per ratio and device
            taps = _TAPS_CACHE[key] = torch.from_numpy(resample_taps(up, down)).to(x.device)
    rows, n_in = x.shape
    n_out = (n_in * up + down - 1) // down
    out = torch.empty((rows, n_out), dtype=torch.float32, device=x.device)
    _lib.check(_lib.lib().al_resample_poly(x.data_ptr(), x.stride(0), out.data_ptr(), out.stride(0), rows, n_in,
                                           n_out, up, down, taps.data_ptr(), taps.numel(), _stream()),
               "al_resample_poly")
    return out


def sub(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    _need_cuda(a, b)
    if a.shape != b.shape or a.dtype != torch.float32 or b.dtype != torch.float32:
        raise ValueError("a, b must be fp32 of equal shape")
    a, b = a.contiguous(), b.contiguous()
    out = torch.empty_like(a)
    _lib.check(_lib.lib().al_sub(a.data_ptr(), b.data_ptr(), out.data_ptr(), a.numel(), _stream()), "al_sub")
    return out
