"""Chunked demix drivers: pad-and-chunk -> al_stft -> mask net -> al_istft -> overlap-add.

One class per separator family of the reference's hot path:

* ``MdxDemixer``      rows a8-a11: ConvTDFNetTrim.stft/istft + Predictor.demix/demix_base
                      (/root/reference/modules/rvc/infer/modules/uvr5/mdxnet.py:41-75, 109-197) and the
                      windowed overlap-add form of upstream MDXSeparator.demix.
* ``RoformerDemixer`` row a12: upstream MDXCSeparator.demix roformer branch + BSRoformer /
                      MelBandRoformer forward.
* ``HTDemucsDemixer`` row a13: DemucsSeparator.demix_demucs -> demucs.apply.apply_model ->
                      HTDemucs.forward.

Everything stays on the device: no ``.cpu()`` per chunk (the reference moves every chunk to the
host for numpy OLA), no host synchronisation inside a demix call.  The
spectral work is done by the sm_100a kernels behind the C ABI; PyTorch is used for device memory,
streams and the mask network's dense layers.
"""
from __future__ import annotations

import math
import random
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
import collections

import torch
import torch.nn.functional as F

from . import spectral as sp
from .configs import HTDemucsConfig, MdxConfig, RoformerConfig


# Offset / multiplicity tables of a chunk schedule are tiny and repeat for every track of the same length: keep the
# device copies (each torch.tensor(list, device=cuda) is a blocking pageable host-to-device copy).
_DEV_TABLES: "collections.OrderedDict" = collections.OrderedDict()
_DEV_TABLES_MAX = 256


def _dev_table(v: Sequence[int], device, dtype) -> torch.Tensor:
    key = (tuple(int(x) for x in v), str(device), dtype)
    hit = _DEV_TABLES.get(key)
    if hit is not None:
        _DEV_TABLES.move_to_end(key)
        return hit
    t = torch.tensor(key[0], dtype=dtype, device=device)
    _DEV_TABLES[key] = t
    while len(_DEV_TABLES) > _DEV_TABLES_MAX:
        _DEV_TABLES.popitem(last=False)
    return t


def _dev_i64(v: Sequence[int], device) -> torch.Tensor:
    return _dev_table(v, device, torch.int64)


def _dev_i32(v: Sequence[int], device) -> torch.Tensor:
    return _dev_table(v, device, torch.int32)


def _check_mix(mix: torch.Tensor, channels: int = 2) -> torch.Tensor:
    if not mix.is_cuda:
        raise RuntimeError("demix needs a CUDA tensor: audiolab_b200 has no CPU fallback")
    if mix.dim() != 2 or mix.shape[0] != channels:
        raise ValueError(f"mix must be [{channels}, n]")
    return mix.contiguous().float()


# ======================================================================================
# MDX-Net
# ======================================================================================
class MdxDemixer:
    """MDX-Net spectral loop around ``model_run(spek[B,4,dim_f,dim_t]) -> spec_pred`` (same shape)."""

    def __init__(self, cfg: MdxConfig, model_run: Callable[[torch.Tensor], torch.Tensor], batch_size: int = 16):
        self.cfg = cfg
        self.model_run = model_run
        self.batch_size = int(batch_size)
        self.plan = sp.StftPlan(cfg.n_fft, cfg.hop)
        self._weights = {}

    # -- net on a batch of chunks, in sub-batches that bound activation memory ----------------------
    def _net(self, spek: torch.Tensor) -> torch.Tensor:
        outs = []
        for i in range(0, spek.shape[0], self.batch_size):
            part = spek[i: i + self.batch_size]
            if self.cfg.denoise:                                  # mdxnet.py:168-172
                pred = self.model_run(part) * 0.5 - self.model_run(-part) * 0.5
            else:
                pred = self.model_run(part)
            outs.append(pred)
        out = outs[0] if len(outs) == 1 else torch.cat(outs)
        return out.contiguous().float()

    def _const_weight(self, value: float, n: int, device) -> Optional[torch.Tensor]:
        if value == 1.0:
            return None
        key = (value, n, str(device))
        if key not in self._weights:
            self._weights[key] = torch.full((n,), value, dtype=torch.float32, device=device)
        return self._weights[key]

    @torch.no_grad()
    def demix_trim_concat(self, mix: torch.Tensor, out: Optional[torch.Tensor] = None,
                          clip: Optional[Tuple[int, int]] = None) -> torch.Tensor:
        """Predictor.demix_base for one segment (mdxnet.py:147-183): [2, n] -> [2, n].

        ``out``/``clip`` let demix_segments place ``tar_signal[:, start:end]`` straight into the
        final buffer (out is then the destination row pointer for local sample ``start``).
        """
        c = self.cfg
        mix = _check_mix(mix)
        n = mix.shape[1]
        trim, gen, chunk = c.trim, c.gen_size, c.chunk_size
        pad = gen - n % gen
        n_chunks = (n + pad) // gen
        spek = self.plan.stft(mix, chunk_len=chunk, n_chunks=n_chunks, off0=-trim, off_step=gen,
                              n_frames=c.dim_t, layout=sp.CAC, n_bins_out=c.dim_f, zero_low_bins=c.zero_low_bins)
        pred = self._net(spek)
        start, end = (0, n) if clip is None else clip
        if out is None:
            out = torch.empty((2, end - start), dtype=torch.float32, device=mix.device)
        self.plan.istft(pred, n_chunks=n_chunks, channels=2, layout=sp.CAC, out_start=c.n_fft // 2 + trim,
                        out_len=gen, weight=self._const_weight(c.compensate, gen, mix.device), dst=out,
                        dst_ch_stride=out.stride(0), dst_chunk_stride=0, dst_off0=-start, dst_off_step=gen,
                        dst_limit=end - start)
        return out

    @torch.no_grad()
    def demix_segments(self, mix: torch.Tensor, chunks: int = 0, margin: int = 44100) -> torch.Tensor:
        """Predictor.demix (mdxnet.py:109-141) + margin strip / concat (:185-194)."""
        mix = _check_mix(mix)
        samples = mix.shape[-1]
        chunk_size = chunks * 44100
        if margin == 0:
            raise ValueError("margin cannot be zero!")
        if margin > chunk_size:
            margin = chunk_size
        if chunks == 0 or samples < chunk_size:
            chunk_size = samples
        segs = []
        counter = -1
        for skip in range(0, samples, chunk_size):
            counter += 1
            s_margin = 0 if counter == 0 else margin
            end = min(skip + chunk_size + margin, samples)
            segs.append((skip - s_margin, end))
            if end == samples:
                break
        # output layout: concat of tar[:, start:end_] per segment
        pieces = []
        for i, (a, b) in enumerate(segs):
            seg_len = b - a
            start = 0 if i == 0 else margin
            stop = seg_len if i == len(segs) - 1 else seg_len - margin
            pieces.append((a, b, start, stop))
        total = sum(stop - start for _, _, start, stop in pieces)
        out = torch.empty((2, total), dtype=torch.float32, device=mix.device)
        pos = 0
        for a, b, start, stop in pieces:
            self.demix_trim_concat(mix[:, a:b], out=out[:, pos:], clip=(start, stop))
            pos += stop - start
        return out

    def _hann_tables(self, lens: List[int], chunk: int, device) -> torch.Tensor:
        tabs = np.zeros((len(lens), chunk), dtype=np.float32)
        for i, l in enumerate(lens):
            tabs[i, :l] = np.hanning(l).astype(np.float32)
        return torch.from_numpy(tabs).to(device)

    @torch.no_grad()
    def demix_windowed(self, mix: torch.Tensor, is_match_mix: bool = False) -> torch.Tensor:
        """Upstream MDXSeparator.demix, windowed overlap-add form: [2, n] -> [2, n]."""
        c = self.cfg
        mix = _check_mix(mix)
        dev = mix.device
        n = mix.shape[1]
        trim, gen, chunk = c.trim, c.gen_size, c.chunk_size
        overlap = 0.02 if is_match_mix else c.overlap
        pad = gen + trim - (n % gen)
        total = trim + n + pad
        step = int((1 - overlap) * chunk)
        offs = list(range(0, total, step))
        lens = [min(chunk, total - o) for o in offs]
        n_chunks = len(offs)
        spek = self.plan.stft(mix, chunk_len=chunk, n_chunks=n_chunks, off0=-trim, off_step=step,
                              n_frames=c.dim_t, layout=sp.CAC, n_bins_out=c.dim_f,
                              zero_low_bins=c.zero_low_bins)   # upstream run_model zeroes the low bins before the
                                                               # is_match_mix branch: the identity pass loses them too
        pred = spek if is_match_mix else self._net(spek)
        waves = self.plan.istft(pred, n_chunks=n_chunks, channels=2, layout=sp.CAC, out_len=chunk)
        wtab = tab_id = None
        if overlap != 0:
            uniq = sorted(set(lens), reverse=True)
            wtab = self._hann_tables(uniq, chunk, dev)
            tab_id = _dev_i32([uniq.index(l) for l in lens], dev)
        scale = 1.0 if is_match_mix else float(c.compensate)
        track = sp.ola_gather(waves.view(n_chunks, 2, chunk), _dev_i64(offs, dev), total, wtab=wtab, tab_id=tab_id,
                              p0=trim, p1=trim + n, eps=1e-30, scale=scale,
                              out=torch.empty((2, total), dtype=torch.float32, device=dev))
        return track[:, trim: trim + n]


    # ---- secondary stem by spectral inversion (invert_using_spec, stem_separator.py:105) -------------------
    _inv_plan = None

    @torch.no_grad()
    def invert_stem(self, raw_mix: torch.Tensor, stem: torch.Tensor) -> torch.Tensor:
        """Upstream ``spec_utils.invert_stem``: STFT (n_fft 2048, hop 1024, zero centre padding -- librosa >= 0.10) of the
        match-mix pass and of the primary stem, the in-tree "invert_p" arithmetic
        (/root/reference/modules/rvc/infer/lib/uvr5_pack/lib_v5/spec_utils.py:614-623)
        ``v = y - max(|X|, |y|) * exp(1j * angle(X))``, iSTFT, sign flipped.  [2, n] x 2 -> [2, n] (the tail past
        hop * (n // hop), which librosa's istft does not produce, is zero)."""
        n_fft, hop = 2048, 1024
        if MdxDemixer._inv_plan is None:
            MdxDemixer._inv_plan = sp.StftPlan(n_fft, hop)
        plan = MdxDemixer._inv_plan
        raw_mix, stem = _check_mix(raw_mix), _check_mix(stem)
        n = raw_mix.shape[1]
        T = 1 + n // hop
        # zero centre padding = a chunk that starts n_fft/2 before the track: K1 reads zeros outside [0, n)
        kw = dict(chunk_len=n + n_fft, n_chunks=1, off0=-(n_fft // 2), off_step=0, n_valid=n, center_pad=0, n_frames=T,
                  layout=sp.FRAME_MAJOR)
        X = plan.stft(raw_mix, **kw)                                   # c64 [2, T, F]
        y = plan.stft(stem, **kw)
        x_mag, y_mag = X.abs(), y.abs()
        max_mag = torch.where(x_mag >= y_mag, x_mag, y_mag)
        unit = torch.where(x_mag > 0, X / x_mag.clamp(min=1e-30), torch.ones_like(X))   # exp(1j * angle(X)); angle(0) = 0
        v = (y - max_mag * unit).contiguous()
        wave = plan.istft(v, n_chunks=1, channels=2, layout=sp.FRAME_MAJOR)              # [1, 1, 2, hop * (T - 1)]
        out = torch.zeros_like(raw_mix)
        m = min(n, wave.shape[-1])
        out[:, :m] = -wave[0, 0, :, :m]
        return out

    @torch.no_grad()
    def secondary_by_inversion(self, mix: torch.Tensor, primary: torch.Tensor) -> torch.Tensor:
        """``invert_stem(demix(mix, is_match_mix=True), primary)`` -- upstream MDXSeparator.separate with invert_using_spec."""
        return self.invert_stem(self.demix_windowed(mix, is_match_mix=True), primary)


# ======================================================================================
# BS-RoFormer / Mel-Band RoFormer
# ======================================================================================
def roformer_schedule(n: int, chunk: int, step: int) -> Tuple[List[int], List[int]]:
    """(offsets, multiplicities) of SURVEY.md A.2's loop; repeated tail-aligned chunks are merged."""
    offs, mult, tail = [], [], 0
    for i in range(0, n, step):
        if i + chunk > n:
            tail += 1
        else:
            offs.append(i)
            mult.append(1)
    if tail:
        if offs and offs[-1] == n - chunk:      # the tail coincides with the last regular chunk
            mult[-1] += tail
        else:
            offs.append(n - chunk)
            mult.append(tail)
    return offs, mult


def hamming_sym(n: int) -> np.ndarray:
    """scipy.signal.windows.hamming(n) (symmetric), float32 like the reference's torch.tensor cast."""
    if n == 1:
        return np.ones(1, dtype=np.float32)
    k = np.arange(n, dtype=np.float64)
    return (0.54 - 0.46 * np.cos(2.0 * np.pi * k / (n - 1))).astype(np.float32)


class RoformerDemixer:
    """mix [s, n] -> stems [num_stems, s, n] with Hamming-weighted overlap-add over 8 s chunks."""

    def __init__(self, cfg: RoformerConfig, net, batch_size: int = 4):
        self.cfg = cfg
        self.net = net                      # audiolab_b200.nets.roformer.RoformerMaskNet on the device
        self.batch_size = int(batch_size)
        if cfg.stft_win_length != cfg.stft_n_fft:
            raise ValueError("win_length != n_fft is not used by any reference model")
        self.plan = sp.StftPlan(cfg.stft_n_fft, cfg.stft_hop_length, normalized=cfg.stft_normalized)
        self._wtab = {}

    def window(self, device) -> torch.Tensor:
        key = str(device)
        if key not in self._wtab:
            self._wtab[key] = torch.from_numpy(hamming_sym(self.cfg.chunk_size)[None]).to(device)
        return self._wtab[key]

    @torch.no_grad()
    def chunk_waves(self, mix: torch.Tensor, offs: Sequence[int], n_valid: Optional[int] = None,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Run STFT -> net -> mask (.) STFT -> iSTFT for the chunks at `offs`: [len(offs), stems*s, C]."""
        c = self.cfg
        C, s, stems = c.chunk_size, c.audio_channels, c.num_stems
        dev = mix.device
        n_frames = 1 + C // c.stft_hop_length
        if out is None:
            out = torch.empty((len(offs), stems * s, C), dtype=torch.float32, device=dev)
        for i in range(0, len(offs), self.batch_size):
            part = list(offs[i: i + self.batch_size])
            b = len(part)
            affine = all(part[k + 1] - part[k] == part[1] - part[0] for k in range(b - 1)) if b > 1 else True
            kw = dict(off0=part[0], off_step=(part[1] - part[0]) if b > 1 else 0) if affine \
                else dict(offsets=_dev_i64(part, dev))
            spec = self.plan.stft(mix, chunk_len=C, n_chunks=b, n_valid=n_valid, n_frames=n_frames,
                                  layout=sp.FRAME_INTERLEAVED, **kw)                   # [b, T, F, s]
            mask = self.net.mask(spec)                                                 # [b, n, T, F, s]
            self.plan.istft(spec, mask=mask.contiguous(), n_chunks=b, channels=s, stems=stems,
                            layout=sp.FRAME_INTERLEAVED, out_len=C, dst=out[i: i + b],
                            dst_ch_stride=C, dst_chunk_stride=stems * s * C, dst_limit=C)
        return out

    @torch.no_grad()
    def demix(self, mix: torch.Tensor) -> torch.Tensor:
        c = self.cfg
        s, stems, C = c.audio_channels, c.num_stems, c.chunk_size
        mix = _check_mix(mix, s)
        dev = mix.device
        n = mix.shape[1]
        if n < C:
            # undefined upstream; zero-pad to one chunk, run once, crop (same rule as the oracle)
            waves = self.chunk_waves(mix, [0], n_valid=n)
            return waves[0].view(stems, s, C)[..., :n].contiguous()
        offs, mult = roformer_schedule(n, C, c.step)
        waves = self.chunk_waves(mix, offs)
        track = sp.ola_gather(waves, _dev_i64(offs, dev), n, mult=_dev_i32(mult, dev), wtab=self.window(dev),
                              eps=1e-10, out=torch.empty((stems * s, n), dtype=torch.float32, device=dev))
        return track.view(stems, s, n)


# ======================================================================================
# HTDemucs
# ======================================================================================
def triangle_weight(segment: int, power: float = 1.0) -> np.ndarray:
    w = np.concatenate([np.arange(1, segment // 2 + 1), np.arange(segment - segment // 2, 0, -1)]).astype(np.float32)
    return (w / w.max()) ** power


class HTDemucsDemixer:
    """DemucsSeparator.demix_demucs around ``core(mag[B,4,F,T], xt[B,2,L]) -> (x_spec[B,S,4,F,T], x_time[B,S,2,L])``."""

    def __init__(self, cfg: HTDemucsConfig, core: Callable, batch_size: int = 8):
        self.cfg = cfg
        self.core = core
        self.batch_size = int(batch_size)
        self.plan = sp.StftPlan(cfg.nfft, cfg.hop, normalized=True)
        self._wtab = {}

    def _weight(self, device) -> torch.Tensor:
        key = str(device)
        if key not in self._wtab:
            w = triangle_weight(self.cfg.segment_samples, self.cfg.transition_power)
            self._wtab[key] = torch.from_numpy(w[None]).to(device)
        return self._wtab[key]

    @torch.no_grad()
    def _segments(self, tensor: torch.Tensor, starts: Sequence[int]) -> torch.Tensor:
        """HTDemucs.forward for the zero-padded windows tensor[:, st : st+segment] -> [B, S, 2, segment]."""
        c = self.cfg
        seg, hl = c.segment_samples, c.hop
        dev = tensor.device
        total = tensor.shape[1]
        le = int(math.ceil(seg / hl))
        pad = hl // 2 * 3
        outs = []
        for i in range(0, len(starts), self.batch_size):
            part = list(starts[i: i + self.batch_size])
            b = len(part)
            mag = self.plan.stft(tensor, chunk_len=seg, n_chunks=b, offsets=_dev_i64(part, dev), center_pad=pad,
                                 n_frames=le, layout=sp.CAC, n_bins_out=c.nfft // 2)      # [b, 4, 2048, le]
            # time-branch input: the same zero-padded windows
            xt = torch.zeros((b, 2, seg), dtype=torch.float32, device=dev)
            for k, st in enumerate(part):
                a, e = max(0, st), min(total, st + seg)
                if e > a:
                    xt[k, :, a - st: e - st] = tensor[:, a:e]
            mean = mag.mean(dim=(1, 2, 3), keepdim=True)
            std = mag.std(dim=(1, 2, 3), keepdim=True)
            meant = xt.mean(dim=(1, 2), keepdim=True)
            stdt = xt.std(dim=(1, 2), keepdim=True)
            x, xt_out = self.core((mag - mean) / (1e-5 + std), (xt - meant) / (1e-5 + stdt))
            x = (x.float() * std[:, None] + mean[:, None]).contiguous()                    # [b, S, 4, F, le]
            S = x.shape[1]
            wav = self.plan.istft(x, n_chunks=b, channels=2, stems=S, layout=sp.CAC, spec_has_stems=True,
                                  frame_pad=2, out_start=c.nfft // 2 + pad, out_len=seg)  # [b, S, 2, seg]
            wav += xt_out.float() * stdt[:, None] + meant[:, None]
            outs.append(wav)
        return outs[0] if len(outs) == 1 else torch.cat(outs)

    @torch.no_grad()
    def apply_split(self, tensor: torch.Tensor, base: int, length: int) -> torch.Tensor:
        """demucs.apply.apply_model(split=True) on tensor[:, base:base+length] -> [S, 2, length]."""
        c = self.cfg
        seg = c.segment_samples
        dev = tensor.device
        stride = int((1 - c.overlap) * seg)
        offs = list(range(0, length, stride))
        starts, shifts_in = [], []
        for off in offs:
            clen = min(length - off, seg)
            delta = seg - clen
            starts.append(base + off - delta // 2)      # TensorChunk.padded: centred, real context
            shifts_in.append(delta // 2)                # center_trim
        waves = self._segments(tensor, starts)          # [B, S, 2, seg]
        B, S = waves.shape[0], waves.shape[1]
        for k, sh in enumerate(shifts_in):
            if sh:
                clen = min(length - offs[k], seg)
                waves[k, ..., :clen] = waves[k, ..., sh: sh + clen].clone()
        track = sp.ola_gather(waves.view(B, S * 2, seg), _dev_i64(offs, dev), length, wtab=self._weight(dev),
                              eps=1e-30, out=torch.empty((S * 2, length), dtype=torch.float32, device=dev))
        return track.view(S, 2, length)

    def shift_offsets(self, seed: int = 0) -> List[int]:
        rng = random.Random(seed)
        max_shift = int(0.5 * self.cfg.samplerate)
        return [rng.randint(0, max_shift) for _ in range(self.cfg.shifts)]

    @torch.no_grad()
    def demix(self, mix: torch.Tensor, seed: int = 0) -> torch.Tensor:
        """mix [2, L] -> sources [S, 2, L]."""
        c = self.cfg
        mix = _check_mix(mix)
        ref = mix.mean(0)
        mean, std = ref.mean(), ref.std()
        x = (mix - mean) / std
        length = x.shape[1]
        if not c.shifts:
            out = self.apply_split(x, 0, length)
        else:
            max_shift = int(0.5 * c.samplerate)
            padded = F.pad(x, (max_shift, max_shift))
            out = None
            for offset in self.shift_offsets(seed):
                sh = self.apply_split(padded, offset, length + max_shift - offset)[..., max_shift - offset:]
                out = sh if out is None else out + sh
            out = out / c.shifts
        return out * std + mean
