"""BS-RoFormer / Mel-Band RoFormer mask networks for the CUDA demix path.

The module tree and parameter names follow upstream (lucidrains / ZFTurbo as vendored by
`audio-separator`), so a released checkpoint's ``state_dict`` loads unchanged.
What differs from upstream's ``forward``: STFT, the complex mask multiply and the iSTFT are
NOT here -- they are the al_stft / al_istft kernels.  ``mask()`` consumes the spectrogram in the
kernels' FRAME_INTERLEAVED layout ``[b, t, f, s]`` complex64 (bit-identical to upstream's
``'b s f t c -> b t (f s c)'`` view) and returns the mask ``[b, n, t, f, s]`` complex64 in the
layout al_istft multiplies in-kernel, so no permute copies surround the network.

Compute dtype: ``torch.bfloat16`` runs the network under autocast (the reference constructs its
Separator with ``use_autocast=True``, /root/reference/modules/separator/stem_separator.py:106);
``torch.float32`` is the parity configuration.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

import os

from ..configs import RoformerConfig

_BAND_ATTN = os.environ.get("AUDIOLAB_B200_BAND_ATTN") == "1"   # opt-in until measured on a B200 (NOTES.md)
_BAND_ATTN_TC = os.environ.get("AUDIOLAB_B200_BAND_ATTN", "1") != "0"   # band-axis attention kernel inside the tc path (default)
# time-axis attention kernel (csrc/al_fattn.cu, tcgen05 flash attention + gate) inside the tc path (default); 0 = cuDNN SDPA +
# gate pass (the comparison path: 4.39 ms per call against 4.27 ms, profiles/r02u_*)
_TIME_ATTN_TC = os.environ.get("AUDIOLAB_B200_TIME_ATTN", "1") != "0"
_GROUPED = os.environ.get("AUDIOLAB_B200_GROUPED", "1") != "0"   # band split / mask estimator as grouped tcgen05 GEMMs
_TC_GEMM = os.environ.get("AUDIOLAB_B200_TC_GEMM", "1") != "0"  # tcgen05 GEMM path (default); 0 = cuBLAS comparison path


class RMSNorm(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.scale = dim ** 0.5
        self.gamma = nn.Parameter(torch.ones(dim))

    def forward(self, x):
        return F.normalize(x, dim=-1) * self.scale * self.gamma


class RotaryEmbedding(nn.Module):
    def __init__(self, dim: int, theta: float = 10000.0):
        super().__init__()
        self.freqs = nn.Parameter(1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim)),
                                  requires_grad=False)
        self._cache = {}

    def tables(self, n: int, device, dtype) -> Tuple[torch.Tensor, torch.Tensor]:
        key = (n, str(device), dtype)
        hit = self._cache.get(key)
        if hit is None:
            pos = torch.arange(n, device=device, dtype=torch.float32)
            ang = torch.repeat_interleave(pos[:, None] * self.freqs.to(device)[None, :], 2, dim=-1)
            hit = (ang.cos().to(dtype), ang.sin().to(dtype))
            self._cache[key] = hit
        return hit

    def rotate(self, t: torch.Tensor) -> torch.Tensor:
        cos, sin = self.tables(t.shape[-2], t.device, t.dtype)
        x = t.unflatten(-1, (-1, 2))
        rot = torch.stack((-x[..., 1], x[..., 0]), dim=-1).flatten(-2)
        return t * cos + rot * sin


class FeedForward(nn.Module):
    def __init__(self, dim: int, mult: int = 4):
        super().__init__()
        inner = int(dim * mult)
        self.net = nn.Sequential(RMSNorm(dim), nn.Linear(dim, inner), nn.GELU(), nn.Dropout(0.0),
                                 nn.Linear(inner, dim), nn.Dropout(0.0))

    def forward(self, x):
        return self.net(x)


class Attention(nn.Module):
    def __init__(self, dim: int, heads: int, dim_head: int, rotary_embed: Optional[RotaryEmbedding]):
        super().__init__()
        self.heads = heads
        inner = heads * dim_head
        self.rotary_embed = rotary_embed
        self.norm = RMSNorm(dim)
        self.to_qkv = nn.Linear(dim, inner * 3, bias=False)
        self.to_gates = nn.Linear(dim, heads)
        self.to_out = nn.Sequential(nn.Linear(inner, dim, bias=False), nn.Dropout(0.0))

    def forward(self, x):
        b, n, _ = x.shape
        x = self.norm(x)
        q, k, v = self.to_qkv(x).view(b, n, 3, self.heads, -1).permute(2, 0, 3, 1, 4)
        if self.rotary_embed is not None:
            q = self.rotary_embed.rotate(q)
            k = self.rotary_embed.rotate(k)
        out = F.scaled_dot_product_attention(q, k, v)
        gates = self.to_gates(x)
        out = out * gates.permute(0, 2, 1).unsqueeze(-1).sigmoid()
        return self.to_out(out.permute(0, 2, 1, 3).reshape(b, n, -1))


class Transformer(nn.Module):
    def __init__(self, dim, depth, heads, dim_head, ff_mult, rotary_embed, norm_output):
        super().__init__()
        self.layers = nn.ModuleList([
            nn.ModuleList([Attention(dim, heads, dim_head, rotary_embed), FeedForward(dim, ff_mult)])
            for _ in range(depth)
        ])
        self.norm = RMSNorm(dim) if norm_output else nn.Identity()

    def forward(self, x):
        for attn, ff in self.layers:
            x = attn(x) + x
            x = ff(x) + x
        return self.norm(x)


class BandSplit(nn.Module):
    def __init__(self, dim: int, dim_inputs: Tuple[int, ...]):
        super().__init__()
        self.dim_inputs = tuple(dim_inputs)
        self.to_features = nn.ModuleList([nn.Sequential(RMSNorm(d), nn.Linear(d, dim)) for d in dim_inputs])

    def forward(self, x):
        parts = x.split(self.dim_inputs, dim=-1)
        return torch.stack([f(p) for p, f in zip(parts, self.to_features)], dim=-2)


def _mlp(dim_in, dim_out, dim_hidden, depth):
    dims = (dim_in,) + (dim_hidden,) * (depth - 1) + (dim_out,)
    net = []
    for i, (a, b) in enumerate(zip(dims[:-1], dims[1:])):
        net.append(nn.Linear(a, b))
        if i != len(dims) - 2:
            net.append(nn.Tanh())
    return nn.Sequential(*net)


class MaskEstimator(nn.Module):
    def __init__(self, dim, dim_inputs, depth, mlp_expansion_factor=4):
        super().__init__()
        self.dim_inputs = tuple(dim_inputs)
        self.to_freqs = nn.ModuleList([
            nn.Sequential(_mlp(dim, d * 2, dim * mlp_expansion_factor, depth), nn.GLU(dim=-1)) for d in dim_inputs
        ])

    def forward(self, x):
        return torch.cat([mlp(b) for b, mlp in zip(x.unbind(dim=-2), self.to_freqs)], dim=-1)


# ---- mel band membership (librosa.filters.mel, Slaney scale / norm; only `> 0` is used) ------------
def _hz_to_mel(f):
    f = np.asarray(f, dtype=np.float64)
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, f / f_sp)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_band_membership(sr: int, n_fft: int, n_mels: int) -> np.ndarray:
    fftfreqs = np.linspace(0, sr / 2.0, 1 + n_fft // 2)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(0.0), _hz_to_mel(sr / 2.0), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    fb = np.zeros((n_mels, 1 + n_fft // 2))
    for i in range(n_mels):
        fb[i] = np.maximum(0, np.minimum(-ramps[i] / fdiff[i], ramps[i + 2] / fdiff[i + 1]))
    fb *= (2.0 / (mel_f[2: n_mels + 2] - mel_f[:n_mels]))[:, None]
    fb = fb.astype(np.float32)
    fb[0, 0] = 1.0
    fb[-1, -1] = 1.0
    member = fb > 0
    if not member.any(axis=0).all():
        raise ValueError("mel bands do not cover every frequency bin")
    return member


class RoformerMaskNet(nn.Module):
    """spec [b, t, f, s] complex64 -> mask [b, n, t, f, s] complex64."""

    def __init__(self, cfg: RoformerConfig):
        super().__init__()
        self.cfg = c = cfg
        mel = c.kind == "mel"
        time_rot, freq_rot = RotaryEmbedding(c.dim_head), RotaryEmbedding(c.dim_head)
        self.layers = nn.ModuleList([
            nn.ModuleList([
                Transformer(c.dim, c.time_transformer_depth, c.heads, c.dim_head, c.ff_mult, time_rot, mel),
                Transformer(c.dim, c.freq_transformer_depth, c.heads, c.dim_head, c.ff_mult, freq_rot, mel),
            ]) for _ in range(c.depth)
        ])
        ch = c.audio_channels
        if mel:
            member = torch.from_numpy(mel_band_membership(c.sample_rate, c.stft_n_fft, c.num_bands))
            freqs = member.shape[1]
            freq_indices = torch.arange(freqs)[None].expand(c.num_bands, freqs)[member]
            if c.stereo:
                freq_indices = (freq_indices[:, None] * 2 + torch.arange(2)[None]).reshape(-1)
            self.register_buffer("freq_indices", freq_indices, persistent=False)
            self.register_buffer("num_bands_per_freq", member.sum(dim=0), persistent=False)
            bands = member.sum(dim=1).tolist()
        else:
            self.final_norm = RMSNorm(c.dim)
            bands = list(c.freqs_per_bands)
            if sum(bands) != c.stft_n_fft // 2 + 1:
                raise ValueError("freqs_per_bands must sum to n_fft/2+1")
        dims = tuple(2 * f * ch for f in bands)
        self.band_split = BandSplit(c.dim, dims)
        self.mask_estimators = nn.ModuleList([
            MaskEstimator(c.dim, dims, c.mask_estimator_depth, c.mlp_expansion_factor) for _ in range(c.num_stems)
        ])
        self.compute_dtype = torch.float32
        self._bf16_cache = {}
        self._rot_cache = {}
        self._fused_dtype = torch.bfloat16      # the kernels are bf16-only; tests of the host logic may override

    # ---- bf16 inference path: fused row-wise kernels (csrc/al_netops.cu) between the contractions -------
    def _bf16(self, p: torch.Tensor) -> torch.Tensor:
        """bf16 copy of a parameter, made once (autocast re-casts every weight on every call)."""
        hit = self._bf16_cache.get(id(p))
        if hit is None or hit[0] != p._version or hit[1].device != p.device:
            hit = (p._version, p.detach().to(self._fused_dtype).contiguous())
            self._bf16_cache[id(p)] = hit
        return hit[1]

    def _cos_sin(self, rot: RotaryEmbedding, n: int, device) -> torch.Tensor:
        key = (id(rot), n, str(device))
        hit = self._rot_cache.get(key)
        if hit is None:
            pos = torch.arange(n, device=device, dtype=torch.float32)
            ang = pos[:, None] * rot.freqs.detach().to(device=device, dtype=torch.float32)[None, :]
            hit = torch.stack((ang.cos(), ang.sin()), dim=-1).contiguous()      # [n, dim_head/2, 2]
            self._rot_cache[key] = hit
        return hit

    def _transformer_fused(self, x2, tr: Transformer, geom, time_axis: bool, pending):
        """x2 [b*t*f, d] bf16 residual stream (token order b, t, f), updated in place.  `pending` is the
        output-projection bias of the previous FeedForward, folded into the next RMSNorm kernel."""
        from .. import netops
        b, t, f = geom
        for attn, ff in tr.layers:
            h, dh = attn.heads, attn.to_qkv.weight.shape[0] // (3 * attn.heads)
            inner = h * dh
            xn = netops.rmsnorm(x2, attn.norm.gamma.detach(), pending)
            pending = None
            w = self._bf16(attn.to_qkv.weight)
            q, k, v = F.linear(xn, w[:inner]), F.linear(xn, w[inner:2 * inner]), F.linear(xn, w[2 * inner:])
            band_kernel = not time_axis and _BAND_ATTN and dh == 64 and f <= 64
            if attn.rotary_embed is not None and not band_kernel:
                if time_axis:
                    netops.rotary_(q, k, self._cos_sin(attn.rotary_embed, t, x2.device), h, dh, f, t)
                else:
                    netops.rotary_(q, k, self._cos_sin(attn.rotary_embed, f, x2.device), h, dh, 1, f)
            # attention over time: batch b, "heads" (band, head); over bands: batch (b, t).  Strided views of the
            # token-major buffers -- no transposition copies around the attention.
            if band_kernel:
                # opt-in (AUDIOLAB_B200_BAND_ATTN=1): our mma.sync kernel for the short band axis (csrc/al_attn.cu),
                # with the rotary embedding folded into its tile staging and the sigmoid gate into its epilogue
                gates = F.linear(xn, self._bf16(attn.to_gates.weight), self._bf16(attn.to_gates.bias))
                cs = None if attn.rotary_embed is None else self._cos_sin(attn.rotary_embed, f, x2.device)
                o2 = netops.band_attention(q, k, v, b * t, f, h, dh, gates=gates, cos_sin=cs)
                x2.addmm_(o2, self._bf16(attn.to_out[0].weight).t())
            else:
                shape = (b, t, f * h, dh) if time_axis else (b * t, f, h, dh)
                o = F.scaled_dot_product_attention(q.view(shape).transpose(1, 2), k.view(shape).transpose(1, 2),
                                                   v.view(shape).transpose(1, 2))
                o = o.transpose(1, 2)
                if not o.is_contiguous():
                    o = o.contiguous()
                o2 = o.view(-1, inner)
                gates = F.linear(xn, self._bf16(attn.to_gates.weight), self._bf16(attn.to_gates.bias))
                netops.gate_sigmoid_(o2, gates, h, dh)
                x2.addmm_(o2, self._bf16(attn.to_out[0].weight).t())
            lin1, lin2 = ff.net[1], ff.net[4]
            xn = netops.rmsnorm(x2, ff.net[0].gamma.detach())
            hid = netops.gelu_(F.linear(xn, self._bf16(lin1.weight), self._bf16(lin1.bias)))
            x2.addmm_(hid, self._bf16(lin2.weight).t())
            pending = lin2.bias.detach()
        if isinstance(tr.norm, RMSNorm):
            x2 = netops.rmsnorm(x2, tr.norm.gamma.detach(), pending)
            pending = None
        return x2, pending

    def _axial_fused(self, x):
        """bf16 twin of `_axial` (+ BS-RoFormer's final norm): same parameters, same arithmetic order, fp32 inside
        each fused kernel."""
        from .. import netops
        b, t, f, d = x.shape
        x2 = x.to(self._fused_dtype).contiguous().view(-1, d)
        pending = None
        for time_transformer, freq_transformer in self.layers:
            x2, pending = self._transformer_fused(x2, time_transformer, (b, t, f), True, pending)
            x2, pending = self._transformer_fused(x2, freq_transformer, (b, t, f), False, pending)
        if self.cfg.kind != "mel":
            x2 = netops.rmsnorm(x2, self.final_norm.gamma.detach(), pending)
        elif pending is not None:
            x2 = x2 + pending.to(x2.dtype)
        return x2.view(b, t, f, d)

    # ---- bf16 inference path on the tcgen05 GEMM (csrc/al_gemm.cu): fp32 residual stream, RMSNorm / rotary / GELU /
    # bias / residual adds folded into the GEMM epilogues ---------------------------------------------------------
    def _tc_supported(self) -> bool:
        """Shapes the fused epilogues cover: 64-wide heads (one rotary head per 64-column epilogue step), q / k / v
        column blocks on N-tile boundaries, residual width on 128-column slabs."""
        c = self.cfg
        inner = c.heads * c.dim_head
        return (_TC_GEMM and c.dim_head == 64 and inner % 256 == 0 and c.heads <= 16 and c.dim % 128 == 0
                and c.dim % 8 == 0 and int(c.dim * c.ff_mult) % 8 == 0)

    def _tc_pack(self, attn: Attention, ff: FeedForward):
        """bf16 operands of one (attention, feed-forward) pair, made once: gamma of the pre-norms folded into the
        columns of to_qkv / to_gates / ff.Linear1 (RMSNorm(x) W^T = diag(rowscale) x (gamma (.) W)^T), to_gates appended
        to to_qkv as 16 extra output rows (8 real + zero padding) with its bias."""
        params = (attn.norm.gamma, attn.to_qkv.weight, attn.to_gates.weight, attn.to_gates.bias, attn.to_out[0].weight,
                  ff.net[0].gamma, ff.net[1].weight, ff.net[1].bias, ff.net[4].weight, ff.net[4].bias)
        key = ("tc", id(attn), self._fused_dtype)
        ver = tuple(p._version for p in params) + (str(params[0].device),)
        hit = self._bf16_cache.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1]
        with torch.no_grad():
            h = attn.heads
            g1 = attn.norm.gamma.detach().float()
            wq = attn.to_qkv.weight.detach().float() * g1[None, :]
            wg = attn.to_gates.weight.detach().float() * g1[None, :]
            n3 = wq.shape[0]
            w_qkvg = torch.zeros(n3 + 16, wq.shape[1], device=wq.device)
            w_qkvg[:n3] = wq
            w_qkvg[n3:n3 + h] = wg
            b_qkvg = torch.zeros(n3 + 16, device=wq.device)
            b_qkvg[n3:n3 + h] = attn.to_gates.bias.detach().float()
            g2 = ff.net[0].gamma.detach().float()
            pack = {
                "w_qkvg": w_qkvg.to(self._fused_dtype).contiguous(), "b_qkvg": b_qkvg.contiguous(),
                "w_out": attn.to_out[0].weight.detach().to(self._fused_dtype).contiguous(),
                "w1": (ff.net[1].weight.detach().float() * g2[None, :]).to(self._fused_dtype).contiguous(),
                "b1": ff.net[1].bias.detach().float().contiguous(),
                "w2": ff.net[4].weight.detach().to(self._fused_dtype).contiguous(),
                "b2": ff.net[4].bias.detach().float().contiguous(),
            }
        self._bf16_cache[key] = (ver, pack)
        return pack

    def _transformer_tc(self, st, tr: Transformer, geom, time_axis: bool):
        from .. import netops
        b, t, f = geom
        x32, xb, ss, q, k, v, gates, hid = st
        d = x32.shape[1]
        scale = float(d) ** 0.5
        for attn, ff in tr.layers:
            h = attn.heads
            dh = attn.to_qkv.weight.shape[0] // (3 * h)
            inner = h * dh
            pk = self._tc_pack(attn, ff)
            rot = attn.rotary_embed
            cs = None
            if rot is not None:
                cs = self._cos_sin(rot, t if time_axis else f, x32.device)
            # to_qkv + to_gates of RMSNorm(x): row scale from the sums of squares the last residual epilogue left
            netops.gemm_bf16(xb, pk["w_qkvg"], [q, k, v, gates], bias=pk["b_qkvg"], row_ss=ss, ss_scale=scale,
                             cos_sin=cs, pos_div=(f if time_axis else 1), pos_mod=(t if time_axis else f),
                             rot_cols=2 * inner if cs is not None else 0, out_split=inner)
            if not time_axis and _BAND_ATTN_TC and f <= 64:
                # band axis (<= 64 tokens per sequence): our kernel (csrc/al_attn.cu), sigmoid gate folded into its epilogue
                o2 = netops.band_attention(q, k, v, b * t, f, h, dh, gates=gates[:, :h])
            elif time_axis and _TIME_ATTN_TC:
                # time axis: tcgen05 flash attention straight on the token-major layout, sigmoid gate in its epilogue
                o2 = netops.time_attention(q, k, v, b, t, f, h, dh, gates=gates[:, :h])
            else:
                shape = (b, t, f * h, dh) if time_axis else (b * t, f, h, dh)
                o = F.scaled_dot_product_attention(q.view(shape).transpose(1, 2), k.view(shape).transpose(1, 2),
                                                   v.view(shape).transpose(1, 2))
                o = o.transpose(1, 2)
                if not o.is_contiguous():
                    o = o.contiguous()
                o2 = o.view(-1, inner)
                netops.gate_sigmoid_(o2, gates[:, :h], h, dh)
            netops.gemm_bf16_residual(o2, pk["w_out"], x32, xb, ss)                       # x += to_out(gated attention)
            netops.gemm_bf16(xb, pk["w1"], hid, bias=pk["b1"], row_ss=ss, ss_scale=scale, act="gelu")
            netops.gemm_bf16_residual(hid, pk["w2"], x32, xb, ss, bias=pk["b2"])          # x += ff(x)
        if isinstance(tr.norm, RMSNorm):
            netops.resid_prepare(x32, x32, xb, ss, gamma=tr.norm.gamma.detach().float())

    # ---- band split / mask estimator as grouped GEMMs (BS-RoFormer; bands of equal width form one grouped call) ----------
    def _band_classes(self):
        """[(f0, f1, d, off)]: maximal runs of consecutive bands with the same input width d, and the column offset of the
        run in the 'b t (f s c)' feature row."""
        dims = list(self.band_split.dim_inputs)
        out, off, f0 = [], 0, 0
        while f0 < len(dims):
            f1 = f0
            while f1 < len(dims) and dims[f1] == dims[f0]:
                f1 += 1
            out.append((f0, f1, dims[f0], off))
            off += (f1 - f0) * dims[f0]
            f0 = f1
        return out

    def _grouped_supported(self) -> bool:
        c = self.cfg
        if c.kind == "mel" or c.mask_estimator_depth != 2 or not self._tc_supported():
            return False
        for f0, f1, d, off in self._band_classes():
            # TMA: 16-byte aligned bases and group strides; a lone band may be padded up to a multiple of 8 columns
            if off % 8 != 0 or d % 4 != 0 or (f1 - f0 > 1 and d % 8 != 0):
                return False
        return True

    def _grouped_pack(self):
        key = ("grouped", self._fused_dtype)
        params = list(self.band_split.parameters()) + list(self.mask_estimators.parameters())
        ver = tuple(p._version for p in params) + (str(params[0].device),)
        hit = self._bf16_cache.get(key)
        if hit is not None and hit[0] == ver:
            return hit[1]
        from .. import netops
        dt, dev = self._fused_dtype, params[0].device
        classes = self._band_classes()
        total = sum((f1 - f0) * d for f0, f1, d, _ in classes)
        with torch.no_grad():
            offs = [0]
            for d in self.band_split.dim_inputs:
                offs.append(offs[-1] + d)
            pack = {
                "total": total, "ld": (total + 7) // 8 * 8 + 8,          # A rows: zero padding behind the last band
                "band_off": torch.tensor(offs, dtype=torch.int32, device=dev),
                "gamma": torch.cat([f[0].gamma.detach().float() for f in self.band_split.to_features]).contiguous(),
                "split": [], "est": [],
            }
            for f0, f1, d, off in classes:
                k = (d + 7) // 8 * 8
                w = torch.zeros((f1 - f0, self.cfg.dim, k), device=dev)
                for j in range(f0, f1):
                    w[j - f0, :, :d] = self.band_split.to_features[j][1].weight.detach().float()
                b = torch.stack([self.band_split.to_features[j][1].bias.detach().float() for j in range(f0, f1)])
                pack["split"].append((f0, f1, k, off, w.to(dt).contiguous(), b.contiguous()))
            for est in self.mask_estimators:
                lin1 = [m[0][0] for m in est.to_freqs]
                lin2 = [m[0][2] for m in est.to_freqs]
                w1 = torch.stack([l.weight.detach().float() for l in lin1]).to(dt).contiguous()        # [F, hidden, dim]
                b1 = torch.stack([l.bias.detach().float() for l in lin1]).contiguous()
                second = []
                for f0, f1, d, off in classes:
                    n = (2 * d + 15) // 16 * 16
                    w2 = torch.zeros((f1 - f0, n, w1.shape[1]), device=dev)
                    b2 = torch.zeros((f1 - f0, n), device=dev)
                    for j in range(f0, f1):
                        w2[j - f0, : 2 * d] = netops.interleave_glu(lin2[j].weight.detach().float())
                        b2[j - f0, : 2 * d] = netops.interleave_glu(lin2[j].bias.detach().float())
                    second.append((f0, f1, d, off, w2.to(dt).contiguous(), b2.contiguous()))
                pack["est"].append((w1, b1, second))
        self._bf16_cache[key] = (ver, pack)
        return pack

    def _band_split_tc(self, feats: torch.Tensor, st) -> None:
        """feats fp32 [b*t, (f s c)] -> the residual stream (x32, xb, ss) of `st`: per-band RMSNorm (al_band_norm), then one
        grouped GEMM per run of equal-width bands whose residual epilogue STARTS the fp32 stream (no read of x32)."""
        from .. import netops
        pk = self._grouped_pack()
        x32, xb, ss = st[0], st[1], st[2]
        bt = feats.shape[0]
        nb, d = len(self.band_split.dim_inputs), x32.shape[1]
        key = ("xn", bt, str(feats.device), self._fused_dtype)
        xn = self._rot_cache.get(key)
        if xn is None:
            xn = torch.zeros((bt, pk["ld"]), device=feats.device, dtype=self._fused_dtype)   # padding columns stay zero
            self._rot_cache[key] = xn
        netops.band_norm(feats, pk["gamma"], pk["band_off"], xn)
        x3, xb3, ss3 = x32.view(bt, nb, d), xb.view(bt, nb, d), ss.view(bt, nb, -1)
        for f0, f1, k, off, w, b in pk["split"]:
            a = xn.as_strided((f1 - f0, bt, k), (k if f1 - f0 > 1 else 0, xn.stride(0), 1), off)
            netops.gemm_bf16_residual(a, w, x3[:, f0:f1].transpose(0, 1), xb3[:, f0:f1].transpose(0, 1),
                                      ss3[:, f0:f1].transpose(0, 1), bias=b, accumulate=False)

    def _mask_estimate_tc(self, xb: torch.Tensor, b: int, t: int) -> torch.Tensor:
        """xb bf16 [b*t*F, dim] (final-normed) -> masks fp32 [b, n, t, (f s c)]: per stem one grouped GEMM for the first
        Linear + tanh of all bands, then one grouped GEMM per run of equal-width bands whose epilogue applies the GLU and
        writes fp32 straight into the mask tensor."""
        from .. import netops
        pk = self._grouped_pack()
        bt = b * t
        nb, d = len(self.band_split.dim_inputs), xb.shape[1]
        a1 = xb.view(bt, nb, d).transpose(0, 1)                                   # [F, bt, dim], strided
        outs = []
        for w1, b1, second in pk["est"]:
            hid = torch.empty((nb, bt, w1.shape[1]), device=xb.device, dtype=self._fused_dtype)
            netops.gemm_bf16(a1, w1, hid, bias=b1, act="tanh")
            m = torch.empty((bt, pk["total"]), device=xb.device, dtype=torch.float32)
            for f0, f1, dd, off, w2, b2 in second:
                out = m.as_strided((f1 - f0, bt, dd), (dd if f1 - f0 > 1 else 0, m.stride(0), 1), off)
                netops.gemm_bf16_glu(hid[f0:f1], w2, out, bias=b2)
            outs.append(m.view(b, t, -1))
        return torch.stack(outs, dim=1) if len(outs) > 1 else outs[0].unsqueeze(1)

    def _tc_state(self, m: int, dev):
        from .. import netops
        c = self.cfg
        d, inner = c.dim, c.heads * c.dim_head
        dt = self._fused_dtype
        x32 = torch.empty((m, d), device=dev, dtype=torch.float32)
        xb = torch.empty((m, d), device=dev, dtype=dt)
        ss = torch.empty((m, d // netops.resid_slab(d)), device=dev, dtype=torch.float32)
        q, k, v = (torch.empty((m, inner), device=dev, dtype=dt) for _ in range(3))
        gates = torch.empty((m, 16), device=dev, dtype=dt)
        hid = torch.empty((m, int(d * c.ff_mult)), device=dev, dtype=dt)
        return (x32, xb, ss, q, k, v, gates, hid)

    def _layers_tc(self, st, geom) -> torch.Tensor:
        """All axial transformer pairs (+ BS-RoFormer's final norm) on a prepared state; returns the bf16 shadow."""
        from .. import netops
        for time_transformer, freq_transformer in self.layers:
            self._transformer_tc(st, time_transformer, geom, True)
            self._transformer_tc(st, freq_transformer, geom, False)
        x32, xb, ss = st[0], st[1], st[2]
        if self.cfg.kind != "mel":
            netops.resid_prepare(x32, x32, xb, ss, gamma=self.final_norm.gamma.detach().float())
        return xb

    def _axial_tc(self, x):
        """bf16-operand twin of `_axial` (+ BS-RoFormer's final norm) on the tcgen05 GEMM.  The residual stream is fp32
        (x32) with a bf16 shadow (xb) as the A operand of the next GEMM; every row-wise operator between two
        contractions except the time-axis attention gate lives in a GEMM epilogue."""
        from .. import netops
        b, t, f, d = x.shape
        st = self._tc_state(b * t * f, x.device)
        netops.resid_prepare(x.reshape(b * t * f, d).float().contiguous(), st[0], st[1], st[2])
        return self._layers_tc(st, (b, t, f)).view(b, t, f, d)

    def set_compute_dtype(self, dtype: torch.dtype) -> "RoformerMaskNet":
        """float32 = parity configuration (module path); bfloat16 / float16 = the tcgen05 path with that 16-bit operand
        format (fp32 accumulation and fp32 residual stream either way).  float16 carries 11 significand bits instead of 8
        at the same tensor-core rate: ~18 dB more SI-SDR against the fp32 oracle, with a saturating
        conversion at +-65504 where bfloat16 has fp32's range."""
        self.compute_dtype = dtype
        if dtype in (torch.bfloat16, torch.float16) and self._fused_dtype in (torch.bfloat16, torch.float16):
            self._fused_dtype = dtype
        return self

    def _axial(self, x):
        for time_transformer, freq_transformer in self.layers:
            b, t, f, d = x.shape
            x = time_transformer(x.permute(0, 2, 1, 3).reshape(b * f, t, d))
            x = freq_transformer(x.view(b, f, t, d).permute(0, 2, 1, 3).reshape(b * t, f, d))
            x = x.view(b, t, f, d)
        return x

    @torch.no_grad()
    def mask(self, spec: torch.Tensor) -> torch.Tensor:
        c = self.cfg
        b, t, f, s = spec.shape
        feats = torch.view_as_real(spec).reshape(b, t, f * s * 2)          # 'b t (f s c)', zero-copy
        ac = self.compute_dtype != torch.float32
        on_dev = spec.is_cuda or self._fused_dtype == torch.float32        # (fp32 "fused dtype": the CPU host-logic tests)
        half = self.compute_dtype in (torch.bfloat16, torch.float16)
        if ac and on_dev and half and _GROUPED and self._grouped_supported():
            # everything on the tcgen05 GEMM: band split -> fp32 residual stream -> transformers -> mask estimator (fp32 out)
            nb = len(self.band_split.dim_inputs)
            st = self._tc_state(b * t * nb, spec.device)
            self._band_split_tc(feats.reshape(b * t, -1), st)
            xb = self._layers_tc(st, (b, t, nb))
            m = self._mask_estimate_tc(xb, b, t)                            # fp32 [b, n, t, (f s c)]
            return torch.view_as_complex(m.reshape(b, m.shape[1], t, f, s, 2))
        with torch.autocast("cuda", dtype=self.compute_dtype, enabled=ac):
            if c.kind == "mel":
                rows = torch.view_as_real(spec).reshape(b, t, f * s, 2)
                x = rows[:, :, self.freq_indices].reshape(b, t, -1)
            else:
                x = feats
            x = self.band_split(x)
            if ac and x.is_cuda and half and (self._tc_supported() or self.compute_dtype == torch.bfloat16):
                with torch.autocast("cuda", enabled=False):
                    x = self._axial_tc(x) if self._tc_supported() else self._axial_fused(x)
            else:
                x = self._axial(x)
                if c.kind != "mel":
                    x = self.final_norm(x)
            m = torch.stack([fn(x) for fn in self.mask_estimators], dim=1)  # b n t (f' s c)
        n = m.shape[1]
        m = m.float()
        if c.kind == "mel":
            m = torch.view_as_complex(m.reshape(b, n, t, -1, 2))
            idx = self.freq_indices[None, None, None, :].expand(b, n, t, -1)
            summed = torch.zeros(b, n, t, f * s, dtype=m.dtype, device=m.device).scatter_add_(3, idx, m)
            denom = torch.repeat_interleave(self.num_bands_per_freq, s).clamp(min=1e-8)
            return (summed / denom).view(b, n, t, f, s)
        return torch.view_as_complex(m.reshape(b, n, t, f, s, 2))
