"""Host wrappers of the fused row-wise bf16 operators (csrc/al_netops.cu) used by the RoFormer mask
network's inference path.  CUDA tensors only; every call enqueues on torch's current stream."""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _check_bf16_rows(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError("audiolab_b200 kernels need CUDA tensors (there is no CPU fallback)")
    if t.dtype != torch.bfloat16 or t.dim() != 2 or not t.is_contiguous():
        raise ValueError(f"{name} must be a contiguous bf16 [rows, cols] tensor")


def rmsnorm(x: torch.Tensor, gamma: torch.Tensor, bias: Optional[torch.Tensor] = None,
            out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = F.normalize(x, dim=-1) * sqrt(dim) * gamma; with `bias`, x += bias happens (in place) first."""
    _check_bf16_rows(x, "x")
    n, d = x.shape
    if gamma.dtype != torch.float32 or gamma.numel() != d or (bias is not None and
                                                              (bias.dtype != torch.float32 or bias.numel() != d)):
        raise ValueError("gamma / bias must be fp32 [dim]")
    if out is None:
        out = torch.empty_like(x)
    else:
        _check_bf16_rows(out, "out")
    _lib.check(_lib.lib().al_rmsnorm_bf16(x.data_ptr(), gamma.data_ptr(), None if bias is None else bias.data_ptr(),
                                          out.data_ptr(), n, d, float(d) ** 0.5, 1e-12, _stream()), "al_rmsnorm_bf16")
    return out


def rotary_(q: torch.Tensor, k: torch.Tensor, cos_sin: torch.Tensor, heads: int, dim_head: int, pos_div: int,
            pos_mod: int) -> None:
    """Rotate q and k [rows, heads*dim_head] in place; row position = (row // pos_div) % pos_mod."""
    _check_bf16_rows(q, "q")
    _check_bf16_rows(k, "k")
    if q.shape != k.shape or q.shape[1] != heads * dim_head:
        raise ValueError("q, k must be [rows, heads*dim_head]")
    if cos_sin.dtype != torch.float32 or tuple(cos_sin.shape) != (pos_mod, dim_head // 2, 2) or not cos_sin.is_contiguous():
        raise ValueError("cos_sin must be contiguous fp32 [pos_mod, dim_head/2, 2]")
    _lib.check(_lib.lib().al_rotary_bf16(q.data_ptr(), k.data_ptr(), cos_sin.data_ptr(), q.shape[0], heads, dim_head,
                                         int(pos_div), int(pos_mod), _stream()), "al_rotary_bf16")


def gate_sigmoid_(o: torch.Tensor, gates: torch.Tensor, heads: int, dim_head: int) -> None:
    """o[row, h, :] *= sigmoid(gates[row, h]) in place."""
    if o.shape[1] != heads * dim_head or tuple(gates.shape) != (o.shape[0], heads):
        raise ValueError("o must be [rows, heads*dim_head] and gates [rows, heads]")
    if (gates.is_cuda and o.is_cuda and o.dtype in HALF_DTYPES and o.dim() == 2 and o.is_contiguous() and gates.stride(1) == 1
            and (gates.stride(0) != heads or o.dtype == torch.float16)):
        # the gate columns of the fused to_qkv + to_gates GEMM output (a strided view), bf16 or fp16
        _lib.check(_lib.lib().al_gate_sigmoid_ld_bf16(o.data_ptr(), gates.data_ptr(), gates.stride(0), o.shape[0], heads,
                                                      dim_head, _half_kind(o, gates), _stream()), "al_gate_sigmoid_ld_bf16")
        return
    _check_bf16_rows(o, "o")
    _check_bf16_rows(gates, "gates")
    _lib.check(_lib.lib().al_gate_sigmoid_bf16(o.data_ptr(), gates.data_ptr(), o.shape[0], heads, dim_head, _stream()),
               "al_gate_sigmoid_bf16")


def gelu_(x: torch.Tensor) -> torch.Tensor:
    """Exact (erf) GELU of a contiguous bf16 tensor, in place."""
    if not x.is_cuda:
        raise RuntimeError("audiolab_b200 kernels need CUDA tensors (there is no CPU fallback)")
    if x.dtype != torch.bfloat16 or not x.is_contiguous() or x.numel() % 8 != 0:
        raise ValueError("x must be a contiguous bf16 tensor with a multiple of 8 elements")
    _lib.check(_lib.lib().al_gelu_bf16(x.data_ptr(), x.numel(), _stream()), "al_gelu_bf16")
    return x


def band_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, n_seq: int, seq_len: int, heads: int,
                   dim_head: int, gates: Optional[torch.Tensor] = None,
                   cos_sin: Optional[torch.Tensor] = None) -> torch.Tensor:
    """softmax(q k^T / sqrt(dim_head)) v per (sequence, head) on token-major [n_seq * seq_len, heads * dim_head] buffers
    (al_attn.cu; seq_len <= 64, dim_head == 64).  With `gates` [n_seq * seq_len, heads] the output is also multiplied by
    sigmoid(gates) per (token, head); with `cos_sin` [seq_len, dim_head/2, 2] fp32, q and k are rotated by their position in
    the sequence first (rotary_ semantics, without modifying q / k).  Returns a new [n_seq * seq_len, heads * dim_head] tensor."""
    for t, name in ((q, "q"), (k, "k"), (v, "v")):
        if not t.is_cuda:
            raise RuntimeError("audiolab_b200 kernels need CUDA tensors (there is no CPU fallback)")
        if t.dtype not in HALF_DTYPES or t.dim() != 2 or not t.is_contiguous():
            raise ValueError(f"{name} must be a contiguous bf16 / fp16 [rows, cols] tensor")
    fp16 = _half_kind(q, k, v)
    if q.shape != k.shape or q.shape != v.shape or tuple(q.shape) != (n_seq * seq_len, heads * dim_head):
        raise ValueError("q, k, v must be [n_seq * seq_len, heads * dim_head]")
    gate_ld = 0
    if gates is not None:
        if (not gates.is_cuda or gates.dtype != q.dtype or tuple(gates.shape) != (q.shape[0], heads)
                or gates.stride(1) != 1):
            raise ValueError("gates must be a CUDA [n_seq * seq_len, heads] tensor of q's dtype with unit column stride")
        gate_ld = gates.stride(0)
    if cos_sin is not None and (cos_sin.dtype != torch.float32 or tuple(cos_sin.shape) != (seq_len, dim_head // 2, 2)
                                or not cos_sin.is_contiguous() or not cos_sin.is_cuda):
        raise ValueError("cos_sin must be a contiguous CUDA fp32 [seq_len, dim_head/2, 2] tensor")
    o = torch.empty_like(q)
    _lib.check(_lib.lib().al_band_attention_bf16(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(),
                                                 None if gates is None else gates.data_ptr(), int(gate_ld),
                                                 None if cos_sin is None else cos_sin.data_ptr(), int(n_seq), int(seq_len),
                                                 int(heads), int(dim_head), float(dim_head) ** -0.5, fp16, _stream()),
               "al_band_attention_bf16")
    return o


def time_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, n_batch: int, seq_len: int, inner: int, heads: int,
                   dim_head: int, gates: Optional[torch.Tensor] = None) -> torch.Tensor:
    """softmax(q k^T / sqrt(dim_head)) v per (batch, inner index, head) along the seq_len axis of token-major
    [n_batch * seq_len * inner, heads * dim_head] buffers -- token (b, t, i) in row (b * seq_len + t) * inner + i -- on the
    tcgen05 tensor cores (al_fattn.cu; dim_head == 64, any seq_len).  With `gates` [rows, heads] the output is also multiplied
    by sigmoid(gates) per (token, head).  Returns a new tensor of q's shape."""
    for t, name in ((q, "q"), (k, "k"), (v, "v")):
        if not t.is_cuda:
            raise RuntimeError("audiolab_b200 kernels need CUDA tensors (there is no CPU fallback)")
        if t.dtype not in HALF_DTYPES or t.dim() != 2 or not t.is_contiguous():
            raise ValueError(f"{name} must be a contiguous bf16 / fp16 [rows, cols] tensor")
    fp16 = _half_kind(q, k, v)
    if q.shape != k.shape or q.shape != v.shape or tuple(q.shape) != (n_batch * seq_len * inner, heads * dim_head):
        raise ValueError("q, k, v must be [n_batch * seq_len * inner, heads * dim_head]")
    gate_ld = 0
    if gates is not None:
        if (not gates.is_cuda or gates.dtype != q.dtype or tuple(gates.shape) != (q.shape[0], heads)
                or gates.stride(1) != 1):
            raise ValueError("gates must be a CUDA [rows, heads] tensor of q's dtype with unit column stride")
        gate_ld = gates.stride(0)
    o = torch.empty_like(q)
    _lib.check(_lib.lib().al_time_attention_bf16(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(),
                                                 None if gates is None else gates.data_ptr(), int(gate_ld), int(n_batch),
                                                 int(seq_len), int(inner), int(heads), int(dim_head), float(dim_head) ** -0.5,
                                                 fp16, _stream()),
               "al_time_attention_bf16")
    return o


# ---- K4: tcgen05 GEMM with fused epilogues (csrc/al_gemm.cu) ---------------------------------------------------
import ctypes as _C


class GemmArgs(_C.Structure):
    """ctypes mirror of `al_gemm_args` (include/audiolab_b200.h)."""
    _fields_ = [
        ("A", _C.c_void_p), ("W", _C.c_void_p), ("M", _C.c_int64),
        ("N", _C.c_int32), ("K", _C.c_int32), ("groups", _C.c_int32),
        ("lda", _C.c_int64), ("a_group_stride", _C.c_int64), ("ldw", _C.c_int64), ("w_group_stride", _C.c_int64),
        ("epi", _C.c_int32), ("act", _C.c_int32),
        ("bias", _C.c_void_p), ("row_ss", _C.c_void_p), ("ss_parts", _C.c_int32),
        ("ss_scale", _C.c_float), ("ss_eps", _C.c_float),
        ("cos_sin", _C.c_void_p), ("pos_div", _C.c_int64), ("pos_mod", _C.c_int32), ("rot_cols", _C.c_int32),
        ("out", _C.c_void_p * 4), ("ldo", _C.c_int64 * 4), ("o_group_stride", _C.c_int64 * 4),
        ("out_split", _C.c_int32),
        ("x32", _C.c_void_p), ("xb", _C.c_void_p),
        ("ldx", _C.c_int64), ("x_group_stride", _C.c_int64), ("ldxb", _C.c_int64), ("xb_group_stride", _C.c_int64),
        ("ss_out", _C.c_void_p), ("max_ctas", _C.c_int32), ("no_accumulate", _C.c_int32),
        ("side_row_stride", _C.c_int64), ("side_group_stride", _C.c_int64), ("operand_fp16", _C.c_int32),
    ]


ACT = {None: 0, "none": 0, "gelu": 1, "tanh": 2}

# bench.py's roofline leg: when set to a list, every al_gemm_bf16 launch is bracketed by CUDA events on its stream and
# appended as (epilogue name, flops, algorithmic bytes, start event, end event).  None (default) = no events.
gemm_timer = None


def _timed_gemm(kind: str, flops: float, nbytes: float, args, what: str) -> None:
    if gemm_timer is None:
        _lib.check(_lib.lib().al_gemm_bf16(_C.byref(args), _stream()), what)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check(_lib.lib().al_gemm_bf16(_C.byref(args), _stream()), what)
    e1.record()
    gemm_timer.append((kind, flops, nbytes, e0, e1))
HALF_DTYPES = (torch.bfloat16, torch.float16)     # 16-bit operand formats of the tcgen05 path (same tensor-core rate)


def _half_kind(*tensors) -> int:
    """0 = bfloat16, 1 = IEEE half; every 16-bit tensor of one call must agree."""
    kinds = {t.dtype for t in tensors}
    if len(kinds) != 1 or next(iter(kinds)) not in HALF_DTYPES:
        raise ValueError(f"the 16-bit tensors of one call must all be bfloat16 or all be float16, got {sorted(map(str, kinds))}")
    return 1 if next(iter(kinds)) == torch.float16 else 0


def _rows2d(t: torch.Tensor, name: str, dtype) -> None:
    if not t.is_cuda:
        raise RuntimeError("audiolab_b200 kernels need CUDA tensors (there is no CPU fallback)")
    ok = t.dtype in dtype if isinstance(dtype, tuple) else t.dtype == dtype
    if not ok or t.dim() not in (2, 3) or t.stride(-1) != 1:
        raise ValueError(f"{name} must be a {dtype} [rows, cols] or [groups, rows, cols] tensor with unit column stride")


def _geom(t: torch.Tensor):
    """(groups, rows, cols, row stride, group stride) of a 2-D / 3-D operand."""
    if t.dim() == 2:
        return 1, t.shape[0], t.shape[1], t.stride(0), t.stride(0) * t.shape[0]
    return t.shape[0], t.shape[1], t.shape[2], t.stride(1), t.stride(0)


def gemm_bf16(a: torch.Tensor, w: torch.Tensor, outs, *, bias: Optional[torch.Tensor] = None,
              row_ss: Optional[torch.Tensor] = None, ss_scale: float = 1.0, ss_eps: float = 1e-12,
              cos_sin: Optional[torch.Tensor] = None, pos_div: int = 1, pos_mod: int = 1, rot_cols: int = 0,
              act: Optional[str] = None, out_split: int = 0, max_ctas: int = 0) -> None:
    """outs[i][..., m, :] = epilogue(a @ w^T) columns [i * out_split, (i+1) * out_split)  (al_gemm_bf16, EPI_BF16).

    a [M, K] / [G, M, K] bf16, w [N, K] / [G, N, K] bf16 (nn.Linear layout), outs: one bf16 tensor or a list of up to 4.
    row_ss [M(, G), parts] fp32 turns on the folded RMSNorm row scale ss_scale / max(sqrt(sum row_ss), ss_eps)."""
    if isinstance(outs, torch.Tensor):
        outs = [outs]
    _rows2d(a, "a", HALF_DTYPES)
    _rows2d(w, "w", HALF_DTYPES)
    g, m, k, lda, ags = _geom(a)
    gw, n, kw, ldw, wgs = _geom(w)
    if gw != g or kw != k:
        raise ValueError("a and w disagree on groups / K")
    args = GemmArgs()
    args.operand_fp16 = _half_kind(a, w, *outs)
    args.A, args.W, args.M, args.N, args.K, args.groups = a.data_ptr(), w.data_ptr(), m, n, k, g
    args.lda, args.a_group_stride, args.ldw, args.w_group_stride = lda, ags, ldw, wgs
    args.epi, args.act = 0, ACT[act]
    if bias is not None:
        if bias.dtype != torch.float32 or bias.numel() != g * n or not bias.is_contiguous():
            raise ValueError("bias must be contiguous fp32 [groups, N]")
        args.bias = bias.data_ptr()
    if row_ss is not None:
        if row_ss.dtype != torch.float32 or not row_ss.is_contiguous() or row_ss.numel() % (g * m) != 0:
            raise ValueError("row_ss must be contiguous fp32 [groups * M, parts]")
        args.row_ss, args.ss_parts = row_ss.data_ptr(), row_ss.numel() // (g * m)
        args.ss_scale, args.ss_eps = float(ss_scale), float(ss_eps)
    if cos_sin is not None:
        if cos_sin.dtype != torch.float32 or tuple(cos_sin.shape) != (pos_mod, 32, 2) or not cos_sin.is_contiguous():
            raise ValueError("cos_sin must be contiguous fp32 [pos_mod, 32, 2] (dim_head 64)")
        args.cos_sin, args.pos_div, args.pos_mod, args.rot_cols = cos_sin.data_ptr(), int(pos_div), int(pos_mod), int(rot_cols)
    if len(outs) > 4:
        raise ValueError("at most 4 outputs")
    for i, o in enumerate(outs):
        _rows2d(o, f"outs[{i}]", HALF_DTYPES)
        go, mo, _, ldo, ogs = _geom(o)
        if go != g or mo != m:
            raise ValueError("outputs disagree with a on groups / M")
        args.out[i], args.ldo[i], args.o_group_stride[i] = o.data_ptr(), ldo, ogs
    args.out_split = int(out_split)
    args.max_ctas = int(max_ctas)
    _timed_gemm("bf16" if act is None else act, 2.0 * g * m * n * k, 2.0 * g * m * (k + n) + 2.0 * g * n * k, args, "al_gemm_bf16")


def resid_slab(n: int) -> int:
    """Columns per partial sum of squares of the residual epilogue (half of its N tile: one epilogue warp's share)."""
    return 128 if n % 256 == 0 else 64


def gemm_bf16_glu(a: torch.Tensor, w: torch.Tensor, out: torch.Tensor, *, bias: Optional[torch.Tensor] = None,
                  max_ctas: int = 0) -> None:
    """out[..., m, i] = (a @ w^T + bias)[2i] * sigmoid((a @ w^T + bias)[2i + 1]) in fp32 (al_gemm_bf16, EPI_GLU): w's rows
    (and bias) are ALREADY interleaved (a_0, b_0, a_1, b_1, ...), see `interleave_glu`.  a [M, K] / [G, M, K] bf16,
    w [N, K] / [G, N, K] bf16, out fp32 [M, N / 2] / [G, M, N / 2] (may be a strided view into the mask tensor)."""
    _rows2d(a, "a", HALF_DTYPES)
    _rows2d(w, "w", HALF_DTYPES)
    _rows2d(out, "out", torch.float32)
    g, m, k, lda, ags = _geom(a)
    gw, n, kw, ldw, wgs = _geom(w)
    go, mo, no, ldo, ogs = _geom(out)
    if gw != g or kw != k or (go, mo) != (g, m) or n % 16 != 0 or not (n // 2 - 8 < no <= n // 2):
        raise ValueError("a / w / out disagree on groups, M, K, or N is not the output width rounded up to a multiple of 16")
    args = GemmArgs()
    args.A, args.W, args.M, args.N, args.K, args.groups = a.data_ptr(), w.data_ptr(), m, n, k, g
    args.lda, args.a_group_stride, args.ldw, args.w_group_stride = lda, ags, ldw, wgs
    args.epi = 2
    args.operand_fp16 = _half_kind(a, w)
    if bias is not None:
        if bias.dtype != torch.float32 or bias.numel() != g * n or not bias.is_contiguous():
            raise ValueError("bias must be contiguous fp32 [groups, N] (interleaved like w's rows)")
        args.bias = bias.data_ptr()
    args.out[0], args.ldo[0], args.o_group_stride[0] = out.data_ptr(), ldo, ogs
    args.out_split = no                 # valid output columns (w may carry zero rows up to a multiple of 16)
    args.max_ctas = int(max_ctas)
    _timed_gemm("glu", 2.0 * g * m * n * k, 2.0 * g * m * k + 4.0 * g * m * no + 2.0 * g * n * k, args, "al_gemm_bf16(glu)")


def interleave_glu(t: torch.Tensor) -> torch.Tensor:
    """Rows (dim -2 of a weight, or the last dim of a bias) [a_0 .. a_{h-1}, b_0 .. b_{h-1}] -> [a_0, b_0, a_1, b_1, ...]:
    nn.GLU's two halves side by side, as the GLU epilogue of the GEMM wants them."""
    if t.dim() == 1:
        h = t.shape[0] // 2
        return torch.stack((t[:h], t[h:]), dim=1).reshape(-1).contiguous()
    h = t.shape[-2] // 2
    return torch.stack((t[..., :h, :], t[..., h:, :]), dim=-2).reshape(*t.shape[:-2], 2 * h, t.shape[-1]).contiguous()


def band_norm(x: torch.Tensor, gamma: torch.Tensor, band_off: torch.Tensor, out: torch.Tensor, eps: float = 1e-12) -> None:
    """Per-band RMSNorm (upstream BandSplit): out[:, off_j:off_{j+1}] = bf16(normalize(x[:, off_j:off_{j+1}]) * sqrt(d_j) *
    gamma[off_j:off_{j+1}]).  x fp32 [rows, >= off_last] (row stride free), out bf16 likewise, band_off int32 device."""
    if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 2 or x.stride(1) != 1:
        raise ValueError("x must be a CUDA fp32 [rows, cols] tensor with unit column stride")
    if out.dtype not in HALF_DTYPES or out.dim() != 2 or out.stride(1) != 1 or out.shape[0] != x.shape[0]:
        raise ValueError("out must be bf16 / fp16 [rows, cols] with unit column stride")
    if band_off.dtype != torch.int32 or not band_off.is_contiguous() or gamma.dtype != torch.float32 or not gamma.is_contiguous():
        raise ValueError("band_off must be contiguous int32, gamma contiguous fp32")
    _lib.check(_lib.lib().al_band_norm(x.data_ptr(), x.stride(0), gamma.data_ptr(), band_off.data_ptr(), band_off.numel() - 1,
                                       out.data_ptr(), out.stride(0), x.shape[0], float(eps), _half_kind(out), _stream()),
               "al_band_norm")


def gemm_bf16_residual(a: torch.Tensor, w: torch.Tensor, x32: torch.Tensor, xb: torch.Tensor, ss_out: torch.Tensor, *,
                       bias: Optional[torch.Tensor] = None, max_ctas: int = 0, accumulate: bool = True) -> None:
    """x32 += a @ w^T + bias (fp32, in place); xb = bf16(x32); ss_out[m, j] = sum of x32[m, S j : S (j+1)]^2 with
    S = resid_slab(N)."""
    _rows2d(a, "a", HALF_DTYPES)
    _rows2d(w, "w", HALF_DTYPES)
    _rows2d(x32, "x32", torch.float32)
    _rows2d(xb, "xb", HALF_DTYPES)
    g, m, k, lda, ags = _geom(a)
    gw, n, kw, ldw, wgs = _geom(w)
    if gw != g or kw != k or n % 128 != 0:
        raise ValueError("a and w disagree on groups / K, or N is not a multiple of 128")
    gx, mx, nx, ldx, xgs = _geom(x32)
    gb, mb, nb, ldxb, xbgs = _geom(xb)
    if (gx, mx, nx) != (g, m, n) or (gb, mb, nb) != (g, m, n):
        raise ValueError("x32 / xb must be [groups, M, N]")
    parts = n // resid_slab(n)
    side_rs = side_gs = 0
    if ss_out.dtype != torch.float32:
        raise ValueError("ss_out must be fp32")
    if ss_out.dim() == 3:
        # [groups, M, parts] view into a token-ordered array (band-grouped calls): strides in units of `parts`
        if tuple(ss_out.shape) != (g, m, parts) or ss_out.stride(2) != 1 or ss_out.stride(0) % parts or ss_out.stride(1) % parts:
            raise ValueError("ss_out view must be [groups, M, parts] with strides that are multiples of parts")
        side_gs, side_rs = ss_out.stride(0) // parts, ss_out.stride(1) // parts
    elif not ss_out.is_contiguous() or ss_out.numel() != g * m * parts:
        raise ValueError("ss_out must be contiguous fp32 [groups * M, N / resid_slab(N)]")
    args = GemmArgs()
    args.A, args.W, args.M, args.N, args.K, args.groups = a.data_ptr(), w.data_ptr(), m, n, k, g
    args.lda, args.a_group_stride, args.ldw, args.w_group_stride = lda, ags, ldw, wgs
    args.epi = 1
    args.operand_fp16 = _half_kind(a, w, xb)
    if bias is not None:
        if bias.dtype != torch.float32 or bias.numel() != g * n or not bias.is_contiguous():
            raise ValueError("bias must be contiguous fp32 [groups, N]")
        args.bias = bias.data_ptr()
    args.x32, args.xb, args.ldx, args.x_group_stride, args.ldxb, args.xb_group_stride = (
        x32.data_ptr(), xb.data_ptr(), ldx, xgs, ldxb, xbgs)
    args.ss_out = ss_out.data_ptr()
    args.max_ctas = int(max_ctas)
    args.no_accumulate = 0 if accumulate else 1
    args.side_row_stride, args.side_group_stride = side_rs, side_gs
    # A read + x32 read (if accumulating) + x32 write + 16-bit shadow write + weights
    nbytes = 2.0 * g * m * k + (8.0 if accumulate else 4.0) * g * m * n + 2.0 * g * m * n + 2.0 * g * n * k
    _timed_gemm("residual", 2.0 * g * m * n * k, nbytes, args, "al_gemm_bf16(residual)")


def resid_prepare(x_in: torch.Tensor, x32: torch.Tensor, xb: torch.Tensor, ss: torch.Tensor, *,
                  bias: Optional[torch.Tensor] = None, gamma: Optional[torch.Tensor] = None, eps: float = 1e-12) -> None:
    """x32 = y, xb = bf16(y), ss = partial sums of y^2 with y = [RMSNorm_gamma](x_in + bias)  (al_resid_prepare)."""
    n, d = x_in.shape
    for t, name in ((x_in, "x_in"), (x32, "x32")):
        if not t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or tuple(t.shape) != (n, d):
            raise ValueError(f"{name} must be a contiguous CUDA fp32 [rows, dim] tensor")
    if not xb.is_cuda or xb.dtype not in HALF_DTYPES or not xb.is_contiguous() or tuple(xb.shape) != (n, d):
        raise ValueError("xb must be a contiguous CUDA bf16 / fp16 [rows, dim] tensor")
    if ss.dtype != torch.float32 or not ss.is_contiguous() or ss.numel() % n != 0:
        raise ValueError("ss must be contiguous fp32 [rows, parts]")
    for v, name in ((bias, "bias"), (gamma, "gamma")):
        if v is not None and (v.dtype != torch.float32 or v.numel() != d or not v.is_contiguous()):
            raise ValueError(f"{name} must be contiguous fp32 [dim]")
    _lib.check(_lib.lib().al_resid_prepare(x_in.data_ptr(), None if bias is None else bias.data_ptr(),
                                           None if gamma is None else gamma.data_ptr(), x32.data_ptr(), xb.data_ptr(),
                                           ss.data_ptr(), n, d, ss.numel() // n, float(eps), _half_kind(xb), _stream()),
               "al_resid_prepare")
