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
    _check_bf16_rows(o, "o")
    _check_bf16_rows(gates, "gates")
    if o.shape[1] != heads * dim_head or tuple(gates.shape) != (o.shape[0], heads):
        raise ValueError("o must be [rows, heads*dim_head] and gates [rows, heads]")
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
        _check_bf16_rows(t, name)
    if q.shape != k.shape or q.shape != v.shape or tuple(q.shape) != (n_seq * seq_len, heads * dim_head):
        raise ValueError("q, k, v must be [n_seq * seq_len, heads * dim_head]")
    if gates is not None:
        _check_bf16_rows(gates, "gates")
        if tuple(gates.shape) != (q.shape[0], heads):
            raise ValueError("gates must be [n_seq * seq_len, heads]")
    if cos_sin is not None and (cos_sin.dtype != torch.float32 or tuple(cos_sin.shape) != (seq_len, dim_head // 2, 2)
                                or not cos_sin.is_contiguous() or not cos_sin.is_cuda):
        raise ValueError("cos_sin must be a contiguous CUDA fp32 [seq_len, dim_head/2, 2] tensor")
    o = torch.empty_like(q)
    _lib.check(_lib.lib().al_band_attention_bf16(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(),
                                                 None if gates is None else gates.data_ptr(),
                                                 None if cos_sin is None else cos_sin.data_ptr(), int(n_seq), int(seq_len),
                                                 int(heads), int(dim_head), float(dim_head) ** -0.5, _stream()),
               "al_band_attention_bf16")
    return o
