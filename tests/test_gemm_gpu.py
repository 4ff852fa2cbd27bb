"""K4 parity: the tcgen05 GEMM (csrc/al_gemm.cu) through the C ABI against an fp32 torch reference of the same
operator, computed from the same bf16-rounded operands.  Tolerances: the accumulation is fp32 on both sides, so
the only differences are summation order (~1e-6 relative) and the final bf16 rounding of the stored result
(half an ulp = 2^-9 relative), hence rtol 2^-8 on bf16 outputs and 2e-5 on the fp32 residual stream."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from audiolab_b200 import netops
    return netops


def _rand(shape, seed, scale=1.0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).cuda()


def _close_bf16(got, ref, what):
    got = got.float()
    err = (got - ref).abs()
    tol = 2.0 ** -8 * ref.abs() + 2e-5 * ref.abs().max().clamp(min=1e-6)
    bad = (err > tol)
    assert not bad.any(), f"{what}: {int(bad.sum())} of {bad.numel()} outside bf16 rounding, max err {float(err.max()):.3e}"


@pytest.mark.parametrize("m,n,k", [(128, 256, 64), (256, 512, 512), (1000, 1536, 512), (4097, 2048, 512),
                                   (777, 512, 2048), (300, 128, 512), (300, 64, 192), (129, 1552, 512)])
def test_gemm_plain(m, n, k):
    netops = _cuda()
    a = _rand((m, k), 1).bfloat16()
    w = _rand((n, k), 2, k ** -0.5).bfloat16()
    out = torch.full((m, n), float("nan"), device="cuda", dtype=torch.bfloat16)
    netops.gemm_bf16(a, w, out)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t()
    _close_bf16(out, ref, f"plain {m}x{n}x{k}")


def test_gemm_many_tiles_persistent():
    """More tiles than CTAs, restricted grid: every CTA walks several tiles through both accumulators."""
    netops = _cuda()
    m, n, k = 128 * 37 + 5, 1024, 512
    a = _rand((m, k), 3).bfloat16()
    w = _rand((n, k), 4, k ** -0.5).bfloat16()
    out = torch.empty((m, n), device="cuda", dtype=torch.bfloat16)
    netops.gemm_bf16(a, w, out, max_ctas=7)
    torch.cuda.synchronize()
    _close_bf16(out, a.float() @ w.float().t(), "persistent")


@pytest.mark.parametrize("act", [None, "gelu", "tanh"])
def test_gemm_rowscale_bias_act(act):
    netops = _cuda()
    m, n, k = 1111, 2048, 512
    x = _rand((m, k), 5, 3.0)
    a = x.bfloat16()
    w = _rand((n, k), 6, k ** -0.5).bfloat16()
    bias = _rand((n,), 7, 0.5)
    ss = torch.stack((x[:, :256].square().sum(-1), x[:, 256:].square().sum(-1)), dim=-1).contiguous()
    out = torch.empty((m, n), device="cuda", dtype=torch.bfloat16)
    netops.gemm_bf16(a, w, out, bias=bias, row_ss=ss, ss_scale=math.sqrt(k), act=act)
    torch.cuda.synchronize()
    rs = math.sqrt(k) / x.square().sum(-1).sqrt().clamp(min=1e-12)
    ref = (a.float() @ w.float().t()) * rs[:, None] + bias
    if act == "gelu":
        ref = torch.nn.functional.gelu(ref)
    elif act == "tanh":
        ref = torch.tanh(ref)
    _close_bf16(out, ref, f"rowscale+bias+{act}")


@pytest.mark.parametrize("time_axis", [True, False])
def test_gemm_qkv_rotary_split(time_axis):
    """to_qkv + to_gates in one call: q, k rotated by the token position, v plain, gates with bias; 4 outputs."""
    netops = _cuda()
    b, t, f, heads, dh, d = 2, 37, 11, 8, 64, 512
    m, inner = b * t * f, heads * dh
    x = _rand((m, d), 8)
    a = x.bfloat16()
    w = _rand((3 * inner + 16, d), 9, d ** -0.5).bfloat16()
    w[3 * inner + heads:] = 0
    bias = torch.zeros(3 * inner + 16, device="cuda")
    bias[3 * inner: 3 * inner + heads] = _rand((heads,), 10)
    n_pos = t if time_axis else f
    pos_div = f if time_axis else 1
    freqs = 1.0 / (10000.0 ** (torch.arange(0, dh, 2).float() / dh))
    ang = torch.arange(n_pos).float()[:, None] * freqs[None, :]
    cos_sin = torch.stack((ang.cos(), ang.sin()), dim=-1).contiguous().cuda()
    q, k_, v = (torch.empty((m, inner), device="cuda", dtype=torch.bfloat16) for _ in range(3))
    gates = torch.empty((m, 16), device="cuda", dtype=torch.bfloat16)
    netops.gemm_bf16(a, w, [q, k_, v, gates], bias=bias, cos_sin=cos_sin, pos_div=pos_div, pos_mod=n_pos,
                     rot_cols=2 * inner, out_split=inner)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t() + bias
    pos = (torch.arange(m, device="cuda") // pos_div) % n_pos
    c = cos_sin[pos, :, 0][:, None, :]
    s = cos_sin[pos, :, 1][:, None, :]

    def rot(z):
        z = z.view(m, heads, dh // 2, 2)
        return torch.stack((z[..., 0] * c - z[..., 1] * s, z[..., 1] * c + z[..., 0] * s), dim=-1).reshape(m, inner)

    _close_bf16(q, rot(ref[:, :inner]), "q")
    _close_bf16(k_, rot(ref[:, inner:2 * inner]), "k")
    _close_bf16(v, ref[:, 2 * inner:3 * inner], "v")
    _close_bf16(gates[:, :heads], ref[:, 3 * inner:3 * inner + heads], "gates")


def test_gemm_strided_views_and_groups():
    """Grouped call on strided operands: the per-band Linear layers (rows f::F of the token grid)."""
    netops = _cuda()
    bt, f, d, n = 333, 5, 512, 256
    x = _rand((bt, f, d), 11).bfloat16()
    w = _rand((f, n, d), 12, d ** -0.5).bfloat16()
    bias = _rand((f, n), 13)
    out = torch.empty((bt, f, n), device="cuda", dtype=torch.bfloat16)
    netops.gemm_bf16(x.transpose(0, 1), w, out.transpose(0, 1), bias=bias, act="tanh")
    torch.cuda.synchronize()
    ref = torch.tanh(torch.einsum("mfk,fnk->mfn", x.float(), w.float()) + bias[None])
    _close_bf16(out, ref, "grouped")


@pytest.mark.parametrize("m,k,with_bias,n", [(128, 512, False, 512), (1000, 512, True, 512), (5000, 2048, True, 512),
                                               (700, 384, True, 384)])
def test_gemm_residual_epilogue(m, k, with_bias, n):
    netops = _cuda()
    a = _rand((m, k), 14).bfloat16()
    w = _rand((n, k), 15, k ** -0.5).bfloat16()
    bias = _rand((n,), 16) if with_bias else None
    x0 = _rand((m, n), 17, 2.0)
    x32 = x0.clone()
    xb = torch.full((m, n), float("nan"), device="cuda", dtype=torch.bfloat16)
    parts = n // netops.resid_slab(n)
    ss = torch.full((m, parts), float("nan"), device="cuda")
    netops.gemm_bf16_residual(a, w, x32, xb, ss, bias=bias, max_ctas=5 if m > 1000 else 0)
    torch.cuda.synchronize()
    ref = x0 + a.float() @ w.float().t() + (bias if with_bias else 0.0)
    assert torch.allclose(x32, ref, rtol=2e-5, atol=2e-5 * float(ref.abs().max())), float((x32 - ref).abs().max())
    assert torch.equal(xb, x32.bfloat16()), "xb must be the bf16 rounding of the stored fp32 row"
    ss_ref = x32.view(m, parts, -1).square().sum(-1)
    assert torch.allclose(ss, ss_ref, rtol=1e-5), float((ss - ss_ref).abs().max())


def test_resid_prepare():
    netops = _cuda()
    m, d = 999, 512
    x = _rand((m, d), 18, 2.0)
    bias = _rand((d,), 19)
    gamma = _rand((d,), 20).abs() + 0.5
    for use_gamma in (False, True):
        x32 = torch.empty_like(x)
        xb = torch.empty((m, d), device="cuda", dtype=torch.bfloat16)
        ss = torch.empty((m, 2), device="cuda")
        netops.resid_prepare(x, x32, xb, ss, bias=bias, gamma=gamma if use_gamma else None)
        torch.cuda.synchronize()
        y = x + bias
        if use_gamma:
            y = torch.nn.functional.normalize(y, dim=-1) * math.sqrt(d) * gamma
        assert torch.allclose(x32, y, rtol=1e-5, atol=1e-5)
        assert torch.equal(xb, x32.bfloat16())
        ss_ref = torch.stack((x32[:, :256].square().sum(-1), x32[:, 256:].square().sum(-1)), dim=-1)
        assert torch.allclose(ss, ss_ref, rtol=1e-5)


@pytest.mark.parametrize("n_out,groups", [(8, 3), (48, 2), (96, 1), (516, 1), (512, 1)])
def test_gemm_glu_epilogue(n_out, groups):
    """Second Linear + GLU of the mask estimator: interleaved (a, b) rows, fp32 output written into a strided view."""
    netops = _cuda()
    m, k = 777, 2048
    n_pad = (2 * n_out + 15) // 16 * 16
    a = _rand((groups, m, k), 30).bfloat16()
    w_full = _rand((groups, 2 * n_out, k), 31, k ** -0.5)
    bias_full = _rand((groups, 2 * n_out), 32)
    w = torch.zeros((groups, n_pad, k), device="cuda")
    b = torch.zeros((groups, n_pad), device="cuda")
    for g in range(groups):
        w[g, : 2 * n_out] = netops.interleave_glu(w_full[g])
        b[g, : 2 * n_out] = netops.interleave_glu(bias_full[g])
    w = w.bfloat16()
    total = groups * n_out + 12
    buf = torch.full((m, total), float("nan"), device="cuda")
    out = buf.as_strided((groups, m, n_out), (n_out if groups > 1 else 0, total, 1), 4)
    netops.gemm_bf16_glu(a, w, out, bias=b)
    torch.cuda.synchronize()
    y = torch.einsum("gmk,gnk->gmn", a.float(), w_full.bfloat16().float()) + bias_full[:, None, :]
    ref = y[..., :n_out] * torch.sigmoid(y[..., n_out:])
    got = buf[:, 4: 4 + groups * n_out].reshape(m, groups, n_out).transpose(0, 1)
    assert torch.allclose(got, ref, rtol=2e-5, atol=2e-5 * float(ref.abs().max())), float((got - ref).abs().max())
    assert torch.isnan(buf[:, :4]).all() and torch.isnan(buf[:, 4 + groups * n_out:]).all(), "wrote outside its columns"


def test_band_norm_and_grouped_stream_start():
    """al_band_norm + the grouped band-split GEMM whose residual epilogue starts the fp32 stream (no read of x32), on the
    strided per-band views of the token-ordered buffers."""
    netops = _cuda()
    bt, dim = 555, 512
    widths = [8, 8, 8, 16, 16, 516]
    offs = [0]
    for d in widths:
        offs.append(offs[-1] + d)
    total = offs[-1]
    x = _rand((bt, total), 40, 2.0)
    gamma = _rand((total,), 41).abs() + 0.5
    ld = (total + 7) // 8 * 8 + 8
    xn = torch.zeros((bt, ld), device="cuda", dtype=torch.bfloat16)
    netops.band_norm(x, gamma, torch.tensor(offs, dtype=torch.int32, device="cuda"), xn)
    torch.cuda.synchronize()
    ref_n = torch.zeros((bt, total), device="cuda")
    for a, b in zip(offs[:-1], offs[1:]):
        ref_n[:, a:b] = torch.nn.functional.normalize(x[:, a:b], dim=-1) * math.sqrt(b - a) * gamma[a:b]
    _close_bf16(xn[:, :total], ref_n, "band_norm")
    assert float(xn[:, total:].abs().max()) == 0.0
    nb = len(widths)
    parts = dim // netops.resid_slab(dim)
    x32 = torch.full((bt, nb, dim), float("nan"), device="cuda")
    xb = torch.empty((bt, nb, dim), device="cuda", dtype=torch.bfloat16)
    ss = torch.empty((bt, nb, parts), device="cuda")
    ref = torch.empty((bt, nb, dim), device="cuda")
    for f0, f1 in ((0, 3), (3, 5), (5, 6)):
        d = widths[f0]
        k = (d + 7) // 8 * 8
        w = torch.zeros((f1 - f0, dim, k), device="cuda")
        w[:, :, :d] = _rand((f1 - f0, dim, d), 42 + f0, d ** -0.5)
        w = w.bfloat16()
        bias = _rand((f1 - f0, dim), 50 + f0)
        a = xn.as_strided((f1 - f0, bt, k), (k if f1 - f0 > 1 else 0, ld, 1), offs[f0])
        netops.gemm_bf16_residual(a, w, x32[:, f0:f1].transpose(0, 1), xb[:, f0:f1].transpose(0, 1),
                                  ss[:, f0:f1].transpose(0, 1), bias=bias, accumulate=False)
        ref[:, f0:f1] = (torch.einsum("gmk,gnk->mgn", a.float(), w.float()) + bias[None])
    torch.cuda.synchronize()
    assert torch.allclose(x32, ref, rtol=2e-5, atol=2e-5 * float(ref.abs().max())), float((x32 - ref).abs().max())
    assert torch.equal(xb, x32.bfloat16())
    assert torch.allclose(ss, x32.view(bt, nb, parts, -1).square().sum(-1), rtol=1e-5)


def test_roformer_mask_all_tc_path_vs_fp32_modules():
    """mask() on the all-tcgen05 path (band split -> transformers -> mask estimator) against the fp32 module path of the same
    network: bf16 operand rounding only (relative L2 error of the mask < 1.5 %)."""
    _cuda()
    from audiolab_b200.configs import RoformerConfig
    from audiolab_b200.nets.roformer import RoformerMaskNet
    torch.manual_seed(4321)
    cfg = RoformerConfig(dim=128, depth=2, heads=4, dim_head=64, chunk_size=441 * 40)
    net = RoformerMaskNet(cfg).cuda().eval()
    assert net._grouped_supported()
    g = torch.Generator(device="cpu").manual_seed(5)
    spec = torch.view_as_complex(torch.randn((3, 41, 1025, 2, 2), generator=g)).cuda()
    ref = net.set_compute_dtype(torch.float32).mask(spec)
    got = net.set_compute_dtype(torch.bfloat16).mask(spec)
    torch.cuda.synchronize()
    assert got.shape == ref.shape
    rel = float((torch.view_as_real(got) - torch.view_as_real(ref)).norm() / torch.view_as_real(ref).norm())
    assert rel < 1.5e-2, rel


# ---- IEEE-half operand mode (al_gemm_args.operand_fp16): same kernels, 11 significand bits -------------------------------
def _close_f16(got, ref, what):
    got = got.float()
    err = (got - ref).abs()
    tol = 2.0 ** -11 * ref.abs() + 2e-5 * ref.abs().max().clamp(min=1e-6)      # half an ulp of binary16 = 2^-12 relative
    bad = err > tol
    assert not bad.any(), f"{what}: {int(bad.sum())} of {bad.numel()} outside fp16 rounding, max err {float(err.max()):.3e}"


def test_fp16_gemm_epilogues():
    netops = _cuda()
    m, n, k = 1111, 2048, 512
    x = _rand((m, k), 60, 3.0)
    a = x.half()
    w = _rand((n, k), 61, k ** -0.5).half()
    bias = _rand((n,), 62, 0.5)
    ss = x.view(m, 4, -1).square().sum(-1).contiguous()
    out = torch.empty((m, n), device="cuda", dtype=torch.float16)
    netops.gemm_bf16(a, w, out, bias=bias, row_ss=ss, ss_scale=math.sqrt(k), act="gelu")
    torch.cuda.synchronize()
    rs = math.sqrt(k) / x.square().sum(-1).sqrt()
    _close_f16(out, torch.nn.functional.gelu((a.float() @ w.float().t()) * rs[:, None] + bias), "fp16 rowscale+bias+gelu")
    with pytest.raises(ValueError):
        netops.gemm_bf16(a, w.bfloat16(), out)                       # mixed 16-bit formats are refused
    # residual epilogue: xb is the half rounding (saturating) of the stored fp32 row
    w2 = _rand((512, n), 63, n ** -0.5).half()
    x0 = _rand((m, 512), 64, 2.0)
    x32, xb, ss2 = x0.clone(), torch.empty((m, 512), device="cuda", dtype=torch.float16), torch.empty((m, 4), device="cuda")
    netops.gemm_bf16_residual(out, w2, x32, xb, ss2, bias=None)
    torch.cuda.synchronize()
    ref = x0 + out.float() @ w2.float().t()
    assert torch.allclose(x32, ref, rtol=2e-5, atol=2e-5 * float(ref.abs().max()))
    assert torch.equal(xb, x32.half())
    # GLU epilogue
    n_out = 48
    w3 = netops.interleave_glu(_rand((2 * n_out, n), 65, n ** -0.5)).half()
    o3 = torch.empty((m, n_out), device="cuda")
    netops.gemm_bf16_glu(out, w3, o3)
    torch.cuda.synchronize()
    y = out.float() @ w3.float().t()
    assert torch.allclose(o3, y[:, 0::2] * torch.sigmoid(y[:, 1::2]), rtol=2e-5, atol=2e-5)
    # saturation instead of inf
    big = torch.full((128, 64), 300.0, device="cuda").half()
    o4 = torch.empty((128, 64), device="cuda", dtype=torch.float16)
    netops.gemm_bf16(big, big[:64], o4)
    torch.cuda.synchronize()
    assert torch.isfinite(o4.float()).all() and float(o4.float().max()) == 65504.0


def test_fp16_rowwise_kernels_and_band_attention():
    netops = _cuda()
    m, d = 999, 512
    x = _rand((m, d), 70, 2.0)
    x32, xb, ss = torch.empty_like(x), torch.empty((m, d), device="cuda", dtype=torch.float16), torch.empty((m, 4), device="cuda")
    gamma = _rand((d,), 71).abs() + 0.5
    netops.resid_prepare(x, x32, xb, ss, gamma=gamma)
    offs = torch.tensor([0, 8, 24, 72, 512], dtype=torch.int32, device="cuda")
    xn = torch.zeros((m, 520), device="cuda", dtype=torch.float16)
    netops.band_norm(x, gamma, offs, xn)
    torch.cuda.synchronize()
    y = torch.nn.functional.normalize(x, dim=-1) * math.sqrt(d) * gamma
    assert torch.allclose(x32, y, rtol=1e-5, atol=1e-5) and torch.equal(xb, x32.half())
    for a, b in zip(offs.tolist()[:-1], offs.tolist()[1:]):
        _close_f16(xn[:, a:b], torch.nn.functional.normalize(x[:, a:b], dim=-1) * math.sqrt(b - a) * gamma[a:b], "band_norm fp16")
    # band attention + gate in half
    n_seq, f, h, dh = 37, 62, 8, 64
    q, k, v = (_rand((n_seq * f, h * dh), 72 + i).half() for i in range(3))
    gates16 = _rand((n_seq * f, 16), 75).half()
    got = netops.band_attention(q, k, v, n_seq, f, h, dh, gates=gates16[:, :h]).float()
    shp = (n_seq, f, h, dh)
    ref = torch.nn.functional.scaled_dot_product_attention(q.view(shp).transpose(1, 2).float(), k.view(shp).transpose(1, 2).float(),
                                                           v.view(shp).transpose(1, 2).float()).transpose(1, 2)
    ref = (ref * torch.sigmoid(gates16[:, :h].float()).view(n_seq, f, h, 1)).reshape(n_seq * f, h * dh)
    assert float((got - ref).abs().max()) <= 4e-3 * float(ref.abs().max())
    o = _rand((n_seq * f, h * dh), 76).half()
    want = (o.float().view(-1, h, dh) * torch.sigmoid(gates16[:, :h].float())[:, :, None]).view(-1, h * dh)
    netops.gate_sigmoid_(o, gates16[:, :h], h, dh)
    torch.cuda.synchronize()
    _close_f16(o, want, "gate fp16")


def test_roformer_mask_fp16_operands_vs_fp32_modules():
    """The all-tcgen05 mask() path with IEEE-half operands: ~8x closer to the fp32 module path than bfloat16 operands."""
    _cuda()
    from audiolab_b200.configs import RoformerConfig
    from audiolab_b200.nets.roformer import RoformerMaskNet
    torch.manual_seed(4321)
    cfg = RoformerConfig(dim=128, depth=2, heads=4, dim_head=64, chunk_size=441 * 40)
    net = RoformerMaskNet(cfg).cuda().eval()
    g = torch.Generator(device="cpu").manual_seed(5)
    spec = torch.view_as_complex(torch.randn((3, 41, 1025, 2, 2), generator=g)).cuda()
    ref = torch.view_as_real(net.set_compute_dtype(torch.float32).mask(spec))
    rel = {}
    for dt in (torch.bfloat16, torch.float16):
        got = torch.view_as_real(net.set_compute_dtype(dt).mask(spec))
        rel[dt] = float((got - ref).norm() / ref.norm())
    print(f"relative L2 error of the mask: bf16 operands {rel[torch.bfloat16]:.2e}, fp16 operands {rel[torch.float16]:.2e}")
    assert rel[torch.float16] < 2.5e-3 and rel[torch.float16] < 0.25 * rel[torch.bfloat16]
