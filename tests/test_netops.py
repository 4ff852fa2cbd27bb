"""Fused row-wise bf16 operators of the RoFormer inference path (csrc/al_netops.cu, audiolab_b200/netops.py).

CPU: the host logic of `RoformerMaskNet._axial_fused` (token-major residual stream, strided attention
views, deferred FeedForward bias) against the upstream-shaped module path, with the three kernels replaced
by their PyTorch definitions below.  GPU: each kernel against the same definitions, and the fused bf16
mask against the module path under bf16 autocast.
"""
import pytest
import torch
import torch.nn.functional as F

from audiolab_b200.configs import RoformerConfig
from audiolab_b200.nets.roformer import RoformerMaskNet


# ---- PyTorch definitions of the three operators (upstream RMSNorm / rotary / gated attention) ----
def ref_rmsnorm(x, gamma, bias=None, out=None):
    if bias is not None:
        x.copy_((x.float() + bias).to(x.dtype))
    y = (F.normalize(x.float(), dim=-1) * (x.shape[-1] ** 0.5) * gamma).to(x.dtype)
    if out is not None:
        out.copy_(y)
        return out
    return y


def ref_rotary_(q, k, cs, heads, dh, pos_div, pos_mod):
    n = q.shape[0]
    pos = (torch.arange(n, device=q.device) // pos_div) % pos_mod
    c, s = cs[pos][:, None, :, 0], cs[pos][:, None, :, 1]       # [n, 1, dh/2]
    for t in (q, k):
        v = t.view(n, heads, dh // 2, 2).float()
        a, b = v[..., 0], v[..., 1]
        t.copy_(torch.stack((a * c - b * s, b * c + a * s), dim=-1).view(n, -1).to(t.dtype))


def ref_gate_(o, gates, heads, dh):
    n = o.shape[0]
    o.copy_((o.view(n, heads, dh).float() * torch.sigmoid(gates.float())[:, :, None]).view(n, -1).to(o.dtype))


def ref_gelu_(x):
    x.copy_(F.gelu(x.float()).to(x.dtype))
    return x


def _small_net(kind):
    torch.manual_seed(0)
    if kind == "bs":
        cfg = RoformerConfig(dim=32, depth=2, heads=2, dim_head=16, chunk_size=441 * 12)
    else:
        cfg = RoformerConfig(kind="mel", dim=32, depth=2, heads=2, dim_head=16, chunk_size=441 * 12, num_bands=20)
    net = RoformerMaskNet(cfg).eval()
    with torch.no_grad():
        for p in net.parameters():          # default init leaves gammas at 1 and biases near 0
            p.add_(0.05 * torch.randn_like(p))
    return cfg, net


@pytest.mark.parametrize("kind", ["bs", "mel"])
def test_fused_axial_host_logic_matches_module_path(kind, monkeypatch):
    import audiolab_b200.netops as netops
    cfg, net = _small_net(kind)
    monkeypatch.setattr(netops, "rmsnorm", ref_rmsnorm)
    monkeypatch.setattr(netops, "rotary_", ref_rotary_)
    monkeypatch.setattr(netops, "gate_sigmoid_", ref_gate_)
    monkeypatch.setattr(netops, "gelu_", ref_gelu_)
    net._fused_dtype = torch.float32
    b, t, f = 2, 13, len(net.band_split.dim_inputs)
    x = torch.randn(b, t, f, cfg.dim)
    with torch.no_grad():
        ref = net._axial(x.clone())
        if kind != "mel":
            ref = net.final_norm(ref)
        got = net._axial_fused(x.clone())
    assert float((got - ref).abs().max()) <= 1e-5 * max(1.0, float(ref.abs().max()))


def ref_band_attention(q, k, v, n_seq, seq_len, heads, dh, gates=None):
    shp = (n_seq, seq_len, heads, dh)
    o = F.scaled_dot_product_attention(q.view(shp).transpose(1, 2).float(), k.view(shp).transpose(1, 2).float(),
                                       v.view(shp).transpose(1, 2).float())
    o = o.transpose(1, 2).reshape(q.shape)
    if gates is not None:
        o = (o.view(-1, heads, dh) * torch.sigmoid(gates.float())[:, :, None]).reshape(q.shape)
    return o.to(q.dtype)


def test_fused_axial_host_logic_with_band_attention_kernel(monkeypatch):
    """The opt-in band-axis attention (AUDIOLAB_B200_BAND_ATTN=1, csrc/al_attn.cu): token-major q / k / v of the frequency
    transformer go to the kernel as [b*t sequences x f bands]; with the kernel replaced by its torch definition the fused
    path must still equal the module path."""
    import audiolab_b200.netops as netops
    import audiolab_b200.nets.roformer as rof
    torch.manual_seed(0)
    cfg = RoformerConfig(dim=32, depth=1, heads=2, dim_head=64, chunk_size=441 * 12)
    net = RoformerMaskNet(cfg).eval()
    with torch.no_grad():
        for p in net.parameters():
            p.add_(0.05 * torch.randn_like(p))
    calls = []

    def spy(q, k, v, n_seq, seq_len, heads, dh, gates=None, cos_sin=None):
        assert gates is not None and tuple(gates.shape) == (q.shape[0], heads)
        assert cos_sin is not None and tuple(cos_sin.shape) == (seq_len, dh // 2, 2)
        calls.append((tuple(q.shape), n_seq, seq_len, heads, dh))
        q, k = q.clone(), k.clone()
        ref_rotary_(q, k, cos_sin, heads, dh, 1, seq_len)
        return ref_band_attention(q, k, v, n_seq, seq_len, heads, dh, gates)

    monkeypatch.setattr(netops, "rmsnorm", ref_rmsnorm)
    monkeypatch.setattr(netops, "rotary_", ref_rotary_)
    monkeypatch.setattr(netops, "gate_sigmoid_", ref_gate_)
    monkeypatch.setattr(netops, "gelu_", ref_gelu_)
    monkeypatch.setattr(netops, "band_attention", spy)
    monkeypatch.setattr(rof, "_BAND_ATTN", True)
    net._fused_dtype = torch.float32
    b, t, f = 2, 13, len(net.band_split.dim_inputs)
    x = torch.randn(b, t, f, cfg.dim)
    with torch.no_grad():
        ref = net.final_norm(net._axial(x.clone()))
        got = net._axial_fused(x.clone())
    assert calls and all(c == ((b * t * f, 128), b * t, f, 2, 64) for c in calls)
    assert float((got - ref).abs().max()) <= 1e-5 * max(1.0, float(ref.abs().max()))


def test_netops_refuse_cpu_tensors():
    import audiolab_b200.netops as netops
    x = torch.zeros(4, 64, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        netops.rmsnorm(x, torch.ones(64))
    with pytest.raises(RuntimeError):
        netops.gate_sigmoid_(x, torch.zeros(4, 4, dtype=torch.bfloat16), 4, 16)


# ---- GPU: kernels vs definitions -----------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("n,d,with_bias", [(1000, 512, False), (777, 512, True), (65, 1024, True), (33, 2048, False),
                                           (50, 136, True)])
def test_rmsnorm_kernel(cuda, n, d, with_bias):
    import audiolab_b200.netops as netops
    g = torch.Generator().manual_seed(n + d)
    x = (torch.randn(n, d, generator=g) * 3).to(torch.bfloat16).to(cuda)
    gamma = (1 + 0.1 * torch.randn(d, generator=g)).to(cuda)
    bias = (0.2 * torch.randn(d, generator=g)).to(cuda) if with_bias else None
    x_ref = x.clone()
    ref = ref_rmsnorm(x_ref, gamma, bias)
    x_k = x.clone()
    got = netops.rmsnorm(x_k, gamma, bias)
    assert torch.equal(x_k, x_ref)                                  # the stored residual (x + bias, bf16)
    assert float((got.float() - ref.float()).abs().max()) <= 2 ** -7 * float(ref.float().abs().max())   # 1 bf16 ulp
    # in place
    x_i = x.clone()
    netops.rmsnorm(x_i, gamma, bias, out=x_i)
    assert torch.equal(x_i, got)


@pytest.mark.gpu
@pytest.mark.parametrize("pos_div,pos_mod", [(7, 11), (1, 7)])
def test_rotary_kernel(cuda, pos_div, pos_mod):
    import audiolab_b200.netops as netops
    heads, dh, n = 8, 64, 3 * 7 * 11
    g = torch.Generator().manual_seed(5)
    q = torch.randn(n, heads * dh, generator=g).to(torch.bfloat16).to(cuda)
    k = torch.randn(n, heads * dh, generator=g).to(torch.bfloat16).to(cuda)
    freqs = 1.0 / (10000.0 ** (torch.arange(0, dh, 2).float() / dh))
    ang = torch.arange(pos_mod).float()[:, None] * freqs[None]
    cs = torch.stack((ang.cos(), ang.sin()), dim=-1).contiguous().to(cuda)
    qr, kr = q.clone(), k.clone()
    ref_rotary_(qr, kr, cs, heads, dh, pos_div, pos_mod)
    netops.rotary_(q, k, cs, heads, dh, pos_div, pos_mod)
    assert float((q.float() - qr.float()).abs().max()) <= 2 ** -7 * float(qr.float().abs().max())
    assert float((k.float() - kr.float()).abs().max()) <= 2 ** -7 * float(kr.float().abs().max())


@pytest.mark.gpu
def test_gate_kernel(cuda):
    import audiolab_b200.netops as netops
    heads, dh, n = 8, 64, 999
    g = torch.Generator().manual_seed(6)
    o = torch.randn(n, heads * dh, generator=g).to(torch.bfloat16).to(cuda)
    gates = (2 * torch.randn(n, heads, generator=g)).to(torch.bfloat16).to(cuda)
    ref = o.clone()
    ref_gate_(ref, gates, heads, dh)
    netops.gate_sigmoid_(o, gates, heads, dh)
    assert float((o.float() - ref.float()).abs().max()) <= 2 ** -7 * float(ref.float().abs().max())


@pytest.mark.gpu
@pytest.mark.parametrize("n", [8, 8 * 1001, 16 * 4096 + 8, 8 * 1024 * 148 * 2 + 8 * 77])
def test_gelu_kernel_is_torch_gelu(cuda, n):
    import audiolab_b200.netops as netops
    g = torch.Generator().manual_seed(n)
    x = (3 * torch.randn(n, generator=g)).to(torch.bfloat16).to(cuda)
    ref = F.gelu(x)                                     # torch: fp32 inside, erf form, rounded to bf16
    got = netops.gelu_(x.clone())
    assert float((got.float() - ref.float()).abs().max()) <= 2 ** -8 * float(ref.float().abs().max())
    assert float((got.float() - ref.float()).abs().mean()) <= 1e-4
    print(f"gelu n={n}: bit-identical to torch.nn.functional.gelu: {torch.equal(got, ref)}")
    with pytest.raises(ValueError):
        netops.gelu_(torch.zeros(12, dtype=torch.bfloat16, device=cuda))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["bs", "mel"])
def test_fused_bf16_mask_matches_autocast_module_path(cuda, kind):
    """The fused path must be at least as close to the fp32 network as the reference's autocast configuration."""
    cfg, net = _small_net(kind)
    net = net.to(cuda)
    t, f = 1 + cfg.chunk_size // cfg.stft_hop_length, cfg.stft_n_fft // 2 + 1
    g = torch.Generator().manual_seed(3)
    spec = torch.view_as_complex(torch.randn(2, t, f, 2, 2, generator=g)).to(cuda)
    m32 = net.set_compute_dtype(torch.float32).mask(spec)
    net.set_compute_dtype(torch.bfloat16)
    fused = net.mask(spec)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):     # the module path under autocast
        x = net.band_split(torch.view_as_real(spec).reshape(2, t, -1) if kind == "bs" else
                           torch.view_as_real(spec).reshape(2, t, f * 2, 2)[:, :, net.freq_indices].reshape(2, t, -1))
        x = net._axial(x)
        if kind != "mel":
            x = net.final_norm(x)
    x_fused = None
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        xb = net.band_split(torch.view_as_real(spec).reshape(2, t, -1) if kind == "bs" else
                            torch.view_as_real(spec).reshape(2, t, f * 2, 2)[:, :, net.freq_indices].reshape(2, t, -1))
    with torch.no_grad():
        x_fused = net._axial_fused(xb)
        x32 = net._axial(xb.float())
        if kind != "mel":
            x32 = net.final_norm(x32)
    err_fused = float((x_fused.float() - x32).norm() / x32.norm())
    err_autocast = float((x.float() - x32).norm() / x32.norm())
    print(f"{kind}: relative error vs fp32 trunk: fused {err_fused:.3e}, autocast modules {err_autocast:.3e}")
    assert err_fused <= 1.25 * err_autocast + 1e-3
    assert fused.shape == m32.shape and torch.isfinite(torch.view_as_real(fused)).all()
    rel = float((torch.view_as_real(fused) - torch.view_as_real(m32)).norm() / torch.view_as_real(m32).norm())
    assert rel <= 0.05


@pytest.mark.gpu
@pytest.mark.skipif(__import__("os").environ.get("AUDIOLAB_B200_BAND_ATTN") != "1",
                    reason="opt-in kernel (csrc/al_attn.cu): emulation-tested, not yet run on a GPU -- set AUDIOLAB_B200_BAND_ATTN=1")
@pytest.mark.parametrize("F_,H,n_seq", [(62, 8, 700), (64, 8, 33), (17, 4, 5)])
def test_band_attention_kernel(cuda, F_, H, n_seq):
    import audiolab_b200.netops as netops
    g = torch.Generator().manual_seed(F_ + H)
    q, k, v = (torch.randn(n_seq * F_, H * 64, generator=g).to(torch.bfloat16).to(cuda) for _ in range(3))
    gates = (2 * torch.randn(n_seq * F_, H, generator=g)).to(torch.bfloat16).to(cuda)
    ang = torch.arange(F_)[:, None].float() * (1.0 / (10000 ** (torch.arange(0, 64, 2).float() / 64)))[None]
    cs = torch.stack((ang.cos(), ang.sin()), dim=-1).contiguous().to(cuda)
    got = netops.band_attention(q, k, v, n_seq, F_, H, 64, gates=gates, cos_sin=cs).float()
    qr, kr = q.clone(), k.clone()
    ref_rotary_(qr, kr, cs, H, 64, 1, F_)
    ref = ref_band_attention(qr, kr, v, n_seq, F_, H, 64, gates).float()
    assert float((got - ref).abs().max()) <= 2 ** -6 * float(ref.abs().max())


def ref_time_attention(q, k, v, n_batch, seq_len, inner, heads, dh, gates=None):
    """fp32 definition of al_time_attention_bf16: attention along t of token-major rows (b, t, i), per (b, i, head)."""
    shp = (n_batch, seq_len, inner * heads, dh)
    o = F.scaled_dot_product_attention(q.view(shp).transpose(1, 2).float(), k.view(shp).transpose(1, 2).float(),
                                       v.view(shp).transpose(1, 2).float())
    o = o.transpose(1, 2).reshape(q.shape)
    if gates is not None:
        o = (o.view(-1, heads, dh) * torch.sigmoid(gates.float())[:, :, None]).reshape(q.shape)
    return o


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("B,T,I,H", [(2, 801, 3, 2), (1, 128, 1, 1), (1, 129, 2, 1), (3, 33, 5, 2), (1, 1, 1, 1), (2, 300, 2, 3),
                                     (1, 1000, 1, 2), (1, 256, 7, 1), (1, 801, 1, 3)])
def test_time_attention_kernel(cuda, B, T, I, H, dtype):
    """csrc/al_fattn.cu (tcgen05 flash attention along the time axis, gate folded in) against fp32 torch SDPA of the same
    16-bit inputs.  Tolerance: the output is rounded to 16 bits (2^-11 / 2^-8 relative) and P is a 16-bit operand."""
    import audiolab_b200.netops as netops
    g = torch.Generator().manual_seed(B * 1000 + T + I + H)
    q, k, v = (torch.randn(B * T * I, H * 64, generator=g).to(dtype).to(cuda) for _ in range(3))
    gates = (2 * torch.randn(B * T * I, 16, generator=g)).to(dtype).to(cuda)[:, :H]      # strided rows, as in the network
    got = netops.time_attention(q, k, v, B, T, I, H, 64, gates=gates).float()
    ref = ref_time_attention(q, k, v, B, T, I, H, 64, gates)
    tol = (2 ** -9 if dtype == torch.float16 else 2 ** -6) * float(ref.abs().max())
    assert torch.isfinite(got).all()
    assert float((got - ref).abs().max()) <= tol
    got2 = netops.time_attention(q, k, v, B, T, I, H, 64).float()                         # no gate
    ref2 = ref_time_attention(q, k, v, B, T, I, H, 64)
    assert float((got2 - ref2).abs().max()) <= tol


@pytest.mark.gpu
def test_time_attention_kernel_rescales_when_the_row_maximum_grows(cuda):
    """Scores that grow along the key axis by far more than 2^8 per key tile: the lazy reference of the online softmax must
    move and the accumulator in tensor memory must be rescaled (the rare path of al_fattn.cu)."""
    import audiolab_b200.netops as netops
    B, T, I, H = 1, 700, 2, 2
    g = torch.Generator().manual_seed(7)
    q = torch.randn(B * T * I, H * 64, generator=g)
    k = torch.randn(B * T * I, H * 64, generator=g)
    v = torch.randn(B * T * I, H * 64, generator=g)
    # key t = q-direction-independent noise + a component along a fixed direction that grows with t; queries have a
    # positive component on that direction, so the maximum of row i sits near the last keys and rises tile after tile
    d = torch.zeros(H * 64)
    d[::2] = 1.0
    t_idx = (torch.arange(B * T * I) // I) % T
    k = 0.3 * k + d[None, :] * (t_idx[:, None].float() / 40.0)
    q = 0.3 * q + d[None, :] * 0.5
    q, k, v = (x.to(torch.float16).to(cuda) for x in (q, k, v))
    got = netops.time_attention(q, k, v, B, T, I, H, 64).float()
    ref = ref_time_attention(q, k, v, B, T, I, H, 64)
    assert torch.isfinite(got).all()
    assert float((got - ref).abs().max()) <= 2 ** -8 * float(ref.abs().max())


# ---- host logic of the tcgen05 path (nets/roformer.py::_axial_tc) with the GEMM replaced by its torch definition ----
def ref_gemm_bf16(a, w, outs, *, bias=None, row_ss=None, ss_scale=1.0, ss_eps=1e-12, cos_sin=None, pos_div=1, pos_mod=1,
                  rot_cols=0, act=None, out_split=0, max_ctas=0):
    if isinstance(outs, torch.Tensor):
        outs = [outs]
    y = a.float() @ w.float().t()
    m, n = y.shape
    if row_ss is not None:
        y = y * (ss_scale / row_ss.view(m, -1).sum(-1).sqrt().clamp(min=ss_eps))[:, None]
    if bias is not None:
        y = y + bias
    if cos_sin is not None:
        assert rot_cols % 64 == 0 and tuple(cos_sin.shape) == (pos_mod, 32, 2)
        pos = (torch.arange(m) // pos_div) % pos_mod
        c, s = cos_sin[pos][:, None, :, 0], cos_sin[pos][:, None, :, 1]
        z = y[:, :rot_cols].reshape(m, rot_cols // 64, 32, 2)
        y = torch.cat((torch.stack((z[..., 0] * c - z[..., 1] * s, z[..., 1] * c + z[..., 0] * s), dim=-1).reshape(m, rot_cols),
                       y[:, rot_cols:]), dim=1)
    if act == "gelu":
        y = F.gelu(y)
    elif act == "tanh":
        y = torch.tanh(y)
    split = out_split or n
    for i, o in enumerate(outs):
        o.copy_(y[:, i * split: i * split + o.shape[1]].to(o.dtype))


def ref_gemm_bf16_residual(a, w, x32, xb, ss_out, *, bias=None, max_ctas=0):
    import audiolab_b200.netops as netops
    y = x32 + a.float() @ w.float().t()
    if bias is not None:
        y = y + bias
    x32.copy_(y)
    xb.copy_(y.to(xb.dtype))
    slab = netops.resid_slab(y.shape[1])
    ss_out.copy_(y.view(y.shape[0], -1, slab).square().sum(-1))


def ref_resid_prepare(x_in, x32, xb, ss, *, bias=None, gamma=None, eps=1e-12):
    y = x_in if bias is None else x_in + bias
    if gamma is not None:
        y = F.normalize(y, dim=-1) * (y.shape[-1] ** 0.5) * gamma
    y = y.clone()
    x32.copy_(y)
    xb.copy_(y.to(xb.dtype))
    ss.copy_(y.view(y.shape[0], ss.shape[1], -1).square().sum(-1))


@pytest.mark.parametrize("kind", ["bs", "mel"])
def test_tc_axial_host_logic_matches_module_path(kind, monkeypatch):
    """`_axial_tc`: gamma folded into the weights + row scale in the consumer's epilogue, to_gates riding on to_qkv,
    rotary in the epilogue, fp32 residual stream -- with the kernels replaced by fp32 torch definitions the result must
    equal the upstream-shaped module path."""
    import audiolab_b200.netops as netops
    torch.manual_seed(0)
    kw = dict(dim=128, depth=2, heads=4, dim_head=64, chunk_size=441 * 12)
    cfg = RoformerConfig(**kw) if kind == "bs" else RoformerConfig(kind="mel", num_bands=20, **kw)
    net = RoformerMaskNet(cfg).eval()
    with torch.no_grad():
        for p in net.parameters():
            p.add_(0.05 * torch.randn_like(p))
    assert net._tc_supported()
    monkeypatch.setattr(netops, "gemm_bf16", ref_gemm_bf16)
    monkeypatch.setattr(netops, "gemm_bf16_residual", ref_gemm_bf16_residual)
    monkeypatch.setattr(netops, "resid_prepare", ref_resid_prepare)
    monkeypatch.setattr(netops, "gate_sigmoid_", ref_gate_)
    monkeypatch.setattr(netops, "band_attention",
                        lambda q, k, v, n_seq, seq_len, heads, dh, gates=None, cos_sin=None:
                        ref_band_attention(q, k, v, n_seq, seq_len, heads, dh, gates))
    monkeypatch.setattr(netops, "time_attention",
                        lambda q, k, v, n_batch, seq_len, inner, heads, dh, gates=None:
                        ref_time_attention(q, k, v, n_batch, seq_len, inner, heads, dh, gates).to(q.dtype))
    net._fused_dtype = torch.float32
    b, t, f = 2, 13, len(net.band_split.dim_inputs)
    x = torch.randn(b, t, f, cfg.dim)
    with torch.no_grad():
        ref = net._axial(x.clone())
        if kind != "mel":
            ref = net.final_norm(ref)
        got = net._axial_tc(x.clone())
    assert float((got - ref).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max()))


def ref_band_norm(x, gamma, band_off, out, eps=1e-12):
    offs = band_off.tolist()
    for a, b in zip(offs[:-1], offs[1:]):
        seg = x[:, a:b]
        out[:, a:b] = (F.normalize(seg, dim=-1, eps=eps) * ((b - a) ** 0.5) * gamma[a:b]).to(out.dtype)


def ref_gemm_bf16_glu(a, w, out, *, bias=None, max_ctas=0):
    y = torch.einsum("gmk,gnk->gmn", a.float(), w.float()) if a.dim() == 3 else a.float() @ w.float().t()
    if bias is not None:
        y = y + (bias[:, None, :] if a.dim() == 3 else bias)
    z = y[..., 0::2] * torch.sigmoid(y[..., 1::2])
    out.copy_(z[..., : out.shape[-1]])


def ref_gemm_any(a, w, outs, **kw):
    if a.dim() == 3:
        bias = kw.pop("bias", None)
        for g in range(a.shape[0]):
            ref_gemm_bf16(a[g], w[g], [o[g] for o in ([outs] if isinstance(outs, torch.Tensor) else outs)],
                          bias=None if bias is None else bias[g], **kw)
    else:
        ref_gemm_bf16(a, w, outs, **kw)


def ref_gemm_residual_any(a, w, x32, xb, ss_out, *, bias=None, max_ctas=0, accumulate=True):
    import audiolab_b200.netops as netops
    if a.dim() == 2:
        a, w, x32, xb, ss_out = a[None], w[None], x32[None], xb[None], ss_out.view(1, x32.shape[0], -1)
        bias = None if bias is None else bias[None]
    slab = netops.resid_slab(x32.shape[-1])
    for g in range(a.shape[0]):
        y = a[g].float() @ w[g].float().t()
        if accumulate:
            y = y + x32[g]
        if bias is not None:
            y = y + bias[g]
        x32[g].copy_(y)
        xb[g].copy_(y.to(xb.dtype))
        ss_out[g].copy_(y.view(y.shape[0], -1, slab).square().sum(-1))


def test_tc_grouped_band_split_and_mask_estimator_host_logic(monkeypatch):
    """The all-tcgen05 `mask()` path of BS-RoFormer: per-band RMSNorm kernel + grouped band-split GEMMs that start the fp32
    stream, grouped mask-estimator GEMMs with the GLU epilogue writing into the mask tensor (interleaved / zero-padded
    weights, strided views per run of equal-width bands) -- against the upstream-shaped module path."""
    import audiolab_b200.netops as netops
    torch.manual_seed(0)
    # band widths 2,2,2,4,4,5 (the lone 5-wide band exercises the K / N padding, like the 129-bin band of the real model)
    cfg = RoformerConfig(dim=128, depth=1, heads=4, dim_head=64, chunk_size=441 * 12, stft_n_fft=36,
                         freqs_per_bands=(2, 2, 2, 4, 4, 5), num_stems=2)
    net = RoformerMaskNet(cfg).eval()
    with torch.no_grad():
        for p in net.parameters():
            p.add_(0.05 * torch.randn_like(p))
    assert net._grouped_supported()
    monkeypatch.setattr(netops, "gemm_bf16", ref_gemm_any)
    monkeypatch.setattr(netops, "gemm_bf16_residual", ref_gemm_residual_any)
    monkeypatch.setattr(netops, "gemm_bf16_glu", ref_gemm_bf16_glu)
    monkeypatch.setattr(netops, "band_norm", ref_band_norm)
    monkeypatch.setattr(netops, "resid_prepare", ref_resid_prepare)
    monkeypatch.setattr(netops, "gate_sigmoid_", ref_gate_)
    monkeypatch.setattr(netops, "band_attention",
                        lambda q, k, v, n_seq, seq_len, heads, dh, gates=None, cos_sin=None:
                        ref_band_attention(q, k, v, n_seq, seq_len, heads, dh, gates))
    monkeypatch.setattr(netops, "time_attention",
                        lambda q, k, v, n_batch, seq_len, inner, heads, dh, gates=None:
                        ref_time_attention(q, k, v, n_batch, seq_len, inner, heads, dh, gates).to(q.dtype))
    b, t, f, s = 2, 7, 19, 2
    spec = torch.randn(b, t, f, s, dtype=torch.complex64)
    ref = net.mask(spec.clone())                                    # module path (compute dtype fp32)
    net._fused_dtype = torch.float32
    net.set_compute_dtype(torch.bfloat16)                           # selects the tc path; the stand-ins compute in fp32
    got = net.mask(spec.clone())
    assert got.shape == ref.shape == (b, 2, t, f, s)
    assert float((got - ref).abs().max()) <= 2e-5 * max(1.0, float(ref.abs().max()))
