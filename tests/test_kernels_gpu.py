"""Kernel-level parity (-m gpu): the CUDA path through the C ABI vs torch.stft / torch.istft on the
CPU and the oracle, on seeded inputs.  Tolerances (fp32):
  spectra : max|err| <= 2e-6 * max|ref|   (fp32 FFT rounding, scale-free)
  waves   : max|err| <= 2e-5              (north star: 1e-4 on stems)
"""
import numpy as np
import pytest
import torch

from oracle import mdx as omdx
from oracle.metrics import max_abs_err, rel_err
from oracle.resample import resample_poly_ref
from oracle.synth import synth_mix, synth_noise

pytestmark = pytest.mark.gpu

SPEC_RTOL = 2e-6
WAVE_ATOL = 2e-5

CASES = [  # n_fft, hop, normalized, chunk_len
    (2048, 441, False, 441 * 40),
    (2048, 512, False, 512 * 24 + 100),
    (4096, 1024, True, 1024 * 20),
    (6144, 1024, False, 1024 * 15),
]


def _plan(n_fft, hop, normalized=False, window=None):
    from audiolab_b200.spectral import StftPlan
    return StftPlan(n_fft, hop, window=window, normalized=normalized)


def _torch_stft(x, n_fft, hop, normalized):
    return torch.stft(x, n_fft, hop, window=torch.hann_window(n_fft), center=True, normalized=normalized,
                      return_complex=True)


@pytest.mark.parametrize("n_fft,hop,normalized,L", CASES)
def test_stft_matches_torch_all_layouts(cuda, n_fft, hop, normalized, L):
    from audiolab_b200 import spectral as sp
    x = torch.tensor(synth_mix(L, seed=n_fft + hop))                  # [2, L]
    ref = _torch_stft(x, n_fft, hop, normalized)                       # [2, F, T]
    plan = _plan(n_fft, hop, normalized)
    xd = x.to(cuda)
    got_bm = plan.stft(xd, chunk_len=L, layout=sp.BIN_MAJOR).cpu()
    assert got_bm.shape == ref.shape
    assert rel_err(got_bm, ref) <= SPEC_RTOL
    got_fm = plan.stft(xd, chunk_len=L, layout=sp.FRAME_MAJOR).cpu()
    assert rel_err(got_fm.transpose(1, 2), ref) <= SPEC_RTOL
    got_cac = plan.stft(xd, chunk_len=L, layout=sp.CAC).cpu()          # [1, 4, F, T]
    ref_cac = torch.view_as_real(ref).permute(0, 3, 1, 2).reshape(1, 4, ref.shape[1], ref.shape[2])
    assert rel_err(got_cac, ref_cac) <= SPEC_RTOL


def test_stft_custom_window_and_crop(cuda):
    from audiolab_b200 import spectral as sp
    n_fft, hop, L = 2048, 300, 9000
    w = torch.hamming_window(n_fft)
    x = torch.tensor(synth_mix(L, seed=3))
    ref = torch.stft(x, n_fft, hop, window=w, center=True, return_complex=True)
    plan = _plan(n_fft, hop, window=w)
    got = plan.stft(x.to(cuda), chunk_len=L, layout=sp.BIN_MAJOR, n_bins_out=700, zero_low_bins=3).cpu()
    ref = ref[:, :700].clone()
    ref[:, :3] = 0
    assert rel_err(got, ref) <= SPEC_RTOL


def test_stft_chunks_with_virtual_zero_padding(cuda):
    """pad-and-chunk in-kernel: chunk offsets before 0 / past the end read zeros (mdxnet.py:155-163)."""
    from audiolab_b200 import spectral as sp
    cfg = omdx.MdxConfig(n_fft=6144, dim_f=3072, dim_t_log2=4)
    n = 20001
    mix = synth_mix(n, seed=11)
    trim, gen, chunk = cfg.trim, cfg.gen_size, cfg.chunk_size
    pad = gen - n % gen
    mix_p = np.concatenate((np.zeros((2, trim)), mix, np.zeros((2, pad)), np.zeros((2, trim))), 1)
    waves = np.stack([mix_p[:, i:i + chunk] for i in range(0, n + pad, gen)]).astype(np.float32)
    ref = omdx.MdxSpec(cfg).stft(torch.tensor(waves))                  # [N, 4, dim_f, dim_t]
    plan = _plan(cfg.n_fft, cfg.hop)
    got = plan.stft(torch.tensor(mix).to(cuda), chunk_len=chunk, n_chunks=waves.shape[0], off0=-trim,
                    off_step=gen, n_frames=cfg.dim_t, layout=sp.CAC, n_bins_out=cfg.dim_f).cpu()
    assert got.shape == ref.shape
    assert rel_err(got, ref) <= SPEC_RTOL


def test_stft_explicit_offsets_array(cuda):
    from audiolab_b200 import spectral as sp
    n_fft, hop, L = 2048, 441, 441 * 30
    mix = torch.tensor(synth_mix(50000, seed=5))
    offs = [0, 7000, 50000 - L]
    ref = torch.stack([_torch_stft(mix[:, o:o + L], n_fft, hop, False) for o in offs])   # [3, 2, F, T]
    plan = _plan(n_fft, hop)
    got = plan.stft(mix.to(cuda), chunk_len=L, n_chunks=3,
                    offsets=torch.tensor(offs, dtype=torch.int64, device=cuda), layout=sp.BIN_MAJOR).cpu()
    assert rel_err(got.reshape(ref.shape), ref) <= SPEC_RTOL


@pytest.mark.parametrize("n_fft,hop,normalized,L", CASES)
def test_istft_matches_torch_random_spectrum(cuda, n_fft, hop, normalized, L):
    """Not a round trip: an arbitrary (non-STFT-consistent) spectrum, incl. Im(DC)/Im(Nyquist) != 0."""
    from audiolab_b200 import spectral as sp
    T = 1 + L // hop
    F = n_fft // 2 + 1
    S = torch.view_as_complex(torch.tensor(synth_noise((2, F, T, 2), seed=n_fft)))
    ref = torch.istft(S, n_fft, hop, window=torch.hann_window(n_fft), center=True, normalized=normalized)
    plan = _plan(n_fft, hop, normalized)
    got = plan.istft(S.to(cuda).contiguous(), n_chunks=1, channels=2, layout=sp.BIN_MAJOR).cpu()
    assert got.shape == (1, 1, 2, ref.shape[-1])
    assert max_abs_err(got[0, 0], ref) <= WAVE_ATOL * max(1.0, float(ref.abs().max()))
    got2 = plan.istft(S.transpose(1, 2).contiguous().to(cuda), n_chunks=1, channels=2, layout=sp.FRAME_MAJOR).cpu()
    assert max_abs_err(got2[0, 0], ref) <= WAVE_ATOL * max(1.0, float(ref.abs().max()))
    assert torch.equal(got, got2)          # layout must not change a single bit of the sum order


@pytest.mark.parametrize("n_fft,hop,normalized,L", CASES)
def test_round_trip_identity(cuda, n_fft, hop, normalized, L):
    from audiolab_b200 import spectral as sp
    x = torch.tensor(synth_mix(L, seed=hop))
    plan = _plan(n_fft, hop, normalized)
    for layout in (sp.FRAME_MAJOR, sp.BIN_MAJOR, sp.CAC):
        S = plan.stft(x.to(cuda), chunk_len=L, layout=layout)
        y = plan.istft(S, n_chunks=1, channels=2, layout=layout).cpu()[0, 0]
        assert max_abs_err(y, x[:, : y.shape[-1]]) <= WAVE_ATOL


def test_istft_fused_complex_mask_and_stems(cuda):
    from audiolab_b200 import spectral as sp
    n_fft, hop, L = 2048, 441, 441 * 50
    T, F = 1 + L // hop, n_fft // 2 + 1
    x = torch.tensor(synth_mix(2 * L, seed=8)).reshape(2, 2, L).transpose(0, 1).contiguous()   # 2 chunks [c, ch, L]
    X = torch.stack([_torch_stft(x[c], n_fft, hop, False) for c in range(2)])                  # [2, 2, F, T]
    M = torch.view_as_complex(torch.tensor(synth_noise((2, 3, 2, F, T, 2), seed=9)))            # [c, stem, ch, F, T]
    Y = X[:, None] * M
    ref = torch.istft(Y.reshape(-1, F, T), n_fft, hop, window=torch.hann_window(n_fft), center=True, length=L)
    ref = ref.reshape(2, 3, 2, L)
    plan = _plan(n_fft, hop)
    got = plan.istft(X.reshape(-1, F, T).contiguous().to(cuda), mask=M.reshape(-1, F, T).contiguous().to(cuda),
                     n_chunks=2, channels=2, stems=3, layout=sp.BIN_MAJOR, out_len=L).cpu()
    assert max_abs_err(got, ref) <= WAVE_ATOL * max(1.0, float(ref.abs().max()))
    got_fm = plan.istft(X.reshape(-1, F, T).transpose(1, 2).contiguous().to(cuda),
                        mask=M.reshape(-1, F, T).transpose(1, 2).contiguous().to(cuda),
                        n_chunks=2, channels=2, stems=3, layout=sp.FRAME_MAJOR, out_len=L).cpu()
    assert torch.equal(got, got_fm)


@pytest.mark.parametrize("hop,T,stems,use_mask", [(441, 301, 2, True), (441, 61, 1, False), (512, 130, 1, True),
                                                  (300, 97, 3, True), (100, 150, 1, True), (256, 77, 2, False),
                                                  (777, 45, 1, True), (1024, 33, 2, True)])
def test_istft_frame_interleaved_packed_path_matches_torch(cuda, hop, T, stems, use_mask):
    """Layout 3 (RoFormer 'b t (f c)'), stereo, n_fft 2048 -> the packed fast path of K2 (al_istft_pk.cu).
    Arbitrary spectrum incl. Im(DC) / Im(Nyquist) != 0, fused mask, stems, ragged out_len, several segments; hops from 100
    (21 frames per position) to 1024, odd and even, below and above the 448 of the register-held hop-block emission."""
    from audiolab_b200 import spectral as sp
    n_fft, F, nch = 2048, 1025, 2
    L = (T - 1) * hop - 37                                                       # out_len not a multiple of hop
    X = torch.view_as_complex(torch.tensor(synth_noise((nch, 2, F, T, 2), seed=hop + T)))        # [c, ch, F, T]
    if use_mask:
        M = torch.view_as_complex(torch.tensor(synth_noise((nch, stems, 2, F, T, 2), seed=T)))   # [c, s, ch, F, T]
        Y = X[:, None] * M
    else:
        stems, M = 1, None
        Y = X[:, None]
    ref = torch.istft(Y.reshape(-1, F, T), n_fft, hop, window=torch.hann_window(n_fft), center=True, length=L)
    ref = ref.reshape(nch, stems, 2, L)
    plan = _plan(n_fft, hop)
    Xi = X.permute(0, 3, 2, 1).contiguous().to(cuda)                             # [c, T, F, ch]
    Mi = None if M is None else M.permute(0, 1, 4, 3, 2).contiguous().to(cuda)   # [c, s, T, F, ch]
    got = plan.istft(Xi, mask=Mi, n_chunks=nch, channels=2, stems=stems, layout=sp.FRAME_INTERLEAVED, out_len=L)
    tol = WAVE_ATOL * max(1.0, float(ref.abs().max()))
    assert max_abs_err(got.cpu(), ref) <= tol
    # tiling independence: one chunk alone is segmented differently, yet must give the same bits
    one = plan.istft(Xi[:1].contiguous(), mask=None if Mi is None else Mi[:1].contiguous(), n_chunks=1, channels=2,
                     stems=stems, layout=sp.FRAME_INTERLEAVED, out_len=L)
    assert torch.equal(one[0], got[0])
    # the generic kernel (bin-major layout) agrees within rounding
    gen = plan.istft(X.reshape(-1, F, T).contiguous().to(cuda),
                     mask=None if M is None else M.reshape(-1, F, T).contiguous().to(cuda),
                     n_chunks=nch, channels=2, stems=stems, layout=sp.BIN_MAJOR, out_len=L)
    assert max_abs_err(got.cpu(), gen.cpu()) <= tol


def test_istft_frame_interleaved_weight_and_placement(cuda):
    """Packed K2 path with the chunk weight, per-chunk placement into a track and the dst_limit clip."""
    from audiolab_b200 import spectral as sp
    n_fft, hop, T, F, nch = 2048, 441, 41, 1025, 3
    L = (T - 1) * hop
    X = torch.view_as_complex(torch.tensor(synth_noise((nch, 2, F, T, 2), seed=77)))
    ref = torch.istft(X.reshape(-1, F, T), n_fft, hop, window=torch.hann_window(n_fft), center=True, length=L)
    ref = ref.reshape(nch, 2, L)
    w = torch.tensor(synth_noise((L,), seed=5)).abs() + 0.1
    plan = _plan(n_fft, hop)
    n_track = 3 * L - 200
    track = torch.zeros((2, n_track), device=cuda)
    offs = torch.tensor([0, L + 100, 2 * L + 300], dtype=torch.int64, device=cuda)  # last chunk clipped by dst_limit
    plan.istft(X.permute(0, 3, 2, 1).contiguous().to(cuda), n_chunks=nch, channels=2, layout=sp.FRAME_INTERLEAVED,
               out_len=L, weight=w.to(cuda), dst=track, dst_ch_stride=n_track, dst_chunk_stride=0, dst_offsets=offs,
               dst_limit=n_track)
    exp = torch.zeros((2, n_track))
    exp[:, :L] = ref[0] * w
    exp[:, L + 100:2 * L + 100] = ref[1] * w
    exp[:, 2 * L + 300:] = (ref[2] * w)[:, : n_track - (2 * L + 300)]
    assert max_abs_err(track.cpu(), exp) <= WAVE_ATOL * max(1.0, float(exp.abs().max()))


def test_istft_cac_freq_pad_and_trim_concat_placement(cuda):
    """mdxnet.py:58-75 (zero freq-pad) + :178-183 (trim, concat, drop pad) fused into one launch."""
    from audiolab_b200 import spectral as sp
    cfg = omdx.MdxConfig(n_fft=6144, dim_f=3072, dim_t_log2=4)
    n_chunks, n = 3, 20001
    spec = torch.tensor(synth_noise((n_chunks, 4, cfg.dim_f, cfg.dim_t), seed=4))
    waves = omdx.MdxSpec(cfg).istft(spec)                                     # [3, 2, chunk]
    trim, gen = cfg.trim, cfg.gen_size
    ref = waves[:, :, trim:-trim].transpose(0, 1).reshape(2, -1)[:, :n]
    plan = _plan(cfg.n_fft, cfg.hop)
    dst = torch.full((2, n), 7.0, device=cuda)
    plan.istft(spec.to(cuda), n_chunks=n_chunks, channels=2, layout=sp.CAC, out_start=cfg.n_fft // 2 + trim,
               out_len=gen, dst=dst, dst_ch_stride=n, dst_chunk_stride=0, dst_off0=0, dst_off_step=gen, dst_limit=n)
    assert max_abs_err(dst.cpu(), ref) <= WAVE_ATOL * max(1.0, float(ref.abs().max()))


def test_istft_frame_pad_and_crop_like_htdemucs(cuda):
    from oracle import htdemucs as oh
    from audiolab_b200 import spectral as sp
    cfg = oh.HTDemucsConfig()
    L = 1024 * 12 + 300
    x = torch.tensor(synth_mix(L, seed=21))[None]                              # [1, 2, L]
    z = oh.spec(x, cfg)                                                        # [1, 2, 2048, le]
    le = z.shape[-1]
    plan = _plan(4096, 1024, normalized=True)
    got_z = plan.stft(x[0].to(cuda), chunk_len=L, center_pad=1536, n_frames=le, layout=sp.BIN_MAJOR,
                      n_bins_out=2048).cpu()
    assert rel_err(got_z, z[0]) <= SPEC_RTOL
    ref = oh.ispec(z, L, cfg)                                                  # [1, 2, L]
    got = plan.istft(z[0].contiguous().to(cuda), n_chunks=1, channels=2, layout=sp.BIN_MAJOR, frame_pad=2,
                     out_start=2048 + 1536, out_len=L).cpu()
    assert max_abs_err(got[0, 0], ref[0]) <= WAVE_ATOL


@pytest.mark.parametrize("n_fft,hop,T,crop,low", [(4096, 1024, 24, 2048, 0), (6144, 1024, 16, 3072, 3), (2048, 441, 40, 1025, 0),
                                                 (6144, 1024, 256, 3072, 0)])
def test_stft_cac_vector_stores_equal_bin_major_bitwise(cuda, n_fft, hop, T, crop, low):
    """AL_LAYOUT_CAC with n_frames % 4 == 0 writes 128-bit rows; the values must be those of the c64 [rows, F, T]
    layout bit for bit (cropped frequency rows, zeroed low bins, two chunks)."""
    from audiolab_b200 import spectral as sp
    L = (T - 1) * hop
    x = torch.tensor(synth_mix(2 * L, seed=T)).to(cuda)
    kw = dict(chunk_len=L, n_chunks=2, off0=0, off_step=L, n_frames=T, n_bins_out=crop, zero_low_bins=low)
    plan = _plan(n_fft, hop)
    a = plan.stft(x, layout=sp.BIN_MAJOR, **kw)                       # c64 [chunks*2, crop, T]
    b = plan.stft(x, layout=sp.CAC, **kw)                             # f32 [chunks, 4, crop, T]
    a4 = torch.view_as_real(a).reshape(2, 2, crop, T, 2).permute(0, 1, 4, 2, 3).reshape(2, 4, crop, T)
    assert torch.equal(a4, b.reshape(2, 4, crop, T))
    ref = _torch_stft(x[:, :L].cpu(), n_fft, hop, False)[:, :crop]    # chunk 0 alone (reflect pad at both chunk ends)
    ref[:, :low] = 0
    assert rel_err(a.reshape(2, 2, crop, T)[0].cpu(), ref) <= SPEC_RTOL


@pytest.mark.parametrize("n_fft,hop,T,frame_pad", [(4096, 1024, 64, 0), (4096, 1024, 336, 2), (6144, 1024, 256, 0),
                                                  (6144, 1024, 20, 2), (2048, 441, 80, 0), (2048, 512, 40, 1)])
def test_istft_cac_vector_path_equals_bin_major_bitwise(cuda, n_fft, hop, T, frame_pad):
    """AL_LAYOUT_CAC with T % 4 == 0 takes the 128-bit stage-A loads (rounds re-aligned to 4 stored frames); the
    result must equal the c64 [rows, F, T] layout bit for bit -- same frames, same ascending sum order -- and
    torch.istft within tolerance.  Several segments per row (16 hops each), cropped frequency rows."""
    from audiolab_b200 import spectral as sp
    F = n_fft // 2 + 1
    Fo = F - 1 if frame_pad else F                                  # HTDemucs drops the Nyquist row
    S = torch.view_as_complex(torch.tensor(synth_noise((2, Fo, T, 2), seed=n_fft + T)))
    Sfull = torch.zeros((2, F, T + 2 * frame_pad), dtype=S.dtype)
    Sfull[:, :Fo, frame_pad:frame_pad + T] = S
    ref = torch.istft(Sfull, n_fft, hop, window=torch.hann_window(n_fft), center=True)
    plan = _plan(n_fft, hop)
    out_len = ref.shape[-1]
    a = plan.istft(S.contiguous().to(cuda), n_chunks=1, channels=2, layout=sp.BIN_MAJOR, frame_pad=frame_pad,
                   out_len=out_len).cpu()
    cac = torch.stack((S.real, S.imag), dim=1).reshape(1, 4, Fo, T).contiguous()   # [chunks, (ch, re/im), F, T]
    b = plan.istft(cac.to(cuda), n_chunks=1, channels=2, layout=sp.CAC, frame_pad=frame_pad, out_len=out_len).cpu()
    assert max_abs_err(a[0, 0], ref) <= WAVE_ATOL * max(1.0, float(ref.abs().max()))
    assert torch.equal(a, b)


def test_ola_gather_matches_numpy_and_is_bitwise_shardable(cuda):
    from audiolab_b200 import spectral as sp
    rs = np.random.RandomState(0)
    C, step, n, rows = 4000, 1000, 13337, 3
    offs = list(range(0, n - C + 1, step)) + [n - C]
    mult = [1] * (len(offs) - 1) + [3]
    chunks = rs.standard_normal((len(offs), rows, C)).astype(np.float32)
    w = np.hamming(C).astype(np.float32)
    res = np.zeros((rows, n), np.float32)
    cnt = np.zeros((rows, n), np.float32)
    for c, (o, m) in enumerate(zip(offs, mult)):
        for _ in range(m):
            res[:, o:o + C] += chunks[c] * w
            cnt[:, o:o + C] += w
    ref = res / np.maximum(cnt, 1e-10)
    d = lambda a, dt: torch.tensor(np.asarray(a), dtype=dt, device=cuda)
    args = dict(mult=d(mult, torch.int32), wtab=d(w[None], torch.float32))
    got = sp.ola_gather(d(chunks, torch.float32), d(offs, torch.int64), n, **args)
    assert max_abs_err(got.cpu(), ref) <= 2e-6
    # two "ranks": left produces raw partial sums for the right's span, right continues from them
    cut = 6000
    k = sum(1 for o in offs if o < cut)
    left = sp.ola_gather(d(chunks[:k], torch.float32), d(offs[:k], torch.int64), n, p0=0, p1=cut,
                         mult=d(mult[:k], torch.int32), wtab=args["wtab"])
    halo = sp.ola_gather(d(chunks[:k], torch.float32), d(offs[:k], torch.int64), n, p0=cut, p1=n, raw_out=True,
                         mult=d(mult[:k], torch.int32), wtab=args["wtab"])[:, cut:].contiguous()
    # the right rank counts the left chunks' weights (all offsets) but holds only its own chunk data
    full_right = sp.ola_gather(d(chunks[k:], torch.float32), d(offs, torch.int64), n, p0=cut, p1=n,
                               halo_in=halo, data_chunk0=k, **args)
    stitched = torch.cat([left[:, :cut], full_right[:, cut:]], dim=1)
    assert torch.equal(stitched, got)


def test_resample_poly_matches_scipy(cuda):
    from audiolab_b200 import spectral as sp
    for n_in in (48000, 12345, 160, 7):
        x = synth_mix(n_in, seed=n_in, sr=48000)
        ref = resample_poly_ref(x)
        got = sp.resample_poly(torch.tensor(x).to(cuda)).cpu().numpy()
        assert got.shape == ref.shape
        assert max_abs_err(got, ref) <= 2e-6
    x = synth_mix(30000, seed=2)
    ref = resample_poly_ref(x, 160, 147)
    got = sp.resample_poly(torch.tensor(x).to(cuda), 160, 147).cpu().numpy()
    assert max_abs_err(got, ref) <= 2e-6


def test_errors_are_loud(cuda):
    from audiolab_b200 import spectral as sp
    plan = _plan(2048, 441)
    with pytest.raises(RuntimeError):
        plan.stft(torch.zeros(2, 100), chunk_len=100)               # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        plan.stft(torch.zeros(2, 100, device=cuda), chunk_len=100)  # chunk_len <= center_pad
    with pytest.raises(RuntimeError):
        sp.StftPlan(1024, 256)                                      # unsupported n_fft
