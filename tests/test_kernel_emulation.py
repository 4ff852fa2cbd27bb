"""Host emulation of the index-logic CUDA kernels (tools/cpu_emul): the same source text as
audiolab_b200/csrc/al_ola.cu / al_resample.cu between the [emul-begin]/[emul-end] markers, compiled
for the host with one std::thread per CUDA thread.  Checks addressing, masking, tiling and the
accumulation order on a box without a GPU; the GPU parity tests (tests/test_kernels_gpu.py) remain
the parity gate."""
import ctypes
import importlib.util
import os

import numpy as np
import pytest

from oracle.resample import resample_poly_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emul():
    spec = importlib.util.spec_from_file_location("build_emul", os.path.join(ROOT, "tools", "cpu_emul", "build_emul.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    lib = ctypes.CDLL(mod.build())
    P, LL, I, F = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_float
    lib.emul_ola_gather.argtypes = [P, I, I, I, I, P, P, P, P, LL, LL, LL, P, I, F, F, P, LL, I]
    lib.emul_ola_gather.restype = None
    lib.emul_resample.argtypes = [P, LL, P, LL, I, LL, LL, I, I, P, I, I]
    lib.emul_resample.restype = I
    return lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _ola(lib, chunks, offs, n, mult=None, wtab=None, p0=0, p1=None, halo_in=None, raw_out=False, data_chunk0=0,
         shift_elems=0):
    n_data, rows, C = chunks.shape
    p1 = n if p1 is None else p1
    offs = np.ascontiguousarray(offs, np.int64)
    mult = None if mult is None else np.ascontiguousarray(mult, np.int32)
    buf = np.zeros(rows * n + 8, np.float32)                       # shift_elems moves the track off 16-byte alignment
    track = buf[shift_elems:shift_elems + rows * n].reshape(rows, n)
    lib.emul_ola_gather(_p(chunks), n_data + data_chunk0, data_chunk0, rows, C, _p(offs), _p(mult), _p(wtab), None, n,
                        p0, p1, _p(halo_in), int(raw_out), 1e-10, 1.0, _p(track), n, 64)
    return track.copy()


@pytest.mark.parametrize("C,step,n,shift,rows", [(4000, 1000, 13337, 0, 3), (4000, 1000, 13337, 1, 2), (1021, 255, 5003, 0, 4),
                                                (64, 64, 640, 0, 1), (4000, 1000, 13337, 0, 2)])
def test_ola_gather_emulated_matches_numpy_and_shards_bitwise(emul, C, step, n, shift, rows):
    """Odd row counts take the one-row-per-thread kernel, even ones the two-rows-per-thread kernel."""
    rs = np.random.RandomState(0)
    offs = list(range(0, n - C + 1, step)) + [n - C]
    mult = [1] * (len(offs) - 1) + [3]
    chunks = rs.standard_normal((len(offs), rows, C)).astype(np.float32)
    w = np.hamming(C).astype(np.float32)
    res = np.zeros((rows, n), np.float64)
    cnt = np.zeros((rows, n), np.float64)
    for c, (o, m) in enumerate(zip(offs, mult)):
        for _ in range(m):
            res[:, o:o + C] += chunks[c].astype(np.float64) * w
            cnt[:, o:o + C] += w
    ref = res / np.maximum(cnt, 1e-10)
    got = _ola(emul, chunks, offs, n, mult=mult, wtab=w, shift_elems=shift)
    assert np.abs(got - ref).max() <= 2e-6
    cut = (n // 2) | 1 if shift else (n // 2) & ~3
    k = sum(1 for o in offs if o < cut)
    left = _ola(emul, chunks[:k], offs[:k], n, mult=mult[:k], wtab=w, p0=0, p1=cut)
    halo = _ola(emul, chunks[:k], offs[:k], n, mult=mult[:k], wtab=w, p0=cut, p1=n, raw_out=True)[:, cut:].copy()
    right = _ola(emul, chunks[k:], offs, n, mult=mult, wtab=w, p0=cut, p1=n, halo_in=halo, data_chunk0=k)
    stitched = np.concatenate([left[:, :cut], right[:, cut:]], axis=1)
    assert np.array_equal(stitched, got)
    assert not left[:, cut:].any() and not right[:, :cut].any()     # nothing written outside [p0, p1)


def _taps(up, down):
    from audiolab_b200.spectral import resample_taps
    return resample_taps(up, down)


@pytest.mark.parametrize("up,down,n_in,rows,grid", [(147, 160, 48000, 2, 3), (147, 160, 12345, 1, 1), (147, 160, 160, 2, 4),
                                                   (147, 160, 7, 1, 2), (160, 147, 30000, 2, 2), (147, 160, 9001, 3, 7)])
def test_resample_rb_emulated_matches_scipy(emul, up, down, n_in, rows, grid):
    rs = np.random.RandomState(n_in)
    x = rs.uniform(-1, 1, size=(rows, n_in)).astype(np.float32)
    ref = resample_poly_ref(x, up, down)
    taps = _taps(up, down)
    n_out = (n_in * up + down - 1) // down
    assert ref.shape == (rows, n_out)
    # unaligned variant: odd row stride disables the 128-bit paths
    for pad in (0, 1) if n_in < 20000 else (0,):
        in_stride = (n_in + 3) // 4 * 4 + pad
        out_stride = (n_out + 3) // 4 * 4 + pad
        xin = np.zeros((rows, in_stride), np.float32)
        xin[:, :n_in] = x
        xin[:, n_in:] = np.nan                                      # reads past n_in must be masked
        out = np.full((rows, out_stride), -77.0, np.float32)
        rc = emul.emul_resample(_p(xin), in_stride, _p(out), out_stride, rows, n_in, n_out, up, down, _p(taps),
                                taps.size, grid)
        assert rc >= 0, "register-blocked plan rejected this ratio"
        assert rc == (3 if pad == 0 else 0)
        assert np.abs(out[:, :n_out] - ref).max() <= 2e-6
        assert (out[:, n_out:] == -77.0).all()                       # nothing written past n_out


# ---- generic STFT / iSTFT kernels (al_stft.cu / al_istft.cu + al_fft.cuh), emulated with warp barriers + shuffles ------
def _plan_tables(n_fft, hop, T_total=None):
    """The tables al_plan_create / env_kernel build (audiolab_b200/csrc/al_capi.cu:60-120, al_istft.cu env_kernel)."""
    N, D = n_fft, n_fft // 1024
    i = np.arange(N)
    raw = (0.5 - 0.5 * np.cos(2.0 * np.pi * i / N)).astype(np.float32)
    wa = raw.copy()
    ws = (raw.astype(np.float64) / N).astype(np.float32)
    k1, n2 = np.meshgrid(np.arange(32), np.arange(32), indexing="ij")
    a = -2.0 * np.pi * (k1 * n2) / 1024.0
    tw = np.stack((np.cos(a), np.sin(a)), axis=-1).astype(np.float32).reshape(1024, 2)
    r, k = np.meshgrid(np.arange(1, D), np.arange(513), indexing="ij")
    a = -2.0 * np.pi * r * k / N
    ctw = np.stack((np.cos(a), np.sin(a)), axis=-1).astype(np.float32).reshape(-1, 2)
    env = None
    if T_total is not None:
        total = (T_total - 1) * hop + N
        acc = np.zeros(total, np.float64)
        for t in range(T_total):
            acc[t * hop:t * hop + N] += raw.astype(np.float64) ** 2
        env = np.where(acc > 1e-11, 1.0 / np.maximum(acc, 1e-30), 0.0).astype(np.float32)
    return wa, ws, np.ascontiguousarray(tw), np.ascontiguousarray(ctw), env


def _bind_fft(lib):
    P, LL, I = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int
    lib.emul_stft.argtypes = [I, I, P, LL, LL, I, LL, LL, I, I, I, I, P, P, P, P, I, I, I]
    lib.emul_stft.restype = I
    lib.emul_istft.argtypes = [I, I, P, P, I, I, I, I, I, I, I, I, I, P, P, P, P, I, I, P, P, LL, LL, LL, LL, LL]
    lib.emul_istft.restype = I


@pytest.mark.parametrize("n_fft,hop,T,layout", [(2048, 441, 9, 1), (4096, 1024, 8, 2), (6144, 1024, 8, 2), (4096, 1024, 7, 2),
                                               (6144, 1024, 5, 0)])
def test_generic_stft_emulated_matches_torch(emul, n_fft, hop, T, layout):
    """layout 0: c64 [rows, T, F]; 1: c64 [rows, F, T]; 2: CaC planes (T % 4 == 0 takes the 128-bit row stores)."""
    import torch
    _bind_fft(emul)
    L = (T - 1) * hop
    rs = np.random.RandomState(T + n_fft)
    x = rs.uniform(-1, 1, size=(2, L)).astype(np.float32)
    F = n_fft // 2 + 1
    ref = torch.stft(torch.tensor(x), n_fft, hop, window=torch.hann_window(n_fft), center=True, return_complex=True)
    ref = torch.view_as_real(ref).numpy()                                    # [2, F, T, 2]
    wa, _, tw, ctw, _ = _plan_tables(n_fft, hop)
    spec = np.full((2, F, T, 2), np.nan, np.float32).ravel()
    rc = emul.emul_stft(n_fft, hop, _p(x), L, L, 2, 0, 0, 1, L, n_fft // 2, T, _p(wa), _p(tw), _p(ctw), _p(spec), layout, F, 0)
    assert rc == 0
    if layout == 0:
        got = spec.reshape(2, T, F, 2).transpose(0, 2, 1, 3)
    elif layout == 1:
        got = spec.reshape(2, F, T, 2)
    else:
        got = spec.reshape(2, 2, F, T).transpose(0, 2, 3, 1)
    assert np.isfinite(got).all()
    assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()


@pytest.mark.parametrize("n_fft,hop,T,frame_pad,layout", [(4096, 1024, 40, 0, 2), (4096, 1024, 36, 2, 2), (6144, 1024, 36, 0, 2),
                                                         (2048, 441, 21, 0, 1), (6144, 1024, 19, 0, 2), (2048, 512, 40, 1, 2)])
def test_generic_istft_emulated_matches_torch(emul, n_fft, hop, T, frame_pad, layout):
    """Arbitrary (non-consistent) spectrum; several 16-hop segments per row, so the halo recomputation, the re-aligned
    first round of the 128-bit CaC loads (T % 4 == 0) and the in-place carry all run."""
    import torch
    _bind_fft(emul)
    F = n_fft // 2 + 1
    Fo = F - 1 if frame_pad else F
    rs = np.random.RandomState(T)
    S = (rs.standard_normal((2, Fo, T)) + 1j * rs.standard_normal((2, Fo, T))).astype(np.complex64)
    Sfull = np.zeros((2, F, T + 2 * frame_pad), np.complex64)
    Sfull[:, :Fo, frame_pad:frame_pad + T] = S
    ref = torch.istft(torch.tensor(Sfull), n_fft, hop, window=torch.hann_window(n_fft), center=True).numpy()
    out_len = ref.shape[-1]
    _, ws, tw, ctw, env = _plan_tables(n_fft, hop, T + 2 * frame_pad)
    if layout == 2:
        spec = np.ascontiguousarray(np.stack((S.real, S.imag), axis=1)).astype(np.float32)     # [2, (re, im), Fo, T]
    else:
        spec = np.ascontiguousarray(S).view(np.float32)                                        # c64 [2, Fo, T]
    dst = np.full((2, out_len), np.nan, np.float32)
    segs = emul.emul_istft(n_fft, hop, _p(spec), None, layout, Fo, T, frame_pad, 1, 1, 2, 0, 0, _p(ws), _p(tw), _p(ctw), _p(env),
                           n_fft // 2, out_len, None, _p(dst), out_len, 2 * out_len, 0, 0, out_len)
    assert segs >= 1
    assert np.isfinite(dst).all()
    assert np.abs(dst - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("hop,T,stems,use_mask,warps,pre,fast", [(441, 21, 1, True, 4, 2, 1), (441, 13, 2, True, 8, 2, 1),
                                                                (512, 17, 1, False, 4, 3, 0), (1024, 9, 1, True, 4, 0, 1),
                                                                (441, 30, 1, True, 4, 0, 1), (441, 30, 1, True, 4, 0, 0),
                                                                (441, 61, 1, False, 4, 0, 1)])
def test_packed_istft_emulated_matches_torch(emul, hop, T, stems, use_mask, warps, pre, fast):
    """istft_pk2_kernel (stereo, n_fft 2048, frame-interleaved spectrum and mask): fused complex mask, packed f32x2
    inverse FFT, register-form overlap-add with the in-place carry, several segments per chunk (n_sm = 3 forces
    ip_tiling to split), both warps-per-CTA variants, and the cross-round load pipelining depths (pre)."""
    import torch
    P, LL, I = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int
    emul.emul_istft_pk.argtypes = [P, P, I, I, I, I, P, P, P, P, I, I, P, P, LL, LL, LL, LL, LL, I, I, I, I, I]
    emul.emul_istft_pk.restype = I
    n_fft, F, n_chunks = 2048, 1025, 2
    rs = np.random.RandomState(hop + T)
    cplx = lambda *shape: (rs.standard_normal(shape) + 1j * rs.standard_normal(shape)).astype(np.complex64)
    spec = cplx(n_chunks, T, F, 2)                                         # [chunk, t, f, channel]
    mask = cplx(n_chunks, stems, T, F, 2) if use_mask else None
    out_len = (T - 1) * hop
    _, ws, tw, _, env = _plan_tables(n_fft, hop, T)
    k = np.arange(1024)
    ctw_full = np.ascontiguousarray(np.stack((np.cos(-2 * np.pi * k / n_fft), np.sin(-2 * np.pi * k / n_fft)), -1).astype(np.float32))
    weight = rs.uniform(0.5, 1.5, out_len).astype(np.float32)
    dst = np.full((n_chunks, stems, 2, out_len), np.nan, np.float32)
    segs = emul.emul_istft_pk(_p(spec), _p(mask), T, stems, 0, hop, _p(ws), _p(tw), _p(ctw_full), _p(env), n_fft // 2, out_len,
                              _p(weight), _p(dst), out_len, stems * 2 * out_len, 0, 0, out_len, n_chunks, warps, 3 if T < 40 else 1, pre, fast)
    assert segs >= 1
    assert np.isfinite(dst).all()
    win = torch.hann_window(n_fft)
    for c in range(n_chunks):
        for s_ in range(stems):
            y = spec[c] * (mask[c, s_] if use_mask else 1.0)               # [t, f, ch]
            ref = torch.istft(torch.tensor(np.ascontiguousarray(y.transpose(2, 1, 0))), n_fft, hop, window=win, center=True)
            ref = ref.numpy() * weight
            assert np.abs(dst[c, s_] - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("hop,T,stems,use_mask,consumers,n_sm", [(441, 21, 1, True, 4, 3), (441, 13, 2, True, 4, 3),
                                                                     (512, 17, 1, False, 4, 3), (1024, 9, 1, True, 4, 1),
                                                                     (441, 61, 1, True, 4, 1), (300, 33, 1, False, 4, 2),
                                                                     (441, 21, 1, True, 5, 3), (441, 13, 2, True, 5, 3),
                                                                     (512, 17, 1, False, 5, 3), (1024, 9, 1, True, 5, 1),
                                                                     (441, 61, 1, True, 5, 1), (300, 33, 1, False, 5, 2)])
def test_token_ordered_packed_istft_emulated_matches_torch(emul, hop, T, stems, use_mask, consumers, n_sm):
    """istft_pk4_kernel (consumers = 4) and istft_pk5_kernel (5): producer warp + ring of row slots, Z[k] / Z[1024 - k] from one
    product pair, the slot as transposition scratch, overlap-add in a circular even / odd accumulator -- by two dedicated warps
    from the parked frame (pk4) or from the consumers' registers in token order (pk5); odd and even hops, hops above 448 = the
    in-place emission loop, hop 300 = seven frames per position, several segments per chunk, flush blocks past the last frame."""
    import torch
    P, LL, I = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int
    emul.emul_istft_pk4.argtypes = [P, P, I, I, I, I, P, P, P, P, I, I, P, P, LL, LL, LL, LL, LL, I, I, I]
    emul.emul_istft_pk4.restype = I
    n_fft, F, n_chunks = 2048, 1025, 2
    rs = np.random.RandomState(hop + T)
    cplx = lambda *shape: (rs.standard_normal(shape) + 1j * rs.standard_normal(shape)).astype(np.complex64)
    spec = cplx(n_chunks, T, F, 2)                                         # [chunk, t, f, channel]
    mask = cplx(n_chunks, stems, T, F, 2) if use_mask else None
    out_len = (T - 1) * hop
    _, ws, tw, _, env = _plan_tables(n_fft, hop, T)
    k = np.arange(1024)
    ctw_full = np.ascontiguousarray(np.stack((np.cos(-2 * np.pi * k / n_fft), np.sin(-2 * np.pi * k / n_fft)), -1).astype(np.float32))
    weight = rs.uniform(0.5, 1.5, out_len).astype(np.float32)
    dst = np.full((n_chunks, stems, 2, out_len), np.nan, np.float32)
    segs = emul.emul_istft_pk4(_p(spec), _p(mask), T, stems, 0, hop, _p(ws), _p(tw), _p(ctw_full), _p(env), n_fft // 2, out_len,
                               _p(weight), _p(dst), out_len, stems * 2 * out_len, 0, 0, out_len, n_chunks, consumers, n_sm)
    assert segs >= 1
    assert np.isfinite(dst).all()
    win = torch.hann_window(n_fft)
    for c in range(n_chunks):
        for s_ in range(stems):
            y = spec[c] * (mask[c, s_] if use_mask else 1.0)               # [t, f, ch]
            ref = torch.istft(torch.tensor(np.ascontiguousarray(y.transpose(2, 1, 0))), n_fft, hop, window=win, center=True)
            ref = ref.numpy() * weight
            assert np.abs(dst[c, s_] - ref).max() <= 2e-5 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("kernel", [4, 5])
def test_streaming_packed_istft_emulated_placement_weight_and_clipping(emul, kernel):
    """istft_pk4_kernel / istft_pk5_kernel: ragged out_len (not a multiple of the hop), chunk weight, per-chunk placement into a
    track (dst_off0 + chunk * dst_off_step) and the dst_limit clip of the last chunk -- the paths that leave the 'interior'
    hop-block emission -- against torch.istft; nothing is written outside the clipped spans."""
    import torch
    P, LL, I = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int
    emul.emul_istft_pk4.argtypes = [P, P, I, I, I, I, P, P, P, P, I, I, P, P, LL, LL, LL, LL, LL, I, I, I]
    emul.emul_istft_pk4.restype = I
    n_fft, F, hop, T, n_chunks = 2048, 1025, 441, 19, 3
    rs = np.random.RandomState(17)
    cplx = lambda *shape: (rs.standard_normal(shape) + 1j * rs.standard_normal(shape)).astype(np.complex64)
    spec = cplx(n_chunks, T, F, 2)
    mask = cplx(n_chunks, 1, T, F, 2)
    L = (T - 1) * hop - 57                                                 # ragged
    _, ws, tw, _, env = _plan_tables(n_fft, hop, T)
    k = np.arange(1024)
    ctw_full = np.ascontiguousarray(np.stack((np.cos(-2 * np.pi * k / n_fft), np.sin(-2 * np.pi * k / n_fft)), -1).astype(np.float32))
    weight = rs.uniform(0.5, 1.5, L).astype(np.float32)
    off0, off_step = 100, L + 37
    n_track = off0 + 2 * off_step + L - 300                                # the last chunk is clipped by dst_limit
    track = np.full((2, n_track + 64), -77.0, np.float32)                  # guard band past dst_limit
    segs = emul.emul_istft_pk4(_p(spec), _p(mask), T, 1, 0, hop, _p(ws), _p(tw), _p(ctw_full), _p(env), n_fft // 2, L,
                               _p(weight), _p(track), n_track + 64, 0, off0, off_step, n_track, n_chunks, kernel, 2)
    assert segs >= 1
    win = torch.hann_window(n_fft)
    exp = np.full_like(track, -77.0)
    for c in range(n_chunks):
        y = spec[c] * mask[c, 0]
        ref = torch.istft(torch.tensor(np.ascontiguousarray(y.transpose(2, 1, 0))), n_fft, hop, window=win, center=True,
                          length=L).numpy() * weight
        a = off0 + c * off_step
        b = min(a + L, n_track)
        exp[:, a:b] = ref[:, : b - a]
    written = exp != -77.0
    assert np.array_equal(track != -77.0, written)
    assert np.abs(track[written] - exp[written]).max() <= 2e-5 * max(1.0, np.abs(exp[written]).max())


@pytest.mark.parametrize("hop,T,layout,crop,low,n_chunks", [(441, 21, 3, 1025, 0, 2), (441, 9, 0, 1025, 0, 1), (512, 12, 3, 1000, 3, 2),
                                                           (441, 26, 3, 1025, 0, 3)])
def test_packed_stft_emulated_matches_torch(emul, hop, T, layout, crop, low, n_chunks):
    """stft_pk2_kernel (stereo, n_fft 2048): producer / consumer stage ring on mbarriers, packed f32x2 FFT, lane-mirror
    separation, bulk row stores; chunks cut out of one track (reflect padding at chunk ends, zero beyond the track),
    a persistent grid smaller than the number of tiles, cropped / zeroed bins."""
    import torch
    P, LL, I = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int
    emul.emul_stft_pk.argtypes = [P, LL, LL, LL, LL, I, I, I, I, I, P, P, P, P, I, I, I, I, I]
    emul.emul_stft_pk.restype = I
    n_fft, F = 2048, 1025
    L = (T - 1) * hop
    n = n_chunks * L - 100                                              # the last chunk runs 100 samples past the track
    rs = np.random.RandomState(T)
    x = np.zeros((2, n_chunks * L + 8), np.float32)
    x[:, :n] = rs.uniform(-1, 1, size=(2, n))
    x[:, n:] = np.nan                                                   # beyond n_valid: must never be read as data
    wa, _, tw, _, _ = _plan_tables(n_fft, hop)
    half = np.zeros((544, 2), np.float32)
    k = np.arange(513)
    half[:513, 0], half[:513, 1] = 0.5 * np.cos(-2 * np.pi * k / n_fft), 0.5 * np.sin(-2 * np.pi * k / n_fft)
    spec = np.full(n_chunks * 2 * T * crop * 2, np.nan, np.float32)
    ns = emul.emul_stft_pk(_p(x), n, x.shape[1], 0, L, n_chunks, L, n_fft // 2, hop, T, _p(wa), _p(tw), _p(half), _p(spec),
                           layout, crop, low, 1, 2)
    assert ns >= 1
    assert np.isfinite(spec).all()
    xz = np.nan_to_num(x[:, :n_chunks * L], nan=0.0)
    for c in range(n_chunks):
        ref = torch.stft(torch.tensor(xz[:, c * L:(c + 1) * L]), n_fft, hop, window=torch.hann_window(n_fft), center=True,
                         return_complex=True)
        ref = torch.view_as_real(ref).numpy()[:, :crop].copy()          # [2, crop, T, 2]
        ref[:, :low] = 0
        if layout == 3:
            got = spec.reshape(n_chunks, T, crop, 2, 2)[c].transpose(2, 1, 0, 3)        # [t, f, ch, ri] -> [ch, f, t, ri]
        else:
            got = spec.reshape(n_chunks, 2, T, crop, 2)[c].transpose(0, 2, 1, 3)        # [ch, t, f, ri] -> [ch, f, t, ri]
        assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()


def test_generic_istft_emulated_mask_stems_weight_and_clipped_placement(emul):
    """Generic al_istft, c64 [rows, T, F] layout: fused complex mask over two stems of one shared spectrum, zeroed low
    bins, chunk weight, out_start past the centre trim, chunks placed every `gen` samples into one track and clipped at
    dst_limit (the MDX trim-and-concat form, mdxnet.py:178-183)."""
    import torch
    _bind_fft(emul)
    n_fft, hop, T, stems, n_chunks, low = 4096, 1024, 24, 2, 2, 3
    F = n_fft // 2 + 1
    rs = np.random.RandomState(11)
    cplx = lambda *shape: (rs.standard_normal(shape) + 1j * rs.standard_normal(shape)).astype(np.complex64)
    spec = cplx(n_chunks, 2, T, F)                                        # [chunk, ch, t, f]   (layout 0)
    mask = cplx(n_chunks, stems, 2, T, F)                                 # [chunk, stem, ch, t, f]
    trim = 700
    gen = (T - 1) * hop - 2 * trim
    n_total = 2 * gen - 333                                               # the second chunk is clipped by dst_limit
    weight = rs.uniform(0.5, 1.5, gen).astype(np.float32)
    _, ws, tw, ctw, env = _plan_tables(n_fft, hop, T)
    dst = np.full((stems * 2, n_total), np.nan, np.float32)
    segs = emul.emul_istft(n_fft, hop, _p(spec), _p(mask), 0, F, T, 0, n_chunks, stems, 2, 0, low, _p(ws), _p(tw), _p(ctw), _p(env),
                           n_fft // 2 + trim, gen, _p(weight), _p(dst), n_total, 0, 0, gen, n_total)
    assert segs >= 1
    assert np.isfinite(dst).all()
    win = torch.hann_window(n_fft)
    for c in range(n_chunks):
        for s_ in range(stems):
            y = spec[c] * mask[c, s_]                                     # [ch, t, f]
            y[:, :, :low] = 0
            ref = torch.istft(torch.tensor(np.ascontiguousarray(y.transpose(0, 2, 1))), n_fft, hop, window=win, center=True).numpy()
            ref = ref[:, trim:trim + gen] * weight
            lo, hi = c * gen, min((c + 1) * gen, n_total)
            got = dst[s_ * 2:s_ * 2 + 2, lo:hi]
            assert np.abs(got - ref[:, :hi - lo]).max() <= 2e-5 * max(1.0, np.abs(ref).max())


def test_generic_stft_emulated_virtual_padding_and_crop(emul):
    """Generic al_stft on chunks cut out of one track with a negative first offset and a tail past the track (zeros
    outside [0, n_valid), reflect padding about the CHUNK ends like torch.stft on the materialised chunk), cropped
    frequency rows and zeroed low bins, CaC output (mdxnet.py:41-56, :152-164)."""
    import torch
    _bind_fft(emul)
    n_fft, hop, T, n_chunks, crop, low = 6144, 1024, 8, 3, 3072, 2
    L = (T - 1) * hop
    n = 2 * L + 500
    rs = np.random.RandomState(5)
    x = rs.uniform(-1, 1, size=(2, n)).astype(np.float32)
    off0, step = -3072, L - 1024
    wa, _, tw, ctw, _ = _plan_tables(n_fft, hop)
    spec = np.full((n_chunks, 4, crop, T), np.nan, np.float32)
    rc = emul.emul_stft(n_fft, hop, _p(x), n, n, 2, off0, step, n_chunks, L, n_fft // 2, T, _p(wa), _p(tw), _p(ctw), _p(spec), 2,
                        crop, low)
    assert rc == 0 and np.isfinite(spec).all()
    padded = np.zeros((2, n + 2 * L + 4096), np.float32)
    padded[:, 3072:3072 + n] = x                                          # track position p lives at padded[p + 3072]
    for c in range(n_chunks):
        a = off0 + c * step + 3072
        chunk = padded[:, a:a + L]
        ref = torch.stft(torch.tensor(chunk), n_fft, hop, window=torch.hann_window(n_fft), center=True, return_complex=True)
        ref = torch.view_as_real(ref).numpy()[:, :crop].copy()           # [2, crop, T, 2]
        ref[:, :low] = 0
        got = spec[c].reshape(2, 2, crop, T).transpose(0, 2, 3, 1)
        assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max()


@pytest.mark.parametrize("tag", ["a", "b"])
def test_emulated_kernels_match_reference_golden_stft_istft(emul, tag):
    """The kernels' source (host-emulated) against outputs of the REFERENCE's own ConvTDFNetTrim.stft / .istft
    (/root/reference/modules/rvc/infer/modules/uvr5/mdxnet.py:41-75, vectors from tests/golden/make_mdx_golden.py):
    pins al_stft / al_istft to the reference on a box without a GPU (the GPU twin is
    tests/test_demix_gpu.py::test_mdx_demix_matches_reference_golden)."""
    from oracle.synth import synth_mix
    _bind_fft(emul)
    g = np.load(os.path.join(ROOT, "tests", "golden", "mdx_stft.npz"))
    n_fft, dim_f, dim_t_log2, seed = (int(v) for v in g[f"{tag}_cfg"])
    hop, T = 1024, 2 ** dim_t_log2
    chunk = hop * (T - 1)
    track = np.ascontiguousarray(synth_mix(2 * chunk, seed=seed))          # [2 ch, 2 chunks back to back]
    wa, ws, tw, ctw, env = _plan_tables(n_fft, hop, T)
    spec = np.full((2, 4, dim_f, T), np.nan, np.float32)
    rc = emul.emul_stft(n_fft, hop, _p(track), 2 * chunk, 2 * chunk, 2, 0, chunk, 2, chunk, n_fft // 2, T, _p(wa), _p(tw), _p(ctw),
                        _p(spec), 2, dim_f, 0)
    assert rc == 0 and np.isfinite(spec).all()
    ref_sub = g[f"{tag}_spek_sub"]
    assert np.abs(spec[:, :, ::7, :] - ref_sub).max() <= 2e-6 * np.abs(ref_sub).max()
    ref_sum = g[f"{tag}_spek_sum"]
    assert abs(float(spec.astype(np.float64).sum()) - ref_sum[0]) <= 1e-5 * ref_sum[1]
    back = np.full((2, 2, chunk), np.nan, np.float32)
    segs = emul.emul_istft(n_fft, hop, _p(spec), None, 2, dim_f, T, 0, 2, 1, 2, 0, 0, _p(ws), _p(tw), _p(ctw), _p(env), n_fft // 2,
                           chunk, None, _p(back), chunk, 2 * chunk, 0, 0, chunk)
    assert segs >= 1 and np.isfinite(back).all()
    ref_back = g[f"{tag}_istft"]
    assert np.abs(back - ref_back).max() <= 1e-4 * max(1.0, np.abs(ref_back).max())


@pytest.mark.parametrize("script,n,seed", [("fuzz_istft.py", 8, 5), ("fuzz_stft.py", 8, 5), ("fuzz_misc.py", 12, 5)])
def test_randomised_emulation_subset(script, n, seed):
    """A fixed-seed slice of the randomised emulation fuzzers (tools/cpu_emul/fuzz_*.py; longer runs: `--n 120`)."""
    import subprocess
    import sys
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "cpu_emul", script), "--n", str(n), "--seed", str(seed)],
                         capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-2000:]


# ---- row-wise bf16 operators of the RoFormer inference path (al_netops.cu), emulated ------------------------------------
def _bf16(t):
    """torch bf16 tensor -> numpy uint16 view (and back with _from_bf16)."""
    import torch
    return t.contiguous().view(torch.int16).numpy().view(np.uint16).copy()


def _from_bf16(a, shape):
    import torch
    return torch.from_numpy(a.view(np.int16).copy()).view(torch.bfloat16).reshape(shape)


def test_netops_emulated_match_their_torch_definitions(emul):
    """rmsnorm (+ deferred bias), rotary, sigmoid gate and the table-driven GELU against the PyTorch definitions used by
    tests/test_netops.py (upstream RMSNorm / rotary_embed / gated attention / nn.GELU)."""
    import torch
    import torch.nn.functional as F
    spec_ = importlib.util.spec_from_file_location("tnet", os.path.join(ROOT, "tests", "test_netops.py"))
    tnet = importlib.util.module_from_spec(spec_)
    spec_.loader.exec_module(tnet)
    ref_gate_, ref_rmsnorm, ref_rotary_ = tnet.ref_gate_, tnet.ref_rmsnorm, tnet.ref_rotary_
    P, LL, I, FL = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_float
    emul.emul_rmsnorm.argtypes = [P, P, P, P, LL, I, FL, FL]
    emul.emul_rotary.argtypes = [P, P, P, LL, I, I, LL, I]
    emul.emul_gate.argtypes = [P, P, LL, I, I]
    emul.emul_gelu.argtypes = [P, LL, I]
    for f in (emul.emul_rmsnorm, emul.emul_rotary, emul.emul_gate, emul.emul_gelu):
        f.restype = None
    g = torch.Generator().manual_seed(3)
    # rmsnorm, with and without the deferred bias, dims of the three register tilings
    for n, d, with_bias in [(37, 512, True), (19, 136, False), (9, 1024, True), (5, 2048, False)]:
        x = torch.randn(n, d, generator=g).to(torch.bfloat16)
        gamma = (1 + 0.1 * torch.randn(d, generator=g)).float()
        bias = (0.1 * torch.randn(d, generator=g)).float() if with_bias else None
        xr = x.clone()
        ref = ref_rmsnorm(xr, gamma, bias)
        xa, out = _bf16(x), np.zeros(n * d, np.uint16)
        ga, ba = gamma.numpy().copy(), (None if bias is None else bias.numpy().copy())
        emul.emul_rmsnorm(_p(xa), _p(ga), _p(ba), _p(out), n, d, float(d) ** 0.5, 1e-12)
        got = _from_bf16(out, (n, d)).float()
        assert float((got - ref.float()).abs().max()) <= 2 ** -7 * float(ref.float().abs().max())
        if with_bias:                                         # x += bias is stored back (the residual stream)
            assert torch.equal(_from_bf16(xa, (n, d)), xr)
    # rotary over a [batch, time, band] token grid
    heads, dh, pos_div, pos_mod, n = 4, 32, 3, 7, 3 * 7 * 2
    q = torch.randn(n, heads * dh, generator=g).to(torch.bfloat16)
    k = torch.randn(n, heads * dh, generator=g).to(torch.bfloat16)
    ang = torch.arange(pos_mod)[:, None].float() * (1.0 / (10000 ** (torch.arange(0, dh, 2).float() / dh)))[None]
    cs = torch.stack((ang.cos(), ang.sin()), dim=-1).contiguous()
    qr, kr = q.clone(), k.clone()
    ref_rotary_(qr, kr, cs, heads, dh, pos_div, pos_mod)
    qa, ka, csa = _bf16(q), _bf16(k), cs.numpy().copy()
    emul.emul_rotary(_p(qa), _p(ka), _p(csa), n, heads, dh, pos_div, pos_mod)
    for got, ref in ((_from_bf16(qa, q.shape), qr), (_from_bf16(ka, k.shape), kr)):
        assert float((got.float() - ref.float()).abs().max()) <= 2 ** -7 * float(ref.float().abs().max())
    # gate
    o = torch.randn(33, heads * dh, generator=g).to(torch.bfloat16)
    gates = (2 * torch.randn(33, heads, generator=g)).to(torch.bfloat16)
    orf = o.clone()
    ref_gate_(orf, gates, heads, dh)
    oa, gta = _bf16(o), _bf16(gates)
    emul.emul_gate(_p(oa), _p(gta), 33, heads, dh)
    assert float((_from_bf16(oa, o.shape).float() - orf.float()).abs().max()) <= 2 ** -7 * float(orf.float().abs().max())
    # GELU: the table must reproduce torch's bf16 gelu for every bf16 bit pattern
    bits = np.arange(65536, dtype=np.uint16)
    xall = _from_bf16(bits, (65536,))
    ref = F.gelu(xall.float()).to(torch.bfloat16)
    xa = bits.copy()
    emul.emul_gelu(_p(xa), 65536, 3)
    got = _from_bf16(xa, (65536,))
    # torch's CPU kernel evaluates x * (1 + erf) before the 0.5 and overflows above 1.7e38; below that the table is
    # identical to torch's bf16 gelu for all but a few dozen inputs
    ok = torch.isfinite(xall.float()) & (xall.float().abs() < 1e38)
    # identical bits except where 1 + erf cancels (x in [-6, -3], |gelu| < 3e-3): there libm's erff (host emulation) and
    # torch's differ in the last bits, which the cancellation amplifies -- compare those few absolutely
    ne = got.view(torch.int16)[ok] != ref.view(torch.int16)[ok]
    assert int(ne.sum()) <= 64
    assert float((got.float()[ok] - ref.float()[ok]).abs().max()) <= 2e-5


def test_invert_stem_composition_on_emulated_kernels(emul):
    """Row a14 end to end on the host: the exact kernel calls MdxDemixer.invert_stem makes (packed al_stft with a chunk that
    starts n_fft/2 before the track and no reflection = librosa's zero centre padding; generic al_istft, frame-major)
    around the in-tree "invert_p" arithmetic, against the oracle's librosa-convention restatement."""
    import torch
    from oracle import mdx as omdx
    from oracle.synth import synth_mix
    _bind_fft(emul)
    P, LL, I = ctypes.c_void_p, ctypes.c_longlong, ctypes.c_int
    emul.emul_stft_pk.argtypes = [P, LL, LL, LL, LL, I, I, I, I, I, P, P, P, P, I, I, I, I, I]
    emul.emul_stft_pk.restype = I
    n_fft, hop = 2048, 1024
    wa, _, tw, _, _ = _plan_tables(n_fft, hop)
    half = np.zeros((544, 2), np.float32)
    k = np.arange(513)
    half[:513, 0], half[:513, 1] = 0.5 * np.cos(-2 * np.pi * k / n_fft), 0.5 * np.sin(-2 * np.pi * k / n_fft)

    def stft_dev(x):
        n = x.shape[1]
        T = 1 + n // hop
        x = np.ascontiguousarray(x)
        out = np.full(2 * T * 1025 * 2, np.nan, np.float32)
        rc = emul.emul_stft_pk(_p(x), n, n, -(n_fft // 2), 0, 1, n + n_fft, 0, hop, T, _p(wa), _p(tw), _p(half), _p(out), 0, 1025, 0,
                               0, 3)
        assert rc >= 1 and np.isfinite(out).all()
        return torch.view_as_complex(torch.tensor(out.reshape(2, T, 1025, 2)))

    for n in (5000, 1024 * 9):
        mix = synth_mix(n, seed=n)
        stem = (0.6 * mix[::-1] + 0.1 * synth_mix(n, seed=n + 1)).astype(np.float32)
        X, y = stft_dev(mix), stft_dev(stem)
        xm, ym = X.abs(), y.abs()
        unit = torch.where(xm > 0, X / xm.clamp(min=1e-30), torch.ones_like(X))
        v = torch.view_as_real((y - torch.where(xm >= ym, xm, ym) * unit).contiguous()).numpy()
        T = v.shape[1]
        out_len = (T - 1) * hop
        _, ws, tw2, ctw2, env = _plan_tables(n_fft, hop, T)
        wave = np.full((2, out_len), np.nan, np.float32)
        segs = emul.emul_istft(n_fft, hop, _p(np.ascontiguousarray(v)), None, 0, 1025, T, 0, 1, 1, 2, 0, 0, _p(ws), _p(tw2), _p(ctw2),
                               _p(env), n_fft // 2, out_len, None, _p(wave), out_len, 2 * out_len, 0, 0, out_len)
        assert segs >= 1 and np.isfinite(wave).all()
        ref = omdx.invert_stem(mix, stem)
        assert ref.shape[1] == out_len
        assert np.abs(-wave - ref).max() <= 2e-5


@pytest.mark.parametrize("F,H,n_seq,gated", [(62, 8, 3, True), (64, 2, 2, False), (17, 4, 2, True)])
def test_band_attention_emulated_matches_torch_sdpa(emul, F, H, n_seq, gated):
    """Opt-in band-axis attention kernel (al_attn.cu, mma.sync m16n8k16 emulated as a warp collective from the PTX fragment
    layouts): softmax(Q K^T / 8) V per (sequence, head) on token-major [n_seq * F, H * 64] bf16 buffers, against
    torch's scaled_dot_product_attention on the same bf16 inputs."""
    import torch
    import torch.nn.functional as Fn
    emul.emul_band_attn.argtypes = [ctypes.c_void_p] * 6 + [ctypes.c_int] * 3 + [ctypes.c_float]
    emul.emul_band_attn.restype = None
    g = torch.Generator().manual_seed(F + H)
    q, k, v = (torch.randn(n_seq * F, H * 64, generator=g).to(torch.bfloat16) for _ in range(3))
    qa, ka, va = _bf16(q), _bf16(k), _bf16(v)
    oa = np.full(q.numel(), 0x7FC0, np.uint16)                        # NaN: every output element must be written
    gates = (2 * torch.randn(n_seq * F, H, generator=g)).to(torch.bfloat16) if gated else None
    ga = None if gates is None else _bf16(gates)
    cs = None
    if gated:                                                         # the gated cases also fold the rotary embedding in
        ang = torch.arange(F)[:, None].float() * (1.0 / (10000 ** (torch.arange(0, 64, 2).float() / 64)))[None]
        cs = torch.stack((ang.cos(), ang.sin()), dim=-1).contiguous()
        spec_ = importlib.util.spec_from_file_location("tnet", os.path.join(ROOT, "tests", "test_netops.py"))
        tnet = importlib.util.module_from_spec(spec_)
        spec_.loader.exec_module(tnet)
        tnet.ref_rotary_(q, k, cs, H, 64, 1, F)                       # reference: rotate (and round to bf16) first
    emul.emul_band_attn(_p(qa), _p(ka), _p(va), _p(oa), _p(ga), _p(None if cs is None else cs.numpy().copy()), n_seq, F, H, 0.125)
    got = _from_bf16(oa, q.shape).float()
    shp = (n_seq, F, H, 64)
    ref = Fn.scaled_dot_product_attention(q.float().view(shp).transpose(1, 2), k.float().view(shp).transpose(1, 2),
                                          v.float().view(shp).transpose(1, 2)).transpose(1, 2).reshape(q.shape)
    if gated:
        ref = (ref.view(n_seq * F, H, 64) * torch.sigmoid(gates.float())[:, :, None]).reshape(q.shape)
    assert torch.isfinite(got).all()
    assert float((got - ref).abs().max()) <= 2 ** -6 * float(ref.abs().max())      # bf16 probabilities and bf16 output
    assert float((got - ref).abs().mean()) <= 2e-3 * float(ref.abs().mean() + 1e-6) + 2e-3
