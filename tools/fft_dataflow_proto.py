"""numpy prototype of the CUDA kernels' FFT data flow (development aid, not product code).

Mirrors audiolab_b200/csrc/al_fft.cuh: a frame of N = D*1024 real samples (D in {2,4,6}) is
split into D decimated real sequences x_r[n] = x[D n + r]; pairs are packed into D/2 complex
1024-point FFTs ("units", one warp each: lane = low digit, register = high digit);
unit spectra are unpacked with the lane-mirror trick and combined with a radix-D butterfly.
"""
import numpy as np


def warp_fft1024(v, inverse=False):
    """v[reg, lane] = z[32*reg + lane]  ->  out[reg, lane] = Z[32*reg + lane]."""
    sgn = +1.0 if inverse else -1.0
    # step 1: 32-point DFT over the register index (n1 -> k1), per lane n2
    n1 = np.arange(32)
    W32 = np.exp(sgn * 2j * np.pi * np.outer(n1, n1) / 32)
    A = W32 @ v                                  # A[k1, n2]
    # step 2: twiddle W_1024^(n2*k1)
    A = A * np.exp(sgn * 2j * np.pi * np.outer(n1, n1) / 1024)   # [k1, n2]
    # step 3: transpose through shared memory: lane k1 gets B[n2] = A[k1, n2]
    B = A.T.copy()                               # B[n2(reg), k1(lane)]
    # step 4: 32-point DFT over the register index (n2 -> k2), per lane k1
    return W32 @ B                               # out[k2, k1] = Z[k1 + 32 k2]


def mirror(Z):
    """P[reg, lane] = Z[(1024 - k) % 1024] for k = 32 reg + lane, via lane (32-l)&31 / reg 31-r."""
    P = np.empty_like(Z)
    for r in range(32):
        src = Z[31 - r]
        P[r] = src[(32 - np.arange(32)) & 31]
        P[r, 0] = Z[(32 - r) & 31, 0]            # lane 0 special case
    return P


def rfft_units(x, D):
    """Forward: real frame x[N] -> X_r[kappa], r < D, kappa in [0, 1024) (full, Hermitian)."""
    Xr = np.empty((D, 1024), dtype=complex)
    for w in range(D // 2):
        z = x[2 * w::D][:1024] + 1j * x[2 * w + 1::D][:1024]
        v = z.reshape(32, 32)                    # v[reg, lane] = z[32 reg + lane]
        Z = warp_fft1024(v)
        P = np.conj(mirror(Z))
        Xr[2 * w] = (0.5 * (Z + P)).reshape(-1)
        Xr[2 * w + 1] = (-0.5j * (Z - P)).reshape(-1)
    return Xr


def combine_fwd(Xr, D):
    """radix-D butterfly: X[kappa + 1024 q] for kappa <= 512, all q < D -> bins 0..N/2."""
    N = D * 1024
    X = np.zeros(N // 2 + 1, dtype=complex)
    for kappa in range(513):
        Y = np.array([np.exp(-2j * np.pi * r * kappa / N) * Xr[r, kappa] for r in range(D)])
        for q in range(D):
            val = sum(np.exp(-2j * np.pi * r * q / D) * Y[r] for r in range(D))
            k = kappa + 1024 * q
            if k <= N // 2:
                X[k] = val
            else:
                X[N - k] = np.conj(val)
    return X


def combine_inv(X, D):
    """inverse radix-D: X[0..N/2] -> X_r[kappa], kappa <= 512 (1/D folded in)."""
    N = D * 1024
    X = X.copy()
    X[0] = X[0].real
    X[N // 2] = X[N // 2].real
    Xr = np.zeros((D, 513), dtype=complex)
    for kappa in range(513):
        V = []
        for q in range(D):
            k = kappa + 1024 * q
            V.append(X[k] if k <= N // 2 else np.conj(X[N - k]))
        for r in range(D):
            s = sum(np.exp(2j * np.pi * r * q / D) * V[q] for q in range(D))
            Xr[r, kappa] = np.exp(2j * np.pi * r * kappa / N) * s / D
    return Xr


def irfft_units(Xr_half, D):
    """X_r[kappa<=512] -> real frame x[N] (1/1024 applied here)."""
    N = D * 1024
    x = np.zeros(N)
    kap = np.arange(1024)
    for w in range(D // 2):
        def full(r):
            h = Xr_half[r]
            return np.where(kap <= 512, h[np.minimum(kap, 512)], np.conj(h[np.minimum(1024 - kap, 512)]))
        Z = full(2 * w) + 1j * full(2 * w + 1)
        z = warp_fft1024(Z.reshape(32, 32), inverse=True).reshape(-1) / 1024
        x[2 * w::D] = z.real
        x[2 * w + 1::D] = z.imag
    return x


if __name__ == "__main__":
    rs = np.random.RandomState(0)
    z = rs.randn(1024) + 1j * rs.randn(1024)
    assert np.allclose(warp_fft1024(z.reshape(32, 32)).reshape(-1), np.fft.fft(z))
    assert np.allclose(warp_fft1024(z.reshape(32, 32), True).reshape(-1), np.fft.ifft(z) * 1024)
    for D in (2, 4, 6):
        N = D * 1024
        x = rs.randn(N)
        Xr = rfft_units(x, D)
        for r in range(D):
            assert np.allclose(Xr[r], np.fft.fft(x[r::D])), (D, r)
        X = combine_fwd(Xr, D)
        assert np.allclose(X, np.fft.rfft(x)), D
        Xh = combine_inv(X, D)
        assert np.allclose(Xh, Xr[:, :513] / 1.0 * 1.0 / 1.0 * (1.0) / 1.0 * 1.0 / 1.0 if False else Xh)
        assert np.allclose(Xh, np.array([np.fft.fft(x[r::D])[:513] for r in range(D)]) * 1.0), D
        xb = irfft_units(Xh, D)
        assert np.allclose(xb, x), D
        print("D", D, "ok")
