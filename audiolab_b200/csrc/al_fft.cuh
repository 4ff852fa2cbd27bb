// Warp-level FFT building blocks for the STFT / iSTFT kernels (sm_100a).
//
// A frame of N = D*1024 real samples (D = 2, 4, 6 for n_fft 2048 / 4096 / 6144) is split into D
// decimated real sequences x_r[n] = x[D n + r].  Pairs (x_2w, x_2w+1) are packed into D/2 complex
// 1024-point FFTs ("units").  One warp owns one unit: lane = low digit, register = high digit of
// the element index (element 32*r + lane lives in register r of `lane`), so the unit FFT is two
// in-register radix-32 butterflies around ONE shared-memory transposition.  Unit spectra are
// separated with a lane-mirror shuffle and merged by a radix-D butterfly (decimation in time).
// tools/fft_dataflow_proto.py is the numpy model of exactly this data flow.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fft32_gen.cuh"

namespace al {

constexpr int kSlotF2 = 1058;   // float2 per unit slot: 33*32 transposition scratch, +2 so that
                                // consecutive slots land on different banks for frame-fastest access
constexpr int kXHalf = 528;     // float2 offset of X_{2w+1} inside a slot (X_{2w} at 0), 513 used
constexpr int kTrStride = 33;   // float2 row stride of the transposition scratch

__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cmul_conj(float2 a, float2 b) {  // a * conj(b)
    return make_float2(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}

// in : re/im[r] = z[32*r + lane]
// out: re/im[r] = Z[32*r + lane],  Z[k] = sum_n z[n] exp(-+2 pi i n k / 1024)   (unnormalised)
// tw : shared, tw[k1*32 + n2] = exp(-2 pi i k1 n2 / 1024); scratch: this warp's slot.
template <bool INV>
__device__ __forceinline__ void warp_fft1024(float (&re)[32], float (&im)[32], float2* scratch,
                                             const float2* tw, int lane) {
    if (INV) fft32_inv(re, im); else fft32_fwd(re, im);
    scratch[lane * kTrStride] = make_float2(re[0], im[0]);
#pragma unroll
    for (int k1 = 1; k1 < 32; ++k1) {
        const float2 w = tw[k1 * 32 + lane];
        const float s = INV ? -w.y : w.y;
        scratch[lane * kTrStride + k1] = make_float2(re[k1] * w.x - im[k1] * s, re[k1] * s + im[k1] * w.x);
    }
    __syncwarp();
#pragma unroll
    for (int n2 = 0; n2 < 32; ++n2) {
        const float2 t = scratch[n2 * kTrStride + lane];
        re[n2] = t.x;
        im[n2] = t.y;
    }
    __syncwarp();
    if (INV) fft32_inv(re, im); else fft32_fwd(re, im);
}

// ---- small DFTs over the decimation index: y[q] = sum_r y[r] exp(-+2 pi i r q / D) ----------------
template <int D, bool INV> struct SmallDft;

template <bool INV> struct SmallDft<2, INV> {
    __device__ __forceinline__ static void run(float2 (&y)[2]) {
        const float2 a = y[0], b = y[1];
        y[0] = make_float2(a.x + b.x, a.y + b.y);
        y[1] = make_float2(a.x - b.x, a.y - b.y);
    }
};

template <bool INV> struct SmallDft<4, INV> {
    __device__ __forceinline__ static void run(float2 (&y)[4]) {
        const float2 s02 = make_float2(y[0].x + y[2].x, y[0].y + y[2].y);
        const float2 d02 = make_float2(y[0].x - y[2].x, y[0].y - y[2].y);
        const float2 s13 = make_float2(y[1].x + y[3].x, y[1].y + y[3].y);
        const float2 d13 = make_float2(y[1].x - y[3].x, y[1].y - y[3].y);
        // forward: -i*d13 = (d13.y, -d13.x); inverse: +i*d13 = (-d13.y, d13.x)
        const float2 j13 = INV ? make_float2(-d13.y, d13.x) : make_float2(d13.y, -d13.x);
        y[0] = make_float2(s02.x + s13.x, s02.y + s13.y);
        y[1] = make_float2(d02.x + j13.x, d02.y + j13.y);
        y[2] = make_float2(s02.x - s13.x, s02.y - s13.y);
        y[3] = make_float2(d02.x - j13.x, d02.y - j13.y);
    }
};

template <bool INV>
__device__ __forceinline__ void dft3(float2 a, float2 b, float2 c, float2& o0, float2& o1, float2& o2) {
    // o_q = a + b w^q + c w^{2q},  w = exp(-+2 pi i / 3) = (-1/2, -+sqrt(3)/2)
    const float kS = INV ? 0.866025404f : -0.866025404f;
    const float2 s = make_float2(b.x + c.x, b.y + c.y);
    const float2 d = make_float2(b.x - c.x, b.y - c.y);
    o0 = make_float2(a.x + s.x, a.y + s.y);
    const float2 m = make_float2(a.x - 0.5f * s.x, a.y - 0.5f * s.y);
    // i*kS*d = (-kS*d.y, kS*d.x)
    const float2 r = make_float2(-kS * d.y, kS * d.x);
    o1 = make_float2(m.x + r.x, m.y + r.y);
    o2 = make_float2(m.x - r.x, m.y - r.y);
}

template <bool INV> struct SmallDft<6, INV> {
    __device__ __forceinline__ static void run(float2 (&y)[6]) {
        // Cooley-Tukey 6 = 3 x 2: r = 2 n1 + n2, q = q1 + 3 q2
        float2 a0, a1, a2, b0, b1, b2;
        dft3<INV>(y[0], y[2], y[4], a0, a1, a2);
        dft3<INV>(y[1], y[3], y[5], b0, b1, b2);
        // twiddle w6^q1, w6 = exp(-+2 pi i / 6) = (1/2, -+sqrt(3)/2)
        const float kS = INV ? 0.866025404f : -0.866025404f;
        const float2 t1 = make_float2(0.5f * b1.x - kS * b1.y, 0.5f * b1.y + kS * b1.x);
        const float2 t2 = make_float2(-0.5f * b2.x - kS * b2.y, -0.5f * b2.y + kS * b2.x);
        y[0] = make_float2(a0.x + b0.x, a0.y + b0.y);
        y[3] = make_float2(a0.x - b0.x, a0.y - b0.y);
        y[1] = make_float2(a1.x + t1.x, a1.y + t1.y);
        y[4] = make_float2(a1.x - t1.x, a1.y - t1.y);
        y[2] = make_float2(a2.x + t2.x, a2.y + t2.y);
        y[5] = make_float2(a2.x - t2.x, a2.y - t2.y);
    }
};

// per-D launch shape of the generic kernels: G frames per round, UW = G*D/2 unit warps per CTA
template <int D> struct Cfg;
template <> struct Cfg<2> { static constexpr int G = 8, UW = 8; };
template <> struct Cfg<4> { static constexpr int G = 8, UW = 16; };   // 8 frames = one 32-byte sector of a T-innermost row
template <> struct Cfg<6> { static constexpr int G = 4, UW = 12; };

// float2 slots reserved in shared memory for the radix-D combine twiddles [(D-1)][513] (a multiple of 32 keeps the unit
// slots that follow bank-aligned)
template <int D> struct CtwPad { static constexpr int value = ((D - 1) * 513 + 31) / 32 * 32; };

// ---- spectrogram addressing --------------------------------------------------------------------
// row = group*channels + ch  (group = chunk, or chunk*stems + stem)
// layout 0: c64 [rows, T, Fo]; 1: c64 [rows, Fo, T]; 2: f32 [rows*2 (re, im planes), Fo, T];
// layout 3: c64 [groups, T, Fo, channels]  (upstream RoFormer 'b t (f s c)')
struct SpecView {
    float* base;
    int layout;
    int T;
    int Fo;
    int channels;
};

__device__ __forceinline__ int64_t spec_index(const SpecView& v, int64_t row, int t, int bin) {
    if (v.layout == 0) return (row * v.T + t) * v.Fo + bin;
    if (v.layout == 1) return (row * v.Fo + bin) * (int64_t)v.T + t;
    if (v.layout == 3) {
        const int64_t g = row / v.channels;
        const int ch = (int)(row - g * v.channels);
        return ((g * v.T + t) * v.Fo + bin) * v.channels + ch;
    }
    return ((row * 2) * v.Fo + bin) * (int64_t)v.T + t;
}

__device__ __forceinline__ void spec_store(const SpecView& v, int64_t row, int t, int bin, float2 val) {
    const int64_t o = spec_index(v, row, t, bin);
    if (v.layout != 2) {
        reinterpret_cast<float2*>(v.base)[o] = val;
    } else {
        v.base[o] = val.x;
        v.base[o + (int64_t)v.Fo * v.T] = val.y;
    }
}

__device__ __forceinline__ float2 spec_load(const SpecView& v, int64_t row, int t, int bin) {
    const int64_t o = spec_index(v, row, t, bin);
    if (v.layout != 2) return __ldg(reinterpret_cast<const float2*>(v.base) + o);
    return make_float2(__ldg(v.base + o), __ldg(v.base + o + (int64_t)v.Fo * v.T));
}

}  // namespace al
