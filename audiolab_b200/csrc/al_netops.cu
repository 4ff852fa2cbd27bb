// Fused row-wise operators of the RoFormer mask network's bf16 inference path (sm_100a).
//
// The dense contractions of the network go to the tensor cores through cuBLAS / cuDNN (library calls);
// everything between two contractions is HBM-bound row-wise work that the reference runs as chains
// of 5-12 elementwise PyTorch kernels under autocast (upstream bs_roformer RMSNorm / rotary /
// gated attention).  Each chain is ONE kernel here: bf16 in, fp32 arithmetic, bf16
// out, one 16-byte access per 8 elements.
//
//   rmsnorm_bf16  : [x += bias;] out = x / max(||x||_2, eps) * sqrt(dim) * gamma      (4 B / element)
//   rotary_bf16   : q, k rotated in place by the position of their token                (8 B / element)
//   gate_bf16     : attention output *= sigmoid(gate of its (token, head)), in place   (4 B / element)
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "al_gemm.h"
#include "al_kernels.h"

namespace al {

// [emul-begin]
__device__ __forceinline__ void bf16x8_to_f32(const uint4 u, float (&f)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = __bfloat1622float2(h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}

__device__ __forceinline__ uint4 f32_to_bf16x8(const float (&f)[8]) {
    uint4 u;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return u;
}

// The same for a 16-bit format chosen at compile time: F16 = IEEE half (saturating to +-65504), else bfloat16.
#ifndef AL_CPU_EMUL
__device__ __forceinline__ float2 half2_bits_to_f32(unsigned u) {
    float2 r;
    asm("{.reg .f16 lo, hi; mov.b32 {lo, hi}, %2; cvt.f32.f16 %0, lo; cvt.f32.f16 %1, hi;}" : "=f"(r.x), "=f"(r.y) : "r"(u));
    return r;
}
__device__ __forceinline__ unsigned f32_to_half2_bits(float lo, float hi) {
    unsigned r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
#else   // the host emulation only runs the bfloat16 instantiations
__device__ __forceinline__ float2 half2_bits_to_f32(unsigned) { return make_float2(0.f, 0.f); }
__device__ __forceinline__ unsigned f32_to_half2_bits(float, float) { return 0u; }
#endif
template <bool F16>
__device__ __forceinline__ float2 h16x2_to_f32(unsigned u) {
    if (F16) return half2_bits_to_f32(u);
    return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xFFFF0000u));
}
template <bool F16>
__device__ __forceinline__ unsigned f32_to_h16x2(float lo, float hi) {
    if (F16) return f32_to_half2_bits(lo, hi);
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const unsigned*>(&h);
}
template <bool F16>
__device__ __forceinline__ void h16x8_to_f32(const uint4 u, float (&f)[8]) {
    const unsigned w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 t = h16x2_to_f32<F16>(w[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}
template <bool F16>
__device__ __forceinline__ uint4 f32_to_h16x8(const float (&f)[8]) {
    return make_uint4(f32_to_h16x2<F16>(f[0], f[1]), f32_to_h16x2<F16>(f[2], f[3]), f32_to_h16x2<F16>(f[4], f[5]),
                      f32_to_h16x2<F16>(f[6], f[7]));
}
template <bool F16>
__device__ __forceinline__ float h16_to_f32(unsigned short v) { return h16x2_to_f32<F16>((unsigned)v).x; }

// One warp per row; the row stays in registers between the reduction and the scaling pass.
template <int MAXC>   // chunks of 256 elements held in registers (dim <= 256 * MAXC)
__global__ void __launch_bounds__(256)
rmsnorm_bf16_kernel(__nv_bfloat16* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ bias,
                    __nv_bfloat16* __restrict__ out, long long n_rows, int dim, float scale, float eps) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    uint4* xr = reinterpret_cast<uint4*>(x + row * dim);
    uint4* orow = reinterpret_cast<uint4*>(out + row * dim);
    float v[MAXC][8];
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        const int e = c * 256 + lane * 8;
        if (e < dim) {
            bf16x8_to_f32(xr[e >> 3], v[c]);
            if (bias) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + e));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + e) + 1);
                v[c][0] += b0.x; v[c][1] += b0.y; v[c][2] += b0.z; v[c][3] += b0.w;
                v[c][4] += b1.x; v[c][5] += b1.y; v[c][6] += b1.z; v[c][7] += b1.w;
                const uint4 u = f32_to_bf16x8(v[c]);     // the residual stream is bf16: norm what is stored
                xr[e >> 3] = u;
                bf16x8_to_f32(u, v[c]);
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) ss = fmaf(v[c][i], v[c][i], ss);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    const float inv = scale / fmaxf(sqrtf(ss), eps);
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        const int e = c * 256 + lane * 8;
        if (e < dim) {
            const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + e));
            const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + e) + 1);
            float r[8];
            r[0] = v[c][0] * inv * g0.x; r[1] = v[c][1] * inv * g0.y; r[2] = v[c][2] * inv * g0.z; r[3] = v[c][3] * inv * g0.w;
            r[4] = v[c][4] * inv * g1.x; r[5] = v[c][5] * inv * g1.y; r[6] = v[c][6] * inv * g1.z; r[7] = v[c][7] * inv * g1.w;
            orow[e >> 3] = f32_to_bf16x8(r);
        }
    }
}

// [emul-end]

cudaError_t launch_rmsnorm_bf16(void* x, const float* gamma, const float* bias, void* out, long long n_rows, int dim,
                                float scale, float eps, cudaStream_t stream) {
    const int wpb = 8;
    const unsigned grid = (unsigned)((n_rows + wpb - 1) / wpb);
    auto* xb = reinterpret_cast<__nv_bfloat16*>(x);
    auto* ob = reinterpret_cast<__nv_bfloat16*>(out);
    if (dim <= 512) rmsnorm_bf16_kernel<2><<<grid, wpb * 32, 0, stream>>>(xb, gamma, bias, ob, n_rows, dim, scale, eps);
    else if (dim <= 1024) rmsnorm_bf16_kernel<4><<<grid, wpb * 32, 0, stream>>>(xb, gamma, bias, ob, n_rows, dim, scale, eps);
    else rmsnorm_bf16_kernel<8><<<grid, wpb * 32, 0, stream>>>(xb, gamma, bias, ob, n_rows, dim, scale, eps);
    count_launch();
    return cudaGetLastError();
}

// [emul-begin]
// q, k: [n_rows, heads * dim_head] bf16, rotated in place.  Element pair (2i, 2i+1) of every head turns by
// the angle pos * freq_i, pos = (row / pos_div) % pos_mod;  cs[pos][i] = (cos, sin).
__global__ void __launch_bounds__(256)
rotary_bf16_kernel(uint4* __restrict__ q, uint4* __restrict__ k, const float2* __restrict__ cs, long long n_vec,
                   int vec_per_row, int dim_head, long long pos_div, int pos_mod) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_vec) return;
    const long long row = i / vec_per_row;
    const int col = (int)(i - row * vec_per_row) * 8;
    const int pos = (int)((row / pos_div) % pos_mod);
    const int half = dim_head >> 1;
    const float4* t = reinterpret_cast<const float4*>(cs + (long long)pos * half + ((col % dim_head) >> 1));
    const float4 t0 = __ldg(t), t1 = __ldg(t + 1);
    const float c[4] = {t0.x, t0.z, t1.x, t1.z}, s[4] = {t0.y, t0.w, t1.y, t1.w};
    float a[8], b[8], ra[8], rb[8];
    bf16x8_to_f32(q[i], a);
    bf16x8_to_f32(k[i], b);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        ra[2 * j] = a[2 * j] * c[j] - a[2 * j + 1] * s[j];
        ra[2 * j + 1] = a[2 * j + 1] * c[j] + a[2 * j] * s[j];
        rb[2 * j] = b[2 * j] * c[j] - b[2 * j + 1] * s[j];
        rb[2 * j + 1] = b[2 * j + 1] * c[j] + b[2 * j] * s[j];
    }
    q[i] = f32_to_bf16x8(ra);
    k[i] = f32_to_bf16x8(rb);
}

// [emul-end]

cudaError_t launch_rotary_bf16(void* q, void* k, const float* cs, long long n_rows, int heads, int dim_head,
                               long long pos_div, int pos_mod, cudaStream_t stream) {
    const int vec_per_row = heads * dim_head / 8;
    const long long n_vec = n_rows * vec_per_row;
    rotary_bf16_kernel<<<(unsigned)((n_vec + 255) / 256), 256, 0, stream>>>(
        reinterpret_cast<uint4*>(q), reinterpret_cast<uint4*>(k), reinterpret_cast<const float2*>(cs), n_vec, vec_per_row,
        dim_head, pos_div, pos_mod);
    count_launch();
    return cudaGetLastError();
}

// [emul-begin]
// o: [n_rows, heads * dim_head] bf16, gates: [n_rows, heads] bf16;  o[row, h, :] *= sigmoid(gates[row, h])
template <bool F16>
__global__ void __launch_bounds__(256)
gate_h16_kernel(uint4* __restrict__ o, const __nv_bfloat16* __restrict__ gates, long long n_vec, int vec_per_row,
                int gate_ld, int dim_head) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_vec) return;
    const long long row = i / vec_per_row;
    const int h = ((int)(i - row * vec_per_row) * 8) / dim_head;
    const float g = h16_to_f32<F16>(reinterpret_cast<const unsigned short*>(gates)[row * gate_ld + h]);
    const float sg = 1.f / (1.f + __expf(-g));
    float a[8];
    h16x8_to_f32<F16>(o[i], a);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] *= sg;
    o[i] = f32_to_h16x8<F16>(a);
}

// [emul-end]

cudaError_t launch_gate_bf16(void* o, const void* gates, long long n_rows, int heads, int dim_head, int gate_ld, int fp16,
                             cudaStream_t stream) {
    const int vec_per_row = heads * dim_head / 8;
    const long long n_vec = n_rows * vec_per_row;
    if (fp16)
        gate_h16_kernel<true><<<(unsigned)((n_vec + 255) / 256), 256, 0, stream>>>(
            reinterpret_cast<uint4*>(o), reinterpret_cast<const __nv_bfloat16*>(gates), n_vec, vec_per_row, gate_ld, dim_head);
    else
        gate_h16_kernel<false><<<(unsigned)((n_vec + 255) / 256), 256, 0, stream>>>(
            reinterpret_cast<uint4*>(o), reinterpret_cast<const __nv_bfloat16*>(gates), n_vec, vec_per_row, gate_ld, dim_head);
    count_launch();
    return cudaGetLastError();
}

// [emul-begin]
// Start / re-normalise the fp32 residual stream of the tcgen05 path (al_gemm.cu EPI_RES keeps it afterwards):
// y = x_in (+ bias) (then RMSNorm if gamma);  x32 = y, xb = bf16(y), ss[row][p] = partial sums of y^2.
// One warp per row; lane l owns elements [c * 256 + 8 l, + 8) of chunk c.
template <int MAXC, bool F16>
__global__ void __launch_bounds__(256)
resid_prepare_kernel(const float* __restrict__ x_in, const float* __restrict__ bias, const float* __restrict__ gamma,
                     float* __restrict__ x32, __nv_bfloat16* __restrict__ xb, float* __restrict__ ss_out,
                     long long n_rows, int dim, int ss_parts, float scale, float eps) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const float4* xr = reinterpret_cast<const float4*>(x_in + row * dim);
    float v[MAXC][8];
    float ss = 0.f;
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        const int e = c * 256 + lane * 8;
        if (e < dim) {
            const float4 a = xr[e >> 2], b = xr[(e >> 2) + 1];
            v[c][0] = a.x; v[c][1] = a.y; v[c][2] = a.z; v[c][3] = a.w;
            v[c][4] = b.x; v[c][5] = b.y; v[c][6] = b.z; v[c][7] = b.w;
            if (bias) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + e));
                const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + e) + 1);
                v[c][0] += b0.x; v[c][1] += b0.y; v[c][2] += b0.z; v[c][3] += b0.w;
                v[c][4] += b1.x; v[c][5] += b1.y; v[c][6] += b1.z; v[c][7] += b1.w;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) ss = fmaf(v[c][i], v[c][i], ss);
        }
    }
    if (gamma) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        const float inv = scale / fmaxf(sqrtf(ss), eps);
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            const int e = c * 256 + lane * 8;
            if (e < dim) {
                const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + e));
                const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + e) + 1);
                v[c][0] *= inv * g0.x; v[c][1] *= inv * g0.y; v[c][2] *= inv * g0.z; v[c][3] *= inv * g0.w;
                v[c][4] *= inv * g1.x; v[c][5] *= inv * g1.y; v[c][6] *= inv * g1.z; v[c][7] *= inv * g1.w;
            }
        }
    }
    // partial sums of squares of the stored row: part p covers elements [p * dim / ss_parts, (p + 1) * dim / ss_parts)
    const int part_len = dim / ss_parts;
    float4* o32 = reinterpret_cast<float4*>(x32 + row * dim);
    uint4* ob = reinterpret_cast<uint4*>(xb + row * dim);
    for (int p = 0; p < ss_parts; ++p) {
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < MAXC; ++c) {
            const int e = c * 256 + lane * 8;
            if (e < dim && e / part_len == p) {
#pragma unroll
                for (int i = 0; i < 8; ++i) s = fmaf(v[c][i], v[c][i], s);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) ss_out[row * ss_parts + p] = s;
    }
#pragma unroll
    for (int c = 0; c < MAXC; ++c) {
        const int e = c * 256 + lane * 8;
        if (e < dim) {
            o32[e >> 2] = make_float4(v[c][0], v[c][1], v[c][2], v[c][3]);
            o32[(e >> 2) + 1] = make_float4(v[c][4], v[c][5], v[c][6], v[c][7]);
            ob[e >> 3] = f32_to_h16x8<F16>(v[c]);
        }
    }
}

// [emul-end]

cudaError_t launch_resid_prepare(const float* x_in, const float* bias, const float* gamma, float* x32, void* xb, float* ss,
                                 long long n_rows, int dim, int ss_parts, float eps, int fp16, cudaStream_t stream) {
    const int wpb = 8;
    const unsigned grid = (unsigned)((n_rows + wpb - 1) / wpb);
    auto* ob = reinterpret_cast<__nv_bfloat16*>(xb);
    const float scale = sqrtf((float)dim);
#define AL_RP_LAUNCH(C, H) resid_prepare_kernel<C, H><<<grid, wpb * 32, 0, stream>>>(x_in, bias, gamma, x32, ob, ss, n_rows, dim, ss_parts, scale, eps)
    if (dim <= 512) { if (fp16) AL_RP_LAUNCH(2, true); else AL_RP_LAUNCH(2, false); }
    else if (dim <= 1024) { if (fp16) AL_RP_LAUNCH(4, true); else AL_RP_LAUNCH(4, false); }
    else { if (fp16) AL_RP_LAUNCH(8, true); else AL_RP_LAUNCH(8, false); }
#undef AL_RP_LAUNCH
    count_launch();
    return cudaGetLastError();
}

// [emul-begin]
// Per-band RMSNorm of the band-split input (upstream BandSplit: RMSNorm(d_j) in front of each band's Linear): one CTA per
// row; the row is read once, each warp normalises whole bands (sum of squares over the band's d_j elements, lanes
// strided), and the bf16 result is the A operand of the grouped band-split GEMM.
template <bool F16>
__global__ void __launch_bounds__(256)
band_norm_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ gamma, const int* __restrict__ band_off,
                 int n_bands, __nv_bfloat16* __restrict__ out, long long ldo, float eps) {
    const long long row = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, n_warps = blockDim.x >> 5;
    const float* xr = x + row * ldx;
    __nv_bfloat16* orow = out + row * ldo;
    for (int j = warp; j < n_bands; j += n_warps) {
        const int a = __ldg(band_off + j), b = __ldg(band_off + j + 1);
        float ss = 0.f;
        for (int i = a + lane; i < b; i += 32) {
            const float v = xr[i];
            ss = fmaf(v, v, ss);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
        const float inv = sqrtf((float)(b - a)) / fmaxf(sqrtf(ss), eps);
        unsigned short* o16 = reinterpret_cast<unsigned short*>(orow);
        for (int i = a + lane; i < b; i += 32) o16[i] = (unsigned short)(f32_to_h16x2<F16>(xr[i] * inv * __ldg(gamma + i), 0.f) & 0xFFFFu);
    }
}
// [emul-end]

cudaError_t launch_band_norm(const float* x, long long ldx, const float* gamma, const int* band_off, int n_bands, void* out,
                             long long ldo, long long n_rows, float eps, int fp16, cudaStream_t stream) {
    auto* o = reinterpret_cast<__nv_bfloat16*>(out);
    if (fp16) band_norm_kernel<true><<<(unsigned)n_rows, 256, 0, stream>>>(x, ldx, gamma, band_off, n_bands, o, ldo, eps);
    else band_norm_kernel<false><<<(unsigned)n_rows, 256, 0, stream>>>(x, ldx, gamma, band_off, n_bands, o, ldo, eps);
    count_launch();
    return cudaGetLastError();
}

// x: [n] bf16, exact (erf) GELU in place -- upstream FeedForward's nn.GELU() between its two Linear layers
// (SURVEY.md A.4).  erff costs ~30 issue slots per element, which makes the straightforward kernel
// issue-bound at half the HBM rate (profiles/r01d).  A bf16 -> bf16 function has only 65536 inputs: the
// kernel keeps the whole function as a 128 KB table in shared memory (built once per process with the
// fp32 formula torch uses, x * 0.5 * (1 + erff(x / sqrt 2)), rounded to bf16), so one element costs one
// LDS.U16 and the result is the exact formula's for every bit pattern.
// [emul-begin]
constexpr int kGeluThreads = 1024;
constexpr int kGeluLutBytes = 65536 * 2;

__global__ void gelu_lut_init_kernel(unsigned short* __restrict__ lut) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 65536) return;
    const float x = __uint_as_float((unsigned)i << 16);
    const float y = x * 0.5f * (1.f + erff(x * 0.70710678118654752440f));
    lut[i] = __bfloat16_as_ushort(__float2bfloat16_rn(y));
}

__device__ __forceinline__ unsigned gelu_lut2(const unsigned short* __restrict__ t, unsigned v) {
    return (unsigned)t[v & 0xFFFFu] | ((unsigned)t[v >> 16] << 16);
}

__global__ void __launch_bounds__(kGeluThreads, 1)
gelu_bf16_kernel(uint4* __restrict__ x, long long n_vec, const uint4* __restrict__ lut_g) {
    AL_DYN_SMEM(unsigned char, gelu_smem);
    uint4* lut4 = reinterpret_cast<uint4*>(gelu_smem);
    for (int i = threadIdx.x; i < kGeluLutBytes / 16; i += kGeluThreads) lut4[i] = __ldg(lut_g + i);
    __syncthreads();
    const unsigned short* __restrict__ t = reinterpret_cast<const unsigned short*>(gelu_smem);
    const long long stride = (long long)gridDim.x * kGeluThreads;
    long long i = (long long)blockIdx.x * kGeluThreads + threadIdx.x;
    // two vectors per iteration: both loads are in flight before the first lookup
    for (; i + stride < n_vec; i += 2 * stride) {
        uint4 a = x[i], b = x[i + stride];
        a.x = gelu_lut2(t, a.x); a.y = gelu_lut2(t, a.y); a.z = gelu_lut2(t, a.z); a.w = gelu_lut2(t, a.w);
        b.x = gelu_lut2(t, b.x); b.y = gelu_lut2(t, b.y); b.z = gelu_lut2(t, b.z); b.w = gelu_lut2(t, b.w);
        x[i] = a;
        x[i + stride] = b;
    }
    if (i < n_vec) {
        uint4 a = x[i];
        a.x = gelu_lut2(t, a.x); a.y = gelu_lut2(t, a.y); a.z = gelu_lut2(t, a.z); a.w = gelu_lut2(t, a.w);
        x[i] = a;
    }
}

// [emul-end]

cudaError_t launch_gelu_bf16(void* x, long long n, cudaStream_t stream) {
    const long long n_vec = n / 8;
    if (n_vec <= 0) return cudaSuccess;
    static unsigned short* lut = nullptr;
    static int lut_dev = -1, n_sm = 148;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (!lut || lut_dev != dev) {   // once per process (one process per GPU)
        unsigned short* d = nullptr;
        e = cudaMalloc((void**)&d, kGeluLutBytes);
        if (e != cudaSuccess) return e;
        gelu_lut_init_kernel<<<256, 256, 0, stream>>>(d);
        e = cudaStreamSynchronize(stream);               // other streams may use the table from now on
        if (e != cudaSuccess) { cudaFree(d); return e; }
        e = cudaFuncSetAttribute(gelu_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kGeluLutBytes);
        if (e != cudaSuccess) { cudaFree(d); return e; }
        cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
        if (n_sm <= 0) n_sm = 148;
        lut = d;
        lut_dev = dev;
    }
    const long long want = (n_vec + kGeluThreads - 1) / kGeluThreads;
    const unsigned grid = (unsigned)(want < n_sm ? want : n_sm);
    gelu_bf16_kernel<<<grid, kGeluThreads, kGeluLutBytes, stream>>>(reinterpret_cast<uint4*>(x), n_vec,
                                                                     reinterpret_cast<const uint4*>(lut));
    count_launch();
    return cudaGetLastError();
}

}  // namespace al
