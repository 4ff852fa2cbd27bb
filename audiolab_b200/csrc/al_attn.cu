// Band-axis attention of the RoFormer mask network (opt-in, AUDIOLAB_B200_BAND_ATTN=1): softmax(Q K^T / sqrt(d)) V over the
// <= 64 frequency bands of one (batch, frame), head by head -- upstream Attention.forward inside the frequency
// transformer (driven by MDXCSeparator.demix behind stem_separator.py:281).
//
// Why a kernel of our own: for 62-token sequences cuDNN picks an sm_80 `wmma` flash kernel that moves q, k, v, o at
// ~2.4 TB/s (profiles/r01j_launches_bench_step.txt, 7 % of the step).  The problem is HBM-bound (0.03 flop/B short of
// nothing: 1 MFLOP per 32 KB), so warp-level mma.sync (m16n8k16, bf16 -> fp32) is enough to follow the memory system;
// tcgen05 tiles of 128 rows would be 52 % padding here.
//
// q, k, v, o: [n_seq * F, H * 64] bf16, token (s, f) in row s * F + f (the token-major layout of the residual stream,
// read and written in place -- no transposition copies).  One CTA = 4 warps = one (s, h); warp w owns query rows
// 16 w .. 16 w + 15.  Q, K, V tiles (64 x 64, rows >= F zero) sit in shared memory with a 72-element row stride
// (conflict-free 32-bit fragment loads); S = Q K^T and O = P V stay in mma accumulator registers, the softmax runs on
// the accumulator fragments (row max / sum over the 4 lanes that share a row), P is re-used as the A operand of the second
// product without leaving registers.  With `gates` the sigmoid gate of upstream's Attention (out * to_gates(x).sigmoid())
// is folded into the final normalisation, and with `cos_sin` the rotary embedding of q and k (position = band index) is
// applied while the tiles are staged -- both remove a separate HBM pass for this axis, with the same roundings as the
// stand-alone rotary / gate kernels (bf16 after the rotation; the gate multiplies before the single output rounding).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "al_async.cuh"
#include "al_kernels.h"

namespace al {

// [emul-begin]
constexpr int kBaD = 64;          // head dimension
constexpr int kBaF = 64;          // padded sequence length (bands)
constexpr int kBaLd = 72;         // shared-memory row stride in bf16 elements (144 B: rows shift by 4 banks)

#ifndef AL_CPU_EMUL
// D (16x8, fp32) += A (16x16, bf16, row) * B (16x8, bf16, col); fragment layouts: PTX ISA "mma.m16n8k16"
//   a0..a3: (row lane/4 [+8 for a1, a3], cols 2 (lane%4) + {0,1} [+8 for a2, a3])
//   b0, b1: (k = 2 (lane%4) + {0,1} [+8 for b1], n = lane/4)
//   d0..d3: (row lane/4 [+8 for d2, d3], cols 2 (lane%4) + {0,1})
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
#endif

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<const uint32_t*>(&h);
}

// F16 = the q / k / v / o / gates tensors are IEEE half instead of bfloat16 (the fp16-operand mode of the network):
// the same kernel with mma.sync ... f16.f16 and half conversions.  The host emulation runs the bfloat16 form only.
#ifndef AL_CPU_EMUL
__device__ __forceinline__ void mma_f16_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t pack_f16x2_sat(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ float f16_bits_to_f32(unsigned short v) {
    float r;
    asm("{.reg .f16 h; mov.b16 h, %1; cvt.f32.f16 %0, h;}" : "=f"(r) : "h"(v));
    return r;
}
#else
__device__ __forceinline__ void mma_f16_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) { mma_bf16_16816(d, a, b); }
__device__ __forceinline__ uint32_t pack_f16x2_sat(float lo, float hi) { return pack_bf16x2(lo, hi); }
__device__ __forceinline__ float f16_bits_to_f32(unsigned short v) { return __uint_as_float((unsigned)v << 16); }
#endif
template <bool F16>
__device__ __forceinline__ void mma_h16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    if (F16) mma_f16_16816(d, a, b); else mma_bf16_16816(d, a, b);
}
template <bool F16>
__device__ __forceinline__ uint32_t pack_h16x2(float lo, float hi) { return F16 ? pack_f16x2_sat(lo, hi) : pack_bf16x2(lo, hi); }
template <bool F16>
__device__ __forceinline__ float h16_bits_to_f32(unsigned short v) {
    return F16 ? f16_bits_to_f32(v) : __uint_as_float((unsigned)v << 16);
}

// 8 bf16 = 4 (even, odd) pairs, each turned by its (cos, sin): fp32 inside, rounded back to bf16 like rotary_bf16_kernel
__device__ __forceinline__ uint4 rotate_bf16x8(uint4 u, const float2* __restrict__ cs) {
    uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float a = __uint_as_float(w[j] << 16), b = __uint_as_float(w[j] & 0xFFFF0000u);
        const float2 t = __ldg(cs + j);
        w[j] = pack_bf16x2(a * t.x - b * t.y, b * t.x + a * t.y);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

template <bool F16>
__global__ void __launch_bounds__(128)
band_attn_bf16_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                      const __nv_bfloat16* __restrict__ v, __nv_bfloat16* __restrict__ o,
                      const __nv_bfloat16* __restrict__ gates, const float2* __restrict__ cos_sin, int F, int H,
                      float scale, int gate_ld) {
    __shared__ __align__(16) __nv_bfloat16 Qs[kBaF * kBaLd];
    __shared__ __align__(16) __nv_bfloat16 Ks[kBaF * kBaLd];
    __shared__ __align__(16) __nv_bfloat16 Vs[kBaF * kBaLd];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s = blockIdx.x / H, h = blockIdx.x - s * H;
    const long long ld = (long long)H * kBaD;                       // elements per token row
    const long long base = (long long)s * F * ld + (long long)h * kBaD;

    // ---- stage Q, K, V: 64 rows x 8 vectors of 16 B each; rows >= F are zero ----------------------------------
    if (cos_sin == nullptr) {
        // the production path (rotary already applied by the GEMM epilogue): asynchronous 16-byte copies, all 12 per thread in
        // flight at once -- a register-staged loop serialises its four iterations on the load latency (35 % of the stall
        // samples sat on the first shared-memory store, profiles/r02x_ncu_full_band_attention.txt)
#pragma unroll
        for (int i = tid; i < kBaF * 8; i += 128) {
            const int row = i >> 3, c8 = (i & 7) * 8;
            if (row < F) {
                const long long g = base + (long long)row * ld + c8;
                al_cp_async16(Qs + row * kBaLd + c8, q + g);
                al_cp_async16(Ks + row * kBaLd + c8, k + g);
                al_cp_async16(Vs + row * kBaLd + c8, v + g);
            } else {
                const uint4 z = make_uint4(0u, 0u, 0u, 0u);
                *reinterpret_cast<uint4*>(Qs + row * kBaLd + c8) = z;
                *reinterpret_cast<uint4*>(Ks + row * kBaLd + c8) = z;
                *reinterpret_cast<uint4*>(Vs + row * kBaLd + c8) = z;
            }
        }
        al_cp_async_commit();
        al_cp_async_wait<0>();
    } else {
        for (int i = tid; i < kBaF * 8; i += 128) {
            const int row = i >> 3, c8 = (i & 7) * 8;
            uint4 vq = make_uint4(0u, 0u, 0u, 0u), vk = vq, vv = vq;
            if (row < F) {
                const long long g = base + (long long)row * ld + c8;
                vq = __ldg(reinterpret_cast<const uint4*>(q + g));
                vk = __ldg(reinterpret_cast<const uint4*>(k + g));
                vv = __ldg(reinterpret_cast<const uint4*>(v + g));
                // upstream rotary_embed.rotate_queries_or_keys on q and k, position = band index (= row)
                const float2* cs = cos_sin + row * (kBaD / 2) + (c8 >> 1);
                vq = rotate_bf16x8(vq, cs);
                vk = rotate_bf16x8(vk, cs);
            }
            *reinterpret_cast<uint4*>(Qs + row * kBaLd + c8) = vq;
            *reinterpret_cast<uint4*>(Ks + row * kBaLd + c8) = vk;
            *reinterpret_cast<uint4*>(Vs + row * kBaLd + c8) = vv;
        }
    }
    __syncthreads();

    const int r0 = 16 * warp + (lane >> 2);                          // this lane's rows: r0 and r0 + 8
    const int c2 = 2 * (lane & 3);
    // ---- S = Q K^T ---------------------------------------------------------------------------------------------
    uint32_t qa[4][4];
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        const __nv_bfloat16* p0 = Qs + r0 * kBaLd + ks * 16 + c2;
        qa[ks][0] = *reinterpret_cast<const uint32_t*>(p0);
        qa[ks][1] = *reinterpret_cast<const uint32_t*>(p0 + 8 * kBaLd);
        qa[ks][2] = *reinterpret_cast<const uint32_t*>(p0 + 8);
        qa[ks][3] = *reinterpret_cast<const uint32_t*>(p0 + 8 * kBaLd + 8);
    }
    float sc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int i = 0; i < 4; ++i) sc[nt][i] = 0.f;
        const __nv_bfloat16* kp = Ks + (nt * 8 + (lane >> 2)) * kBaLd + c2;   // B[k][n] = K[n][k]
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            uint32_t b[2];
            b[0] = *reinterpret_cast<const uint32_t*>(kp + ks * 16);
            b[1] = *reinterpret_cast<const uint32_t*>(kp + ks * 16 + 8);
            mma_h16<F16>(sc[nt], qa[ks], b);
        }
    }
    // ---- softmax over the keys (columns); keys >= F are masked out ------------------------------------------------
    float mx0 = -3.0e38f, mx1 = -3.0e38f;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int key = nt * 8 + c2 + (i & 1);
            sc[nt][i] = key < F ? sc[nt][i] * scale : -3.0e38f;
        }
        mx0 = fmaxf(mx0, fmaxf(sc[nt][0], sc[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(sc[nt][2], sc[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    float sum0 = 0.f, sum1 = 0.f;
    uint32_t pa[4][4];                                               // P as the A operand of P V: k-step = 16 keys = 2 n-tiles
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        const float e0 = __expf(sc[nt][0] - mx0), e1 = __expf(sc[nt][1] - mx0);
        const float e2 = __expf(sc[nt][2] - mx1), e3 = __expf(sc[nt][3] - mx1);
        sum0 += e0 + e1;
        sum1 += e2 + e3;
        pa[nt >> 1][(nt & 1) * 2 + 0] = pack_h16x2<F16>(e0, e1);         // a0 / a2: row r0,     cols (+8 for the odd tile)
        pa[nt >> 1][(nt & 1) * 2 + 1] = pack_h16x2<F16>(e2, e3);         // a1 / a3: row r0 + 8
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    float inv0 = 1.f / sum0, inv1 = 1.f / sum1;
    if (gates) {   // upstream Attention: out * to_gates(x).sigmoid(), gates [n_seq * F, H]; folded into the normalisation
        const long long t0 = (long long)s * F + r0;
        const long long gld = gate_ld > 0 ? gate_ld : H;                 // row stride of the gate matrix
        const unsigned short* g16 = reinterpret_cast<const unsigned short*>(gates);
        if (r0 < F) inv0 *= 1.f / (1.f + __expf(-h16_bits_to_f32<F16>(g16[t0 * gld + h])));
        if (r0 + 8 < F) inv1 *= 1.f / (1.f + __expf(-h16_bits_to_f32<F16>(g16[(t0 + 8) * gld + h])));
    }
    // ---- O = P V -------------------------------------------------------------------------------------------------
    const unsigned short* vs16 = reinterpret_cast<const unsigned short*>(Vs);
    __nv_bfloat16* orow = Qs;                                        // the warp's own 16 Q rows become its output rows
    __syncwarp();
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {                                 // nt = tile of 8 head-dimension columns
        float oc[4] = {0.f, 0.f, 0.f, 0.f};
        const int n = nt * 8 + (lane >> 2);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {                             // B[k][n] = V[key k][n]: two keys per register
            const int k0 = ks * 16 + c2;
            uint32_t b[2];
            b[0] = (uint32_t)vs16[k0 * kBaLd + n] | ((uint32_t)vs16[(k0 + 1) * kBaLd + n] << 16);
            b[1] = (uint32_t)vs16[(k0 + 8) * kBaLd + n] | ((uint32_t)vs16[(k0 + 9) * kBaLd + n] << 16);
            mma_h16<F16>(oc, pa[ks], b);
        }
        *reinterpret_cast<uint32_t*>(orow + r0 * kBaLd + nt * 8 + c2) = pack_h16x2<F16>(oc[0] * inv0, oc[1] * inv0);
        *reinterpret_cast<uint32_t*>(orow + (r0 + 8) * kBaLd + nt * 8 + c2) = pack_h16x2<F16>(oc[2] * inv1, oc[3] * inv1);
    }
    __syncwarp();
    // ---- the warp's 16 rows leave as 16-byte vectors ---------------------------------------------------------------
#pragma unroll
    for (int i = lane; i < 16 * 8; i += 32) {
        const int row = 16 * warp + (i >> 3), c8 = (i & 7) * 8;
        if (row < F)
            *reinterpret_cast<uint4*>(o + base + (long long)row * ld + c8) = *reinterpret_cast<const uint4*>(orow + row * kBaLd + c8);
    }
}
// [emul-end]

cudaError_t launch_band_attn_bf16(const void* q, const void* k, const void* v, void* o, const void* gates, const float* cos_sin,
                                  long long n_seq, int F, int heads, float scale, int gate_ld, int fp16, cudaStream_t stream) {
    if (n_seq <= 0) return cudaSuccess;
    const long long ctas = n_seq * heads;
    if (ctas > 0x7fffffffLL) return cudaErrorInvalidValue;
    if (fp16 && cos_sin) return cudaErrorInvalidValue;   // the in-kernel rotary staging is bfloat16-only (the GEMM epilogue rotates)
    auto* qp = reinterpret_cast<const __nv_bfloat16*>(q);
    auto* kp = reinterpret_cast<const __nv_bfloat16*>(k);
    auto* vp = reinterpret_cast<const __nv_bfloat16*>(v);
    auto* op = reinterpret_cast<__nv_bfloat16*>(o);
    auto* gp = reinterpret_cast<const __nv_bfloat16*>(gates);
    auto* cs = reinterpret_cast<const float2*>(cos_sin);
    if (fp16) band_attn_bf16_kernel<true><<<(unsigned)ctas, 128, 0, stream>>>(qp, kp, vp, op, gp, cs, F, heads, scale, gate_ld);
    else band_attn_bf16_kernel<false><<<(unsigned)ctas, 128, 0, stream>>>(qp, kp, vp, op, gp, cs, F, heads, scale, gate_ld);
    count_launch();
    return cudaGetLastError();
}

}  // namespace al
