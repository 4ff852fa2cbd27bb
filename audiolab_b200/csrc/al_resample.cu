// K3: polyphase FIR resampler with scipy.signal.resample_poly semantics (zero-phase, zero ends):
//   out[m] = sum_k taps[k] * x_up[m*down + half - k],  x_up[j*up] = x[j]
//          = sum_i taps[phi + up*i] * x[b - i],  c = m*down + half, phi = c mod up, b = c div up.
// Replaces librosa.load(sr=44100) / res_type "polyphase" (reference:
// modules/separator/stem_separator.py:865; modules/rvc/infer/lib/uvr5_pack/lib_v5/model_param_init.py:22).
//
// HBM-bound by bytes (4*down/up B read + 4 B written per output sample) but ~22 FMAs per output: the
// kernel has to keep both operands of most FMAs out of shared memory to get near the byte rate.
//
// resample_rb_kernel (register-blocked, the path for 147/160 and 160/147):
//   outputs m = P*up + r (P = "period", r = residue) have phase and input offset that depend on r only:
//   phi = (r*down + half) mod up, b = P*down + (r*down + half) div up.  A thread owns kRbR = 4 adjacent
//   residues: their 4 x tpp taps stay in registers for the thread's whole life (as 4 rows of a
//   kRbW-wide window, zero where a row does not reach), and per period it reads ONE shared window of
//   kRbW = 26 input samples for 4 x 26 FMAs (0.25 shared-memory loads per FMA instead of 2).
//   A persistent CTA walks tiles of QT periods; the input span of the next tile is prefetched with
//   cp.async (16 B) into the other half of a double buffer while this tile is computed; outputs go
//   through shared memory (bank-rotated stores) and leave as aligned 128-bit rows.
//   Each output accumulates from the oldest input sample to the newest, as scipy's upfirdn does.
// resample_generic_kernel: any other ratio (taps and input span staged in shared memory).
#include <limits.h>
#include <stdlib.h>

#include "al_kernels.h"
#include "al_async.cuh"

namespace al {

constexpr int kResTile = 2048;   // outputs per CTA

__global__ void __launch_bounds__(256)
resample_generic_kernel(const float* __restrict__ in, long long in_stride, float* __restrict__ out,
                long long out_stride, long long n_in, long long n_out, int up, int down,
                const float* __restrict__ taps, int n_taps, int taps_per_phase, int span_cap) {
    extern __shared__ __align__(16) float s_res[];
    float* s_taps = s_res;                              // [taps_per_phase][up]
    float* s_x = s_res + taps_per_phase * up;           // [span_cap]
    const int row = blockIdx.y;
    const long long m0 = (long long)blockIdx.x * kResTile;
    const int half = (n_taps - 1) / 2;
    for (int i = threadIdx.x; i < taps_per_phase * up; i += blockDim.x) {
        const int ii = i / up, phi = i - ii * up;
        const int k = phi + up * ii;
        s_taps[i] = k < n_taps ? __ldg(taps + k) : 0.f;
    }
    // input span needed by outputs [m0, m0 + tile): b ranges over [c0/up - (tpp-1), c1/up]
    const long long c0 = m0 * down + half;
    const long long b_lo = c0 / up - (taps_per_phase - 1);
    const float* __restrict__ src = in + (long long)row * in_stride;
    for (int i = threadIdx.x; i < span_cap; i += blockDim.x) {
        const long long j = b_lo + i;
        s_x[i] = (j >= 0 && j < n_in) ? __ldg(src + j) : 0.f;
    }
    __syncthreads();
    for (int o = threadIdx.x; o < kResTile; o += blockDim.x) {
        const long long m = m0 + o;
        if (m >= n_out) break;
        const long long c = m * down + half;
        const long long b = c / up;
        const int phi = (int)(c - b * up);
        const float* xp = s_x + (int)(b - b_lo);
        const float* tp = s_taps + phi;
        float acc = 0.f;
        // scipy's upfirdn accumulates from the oldest input sample to the newest
        for (int i = taps_per_phase - 1; i >= 0; --i) acc += tp[i * up] * xp[-i];
        out[(long long)row * out_stride + m] = acc;
    }
}

// [emul-begin]
constexpr int kRbR = 4;          // residues (outputs per period) per thread
constexpr int kRbW = 26;         // input window per thread and period
constexpr int kRbThreads = 128;
constexpr int kRbQI = 8;         // periods per thread and tile

struct ResampleRbParams {
    const float* in;
    long long in_stride;
    float* out;
    long long out_stride;
    long long n_in, n_out;
    int up, down;
    const float* taps;
    int n_taps, tpp, half;
    int G;               // residue groups = ceil(up / kRbR)
    int NQ;              // period lanes per CTA = kRbThreads / G
    int QT;              // periods per tile = NQ * kRbQI (multiple of 4)
    int s_min;           // smallest window start relative to a period's base input index (may be < 0)
    int span;            // floats per staged input span (multiple of 4)
    int tiles_per_row;
    long long total_tiles;
    int vec_in, vec_out;
};

__device__ __forceinline__ void rb_stage(const ResampleRbParams& p, long long tile, float* __restrict__ xs) {
    const int row = (int)(tile / p.tiles_per_row);
    const long long P0 = (tile - (long long)row * p.tiles_per_row) * p.QT;
    const long long x0 = (P0 * p.down + p.s_min) & ~3LL;       // floor to a multiple of 4 (also when negative)
    const float* __restrict__ src = p.in + (long long)row * p.in_stride;
    if (p.vec_in && x0 >= 0 && x0 + p.span <= p.n_in) {         // interior tile: no bounds checks (CTA-uniform)
        const float* __restrict__ s4 = src + x0;
        for (int v = threadIdx.x * 4; v < p.span; v += kRbThreads * 4) al_cp_async16(xs + v, s4 + v);
    } else {
        for (int v = threadIdx.x * 4; v < p.span; v += kRbThreads * 4) {
            const long long j = x0 + v;
            if (p.vec_in && j >= 0 && j + 4 <= p.n_in) {
                al_cp_async16(xs + v, src + j);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) xs[v + e] = (j + e >= 0 && j + e < p.n_in) ? __ldg(src + j + e) : 0.f;
            }
        }
    }
    al_cp_async_commit();
}

__global__ void __launch_bounds__(kRbThreads, 3)
resample_rb_kernel(const ResampleRbParams p) {
    AL_DYN_SMEM(float, smem);
    float* outs = smem + 2 * p.span;                           // [QT * up], 16-byte aligned
    const int tid = threadIdx.x;
    const int ql = tid / p.G, g = tid - ql * p.G;
    const bool active = ql < p.NQ;
    const int r0 = kRbR * g;
    const int cb0 = (r0 * p.down + p.half) / p.up;

    // this thread's taps as kRbR rows over the shared window: t[e][w] multiplies x[P*down + s_g + w]
    float t[kRbR][kRbW];
#pragma unroll
    for (int e = 0; e < kRbR; ++e) {
        const int r = r0 + e;
        const int ce = r * p.down + p.half;
        const int cbe = ce / p.up;
        const int phi = ce - cbe * p.up;
        const int d = cbe - cb0;
#pragma unroll
        for (int w = 0; w < kRbW; ++w) {
            const int i = d + p.tpp - 1 - w;
            const int k = phi + p.up * i;
            t[e][w] = (active && r < p.up && i >= 0 && i < p.tpp && k < p.n_taps) ? __ldg(p.taps + k) : 0.f;
        }
    }
    const int s_g = cb0 - (p.tpp - 1) - p.s_min;              // >= 0
    // outputs leave through shared memory; round k stores element e_k = (k + rot) & 3 so that the 32 lanes of
    // a warp hit 32 different banks.  The rotation is a two-stage select network on thread-constant bits.
    const int rot = (g >> 3) & 3;
    const bool rot1 = rot & 1, rot2 = rot & 2;
    int so[kRbR];                                              // outs offset of round k, -1 if that residue does not exist
#pragma unroll
    for (int k = 0; k < kRbR; ++k) {
        const int e = (k + rot) & 3;
        so[k] = (active && r0 + e < p.up) ? r0 + e : -1;
    }

    long long tile = blockIdx.x;
    int buf = 0;
    if (tile < p.total_tiles) rb_stage(p, tile, smem);
    for (; tile < p.total_tiles; tile += gridDim.x, buf ^= 1) {
        const long long next = tile + gridDim.x;
        if (next < p.total_tiles) { rb_stage(p, next, smem + (buf ^ 1) * p.span); al_cp_async_wait<1>(); }
        else al_cp_async_wait<0>();
        __syncthreads();
        const int row = (int)(tile / p.tiles_per_row);
        const long long P0 = (tile - (long long)row * p.tiles_per_row) * p.QT;
        const long long xs0 = P0 * p.down + p.s_min;
        const int lead = (int)(xs0 - (xs0 & ~3LL));
        if (active) {
            const float* __restrict__ xw = smem + buf * p.span + lead + s_g;
#pragma unroll 1
            for (int qi = 0; qi < kRbQI; ++qi) {
                const int q = ql + p.NQ * qi;
                const float* __restrict__ xp = xw + q * p.down;
                float x[kRbW];
#pragma unroll
                for (int w = 0; w < kRbW; ++w) x[w] = xp[w];
                float acc[kRbR];
#pragma unroll
                for (int e = 0; e < kRbR; ++e) acc[e] = 0.f;
#pragma unroll
                for (int w = 0; w < kRbW; ++w) {               // oldest input sample first, like upfirdn
#pragma unroll
                    for (int e = 0; e < kRbR; ++e) acc[e] += t[e][w] * x[w];
                }
                float b0 = rot1 ? acc[1] : acc[0], b1 = rot1 ? acc[2] : acc[1], b2 = rot1 ? acc[3] : acc[2],
                      b3 = rot1 ? acc[0] : acc[3];
                const float c[kRbR] = {rot2 ? b2 : b0, rot2 ? b3 : b1, rot2 ? b0 : b2, rot2 ? b1 : b3};
                float* __restrict__ o = outs + q * p.up;
#pragma unroll
                for (int k = 0; k < kRbR; ++k)
                    if (so[k] >= 0) o[so[k]] = c[k];
            }
        }
        __syncthreads();
        // tile rows are contiguous in the output: m = P0*up + f, f in [0, QT*up)
        const long long m0 = P0 * p.up;
        long long left = p.n_out - m0;
        const int n_valid = (int)(left < (long long)p.QT * p.up ? (left > 0 ? left : 0) : (long long)p.QT * p.up);
        float* __restrict__ dst = p.out + (long long)row * p.out_stride + m0;
        if (p.vec_out) {
            const int n4 = n_valid & ~3;
            for (int f = tid * 4; f < n4; f += kRbThreads * 4)
                *reinterpret_cast<float4*>(dst + f) = *reinterpret_cast<const float4*>(outs + f);
            if (tid < n_valid - n4) dst[n4 + tid] = outs[n4 + tid];
        } else {
            for (int f = tid; f < n_valid; f += kRbThreads) dst[f] = outs[f];
        }
    }
}
static bool resample_rb_plan(ResampleRbParams& p, int rows, size_t& smem) {
    p.tpp = (p.n_taps + p.up - 1) / p.up;
    p.half = (p.n_taps - 1) / 2;
    p.G = (p.up + kRbR - 1) / kRbR;
    if (p.G > kRbThreads) return false;
    p.NQ = kRbThreads / p.G;
    p.QT = p.NQ * kRbQI;
    // every row of a thread's window must fit: tpp + (largest input offset between its residues) <= kRbW
    int s_max = INT_MIN;
    p.s_min = INT_MAX;
    for (int g = 0; g < p.G; ++g) {
        const int cb0 = (kRbR * g * p.down + p.half) / p.up;
        for (int e = 0; e < kRbR && kRbR * g + e < p.up; ++e) {
            const int d = ((kRbR * g + e) * p.down + p.half) / p.up - cb0;
            if (p.tpp + d > kRbW) return false;
        }
        const int s = cb0 - (p.tpp - 1);
        p.s_min = s < p.s_min ? s : p.s_min;
        s_max = s > s_max ? s : s_max;
    }
    // int range of (r*down + half) and of the tile-local offsets
    if ((long long)p.up * p.down + p.half > INT_MAX / 2) return false;
    const int needed = (p.QT - 1) * p.down + (s_max - p.s_min) + kRbW;
    p.span = (needed + 3 + 3) & ~3;
    const int outs_floats = (p.QT * p.up + 3) & ~3;
    smem = ((size_t)2 * p.span + outs_floats) * sizeof(float);
    if (smem > 72 * 1024) return false;
    const long long periods = (p.n_out + p.up - 1) / p.up;
    p.tiles_per_row = (int)((periods + p.QT - 1) / p.QT);
    p.total_tiles = (long long)p.tiles_per_row * rows;
    p.vec_in = ((reinterpret_cast<uintptr_t>(p.in) & 15) == 0 && (p.in_stride & 3) == 0) ? 1 : 0;
    p.vec_out = ((reinterpret_cast<uintptr_t>(p.out) & 15) == 0 && (p.out_stride & 3) == 0 &&
                 ((long long)p.QT * p.up) % 4 == 0) ? 1 : 0;
    return true;
}
// [emul-end]

cudaError_t launch_resample(const float* in, long long in_stride, float* out, long long out_stride,
                            int rows, long long n_in, long long n_out, int up, int down,
                            const float* taps, int n_taps, cudaStream_t stream) {
    if (rows <= 0 || n_out <= 0) return cudaSuccess;
    static const bool force_generic = getenv("AL_FORCE_GENERIC") != nullptr;
    ResampleRbParams q{};
    q.in = in; q.in_stride = in_stride; q.out = out; q.out_stride = out_stride;
    q.n_in = n_in; q.n_out = n_out; q.up = up; q.down = down; q.taps = taps; q.n_taps = n_taps;
    size_t rb_smem = 0;
    if (!force_generic && resample_rb_plan(q, rows, rb_smem)) {
        static int ctas = 0;
        if (!ctas) {
            cudaError_t e = cudaFuncSetAttribute(resample_rb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
            if (e != cudaSuccess) return e;
            int dev = 0, sms = 0;
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            ctas = 3 * (sms > 0 ? sms : 148);
        }
        const unsigned grid = (unsigned)(q.total_tiles < ctas ? q.total_tiles : ctas);
        resample_rb_kernel<<<grid, kRbThreads, rb_smem, stream>>>(q);
        count_launch();
        return cudaGetLastError();
    }
    const int tpp = (n_taps + up - 1) / up;
    const int span_cap = (int)(((long long)kResTile * down) / up + tpp + 4);
    const size_t smem = ((size_t)tpp * up + span_cap) * sizeof(float);
    if (smem > 160 * 1024) return cudaErrorInvalidValue;
    static PerDeviceOnce attr_set;
    if (attr_set.needed()) {
        cudaError_t e = cudaFuncSetAttribute(resample_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        if (e != cudaSuccess) return e;
        attr_set.mark();
    }
    dim3 grid((unsigned)((n_out + kResTile - 1) / kResTile), (unsigned)rows);
    resample_generic_kernel<<<grid, 256, smem, stream>>>(in, in_stride, out, out_stride, n_in, n_out, up, down, taps,
                                                 n_taps, tpp, span_cap);
    count_launch();
    return cudaGetLastError();
}

}  // namespace al
