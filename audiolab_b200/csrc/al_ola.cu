// K2b: deterministic windowed overlap-add of chunk outputs into a track (gather form), and the
// elementwise complement stem.  Replaces the `result += x*window; counter += window; result/counter`
// loops of upstream MDXSeparator.demix / MDXCSeparator.demix / demucs.apply.
//
// HBM-bound: reads every chunk sample once (4 B) and writes every track sample once.
// One thread owns four consecutive positions of a row (128-bit loads and stores where the addresses
// allow it); every position adds its covering chunks in ascending chunk order, so the sum is
// bit-identical however the chunks were batched or sharded.  The covering chunks are walked in groups of
// four whose 128-bit loads (chunk + weight table) are all issued before the first dependent add.  The weight sum ("counter") is
// recomputed from the same tables instead of being stored.
#include <stdlib.h>

#include "al_kernels.h"

namespace al {

// [emul-begin]
constexpr int kOlaVec = 4;      // consecutive positions per thread
constexpr int kOlaGroup = 4;    // chunks whose loads are issued together

__device__ __forceinline__ bool al_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

// Loads v[e] = src[e] for the elements selected by `m` (bit e), 128-bit when all four are wanted and
// the address allows it; unselected elements read as 0.
__device__ __forceinline__ void ola_load4(const float* __restrict__ src, unsigned m, float (&v)[kOlaVec]) {
    if (m == 0xFu && al_aligned16(src)) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(src));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
        for (int e = 0; e < kOlaVec; ++e) v[e] = ((m >> e) & 1u) ? __ldg(src + e) : 0.f;
    }
}

// RB = rows handled by one thread: the offsets, masks, weights and the weight sum of a position are the same
// for every row (channel / stem), so they are computed once and only the chunk samples are loaded per row.
template <int RB>
__global__ void __launch_bounds__(256, RB == 1 ? 3 : 2)
ola_gather_kernel(const float* __restrict__ chunks, int n_chunks, int data_chunk0, int rows, int chunk_len,
                  const long long* __restrict__ offsets, const int* __restrict__ mult,
                  const float* __restrict__ wtab, const int* __restrict__ tab_id, long long n_total,
                  long long p0, long long p1, const float* __restrict__ halo_in, int raw_out, float eps,
                  float scale, float* __restrict__ track, long long track_stride) {
    // the CTA's first position decides where the chunk walk starts (uniform loads, one search per CTA
    // instead of one per sample); chunks that end before a thread's own positions are skipped below
    const long long pb = p0 + (long long)blockIdx.x * blockDim.x * kOlaVec;
    const long long p = pb + (long long)threadIdx.x * kOlaVec;
    if (p >= p1) return;
    int lo = 0, hi = n_chunks;
    while (lo < hi) {   // offsets ascending: first chunk with off_c + chunk_len > pb
        const int mid = (lo + hi) >> 1;
        if (__ldg(offsets + mid) + chunk_len > pb) hi = mid; else lo = mid + 1;
    }
    const int c_first = lo;
    const unsigned own = p + kOlaVec <= p1 ? 0xFu : ((1u << (int)(p1 - p)) - 1u);   // positions inside [p0, p1)
    for (int r0 = blockIdx.y * RB; r0 < rows; r0 += gridDim.y * RB) {               // rows is a multiple of RB
        float acc[RB][kOlaVec], wsum[kOlaVec];
#pragma unroll
        for (int b = 0; b < RB; ++b) {
            if (halo_in) ola_load4(halo_in + (long long)(r0 + b) * (p1 - p0) + (p - p0), own, acc[b]);
            else {
#pragma unroll
                for (int e = 0; e < kOlaVec; ++e) acc[b][e] = 0.f;
            }
        }
#pragma unroll
        for (int e = 0; e < kOlaVec; ++e) wsum[e] = 0.f;
        // covering chunks are consecutive (offsets ascending): walk them in groups of kOlaGroup, issuing every
        // load of a group before its first dependent add
        for (int c0 = c_first; c0 < n_chunks; c0 += kOlaGroup) {
            if (__ldg(offsets + c0) > p + (kOlaVec - 1)) break;      // ascending: nothing further covers us
            float x[kOlaGroup][RB][kOlaVec], w[kOlaGroup][kOlaVec];
            unsigned msk[kOlaGroup];
            int mm[kOlaGroup];
#pragma unroll
            for (int g = 0; g < kOlaGroup; ++g) {
                const int c = c0 + g;
                msk[g] = 0;
                mm[g] = 1;
                if (c >= n_chunks) continue;
                const long long off = __ldg(offsets + c);
                const long long j = p - off;                         // may be negative (chunk starts inside the group)
                const long long len = min((long long)chunk_len, n_total - off);
                if (off > p + (kOlaVec - 1) || j >= len) continue;   // starts after / ends before this group
                unsigned m4 = own;
                if (j < 0 || j + kOlaVec > len) {
#pragma unroll
                    for (int e = 0; e < kOlaVec; ++e)
                        if (j + e < 0 || j + e >= len) m4 &= ~(1u << e);
                }
                msk[g] = m4;
                if (mult) mm[g] = __ldg(mult + c);                   // the reference re-adds a tail chunk m times
                if (wtab) ola_load4(wtab + (long long)(tab_id ? __ldg(tab_id + c) : 0) * chunk_len + j, m4, w[g]);
                else {
#pragma unroll
                    for (int e = 0; e < kOlaVec; ++e) w[g][e] = 1.f;
                }
                // chunks before data_chunk0 belong to the left neighbour: their partial sums arrive through
                // halo_in, only their weights are counted here
#pragma unroll
                for (int b = 0; b < RB; ++b) {
                    if (c >= data_chunk0)
                        ola_load4(chunks + ((long long)(c - data_chunk0) * rows + r0 + b) * chunk_len + j, m4, x[g][b]);
                    else {
#pragma unroll
                        for (int e = 0; e < kOlaVec; ++e) x[g][b][e] = 0.f;
                    }
                }
            }
#pragma unroll
            for (int g = 0; g < kOlaGroup; ++g) {
                if (!msk[g]) continue;
                for (int k = 0; k < mm[g]; ++k) {                    // keep the reference's rounding: m separate adds
#pragma unroll
                    for (int e = 0; e < kOlaVec; ++e) {
                        if (msk[g] == 0xFu || ((msk[g] >> e) & 1u)) {
#pragma unroll
                            for (int b = 0; b < RB; ++b) acc[b][e] += x[g][b][e] * w[g][e];
                            wsum[e] += w[g][e];
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int b = 0; b < RB; ++b) {
            float v[kOlaVec];
#pragma unroll
            for (int e = 0; e < kOlaVec; ++e) v[e] = raw_out ? acc[b][e] : scale * acc[b][e] / fmaxf(wsum[e], eps);
            float* dst = track + (long long)(r0 + b) * track_stride + p;
            if (own == 0xFu && al_aligned16(dst)) *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
            else {
#pragma unroll
                for (int e = 0; e < kOlaVec; ++e)
                    if ((own >> e) & 1u) dst[e] = v[e];
            }
        }
    }
}
// [emul-end]

cudaError_t launch_ola_gather(const float* chunks, int n_chunks, int data_chunk0, int rows, int chunk_len,
                              const long long* offsets, const int* mult, const float* wtab,
                              const int* tab_id, long long n_total, long long p0, long long p1,
                              const float* halo_in, int raw_out, float eps, float scale, float* track,
                              long long track_stride, cudaStream_t stream) {
    if (p1 <= p0) return cudaSuccess;
    const long long n = p1 - p0;
    const long long per_cta = 256LL * kOlaVec;
    static const bool one_row = getenv("AL_OLA_RB1") != nullptr;
#define AL_OLA_LAUNCH(RB_)                                                                                         \
    do {                                                                                                           \
        dim3 grid((unsigned)((n + per_cta - 1) / per_cta), (unsigned)min(rows / RB_, 8));                          \
        ola_gather_kernel<RB_><<<grid, 256, 0, stream>>>(chunks, n_chunks, data_chunk0, rows, chunk_len, offsets, mult, wtab, \
                                                         tab_id, n_total, p0, p1, halo_in, raw_out, eps, scale, track,       \
                                                         track_stride);                                            \
    } while (0)
    if (rows % 2 == 0 && !one_row) AL_OLA_LAUNCH(2); else AL_OLA_LAUNCH(1);
#undef AL_OLA_LAUNCH
    count_launch();
    return cudaGetLastError();
}

__global__ void __launch_bounds__(256)
sub_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, long long n) {
    const long long i4 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i4 + 3 < n) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(a + i4));
        const float4 y = __ldg(reinterpret_cast<const float4*>(b + i4));
        *reinterpret_cast<float4*>(out + i4) = make_float4(x.x - y.x, x.y - y.y, x.z - y.z, x.w - y.w);
    } else {
        for (long long i = i4; i < n; ++i) out[i] = a[i] - b[i];
    }
}

cudaError_t launch_sub(const float* a, const float* b, float* out, long long n, cudaStream_t stream) {
    if (n <= 0) return cudaSuccess;
    if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(out)) & 15)
        return cudaErrorMisalignedAddress;
    sub_kernel<<<(unsigned)((n + 1023) / 1024), 256, 0, stream>>>(a, b, out, n);
    count_launch();
    return cudaGetLastError();
}

}  // namespace al
