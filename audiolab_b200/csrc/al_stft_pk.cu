// K1 fast path: stereo, n_fft = 2048, bin-innermost output layouts (FRAME_MAJOR / FRAME_INTERLEAVED).
// Fused pad-and-chunk + reflect-pad + framing + window + R2C FFT + consumer layout, like stft_kernel<2>
// (reference: modules/rvc/infer/modules/uvr5/mdxnet.py:41-56; upstream BSRoformer.forward stft,
// SURVEY.md A.0 / A.2), restructured around the instruction budget measured in profiles/r01a:
//
//  * both channels ride in one packed-fp32 FFT (al_fftp.cuh): half the issue slots per frame;
//  * one warp owns one frame end to end: samples -> window -> 1024-point FFT -> lane-mirror separation
//    -> radix-2 combine in registers -> 128-bit stores of (L, R) bins.  No CTA-wide barrier, no
//    spectrum round trip through shared memory;
//  * a producer warp stages the sample span of the next tile (8 frames) into shared memory with
//    128-bit loads while the 8 compute warps work (mbarrier full / empty), so compute warps never
//    wait on HBM; edge tiles (reflection, zero padding, unaligned rows) take an element-wise path in
//    the producer only;
//  * the 16 KB spectrum row of a frame is staged in the warp's transposition scratch and leaves through
//    one bulk asynchronous store (cp.async.bulk shared -> global): plain STG.128 streams drain at the
//    SM's ~20 B/clk store rate and stall every other LDS/STS of the SM behind them (profiles/r01b);
//  * persistent grid: one CTA per SM, tables loaded once.
//
// Algorithmic bytes per frame (both channels): 2 * (hop*4 + n_bins_out*8).
#include "al_fftp.cuh"
#include "al_kernels.h"

namespace al {

// [emul-begin]
constexpr int kPkWarps = 8;          // compute warps = frames per tile (2 per SM sub-partition) + 1 producer
constexpr int kPkProd = 4;           // producer warps (memory-level parallelism of the stage fill)
constexpr int kPkThreads = (kPkWarps + kPkProd) * 32;
constexpr int kPkMaxStages = 4;
constexpr int kPkCtw = 544;          // combine twiddles (513 used, padded for the r = 16 row)

template <int LAYOUT, bool FULL>
__global__ void __launch_bounds__(kPkThreads, 1)
stft_pk2_kernel(const StftPkParams p) {
    AL_DYN_SMEM(unsigned char, smem_raw);
    float2* s_tw = reinterpret_cast<float2*>(smem_raw);          // [1024]
    float2* s_win = s_tw + 1024;                                  // [1024] window pairs (w[2m], w[2m+1])
    float2* s_ctw = s_win + 1024;                                 // [kPkCtw] 0.5 * exp(-2 pi i k / 2048)
    float4* s_scr = reinterpret_cast<float4*>(s_ctw + kPkCtw);    // [kPkWarps][kScrF4] transposition / output row
    float2* s_stage = reinterpret_cast<float2*>(s_scr + kPkWarps * kScrF4);                  // [n_stages][2][sp]  (L, R) sample pairs
    uint64_t* s_full = reinterpret_cast<uint64_t*>(s_stage + (size_t)p.n_stages * 2 * p.sp);
    uint64_t* s_empty = s_full + kPkMaxStages;
    int* s_shift = reinterpret_cast<int*>(s_empty + kPkMaxStages);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int NS = p.n_stages, T = p.n_frames, hop = p.hop, sp = p.sp;

    for (int i = tid; i < 1024; i += kPkThreads) {
        s_tw[i] = p.tw[i];
        s_win[i] = reinterpret_cast<const float2*>(p.window)[i];
    }
    for (int i = tid; i < kPkCtw; i += kPkThreads) s_ctw[i] = p.ctw_half[i];
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) {
            mbar_init(&s_full[s], 32 * kPkProd);
            mbar_init(&s_empty[s], kPkWarps * 32);
        }
    }
    __syncthreads();

    if (warp >= kPkWarps) {
        const int pl = (warp - kPkWarps) * 32 + lane;   // producer lane id
        constexpr int PL = 32 * kPkProd;
        // ================= producer: stage the sample span of each tile =================
        const float* __restrict__ src0 = p.track;
        const float* __restrict__ src1 = p.track + p.ch_stride;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
            const int s = it % NS;
            mbar_wait(&s_empty[s], ((it / NS) & 1) ^ 1);
            const int chunk = tile / p.tiles_per_chunk;
            const int t0 = (tile - chunk * p.tiles_per_chunk) * kPkWarps;
            const int nf = min(kPkWarps, T - t0);
            const long long coff = p.chunk_offsets ? p.chunk_offsets[chunk] : p.off0 + (long long)chunk * p.off_step;
            const long long a = (long long)t0 * hop - p.center;                    // chunk-local span [a, b)
            const long long b = (long long)(t0 + nf - 1) * hop - p.center + 2048;
            float2* st = s_stage + (size_t)s * 2 * sp;
            const bool fast = p.aligned && a >= 0 && b <= p.chunk_len && coff + a >= 0 &&
                              ((coff + b + 3) & ~3LL) <= p.n_valid;
            int shift = 0;
            if (fast) {
                const long long ga = (coff + a) & ~3LL;
                shift = (int)(coff + a - ga);
                const int ng = (shift + (int)(b - a) + 3) >> 2;
                const float4* __restrict__ g0 = reinterpret_cast<const float4*>(src0 + ga);
                const float4* __restrict__ g1 = reinterpret_cast<const float4*>(src1 + ga);
                float4* ev = reinterpret_cast<float4*>(st);
                float4* od = reinterpret_cast<float4*>(st + sp);
                constexpr int U = 11;
                for (int j0 = 0; j0 < ng; j0 += PL * U) {
                    float4 v0[U], v1[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int j = j0 + u * PL + pl;
                        if (j < ng) {
                            v0[u] = __ldg(g0 + j);
                            v1[u] = __ldg(g1 + j);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int j = j0 + u * PL + pl;
                        if (j < ng) {
                            ev[j] = make_float4(v0[u].x, v1[u].x, v0[u].z, v1[u].z);
                            od[j] = make_float4(v0[u].y, v1[u].y, v0[u].w, v1[u].w);
                        }
                    }
                }
            } else {
                // reflect about the chunk (torch.stft center=True), zero outside the track; loads are
                // issued in batches so that an edge tile costs a few memory round trips, not one per element
                const int span = (int)(b - a);
                constexpr int U = 12;
                for (int i0 = 0; i0 < span; i0 += PL * U) {
                    float2 v[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int i = i0 + u * PL + pl;
                        long long j = a + i;
                        if (j < 0) j = -j;
                        if (j >= p.chunk_len) j = 2LL * (p.chunk_len - 1) - j;
                        const long long g = coff + j;
                        v[u] = make_float2(0.f, 0.f);
                        if (i < span && j >= 0 && j < p.chunk_len && g >= 0 && g < p.n_valid)
                            v[u] = make_float2(__ldg(src0 + g), __ldg(src1 + g));
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int i = i0 + u * PL + pl;
                        if (i < span) st[(i & 1) * sp + (i >> 1)] = v[u];
                    }
                }
            }
            if (pl == 0) s_shift[s] = shift;
            mbar_arrive(&s_full[s]);
        }
        return;
    }

    // ================= compute warps: one frame (both channels) per tile =================
    float4* scr = s_scr + warp * kScrF4;
    bool store_pending = false;
    const int Fo = p.n_bins_out;
    const int ml = (32 - lane) & 31;
    int it = 0;
#ifdef AL_PK_PROF
    long long pf[5] = {0, 0, 0, 0, 0};
#define PK_MARK(k) { const long long now_ = clock64(); pf[k] += now_ - tprev_; tprev_ = now_; }
    long long tprev_ = clock64();
#else
#define PK_MARK(k)
#endif
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int s = it % NS;
        const int chunk = tile / p.tiles_per_chunk;
        const int t = (tile - chunk * p.tiles_per_chunk) * kPkWarps + warp;
        PK_MARK(4)
        mbar_wait(&s_full[s], (it / NS) & 1);
        PK_MARK(0)
        float2 re[32], im[32];
        const bool live = t < T;   // warp-uniform
        if (live) {
            const float2* st = s_stage + (size_t)s * 2 * sp;
            const int i0 = s_shift[s] + warp * hop;
            const float2* A = st + (i0 & 1) * sp + (i0 >> 1) + lane;            // samples i0 + 2 m
            const float2* B = st + ((i0 & 1) ^ 1) * sp + ((i0 + 1) >> 1) + lane;   // samples i0 + 2 m + 1
#pragma unroll
            for (int r = 0; r < 32; ++r) {
#ifdef AL_PK_ABL_NOWIN
                const float2 w = make_float2(0.5f, 0.25f);
#else
                const float2 w = s_win[32 * r + lane];
#endif
                re[r] = pscale(A[32 * r], w.x);
                im[r] = pscale(B[32 * r], w.y);
            }
        }
        mbar_arrive(&s_empty[s]);
        if (!live) continue;
        PK_MARK(1)

        if (LAYOUT == 3 && store_pending) {   // the previous row must have left the scratch
            if (lane == 0) bulk_wait_read_all();
            __syncwarp();
        }
        warp_fft1024p_wide<false>(re, im, scr, s_tw, lane);
        PK_MARK(2)

        // Z = FFT(x_even + i x_odd): separate the two real spectra with the lane-mirror partner
        // Z[1024 - kappa] and merge them (radix 2):  X[kappa] = Xe + W^kappa Xo,  X[1024 - kappa] = conj(Xe - W^kappa Xo)
        float4* o4 = nullptr;
        float2 *o0 = nullptr, *o1 = nullptr;
        if (LAYOUT == 3) {
            o4 = reinterpret_cast<float4*>(p.spec) + ((long long)chunk * T + t) * Fo;
        } else {
            o0 = reinterpret_cast<float2*>(p.spec) + ((long long)(chunk * 2) * T + t) * Fo;
            o1 = o0 + (long long)T * Fo;
        }
#pragma unroll
        for (int r = 0; r <= 16; ++r) {
            const int ra = 31 - r, rb = (32 - r) & 31;
            const float2 sr = lane == 0 ? re[rb] : re[ra];
            const float2 si = lane == 0 ? im[rb] : im[ra];
            const float2 pr = shfl2(sr, ml), pi = shfl2(si, ml);
            const float2 er = padd(re[r], pr), ei = psub(im[r], pi);     // 2 Xe
            const float2 orr = padd(im[r], pi), oi = psub(pr, re[r]);    // 2 Xo
            const float2 w = s_ctw[32 * r + lane];                       // 0.5 W^kappa
            const float2 tr = pfma(oi, -w.y, pscale(orr, w.x));
            const float2 ti = pfma(oi, w.x, pscale(orr, w.y));
            const int kappa = 32 * r + lane, mb = 1024 - kappa;
            float4 v = make_float4(fmaf(0.5f, er.x, tr.x), fmaf(0.5f, ei.x, ti.x),
                                   fmaf(0.5f, er.y, tr.y), fmaf(0.5f, ei.y, ti.y));
            float4 m = make_float4(fmaf(0.5f, er.x, -tr.x), fmaf(-0.5f, ei.x, ti.x),
                                   fmaf(0.5f, er.y, -tr.y), fmaf(-0.5f, ei.y, ti.y));
            bool st_v = (r < 16) || lane == 0;
            bool st_m = (r < 16);
            if (!FULL) {
                st_v = st_v && kappa < Fo;
                st_m = st_m && mb < Fo;
                if (kappa < p.zero_low_bins) v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (mb < p.zero_low_bins) m = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (LAYOUT == 3) {
                if (st_v) scr[kappa] = v;
                if (st_m) scr[mb] = m;
            } else {
                if (st_v) { o0[kappa] = make_float2(v.x, v.y); o1[kappa] = make_float2(v.z, v.w); }
                if (st_m) { o0[mb] = make_float2(m.x, m.y); o1[mb] = make_float2(m.z, m.w); }
            }
        }
        if (LAYOUT == 3) {
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                bulk_store_s2g(o4, scr, (uint32_t)Fo * 16u);
                bulk_commit();
            }
            store_pending = true;
        }
        PK_MARK(3)
    }
    if (LAYOUT == 3 && store_pending && lane == 0) bulk_wait_all();
#ifdef AL_PK_PROF
    if (lane == 0 && p.prof) for (int k = 0; k < 5; ++k) p.prof[(blockIdx.x * kPkWarps + warp) * 5 + k] = pf[k];
#endif
}


static int pk_sp(int hop) {
    const int span = (kPkWarps - 1) * hop + 2048 + 3;
    return (((span + 1) / 2 + 2) + 1) & ~1;
}

// launch shape of stft_pk2_kernel: fills sp / tiles / n_stages, returns the dynamic shared memory size (0 = does not fit)
static size_t pk_launch_shape(StftPkParams& p) {
    p.sp = pk_sp(p.hop);
    p.tiles_per_chunk = (p.n_frames + kPkWarps - 1) / kPkWarps;
    p.total_tiles = p.tiles_per_chunk * p.n_chunks;
    const size_t fixed = (size_t)(1024 + 1024 + kPkCtw) * sizeof(float2) + (size_t)kPkWarps * kScrF4 * sizeof(float4) +
                         2 * kPkMaxStages * sizeof(uint64_t) + kPkMaxStages * sizeof(int) + 16;
    const size_t per_stage = (size_t)2 * p.sp * sizeof(float2);
    const size_t cap = 227 * 1024;
    int ns = (int)((cap - fixed) / per_stage);
    if (ns < 1) return 0;
    if (ns > 3) ns = 3;
    p.n_stages = ns;
    return fixed + ns * per_stage;
}
// [emul-end]

cudaError_t launch_stft_pk(const StftPkParams& p0, cudaStream_t stream) {
    StftPkParams p = p0;
    const size_t cap = 227 * 1024;
    const size_t smem = pk_launch_shape(p);
    if (smem == 0) return cudaErrorInvalidValue;
    const int n_sm = sm_count();
    const bool full = p.n_bins_out == 1025 && p.zero_low_bins == 0;
    const unsigned grid = (unsigned)(p.total_tiles < n_sm ? p.total_tiles : n_sm);
#define AL_PK_LAUNCH(L, F)                                                                                   \
    do {                                                                                                      \
        static PerDeviceOnce attr;                                                                             \
        if (attr.needed()) {                                                                                          \
            cudaError_t e = cudaFuncSetAttribute(stft_pk2_kernel<L, F>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                 (int)cap);                                                   \
            if (e != cudaSuccess) return e;                                                                   \
            attr.mark();                                                                                      \
        }                                                                                                     \
        stft_pk2_kernel<L, F><<<grid, kPkThreads, smem, stream>>>(p);                                        \
    } while (0)
    if (p.layout == 3) { if (full) AL_PK_LAUNCH(3, true); else AL_PK_LAUNCH(3, false); }
    else               { if (full) AL_PK_LAUNCH(0, true); else AL_PK_LAUNCH(0, false); }
#undef AL_PK_LAUNCH
    count_launch();
    return cudaGetLastError();
}

}  // namespace al
