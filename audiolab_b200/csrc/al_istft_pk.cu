// K2 fast path: stereo, n_fft = 2048, FRAME_INTERLEAVED spectrum (and mask) -- the RoFormer shape.
// Fused complex mask (.) spec + Hermitian extension + C2R iFFT + synthesis window + overlap-add over
// frames + / sum(window^2) + centre trim (+ chunk weight, + placement), like istft_kernel<2>
// (reference: modules/rvc/infer/modules/uvr5/mdxnet.py:58-75; upstream BSRoformer.forward
// "stft_repr * mask -> torch.istft"), restructured like the K1 fast path:
//
//  * both channels ride in one packed-fp32 inverse FFT (al_fftp.cuh): a spectrum / mask element of
//    layout 3 is one 16-byte (L.re, L.im, R.re, R.im) load;
//  * one warp owns one frame from the HBM row to the windowed time-domain frame.  The packed
//    1024-point input  Z[k] = A[k] + i B[k],  A = Y[k] + conj(Y[1024-k]),
//    B = (Y[k] - conj(Y[1024-k])) conj(W^k)  (Y = mask (.) spec, W = exp(-2 pi i / 2048)) is built
//    straight from global memory: the lane that owns Z[k] loads Y[k] and Y[1024-k] itself (the two
//    loads of a register pair (r, 31-r) fall on the same 128-byte lines and are issued back to
//    back), so there is no shuffle, no lane-0 special case and no spectrum staging in shared memory;
//  * HBM latency is covered by bulk L2 prefetches (cp.async.bulk.prefetch.L2) of the rows of the next
//    a round ahead and by a four-deep register pipeline of loads (optionally, AL_IP_PRE, also by issuing the
//    first stages of the NEXT round's frame before this round's overlap-add -- measured slower, off by default);
//  * the windowed frame is parked in the warp's own transposition scratch; after one CTA barrier all
//    threads gather-sum the round's frames (+ the carry of earlier rounds) in ascending frame
//    order -- the deterministic order of istft_kernel -- and emit both channels.
//
// Algorithmic bytes per frame (both channels): 2 * (1025*8 [+ 1025*8 mask] + hop*4).
#include <stdlib.h>

#include "al_fftp.cuh"
#include "al_kernels.h"

namespace al {

// [emul-begin]
// W = warps = frames per round.  W = 4 (default): two CTAs share an SM, so one's row loads overlap the
// other's FFT / overlap-add; W = 8: one CTA per SM, half the halo recomputation (AL_IP_WARPS=8 selects it).
constexpr int kIpKMax = 5;               // frames covering one position in the register form of the overlap-add (hop >= 410)
constexpr int kIpSmWarps = 8;             // 255 registers per thread: 8 warps fill the register file
constexpr int kIpN = 2048;
constexpr int kIpBins = 1025;
#ifndef AL_IP_DEPTH
#define AL_IP_DEPTH 4
#endif
// PRE (template) = stages of the NEXT round's frame loaded before this round's overlap-add (AL_IP_PRE selects; default 0:
// holding 64-96 load registers across the overlap-add costs more -- ptxas can no longer hoist the round's own loads --
// than the hidden latency gains, profiles/r01l_kernel_bench_pre*.jsonl)
constexpr int kIpDepth = AL_IP_DEPTH;    // register pipeline depth of the row loads (stages of 4 + 4 x 16 B per lane)

__device__ __forceinline__ void prefetch_l2_bulk(const void* g, uint32_t bytes) {
#ifndef AL_CPU_EMUL
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(g), "r"(bytes) : "memory");
#else
    (void)g; (void)bytes;
#endif
}

// streaming read-only load that does not take a line of the (few KB of) L1 the window table lives in
__device__ __forceinline__ float ldg_stream(const float* g) {
#ifndef AL_CPU_EMUL
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(g));
    return v;
#else
    return *g;
#endif
}
__device__ __forceinline__ void prefetch_l2(const void* g) {
#ifndef AL_CPU_EMUL
    asm volatile("prefetch.global.L2 [%0];" ::"l"(g));
#else
    (void)g;
#endif
}
__device__ __forceinline__ void prefetch_l1(const void* g) {
#ifndef AL_CPU_EMUL
    asm volatile("prefetch.global.L1 [%0];" ::"l"(g));
#else
    (void)g;
#endif
}

// Y = x * m per channel; returns packed (L, R) real and imaginary parts
template <bool MASK>
__device__ __forceinline__ void ip_product(const float4 x, const float4 m, float2& yr, float2& yi) {
    if (MASK) {
        yr = make_float2(fmaf(-x.y, m.y, x.x * m.x), fmaf(-x.w, m.w, x.z * m.z));
        yi = make_float2(fmaf(x.y, m.x, x.x * m.y), fmaf(x.w, m.z, x.z * m.w));
    } else {
        yr = make_float2(x.x, x.z);
        yi = make_float2(x.y, x.w);
    }
}

// Z[k] from P = Y[k], Q = Y[1024 - k], w = W^k
__device__ __forceinline__ void ip_combine(float2 pr, float2 pi, float2 qr, float2 qi, float2 w, float2& zr,
                                           float2& zi) {
    const float2 ar = padd(pr, qr), ai = psub(pi, qi);     // A = P + conj(Q)
    const float2 dr = psub(pr, qr), di = padd(pi, qi);     // D = P - conj(Q)
    zr = pfma(dr, w.y, pfma(di, -w.x, ar));                 // A.re - Im(D conj(w))
    zi = pfma(di, w.y, pfma(dr, w.x, ai));                  // A.im + Re(D conj(w))
}

template <bool MASK, int W, int PRE>
__global__ void __launch_bounds__(W * 32, kIpSmWarps / W)
istft_pk2_kernel(const IstftPkParams p) {
    constexpr int kIpWarps = W, kIpThreads = W * 32, kIpPre = PRE;
    static_assert(PRE <= kIpDepth - 1, "preloaded stages must fit the register ring");
    AL_DYN_SMEM(unsigned char, smem_raw);
    float2* s_tw = reinterpret_cast<float2*>(smem_raw);           // [1024]
    float2* s_win = s_tw + 1024;                                   // [1024] (w[2k], w[2k+1])
    float2* s_ctw = s_win + 1024;                                  // [1024] W^k
    float4* s_scr = reinterpret_cast<float4*>(s_ctw + 1024);       // [kIpWarps][kScrF4]
    float2* s_carry = reinterpret_cast<float2*>(s_scr + kIpWarps * kScrF4);   // [2048 - hop] (L, R), updated in place

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = blockIdx.x / p.segs, seg = blockIdx.x - g * p.segs;   // g = chunk*stems + stem
    const int chunk = g / p.stems, stem = g - chunk * p.stems;
    const int hop = p.hop, T = p.n_frames;
    const int carry_len = kIpN - hop;
    const int kOla_K = (kIpN + hop - 1) / hop;          // frames that can cover one position

    // owned untrimmed overlap-add positions [Pa, Pb)
    const long long Pa = (long long)p.out_start + (long long)seg * p.hops_per_cta * hop;
    const long long Pend = (long long)p.out_start + p.out_len;
    const long long Pb = min(Pa + (long long)p.hops_per_cta * hop, Pend);
    if (Pa >= Pb) return;
    int ta = (int)((Pa - kIpN) / hop) + 1;              // first frame touching Pa
    if (Pa < kIpN) ta = 0;
    ta = max(ta, 0);
    const int tb = min((int)((Pb - 1) / hop), T - 1);   // last frame touching Pb - 1
    const int t_last = (int)((Pb - 1) / hop);           // rounds run until the carry is flushed up to Pb

    const long long place = p.dst_offsets ? p.dst_offsets[chunk] : p.dst_off0 + (long long)chunk * p.dst_off_step;
    float* __restrict__ dst0 = p.dst + ((long long)stem * 2) * p.dst_ch_stride + (long long)chunk * p.dst_chunk_stride + place;
    float* __restrict__ dst1 = dst0 + p.dst_ch_stride;

    const float4* __restrict__ X = p.spec + (long long)(p.spec_has_stems ? g : chunk) * T * kIpBins;
    const float4* __restrict__ M = MASK ? p.mask + (long long)g * T * kIpBins : nullptr;

    // pair r: k1 = 32 r + lane (reg r), k2 = 32 (31 - r) + lane (reg 31 - r); mirrors q = 1024 - k
    float4 bx[kIpDepth][4], bm[kIpDepth][4];
#define IP_ISSUE(xrow_, mrow_, r_, b_)                                            \
    do {                                                                          \
        const int k1_ = 32 * (r_) + lane, k2_ = 32 * (31 - (r_)) + lane;          \
        bx[b_][0] = __ldg((xrow_) + k1_);                                         \
        bx[b_][1] = __ldg((xrow_) + (1024 - k2_));                                \
        bx[b_][2] = __ldg((xrow_) + k2_);                                         \
        bx[b_][3] = __ldg((xrow_) + (1024 - k1_));                                \
        if (MASK) {                                                               \
            bm[b_][0] = __ldg((mrow_) + k1_);                                     \
            bm[b_][1] = __ldg((mrow_) + (1024 - k2_));                            \
            bm[b_][2] = __ldg((mrow_) + k2_);                                     \
            bm[b_][3] = __ldg((mrow_) + (1024 - k1_));                            \
        }                                                                         \
    } while (0)
    if (ta + warp <= tb) {   // the first round's frame: its first stages leave now, the second round's rows go to L2
        const float4* __restrict__ xr = X + (long long)(ta + warp) * kIpBins;
        const float4* __restrict__ mr = MASK ? M + (long long)(ta + warp) * kIpBins : nullptr;
#pragma unroll
        for (int r = 0; r < kIpPre; ++r) IP_ISSUE(xr, mr, r, r);
        if (lane == 0 && p.l2_prefetch) {
            if (kIpPre == 0) {
                prefetch_l2_bulk(xr, kIpBins * 16);
                if (MASK) prefetch_l2_bulk(mr, kIpBins * 16);
            } else if (ta + warp + kIpWarps <= tb) {
                prefetch_l2_bulk(xr + (long long)kIpWarps * kIpBins, kIpBins * 16);
                if (MASK) prefetch_l2_bulk(mr + (long long)kIpWarps * kIpBins, kIpBins * 16);
            }
        }
    }
    for (int i = tid; i < 1024; i += kIpThreads) {
        s_tw[i] = p.tw[i];
        s_win[i] = reinterpret_cast<const float2*>(p.window)[i];
        s_ctw[i] = p.ctw[i];
    }
    for (int i = tid; i < carry_len; i += kIpThreads) s_carry[i] = make_float2(0.f, 0.f);
    __syncthreads();

    float4* scr = s_scr + warp * kScrF4;
    for (int tr = ta; tr <= t_last; tr += kIpWarps) {
        const int nf = max(0, min(kIpWarps, tb - tr + 1));   // live frames of this round (CTA-uniform)
        const int t = tr + warp;
        const bool live = warp < nf;                          // warp-uniform
        float2 re[32], im[32];
        if (live) {
            constexpr int kAhead = kIpPre ? 2 : 1;   // rounds ahead of the L2 prefetch
            if (lane == 0 && p.l2_prefetch && t + kAhead * kIpWarps <= tb) {
                prefetch_l2_bulk(X + (long long)(t + kAhead * kIpWarps) * kIpBins, kIpBins * 16);
                if (MASK) prefetch_l2_bulk(M + (long long)(t + kAhead * kIpWarps) * kIpBins, kIpBins * 16);
            }
            // ---- build Z from the spectrum (and mask) rows: register pairs (r, 31 - r) --------------
            // stages 0 .. kIpPre - 1 were issued before the previous round's overlap-add
            const float4* __restrict__ xrow = X + (long long)t * kIpBins;
            const float4* __restrict__ mrow = MASK ? M + (long long)t * kIpBins : nullptr;
#pragma unroll
            for (int r = kIpPre; r < kIpDepth - 1; ++r) IP_ISSUE(xrow, mrow, r, r);
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const int b = r % kIpDepth;
                if (r + kIpDepth - 1 < 16) IP_ISSUE(xrow, mrow, r + kIpDepth - 1, (r + kIpDepth - 1) % kIpDepth);
                float2 p1r, p1i, q2r, q2i, p2r, p2i, q1r, q1i;
                ip_product<MASK>(bx[b][0], bm[b][0], p1r, p1i);
                ip_product<MASK>(bx[b][1], bm[b][1], q2r, q2i);
                ip_product<MASK>(bx[b][2], bm[b][2], p2r, p2i);
                ip_product<MASK>(bx[b][3], bm[b][3], q1r, q1i);
                if (r == 0 && lane == 0) {   // k1 = 0: C2R ignores Im of the DC and Nyquist bins
                    p1i = make_float2(0.f, 0.f);
                    q1i = make_float2(0.f, 0.f);
                }
                const float2 w1 = s_ctw[32 * r + lane], w2 = s_ctw[32 * (31 - r) + lane];
                ip_combine(p1r, p1i, q1r, q1i, w1, re[r], im[r]);
                ip_combine(p2r, p2i, q2r, q2i, w2, re[31 - r], im[31 - r]);
            }
        }
        __syncthreads();   // the previous round's overlap-add has finished reading the scratches
        if (live) {
            warp_fft1024p_wide<true>(re, im, scr, s_tw, lane);
            // z[k] = (x[2k], x[2k+1]) * n_fft; window (carries 1 / n_fft) and park the frame: scr[k] = samples 2k, 2k+1
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                const float2 w = s_win[32 * r + lane];
                const float2 ev = pscale(re[r], w.x), od = pscale(im[r], w.y);
                scr[32 * r + lane] = make_float4(ev.x, ev.y, od.x, od.y);
            }
        }
        if (kIpPre > 0 && t + kIpWarps <= tb) {   // next round's frame: first stages fly during the overlap-add
            const float4* __restrict__ xr = X + (long long)(t + kIpWarps) * kIpBins;
            const float4* __restrict__ mr = MASK ? M + (long long)(t + kIpWarps) * kIpBins : nullptr;
#pragma unroll
            for (int r = 0; r < kIpPre; ++r) IP_ISSUE(xr, mr, r, r);
        }
        __syncthreads();   // all frames of the round are parked

        // ---- overlap-add: span [S, S + kIpWarps hop + carry_len), ascending frame order ------------------------
        // A thread owns offsets j (< hop) and walks the hop-blocks h of the span: position i = h hop + j
        // receives frame f at sample (h - f) hop + j, f = max(0, h - kj) .. min(nf - 1, h), where
        // kj = (2047 - j) / hop is fixed per thread -- no per-sample division, unit-stride LDS.64.
        const long long S = (long long)tr * hop;
        // the carry is updated in place: a thread writes carry[(h - kIpWarps) hop + j] only after it has
        // read that same element (at hop-block h - kIpWarps), and no other thread touches offsets = j mod hop
        const float2* cin = s_carry;
        float2* cout = s_carry;
        const int emit = kIpWarps * hop;
        const int span = emit + carry_len;
        const int fstride = 2 * kScrF4 - hop;                       // float2 step from (f, h) to (f + 1, h)
        // Interior round (CTA-uniform): every frame is live, every emitted position is owned and lands inside the
        // destination -- the register form without per-position predicates and with 32-bit indexing from four bases.
        const long long rel = S - p.out_start;
        if (p.ola_fast && kOla_K <= kIpKMax && nf == kIpWarps && S >= Pa && S + emit <= Pb && place + rel >= 0 &&
            place + rel + emit <= p.dst_limit) {
            constexpr int NH = kIpWarps + kIpKMax - 1;
            const float* __restrict__ envp = p.inv_env + S;
            const float* __restrict__ wgp = p.weight ? p.weight + rel : nullptr;
            float* __restrict__ d0 = dst0 + rel;
            float* __restrict__ d1 = dst1 + rel;
            for (int j = tid; j < hop; j += kIpThreads) {
                const int kj = (kOla_K - 1) * hop + j < kIpN ? kOla_K - 1 : kOla_K - 2;
                const float2* __restrict__ sj = reinterpret_cast<const float2*>(s_scr) + j;
                float ev[kIpWarps], wg[kIpWarps];
#pragma unroll
                for (int h = 0; h < kIpWarps; ++h) {
                    ev[h] = __ldg(envp + h * hop + j);
                    wg[h] = wgp ? __ldg(wgp + h * hop + j) : 1.f;
                }
                float2 out[NH];
#pragma unroll
                for (int h = 0; h < NH; ++h) {
                    const int i = h * hop + j;
                    out[h] = (i < carry_len) ? cin[i] : make_float2(0.f, 0.f);
                }
#pragma unroll
                for (int f = 0; f < kIpWarps; ++f) {
                    const float2* __restrict__ sf = sj + f * (2 * kScrF4);
#pragma unroll
                    for (int k = 0; k < kIpKMax; ++k) {
                        if (k <= kj) {
                            const float2 v = sf[k * hop];
                            out[f + k].x += v.x;
                            out[f + k].y += v.y;
                        }
                    }
                }
#pragma unroll
                for (int h = 0; h < kIpWarps; ++h) {
                    float v0 = out[h].x * ev[h], v1 = out[h].y * ev[h];
                    if (wgp) {
                        v0 *= wg[h];
                        v1 *= wg[h];
                    }
                    d0[h * hop + j] = v0;
                    d1[h * hop + j] = v1;
                }
#pragma unroll
                for (int h = kIpWarps; h < NH; ++h) {
                    const int i = h * hop + j;
                    if (i < span) cout[i - emit] = out[h];
                }
            }
            continue;
        }
        for (int j = tid; j < hop; j += kIpThreads) {
            const int kj = (kOla_K - 1) * hop + j < kIpN ? kOla_K - 1 : kOla_K - 2;
            const float2* __restrict__ sj = reinterpret_cast<const float2*>(s_scr) + j;
            // 1 / envelope and chunk weight of the thread's emitted positions: all loads go out before the
            // first dependent use (one L2 latency per j instead of one per position)
            float ev[kIpWarps], wg[kIpWarps];
#pragma unroll
            for (int h = 0; h < kIpWarps; ++h) {
                const long long P = S + h * hop + j;
                const bool mine = P >= Pa && P < Pb;
                ev[h] = mine ? __ldg(p.inv_env + P) : 0.f;
                wg[h] = (mine && p.weight) ? __ldg(p.weight + (P - p.out_start)) : 1.f;
            }
            if (kOla_K <= kIpKMax) {
                // Register form (CTA-uniform branch; hop >= 410): the thread's kIpWarps + K - 1 hop-blocks are
                // accumulators, frames are added in ascending order f = 0 .. nf - 1, frame f feeding blocks
                // f .. f + kj.  All indices are compile-time, the W x K loads are independent.
                constexpr int NH = kIpWarps + kIpKMax - 1;
                float2 out[NH];
#pragma unroll
                for (int h = 0; h < NH; ++h) {
                    const int i = h * hop + j;
                    out[h] = (i < carry_len) ? cin[i] : make_float2(0.f, 0.f);
                }
#pragma unroll
                for (int f = 0; f < kIpWarps; ++f) {
                    if (f < nf) {
                        const float2* __restrict__ sf = sj + f * (2 * kScrF4);
#pragma unroll
                        for (int k = 0; k < kIpKMax; ++k) {
                            if (k <= kj) {
                                const float2 v = sf[k * hop];
                                out[f + k].x += v.x;
                                out[f + k].y += v.y;
                            }
                        }
                    }
                }
#pragma unroll
                for (int h = 0; h < kIpWarps; ++h) {
                    const long long P = S + h * hop + j;
                    if (P >= Pa && P < Pb) {
                        const long long pp = P - p.out_start;
                        float v0 = out[h].x * ev[h], v1 = out[h].y * ev[h];
                        if (p.weight) {
                            v0 *= wg[h];
                            v1 *= wg[h];
                        }
                        const long long qd = place + pp;
                        if (qd >= 0 && qd < p.dst_limit) {
                            dst0[pp] = v0;
                            dst1[pp] = v1;
                        }
                    }
                }
#pragma unroll
                for (int h = kIpWarps; h < NH; ++h) {
                    const int i = h * hop + j;
                    if (i < span) cout[i - emit] = out[h];
                }
                continue;
            }
            // emitted hop-blocks
#pragma unroll
            for (int h = 0; h < kIpWarps; ++h) {
                const int i = h * hop + j;
                float2 acc = (i < carry_len) ? cin[i] : make_float2(0.f, 0.f);
                const int f_lo = max(0, h - kj);
                const int f_hi = min(nf - 1, h);
                const float2* __restrict__ q = sj + f_lo * fstride + h * hop;   // = scr_f [(h - f) hop + j]
                for (int f = f_lo; f <= f_hi; ++f, q += fstride) {
                    const float2 v = *q;
                    acc.x += v.x;
                    acc.y += v.y;
                }
                const long long P = S + i;
                if (P >= Pa && P < Pb) {
                    const long long pp = P - p.out_start;
                    float v0 = acc.x * ev[h], v1 = acc.y * ev[h];
                    if (p.weight) {
                        v0 *= wg[h];
                        v1 *= wg[h];
                    }
                    const long long qd = place + pp;
                    if (qd >= 0 && qd < p.dst_limit) {
                        dst0[pp] = v0;
                        dst1[pp] = v1;
                    }
                }
            }
            // hop-blocks that stay in the carry
            for (int h = kIpWarps, i = emit + j; i < span; ++h, i += hop) {
                float2 acc = (i < carry_len) ? cin[i] : make_float2(0.f, 0.f);
                const int f_lo = max(0, h - kj);
                const int f_hi = min(nf - 1, h);
                const float2* __restrict__ q = sj + f_lo * fstride + h * hop;
                for (int f = f_lo; f <= f_hi; ++f, q += fstride) {
                    const float2 v = *q;
                    acc.x += v.x;
                    acc.y += v.y;
                }
                cout[i - emit] = acc;
            }
        }
    }
}

#undef IP_ISSUE

// ---- ring-buffered form (default) ---------------------------------------------------------------------------------------
// The kernel above keeps HBM latency away with a four-deep REGISTER pipeline of row loads: 128 of its 255 registers, 8 warps
// per SM, and its three phases (rows -> Z, iFFT, overlap-add) alternate on CTA barriers -- 0.39 of the HBM roofline, latency
// bound at 11 % of the warp slots (profiles/r01j_ncu_full_istft_in_bench.txt).  Here ONE producer warp streams the spectrum
// and mask rows of the CTA's frames, in frame order, into a ring of kRgTeam shared-memory slots with bulk asynchronous
// copies (cp.async.bulk, bytes counted on an mbarrier), and TWO teams of kRgTeam consumer warps take alternate rounds of
// kRgTeam frames: a warp builds Z from its slot (LDS.128, no global-load registers), hands the slot back, runs the iFFT,
// windows and parks its frame; then its team alone does the overlap-add of the round.  While team A transforms and
// overlap-adds round i, the rows of round i + 1 arrive and team B consumes them: the load latency of a round hides behind
// the arithmetic of the round before.  The overlap-adds stay in round order (a token passed between the teams on named
// barriers), so the carry between rounds and the summation order are those of istft_pk2_kernel.
constexpr int kRgTeam = 3;                // consumer warps per team = frames per round = ring slots
constexpr int kRgWarps = 2 * kRgTeam;     // consumer warps
constexpr int kRgThreads = (kRgWarps + 1) * 32;
constexpr int kRgSlotF4 = 2 * kIpBins;    // float4 per slot: spectrum row, mask row

__device__ __forceinline__ void named_bar_sync(int id, int n) {
#ifndef AL_CPU_EMUL
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory");
#else
    (void)id; (void)n;
    __syncthreads();
#endif
}
__device__ __forceinline__ void named_bar_arrive(int id, int n) {
#ifndef AL_CPU_EMUL
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory");
#else
    (void)id; (void)n;
#endif
}

template <bool MASK>
__global__ void __launch_bounds__(kRgThreads, 1)
istft_pk3_kernel(const IstftPkParams p) {
    constexpr int W = kRgTeam;
    AL_DYN_SMEM(unsigned char, smem_raw);
    float2* s_tw = reinterpret_cast<float2*>(smem_raw);                       // [1024]
    float2* s_ctw = s_tw + 1024;                                               // [1024] W^k
    float4* s_scr = reinterpret_cast<float4*>(s_ctw + 1024);                   // [kRgWarps][kScrF4]
    float4* s_ring = s_scr + kRgWarps * kScrF4;                                // [kRgTeam][kRgSlotF4]
    float2* s_carry = reinterpret_cast<float2*>(s_ring + kRgTeam * kRgSlotF4);   // [2048 - hop] (L, R)
    // a parity wait can only tell the current phase from the one before, so every waiter must see EVERY phase of its barrier:
    // one "full" barrier per (team, slot) -- it advances once per round of that team -- and one "empty" per slot for the producer
    uint64_t* s_full = reinterpret_cast<uint64_t*>(s_carry + (kIpN - p.hop));  // [2][kRgTeam]
    uint64_t* s_empty = s_full + 2 * kRgTeam;                                   // [kRgTeam]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = blockIdx.x / p.segs, seg = blockIdx.x - g * p.segs;   // g = chunk*stems + stem
    const int chunk = g / p.stems, stem = g - chunk * p.stems;
    const int hop = p.hop, T = p.n_frames;
    const int carry_len = kIpN - hop;
    const int kOla_K = (kIpN + hop - 1) / hop;          // frames that can cover one position (<= kIpKMax: the launcher checks)

    const long long Pa = (long long)p.out_start + (long long)seg * p.hops_per_cta * hop;
    const long long Pend = (long long)p.out_start + p.out_len;
    const long long Pb = min(Pa + (long long)p.hops_per_cta * hop, Pend);
    if (Pa >= Pb) return;
    int ta = (int)((Pa - kIpN) / hop) + 1;              // first frame touching Pa
    if (Pa < kIpN) ta = 0;
    ta = max(ta, 0);
    const int tb = min((int)((Pb - 1) / hop), T - 1);   // last frame touching Pb - 1
    const int t_last = (int)((Pb - 1) / hop);           // rounds run until the carry is flushed up to Pb

    const long long place = p.dst_offsets ? p.dst_offsets[chunk] : p.dst_off0 + (long long)chunk * p.dst_off_step;
    float* __restrict__ dst0 = p.dst + ((long long)stem * 2) * p.dst_ch_stride + (long long)chunk * p.dst_chunk_stride + place;
    float* __restrict__ dst1 = dst0 + p.dst_ch_stride;
    const float4* __restrict__ X = p.spec + (long long)(p.spec_has_stems ? g : chunk) * T * kIpBins;
    const float4* __restrict__ M = MASK ? p.mask + (long long)g * T * kIpBins : nullptr;

    if (tid == 0) {
        for (int s = 0; s < kRgTeam; ++s) {
            mbar_init(&s_full[s], 1);
            mbar_init(&s_full[kRgTeam + s], 1);
            mbar_init(&s_empty[s], 1);
        }
#ifndef AL_CPU_EMUL
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
        fence_async_smem();
    }
    for (int i = tid; i < 1024; i += kRgThreads) {
        s_tw[i] = p.tw[i];
        s_ctw[i] = p.ctw[i];
    }
    for (int i = tid; i < carry_len; i += kRgThreads) s_carry[i] = make_float2(0.f, 0.f);
    __syncthreads();

    if (warp == kRgWarps) {
        // ================================= producer =================================
        if (lane == 0) {
            for (int t = ta; t <= tb; ++t) {
                const int it = t - ta, s = it % kRgTeam, rnd = it / kRgTeam;
                mbar_wait(&s_empty[s], (rnd & 1) ^ 1);
                float4* slot = s_ring + s * kRgSlotF4;
                uint64_t* full = &s_full[(rnd & 1) * kRgTeam + s];        // the barrier of the team that takes this round
                bulk_load_g2s(slot, X + (long long)t * kIpBins, kIpBins * 16, full);
                if (MASK) bulk_load_g2s(slot + kIpBins, M + (long long)t * kIpBins, kIpBins * 16, full);
                mbar_arrive_expect_tx(full, (MASK ? 2u : 1u) * kIpBins * 16u);
            }
        }
        return;
    }

    // ================================= consumers: team = warp / kRgTeam, member = warp % kRgTeam =================================
    const int team = warp / kRgTeam, member = warp - team * kRgTeam;
    const int ttid = tid - team * (kRgTeam * 32);       // 0 .. 95 inside the team
    constexpr int kTeamThreads = kRgTeam * 32;
    const float2* __restrict__ g_win = reinterpret_cast<const float2*>(p.window);   // (w[2k], w[2k+1]), through L1
    float4* scr = s_scr + warp * kScrF4;
    const float4* team_scr = s_scr + team * (kRgTeam * kScrF4);
    // named barriers: 1 + team = the team's frames are parked / its scratches are free again; 3 = token "the overlap-add
    // before team A's next one is done", 4 = the same for team B.  Team B hands team A the first token.
    if (team == 1) named_bar_arrive(3, 2 * kTeamThreads);
    int round = 0;
    for (int tr = ta; tr <= t_last; tr += W, ++round) {
        if ((round & 1) != team) continue;              // the teams take alternate rounds
        const int nf = max(0, min(W, tb - tr + 1));   // live frames of this round (team-uniform)
        const int t = tr + member;
        const bool live = member < nf;                  // warp-uniform
        if (live) {
            float2 re[32], im[32];
            const int it = t - ta, s = it % kRgTeam;    // = member: every round uses each slot once
            mbar_wait(&s_full[team * kRgTeam + s], (round >> 1) & 1);    // this team's (round / 2)-th use of the slot
            const float4* __restrict__ xs = s_ring + s * kRgSlotF4;
            const float4* __restrict__ ms = xs + kIpBins;
            // ---- Z from the rows in the slot: register pairs (r, 31 - r), k1 = 32 r + lane, k2 = 32 (31 - r) + lane
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const int k1 = 32 * r + lane, k2 = 32 * (31 - r) + lane;
                float2 p1r, p1i, q2r, q2i, p2r, p2i, q1r, q1i;
                const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
                ip_product<MASK>(xs[k1], MASK ? ms[k1] : z, p1r, p1i);
                ip_product<MASK>(xs[1024 - k2], MASK ? ms[1024 - k2] : z, q2r, q2i);
                ip_product<MASK>(xs[k2], MASK ? ms[k2] : z, p2r, p2i);
                ip_product<MASK>(xs[1024 - k1], MASK ? ms[1024 - k1] : z, q1r, q1i);
                if (r == 0 && lane == 0) {   // k1 = 0: C2R ignores Im of the DC and Nyquist bins
                    p1i = make_float2(0.f, 0.f);
                    q1i = make_float2(0.f, 0.f);
                }
                const float2 w1 = s_ctw[k1], w2 = s_ctw[k2];
                ip_combine(p1r, p1i, q1r, q1i, w1, re[r], im[r]);
                ip_combine(p2r, p2i, q2r, q2i, w2, re[31 - r], im[31 - r]);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_empty[s]);   // the slot goes back to the producer: the next round's row may land
            warp_fft1024p_wide<true>(re, im, scr, s_tw, lane);
            // z[k] = (x[2k], x[2k+1]) * n_fft; window (carries 1 / n_fft) and park the frame: scr[k] = samples 2k, 2k+1
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                const float2 w = __ldg(g_win + 32 * r + lane);
                const float2 ev = pscale(re[r], w.x), od = pscale(im[r], w.y);
                scr[32 * r + lane] = make_float4(ev.x, ev.y, od.x, od.y);
            }
        }
        named_bar_sync(1 + team, kTeamThreads);         // the team's frames are parked
        named_bar_sync(3 + team, 2 * kTeamThreads);     // ... and the overlap-add of the round before is done (carry)

        // ---- overlap-add of the round (the register form of istft_pk2_kernel) by the team ------------------
        const long long S = (long long)tr * hop;
        const float2* cin = s_carry;
        float2* cout = s_carry;
        const int emit = W * hop;
        const int span = emit + carry_len;
        constexpr int NH = W + kIpKMax - 1;
        const long long rel = S - p.out_start;
        const bool interior = p.ola_fast && nf == W && S >= Pa && S + emit <= Pb && place + rel >= 0 &&
                              place + rel + emit <= p.dst_limit;
        for (int j = ttid; j < hop; j += kTeamThreads) {
            const int kj = (kOla_K - 1) * hop + j < kIpN ? kOla_K - 1 : kOla_K - 2;
            const float2* __restrict__ sj = reinterpret_cast<const float2*>(team_scr) + j;
            float ev[W], wg[W];
#pragma unroll
            for (int h = 0; h < W; ++h) {
                const long long P = S + h * hop + j;
                const bool mine = interior || (P >= Pa && P < Pb);
                ev[h] = mine ? __ldg(p.inv_env + P) : 0.f;
                wg[h] = (mine && p.weight) ? __ldg(p.weight + (P - p.out_start)) : 1.f;
            }
            float2 out[NH];
#pragma unroll
            for (int h = 0; h < NH; ++h) {
                const int i = h * hop + j;
                out[h] = (i < carry_len) ? cin[i] : make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int f = 0; f < W; ++f) {
                if (f < nf) {
                    const float2* __restrict__ sf = sj + f * (2 * kScrF4);
#pragma unroll
                    for (int k = 0; k < kIpKMax; ++k) {
                        if (k <= kj) {
                            const float2 v = sf[k * hop];
                            out[f + k].x += v.x;
                            out[f + k].y += v.y;
                        }
                    }
                }
            }
#pragma unroll
            for (int h = 0; h < W; ++h) {
                const long long P = S + h * hop + j;
                if (interior || (P >= Pa && P < Pb)) {
                    const long long pp = P - p.out_start;
                    float v0 = out[h].x * ev[h], v1 = out[h].y * ev[h];
                    if (p.weight) {
                        v0 *= wg[h];
                        v1 *= wg[h];
                    }
                    const long long qd = place + pp;
                    if (interior || (qd >= 0 && qd < p.dst_limit)) {
                        dst0[pp] = v0;
                        dst1[pp] = v1;
                    }
                }
            }
#pragma unroll
            for (int h = W; h < NH; ++h) {
                const int i = h * hop + j;
                if (i < span) cout[i - emit] = out[h];
            }
        }
        named_bar_sync(1 + team, kTeamThreads);         // the team's scratches may take the frames of its next round
        named_bar_arrive(4 - team, 2 * kTeamThreads);   // token: the other team may overlap-add
    }
}

// ---- streaming form (AL_IP_RING=2) ------------------------------------------------------------------------------------
// The ring kernel above still runs its consumers in lock step: a team parks three frames, meets on a barrier, overlap-adds,
// meets again.  Its profile (profiles/r02zb_*): 45 % of the stall samples sit in the overlap-add phase (predicated register
// form, dependent loads of 1 / sum(w^2), two named barriers per round), every warp waits for the slowest of its team twice
// per frame, the parked frames pin 16.5 KB of scratch per warp (so only 6 consumer warps fit), and the Hermitian partner
// Y[1024 - k] of every bin is loaded and multiplied by its mask a second time by the lane that owns Z[1024 - k].  Here
//  * the overlap-add is a circular ACCUMULATOR of 2048 stereo positions in shared memory, even and odd positions in two
//    arrays, each owned by ONE dedicated warp (position parity = warp): frame t adds its sample n to position t hop + n, and
//    after frame t the hop-block [t hop, (t+1) hop) is final -- scaled by 1 / sum(w^2) (and the chunk weight), stored,
//    cleared.  The two overlap-add warps run a short loop that stays in the instruction cache and take the frames in
//    ascending order, so every sample is the same left-to-right sum as in istft_pk2_kernel / istft_pk3_kernel.  (A first
//    version passed a token between the CONSUMER warps instead: three times slower -- each warp entered the serial section
//    with cold instructions, 67 % of its stall samples there were instruction fetches, profiles/r02zd_*.)
//  * consumer warps never meet: row slot -> Z -> iFFT with the SLOT as transposition scratch -> window -> the frame parked
//    in the same slot, per position parity and already rotated to the accumulator's entries -> "ready" -> next frame.  No
//    per-warp scratch, so 8 spectrum-row slots + 4 mask-row slots (197 KB) are in flight / in use and 9 consumer warps of 168
//    registers run;
//  * Z[k] and Z[1024 - k] come from ONE product pair: Z[1024 - k] = conj(A) + i conj(B) for Z[k] = A + i B -- half the LDS.128
//    and half the mask multiplications of the row -> Z phase -- in a ROLLED loop that writes both over the spectrum bins they
//    were made from (the lane that forms Z[1024 - k] does not own it; 16 unrolled register pairs + shuffles would save the
//    round trip through the slot but not fit the instruction cache, see below);
//  * the loop body all nine consumers stream through is ~1 600 SASS lines (+ 650 of the overlap-add warps): with the Z pairs
//    unrolled and two inlined FFT passes (6 000 - 7 000 lines) 25 - 34 % of the stall samples were instruction fetches and the
//    kernel ran 2 - 3 x slower than istft_pk3_kernel; warp_fft1024p_wide_rolled has ONE copy of the 32-point butterfly.
// Barriers (a parity wait can only tell the current phase from the one before, so every waiter sees every phase of its
// barrier): full[w] per consumer warp (its j-th frame = phase j), ready[s] per slot (the two overlap-add warps take every
// frame), empty[s] per slot (two arrivals, the producer waits), mempty[s] per mask slot (the consumer's arrival after its Z loop).
constexpr int kTkC = 9;                    // consumer warps
// TWO rings: the spectrum row's slot lives long (row -> Z -> transposition tile -> parked frame -> overlap-add), the mask row is
// dead after the Z loop.  With one 33 KB slot for both, 6 frames were all that fitted and the consumers waited for rows a third
// of the time; 8 long-lived slots of 16.5 KB + 3 short-lived mask slots fit in the same shared memory.
constexpr int kTkXDefault = 8;             // X ring: spectrum row, later Z / the transposition tile / the parked frame
constexpr int kTkSlotF4 = kScrF4;          // float4 per X slot (16 896 B >= the 1025-bin row)
constexpr int kTkMDefault = 4;             // M ring: mask rows (3 left the producer waiting for a mask slot, profiles/r03e_*;
                                           // the depths are compile-time: as run-time parameters the kernel lost 24 %, profiles/r03g_*)
constexpr int kTkMSlotF4 = 1032;           // float4 per M slot (the row rounded to 128 bytes)
constexpr int kTkThreads = (kTkC + 3) * 32;   // + two overlap-add warps + the producer
static_assert(kTkSlotF4 >= kIpBins && kTkMSlotF4 >= kIpBins, "slot too small");
// (istft_pk5_kernel keeps the single ring of 6 row + mask slots)
constexpr int kTkMaskOff = 1032;           // pk5: the mask row starts on a 128-byte line of the slot
constexpr int kSrSlotF4 = 2064;            // pk5: float4 per slot: spectrum row, mask row at kTkMaskOff, rounded to 128 bytes
constexpr int kTkU = 7;                    // block positions per lane of an overlap-add warp held in registers (hop <= 448)

template <bool MASK, int kTkSlots = kTkXDefault, int kTkMSlots = kTkMDefault>
__global__ void __launch_bounds__(kTkThreads, 1)
istft_pk4_kernel(const IstftPkParams p) {
    AL_DYN_SMEM(unsigned char, smem_raw);
    float2* s_tw = reinterpret_cast<float2*>(smem_raw);                       // [1024]
    float2* s_acc = s_tw + 1024;                                               // [2][1024] even / odd positions, (L, R)
    const float2* __restrict__ g_ctw = p.ctw;                                  // [1024] W^k through L1 (8 KB of shared memory = half a mask slot)
    float4* s_ring = reinterpret_cast<float4*>(s_acc + 2048);                  // [kTkSlots][kTkSlotF4]
    float4* s_mring = s_ring + kTkSlots * kTkSlotF4;                           // [kTkMSlots][kTkMSlotF4]
    uint64_t* s_full = reinterpret_cast<uint64_t*>(s_mring + kTkMSlots * kTkMSlotF4);   // [kTkC]
    uint64_t* s_ready = s_full + kTkC;                                          // [kTkSlots]
    uint64_t* s_empty = s_ready + kTkSlots;                                     // [kTkSlots]
    uint64_t* s_mempty = s_empty + kTkSlots;                                    // [kTkMSlots]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = blockIdx.x / p.segs, seg = blockIdx.x - g * p.segs;   // g = chunk*stems + stem
    const int chunk = g / p.stems, stem = g - chunk * p.stems;
    const int hop = p.hop, T = p.n_frames;

    const long long Pa = (long long)p.out_start + (long long)seg * p.hops_per_cta * hop;
    const long long Pend = (long long)p.out_start + p.out_len;
    const long long Pb = min(Pa + (long long)p.hops_per_cta * hop, Pend);
    if (Pa >= Pb) return;
    int ta = (int)((Pa - kIpN) / hop) + 1;              // first frame touching Pa
    if (Pa < kIpN) ta = 0;
    ta = max(ta, 0);
    const int tb = min((int)((Pb - 1) / hop), T - 1);   // last frame touching Pb - 1
    const int t_last = (int)((Pb - 1) / hop);           // hop-blocks are emitted up to the one holding Pb - 1

    if (tid == 0) {
        for (int w = 0; w < kTkC; ++w) mbar_init(&s_full[w], 1);
        for (int s = 0; s < kTkSlots; ++s) {
            mbar_init(&s_ready[s], 1);
            mbar_init(&s_empty[s], 2);
        }
        for (int s = 0; s < kTkMSlots; ++s) mbar_init(&s_mempty[s], 1);
#ifndef AL_CPU_EMUL
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
        fence_async_smem();
    }
    for (int i = tid; i < 1024; i += kTkThreads) s_tw[i] = p.tw[i];
    for (int i = tid; i < 2048; i += kTkThreads) s_acc[i] = make_float2(0.f, 0.f);
    __syncthreads();

    if (warp == kTkC + 2) {
        // ================================= producer =================================
        if (lane == 0) {
            const float4* __restrict__ X = p.spec + (long long)(p.spec_has_stems ? g : chunk) * T * kIpBins;
            const float4* __restrict__ M = MASK ? p.mask + (long long)g * T * kIpBins : nullptr;
            for (int t = ta; t <= tb; ++t) {
                const int it = t - ta, s = it % kTkSlots;
                mbar_wait(&s_empty[s], ((it / kTkSlots) & 1) ^ 1);
                float4* slot = s_ring + s * kTkSlotF4;
                uint64_t* full = &s_full[it % kTkC];
                bulk_load_g2s(slot, X + (long long)t * kIpBins, kIpBins * 16, full);
                if (MASK) {
                    const int sm = it % kTkMSlots;
                    mbar_wait(&s_mempty[sm], ((it / kTkMSlots) & 1) ^ 1);
                    bulk_load_g2s(s_mring + sm * kTkMSlotF4, M + (long long)t * kIpBins, kIpBins * 16, full);
                }
                mbar_arrive_expect_tx(full, (MASK ? 2u : 1u) * kIpBins * 16u);
            }
        }
        return;
    }

    if (warp >= kTkC) {
        // ================================= overlap-add warps: par = position parity =================================
        const int par = warp - kTkC;
        const long long place = p.dst_offsets ? p.dst_offsets[chunk] : p.dst_off0 + (long long)chunk * p.dst_off_step;
        float* __restrict__ dst0 = p.dst + ((long long)stem * 2) * p.dst_ch_stride + (long long)chunk * p.dst_chunk_stride + place;
        float* __restrict__ dst1 = dst0 + p.dst_ch_stride;
        float2* __restrict__ acc = s_acc + par * 1024;
        const bool small_hop = hop <= 64 * kTkU;        // CTA-uniform
        // 32-bit positions relative to out_start (a chunk is far below 2^31 samples)
        const int ra = (int)(Pa - p.out_start), rb = (int)(Pb - p.out_start);
        const long long lim_lo = -place, lim_hi = p.dst_limit - place;      // stores need lim_lo <= pp < lim_hi
        const float* __restrict__ envp = p.inv_env + p.out_start;
        const float* __restrict__ wgp = p.weight;
#pragma unroll 1
        for (int it = 0; ta + it <= t_last; ++it) {
            const int t = ta + it;
            const int P0 = t * hop;                     // < 2^31: the launcher checks (T + 4) * hop
            const int odd = P0 & 1, beta = P0 >> 1;
            // this warp's positions of the frame: samples 2k (+1 when the parities differ) at entries (off + k) & 1023, and of the
            // hop-block: P = P0 + j0 + 2 m at entries (off + m) & 1023
            const bool same = par == odd;
            const int off = same ? beta : beta + odd;
            const int j0 = par ^ odd;
            const int n_blk = (hop - j0 + 1) >> 1;
            const int r0p = P0 + j0 - p.out_start;      // relative position of m = 0
            // interior block (warp-uniform): every position is owned and lands inside the destination
            const bool interior = small_hop && r0p >= ra && r0p + 2 * n_blk <= rb && r0p >= lim_lo && r0p + 2 * n_blk <= lim_hi;
            // 1 / sum(w^2) and chunk weight of the block, fetched before the wait
            float ev[kTkU], wg[kTkU];
            if (interior) {
#pragma unroll
                for (int u = 0; u < kTkU; ++u) {
                    const int m = lane + 32 * u;
                    const bool mine = m < n_blk;
                    ev[u] = mine ? ldg_stream(envp + r0p + 2 * m) : 0.f;
                    wg[u] = (mine && wgp) ? ldg_stream(wgp + r0p + 2 * m) : 1.f;
                }
            }
            if (t <= tb) {                              // frames past tb only flush their hop-block
                const int s = it % kTkSlots;
                mbar_wait(&s_ready[s], (it / kTkSlots) & 1);
                // the consumer parked the frame already rotated to the accumulator's entries: entry e of this parity at [par][e]
                const float2* __restrict__ src = reinterpret_cast<const float2*>(s_ring + s * kTkSlotF4) + par * 1024 + lane;
                float2* __restrict__ ac = acc + lane;
                {   // all 64 loads in flight before the first add: the warp is paced by shared-memory round trips, not by issue slots
                    float2 va[32], vs[32];
#pragma unroll
                    for (int q = 0; q < 32; ++q) {
                        vs[q] = src[32 * q];
                        va[q] = ac[32 * q];
                    }
#pragma unroll
                    for (int q = 0; q < 32; ++q) ac[32 * q] = padd(va[q], vs[q]);
                }
                fence_async_smem();                     // generic-proxy accesses of the slot before the next bulk copy into it
                __syncwarp();
                if (lane == 0) mbar_arrive(&s_empty[s]);
            }
            // ---- the hop-block [P0, P0 + hop) is final: take it out of the accumulator, clear it, scale, store
            if (interior) {
                float2 blk[kTkU];
#pragma unroll
                for (int u = 0; u < kTkU; ++u) {
                    const int m = lane + 32 * u;
                    float2* a = acc + ((off + m) & 1023);
                    blk[u] = *a;                        // (entries past the block belong to later blocks: read, not cleared)
                    if (m < n_blk) *a = make_float2(0.f, 0.f);
                }
                float* __restrict__ d0 = dst0 + r0p;
                float* __restrict__ d1 = dst1 + r0p;
#pragma unroll
                for (int u = 0; u < kTkU; ++u) {
                    const int m = lane + 32 * u;
                    if (m < n_blk) {
                        float v0 = blk[u].x * ev[u], v1 = blk[u].y * ev[u];
                        if (wgp) {
                            v0 *= wg[u];
                            v1 *= wg[u];
                        }
                        d0[2 * m] = v0;
                        d1[2 * m] = v1;
                    }
                }
            } else {
#pragma unroll 1
                for (int m = lane; m < n_blk; m += 32) {
                    const int pp = r0p + 2 * m;
                    float2* a = acc + ((off + m) & 1023);
                    const float2 v = *a;
                    *a = make_float2(0.f, 0.f);
                    if (pp >= ra && pp < rb && pp >= lim_lo && pp < lim_hi) {
                        const float e = ldg_stream(envp + pp);
                        float v0 = v.x * e, v1 = v.y * e;
                        if (wgp) {
                            const float wgt = ldg_stream(wgp + pp);
                            v0 *= wgt;
                            v1 *= wgt;
                        }
                        dst0[pp] = v0;
                        dst1[pp] = v1;
                    }
                }
            }
            __syncwarp();                               // the cleared block before the next frame's adds
        }
        return;
    }

    // ================================= consumers: frame it = warp, warp + kTkC, ... =================================
    const float2* __restrict__ g_win = reinterpret_cast<const float2*>(p.window);   // (w[2k], w[2k+1]), through L1
#pragma unroll 1
    for (int it = warp; ta + it <= tb; it += kTkC) {
        const int s = it % kTkSlots;
        float4* slot = s_ring + s * kTkSlotF4;
        float2 re[32], im[32];
        mbar_wait(&s_full[warp], (it / kTkC) & 1);
        {
            float4* __restrict__ xs = slot;
            const float4* __restrict__ ms = s_mring + (it % kTkMSlots) * kTkMSlotF4;
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            // ---- Z[k1] and Z[1024 - k1] from ONE pair P = Y[k1], Q = Y[1024 - k1], written over the two spectrum bins they came
            // from (no other lane reads those): a rolled loop instead of 16 unrolled register pairs
            {   // Z[512]: its own Hermitian partner (every lane forms it, the stores coincide)
                float2 pr, pi, zr, zi;
                ip_product<MASK>(xs[512], MASK ? ms[512] : z4, pr, pi);
                ip_combine(pr, pi, pr, pi, __ldg(g_ctw + 512), zr, zi);
                __syncwarp();
                xs[512] = make_float4(zr.x, zr.y, zi.x, zi.y);
            }
#pragma unroll 4
            for (int k1 = lane; k1 < 512; k1 += 32) {
                float2 pr, pi, qr, qi;
                ip_product<MASK>(xs[k1], MASK ? ms[k1] : z4, pr, pi);
                ip_product<MASK>(xs[1024 - k1], MASK ? ms[1024 - k1] : z4, qr, qi);
                if (k1 == 0) {               // C2R ignores Im of the DC and Nyquist bins
                    pi = make_float2(0.f, 0.f);
                    qi = make_float2(0.f, 0.f);
                }
                const float2 w = __ldg(g_ctw + k1);
                const float2 ar = padd(pr, qr), ai = psub(pi, qi);     // A = P + conj(Q)
                const float2 dr = psub(pr, qr), di = padd(pi, qi);     // D = P - conj(Q)
                const float2 zr = pfma(dr, w.y, pfma(di, -w.x, ar));    // Z[k1] = A + i D conj(w)
                const float2 zi = pfma(di, w.y, pfma(dr, w.x, ai));
                // Z[1024 - k1] = conj(A) + i conj(D conj(w)): the arithmetic ip_combine(Q, P, -conj(w)) would do
                const float2 ndr = make_float2(-dr.x, -dr.y), nai = make_float2(-ai.x, -ai.y);
                const float2 mr = pfma(ndr, w.y, pfma(di, w.x, ar));
                const float2 mi = pfma(di, w.y, pfma(dr, w.x, nai));
                xs[k1] = make_float4(zr.x, zr.y, zi.x, zi.y);
                xs[1024 - k1] = make_float4(mr.x, mr.y, mi.x, mi.y);   // (k1 = 0 lands on the dead Nyquist bin)
            }
            __syncwarp();
            if (MASK && lane == 0) mbar_arrive(&s_mempty[it % kTkMSlots]);   // the mask row is dead: its slot takes the next one
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                const float4 v = xs[32 * r + lane];
                re[r] = make_float2(v.x, v.y);
                im[r] = make_float2(v.z, v.w);
            }
        }
        __syncwarp();                                   // every lane holds its Z: the slot is scratch now
        warp_fft1024p_wide_rolled<true>(re, im, slot, s_tw, lane);
        // z[k] = (x[2k], x[2k+1]) * n_fft; window (carries 1 / n_fft) and park the frame in the slot, per position parity and
        // ALREADY ROTATED to the accumulator: sample 2k sits at position P0 + 2k = entry (beta + k) & 1023 of parity P0 & 1, sample
        // 2k + 1 at entry (beta + k + odd) & 1023 of the other parity, so the overlap-add warps add entry to entry
        const int P0 = (ta + it) * hop;
        const int odd = P0 & 1, beta = P0 >> 1;
        float2* pA = reinterpret_cast<float2*>(slot) + (odd ? 1024 : 0);
        float2* pB = reinterpret_cast<float2*>(slot) + (odd ? 0 : 1024);
        const int ea = beta + lane, eb = beta + odd + lane;
#pragma unroll
        for (int r = 0; r < 32; ++r) {
            const float2 w = __ldg(g_win + 32 * r + lane);
            pA[(ea + 32 * r) & 1023] = pscale(re[r], w.x);
            pB[(eb + 32 * r) & 1023] = pscale(im[r], w.y);
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_ready[s]);
    }
}

// ---- streaming form, overlap-add from registers (AL_IP_RING=3) -----------------------------------------------------------
// istft_pk4_kernel is paced by the shared-memory pipe (profiles/r02zi_*: 72 % busy in the steady state, the same 21.5 M
// wavefronts and the same 0.127 ms for three different overlap-add codes): per frame 256 wavefronts of rows written by the
// bulk copies + 1 540 by the warps, of which 256 are Z written to the slot and read back (the price of the rolled Z loop), 256
// the parked frame written and read back, 256 the accumulator.  This form drops both round trips:
//  * Z is formed in registers (16 unrolled register pairs, the mirror through four shuffles) -- affordable now that the FFT
//    has one copy of the butterfly: the consumer loop stays near 32 KB of instructions;
//  * the windowed frame goes from the registers straight into the accumulator.  Frames must be added in ascending order, so
//    the consumer warps pass TOKENS (one mbarrier per warp and accumulator parity): a warp waits for the token of parity
//    P0 & 1, adds its even samples, takes the final hop-block entries of that parity out, passes the token on, scales and
//    stores them, then does the same with its odd samples on the other parity -- two chains, so two frames are being added at
//    any time.  (The first token kernel of this round did this with 112 KB of code and lost to instruction fetches.)
//  * the slot goes back to the producer as soon as the transposition has been read back: slots are held for rows -> Z ->
//    half an iFFT only, 11 consumer warps.
constexpr int kSrC = 11;                   // consumer warps
constexpr int kSrSlots = 6;
constexpr int kSrThreads = (kSrC + 1) * 32;

template <bool MASK>
__global__ void __launch_bounds__(kSrThreads, 1)
istft_pk5_kernel(const IstftPkParams p) {
    AL_DYN_SMEM(unsigned char, smem_raw);
    float2* s_tw = reinterpret_cast<float2*>(smem_raw);                       // [1024]
    float2* s_ctw = s_tw + 1024;                                               // [1024] W^k
    float2* s_acc = s_ctw + 1024;                                              // [2][1024] even / odd positions, (L, R)
    float4* s_ring = reinterpret_cast<float4*>(s_acc + 2048);                  // [kSrSlots][kSrSlotF4]
    uint64_t* s_full = reinterpret_cast<uint64_t*>(s_ring + kSrSlots * kSrSlotF4);   // [kSrC]
    uint64_t* s_empty = s_full + kSrC;                                          // [kSrSlots]
    uint64_t* s_tok = s_empty + kSrSlots;                                       // [2][kSrC]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = blockIdx.x / p.segs, seg = blockIdx.x - g * p.segs;   // g = chunk*stems + stem
    const int chunk = g / p.stems, stem = g - chunk * p.stems;
    const int hop = p.hop, T = p.n_frames;

    const long long Pa = (long long)p.out_start + (long long)seg * p.hops_per_cta * hop;
    const long long Pend = (long long)p.out_start + p.out_len;
    const long long Pb = min(Pa + (long long)p.hops_per_cta * hop, Pend);
    if (Pa >= Pb) return;
    int ta = (int)((Pa - kIpN) / hop) + 1;              // first frame touching Pa
    if (Pa < kIpN) ta = 0;
    ta = max(ta, 0);
    const int tb = min((int)((Pb - 1) / hop), T - 1);   // last frame touching Pb - 1
    const int t_last = (int)((Pb - 1) / hop);           // hop-blocks are emitted up to the one holding Pb - 1

    if (tid == 0) {
        for (int w = 0; w < kSrC; ++w) {
            mbar_init(&s_full[w], 1);
            mbar_init(&s_tok[w], 1);
            mbar_init(&s_tok[kSrC + w], 1);
        }
        for (int s = 0; s < kSrSlots; ++s) mbar_init(&s_empty[s], 1);
#ifndef AL_CPU_EMUL
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
#endif
        fence_async_smem();
        mbar_arrive(&s_tok[0]);                         // the first frame has no predecessor on either parity
        mbar_arrive(&s_tok[kSrC]);
    }
    for (int i = tid; i < 1024; i += kSrThreads) {
        s_tw[i] = p.tw[i];
        s_ctw[i] = p.ctw[i];
    }
    for (int i = tid; i < 2048; i += kSrThreads) s_acc[i] = make_float2(0.f, 0.f);
    __syncthreads();

    if (warp == kSrC) {
        // ================================= producer =================================
        if (lane == 0) {
            const float4* __restrict__ X = p.spec + (long long)(p.spec_has_stems ? g : chunk) * T * kIpBins;
            const float4* __restrict__ M = MASK ? p.mask + (long long)g * T * kIpBins : nullptr;
            for (int t = ta; t <= tb; ++t) {
                const int it = t - ta, s = it % kSrSlots;
                mbar_wait(&s_empty[s], ((it / kSrSlots) & 1) ^ 1);
                float4* slot = s_ring + s * kSrSlotF4;
                uint64_t* full = &s_full[it % kSrC];
                bulk_load_g2s(slot, X + (long long)t * kIpBins, kIpBins * 16, full);
                if (MASK) bulk_load_g2s(slot + kTkMaskOff, M + (long long)t * kIpBins, kIpBins * 16, full);
                mbar_arrive_expect_tx(full, (MASK ? 2u : 1u) * kIpBins * 16u);
            }
        }
        return;
    }

    // ================================= consumers: frame it = warp, warp + kSrC, ... =================================
    const long long place = p.dst_offsets ? p.dst_offsets[chunk] : p.dst_off0 + (long long)chunk * p.dst_off_step;
    float* __restrict__ dst0 = p.dst + ((long long)stem * 2) * p.dst_ch_stride + (long long)chunk * p.dst_chunk_stride + place;
    float* __restrict__ dst1 = dst0 + p.dst_ch_stride;
    const float2* __restrict__ g_win = reinterpret_cast<const float2*>(p.window);   // (w[2k], w[2k+1]), through L1
    const int nxt = (warp + 1) % kSrC;
    // 32-bit positions relative to out_start (the launcher checks the range)
    const int ra = (int)(Pa - p.out_start), rb = (int)(Pb - p.out_start);
    const long long lim_lo = -place, lim_hi = p.dst_limit - place;          // stores need lim_lo <= pp < lim_hi
    const float* __restrict__ envp = p.inv_env + p.out_start;
    const float* __restrict__ wgp = p.weight;
#pragma unroll 1
    for (int it = warp; ta + it <= t_last; it += kSrC) {
        const int t = ta + it;
        const bool live = t <= tb;                      // warp-uniform; frames past tb only flush their hop-block
        const uint32_t ph = (uint32_t)(it / kSrC) & 1u;
        {   // 1 / sum(w^2) and chunk weight of this frame's hop-block: into L2 now, read around the tokens
            const int pp = t * hop - p.out_start + 32 * lane;
            if (32 * lane < hop + 32 && pp >= ra && pp < rb) {
                prefetch_l2(envp + pp);
                if (wgp) prefetch_l2(wgp + pp);
            }
        }
        float2 re[32], im[32];
        if (live) {
            const int s = it % kSrSlots;
            float4* slot = s_ring + s * kSrSlotF4;
            mbar_wait(&s_full[warp], ph);
            {
                float4* __restrict__ xs = slot;
                const float4* __restrict__ ms = slot + kTkMaskOff;
                const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
                // ---- Z[k1] and Z[1024 - k1] from ONE pair P = Y[k1], Q = Y[1024 - k1], written over the two spectrum bins they came
                // from (no other lane reads those): a rolled loop instead of 16 unrolled register pairs
                {   // Z[512]: its own Hermitian partner (every lane forms it, the stores coincide)
                    float2 pr, pi, zr, zi;
                    ip_product<MASK>(xs[512], MASK ? ms[512] : z4, pr, pi);
                    ip_combine(pr, pi, pr, pi, s_ctw[512], zr, zi);
                    __syncwarp();
                    xs[512] = make_float4(zr.x, zr.y, zi.x, zi.y);
                }
    #pragma unroll 4
                for (int k1 = lane; k1 < 512; k1 += 32) {
                    float2 pr, pi, qr, qi;
                    ip_product<MASK>(xs[k1], MASK ? ms[k1] : z4, pr, pi);
                    ip_product<MASK>(xs[1024 - k1], MASK ? ms[1024 - k1] : z4, qr, qi);
                    if (k1 == 0) {               // C2R ignores Im of the DC and Nyquist bins
                        pi = make_float2(0.f, 0.f);
                        qi = make_float2(0.f, 0.f);
                    }
                    const float2 w = s_ctw[k1];
                    const float2 ar = padd(pr, qr), ai = psub(pi, qi);     // A = P + conj(Q)
                    const float2 dr = psub(pr, qr), di = padd(pi, qi);     // D = P - conj(Q)
                    const float2 zr = pfma(dr, w.y, pfma(di, -w.x, ar));    // Z[k1] = A + i D conj(w)
                    const float2 zi = pfma(di, w.y, pfma(dr, w.x, ai));
                    // Z[1024 - k1] = conj(A) + i conj(D conj(w)): the arithmetic ip_combine(Q, P, -conj(w)) would do
                    const float2 ndr = make_float2(-dr.x, -dr.y), nai = make_float2(-ai.x, -ai.y);
                    const float2 mr = pfma(ndr, w.y, pfma(di, w.x, ar));
                    const float2 mi = pfma(di, w.y, pfma(dr, w.x, nai));
                    xs[k1] = make_float4(zr.x, zr.y, zi.x, zi.y);
                    xs[1024 - k1] = make_float4(mr.x, mr.y, mi.x, mi.y);   // (k1 = 0 lands on the dead Nyquist bin)
                }
                __syncwarp();
    #pragma unroll
                for (int r = 0; r < 32; ++r) {
                    const float4 v = xs[32 * r + lane];
                    re[r] = make_float2(v.x, v.y);
                    im[r] = make_float2(v.z, v.w);
                }
            }
            __syncwarp();                               // every lane holds its Z: the slot is scratch now
            warp_fft1024p_wide_rolled<true>(re, im, slot, s_tw, lane, &s_empty[s]);
            // z[k] = (x[2k], x[2k+1]) * n_fft; the window carries 1 / n_fft.  The stages below take the accumulator parities in a
            // FIXED order (even positions, then odd positions: with the sample order instead, an odd hop makes frame f's second
            // stage and frame f + 1's first stage meet on the same parity and the two chains collapse into one that also holds
            // the stores in between -- profiles/r02zk_*): re[] gets the samples of the even positions, im[] those of the odd ones
            if ((t * hop) & 1) {
#pragma unroll
                for (int r = 0; r < 32; ++r) {
                    const float2 w = __ldg(g_win + 32 * r + lane);
                    const float2 ev = pscale(re[r], w.x);
                    re[r] = pscale(im[r], w.y);
                    im[r] = ev;
                }
            } else {
#pragma unroll
                for (int r = 0; r < 32; ++r) {
                    const float2 w = __ldg(g_win + 32 * r + lane);
                    re[r] = pscale(re[r], w.x);
                    im[r] = pscale(im[r], w.y);
                }
            }
        }
        const int P0 = t * hop;
        const int odd = P0 & 1, beta = P0 >> 1;
        // ---- two stages: parity 0 (even positions), parity 1.  ONE copy of the stage in the instruction stream: the second trip
        // finds its samples moved into re[]
#pragma unroll 1
        for (int arr = 0; arr < 2; ++arr) {
            const int j0 = arr ^ odd;                   // the block's positions of this parity: P0 + j0 + 2 m
            const int off = beta + (arr ? 0 : odd);     // the frame's k-th sample of this parity -> entry (off + k) & 1023; block entry m -> (off + m) & 1023
            const int n_blk = (hop - j0 + 1) >> 1;
            const int r0p = P0 + j0 - p.out_start;      // relative position of m = 0
            // interior block (warp-uniform): every position is owned and lands inside the destination
            const bool interior = hop <= 64 * kTkU && r0p >= ra && r0p + 2 * n_blk <= rb && r0p >= lim_lo && r0p + 2 * n_blk <= lim_hi;
            float2* __restrict__ acc = s_acc + arr * 1024;
            mbar_wait_hint(&s_tok[arr * kSrC + warp], ph, (uint32_t)p.tok_hint_ns);
            if (live) {
                const int e0 = off + lane;
#pragma unroll
                for (int r0 = 0; r0 < 32; r0 += 8) {
                    float2 va[8];
#pragma unroll
                    for (int q = 0; q < 8; ++q) va[q] = acc[(e0 + 32 * (r0 + q)) & 1023];
#pragma unroll
                    for (int q = 0; q < 8; ++q) acc[(e0 + 32 * (r0 + q)) & 1023] = padd(va[q], re[r0 + q]);
                }
                __syncwarp();
            }
            // the entries of this parity in the hop-block [P0, P0 + hop) are final: take them out, clear them
            if (interior) {
                float2 blk[kTkU];
#pragma unroll
                for (int u = 0; u < kTkU; ++u) {
                    const int m = lane + 32 * u;
                    float2* a = acc + ((off + m) & 1023);
                    blk[u] = *a;                        // (entries past the block belong to later blocks: read, not cleared)
                    if (m < n_blk) *a = make_float2(0.f, 0.f);
                }
                __syncwarp();
                if (lane == 0) { if (p.tok_relaxed) mbar_arrive_relaxed(&s_tok[arr * kSrC + nxt]); else mbar_arrive(&s_tok[arr * kSrC + nxt]); }
                // scale and store outside the chain
                const float* __restrict__ ep = envp + r0p + 2 * lane;
                const float* __restrict__ wp = wgp ? wgp + r0p + 2 * lane : nullptr;
                float* __restrict__ d0 = dst0 + r0p + 2 * lane;
                float* __restrict__ d1 = dst1 + r0p + 2 * lane;
                float ev[kTkU], wg[kTkU];              // all loads in flight before the first store (one latency, not seven)
#pragma unroll
                for (int u = 0; u < kTkU; ++u) {
                    const bool mine = lane + 32 * u < n_blk;
                    ev[u] = mine ? ldg_stream(ep + 64 * u) : 0.f;
                    wg[u] = (mine && wp) ? ldg_stream(wp + 64 * u) : 1.f;
                }
#pragma unroll
                for (int u = 0; u < kTkU; ++u) {
                    if (lane + 32 * u < n_blk) {
                        const float e = ev[u] * wg[u];
                        d0[64 * u] = blk[u].x * e;
                        d1[64 * u] = blk[u].y * e;
                    }
                }
            } else {
#pragma unroll 1
                for (int m = lane; m < n_blk; m += 32) {
                    const int pp = r0p + 2 * m;
                    float2* a = acc + ((off + m) & 1023);
                    const float2 v = *a;
                    *a = make_float2(0.f, 0.f);
                    if (pp >= ra && pp < rb && pp >= lim_lo && pp < lim_hi) {
                        float e = ldg_stream(envp + pp);
                        if (wgp) e *= ldg_stream(wgp + pp);
                        dst0[pp] = v.x * e;
                        dst1[pp] = v.y * e;
                    }
                }
                __syncwarp();
                if (lane == 0) { if (p.tok_relaxed) mbar_arrive_relaxed(&s_tok[arr * kSrC + nxt]); else mbar_arrive(&s_tok[arr * kSrC + nxt]); }
            }
            if (arr == 0 && live) {
#pragma unroll
                for (int r = 0; r < 32; ++r) re[r] = im[r];
            }
        }
    }
}

// segments per row: minimise waves * rounds per segment (a round = kIpWarps frames; each segment re-computes the
// ceil((2048 - hop) / hop) frames that precede its first owned sample)
static void ip_tiling(int rows, int total_hops, int hop, int n_sm, int kIpWarps, int* hpc_out, int* segs_out) {
    const int halo = (kIpN - hop + hop - 1) / hop;
    long long best = -1;
    int best_segs = 1;
    const int max_segs = total_hops / kIpWarps > 1 ? total_hops / kIpWarps : 1;
    for (int segs = 1; segs <= max_segs; ++segs) {
        const int hpc = (total_hops + segs - 1) / segs;
        const int real_segs = (total_hops + hpc - 1) / hpc;
        const long long waves = ((long long)rows * real_segs + n_sm - 1) / n_sm;
        const long long rounds = (hpc + halo + kIpWarps - 1) / kIpWarps + 1;
        const long long cost = waves * rounds;
        if (best < 0 || cost < best) {
            best = cost;
            best_segs = segs;
        }
    }
    const int hpc = (total_hops + best_segs - 1) / best_segs;
    *hpc_out = hpc;
    *segs_out = (total_hops + hpc - 1) / hpc;
}

// launch shape of istft_pk2_kernel<., W>: fills hops_per_cta / segs, returns the dynamic shared memory size
static size_t ip_launch_shape(IstftPkParams& p, int n_chunks, int n_sm, int W) {
    const int rows = n_chunks * p.stems;
    const int total_hops = (p.out_len + p.hop - 1) / p.hop;
    ip_tiling(rows, total_hops, p.hop, n_sm * (kIpSmWarps / W), W, &p.hops_per_cta, &p.segs);
    return (size_t)3 * 1024 * sizeof(float2) + (size_t)W * kScrF4 * sizeof(float4) + (size_t)(kIpN - p.hop) * sizeof(float2);
}
// launch shape of istft_pk3_kernel: one CTA per SM
static size_t rg_launch_shape(IstftPkParams& p, int n_chunks, int n_sm) {
    const int rows = n_chunks * p.stems;
    const int total_hops = (p.out_len + p.hop - 1) / p.hop;
    ip_tiling(rows, total_hops, p.hop, n_sm, kRgTeam, &p.hops_per_cta, &p.segs);
    return (size_t)2 * 1024 * sizeof(float2) + (size_t)kRgWarps * kScrF4 * sizeof(float4) +
           (size_t)kRgTeam * kRgSlotF4 * sizeof(float4) + (size_t)(kIpN - p.hop) * sizeof(float2) + 3 * kRgTeam * sizeof(uint64_t);
}
// launch shape of istft_pk4_kernel: one CTA per SM; a segment costs its hops + the halo frames it recomputes + the fill of the warp pipeline
static size_t tk_launch_shape(IstftPkParams& p, int n_chunks, int n_sm, int kTkSlots = kTkXDefault, int kTkMSlots = kTkMDefault) {
    const int rows = n_chunks * p.stems;
    const int total_hops = (p.out_len + p.hop - 1) / p.hop;
    const int halo = (kIpN - 1) / p.hop;
    long long best = -1;
    int best_segs = 1;
    const int max_segs = total_hops / 8 > 1 ? total_hops / 8 : 1;
    for (int segs = 1; segs <= max_segs; ++segs) {
        const int hpc = (total_hops + segs - 1) / segs;
        const int real_segs = (total_hops + hpc - 1) / hpc;
        const long long waves = ((long long)rows * real_segs + n_sm - 1) / n_sm;
        const long long cost = waves * (hpc + halo + kTkC);
        if (best < 0 || cost < best) {
            best = cost;
            best_segs = segs;
        }
    }
    p.hops_per_cta = (total_hops + best_segs - 1) / best_segs;
    p.segs = (total_hops + p.hops_per_cta - 1) / p.hops_per_cta;
    return (size_t)(1024 + 2048) * sizeof(float2) + (size_t)(kTkSlots * kTkSlotF4 + kTkMSlots * kTkMSlotF4) * sizeof(float4) +
           (size_t)(kTkC + 2 * kTkSlots + kTkMSlots) * sizeof(uint64_t);
}
// launch shape of istft_pk5_kernel: the segments of istft_pk4_kernel
static size_t sr_launch_shape(IstftPkParams& p, int n_chunks, int n_sm) {
    tk_launch_shape(p, n_chunks, n_sm);
    return (size_t)(2 * 1024 + 2048) * sizeof(float2) + (size_t)kSrSlots * kSrSlotF4 * sizeof(float4) +
           (size_t)(3 * kSrC + kSrSlots) * sizeof(uint64_t);
}
// [emul-end]

cudaError_t launch_istft_pk(const IstftPkParams& p0, int n_chunks, cudaStream_t stream) {
    IstftPkParams p = p0;
    const int n_sm = sm_count();
    const int rows = n_chunks * p.stems;
    // the ring-buffered kernel (default; AL_IP_RING=0 selects the register-pipelined one): hop >= 410 (register-form overlap-add)
    // AL_IP_RING: 2 (default) = the streaming kernel, 1 = the ring kernel with teams, 0 = the register-pipelined kernel
    static const int ring = getenv("AL_IP_RING") ? atoi(getenv("AL_IP_RING")) : 2;
    if (ring == 3 && (long long)(p.n_frames + 8) * p.hop < 0x7fffffffLL && (long long)p.out_start + p.out_len < 0x7fffffffLL) {
        p.ola_fast = 0;
        p.l2_prefetch = 0;
        static const int hint5 = getenv("AL_IP_HINT") ? atoi(getenv("AL_IP_HINT")) : 0;
        static const int relaxed5 = getenv("AL_IP_RELAXED") ? atoi(getenv("AL_IP_RELAXED")) : 0;
        p.tok_hint_ns = hint5;
        p.tok_relaxed = relaxed5;
        const size_t smem5 = sr_launch_shape(p, n_chunks, n_sm);
        static PerDeviceOnce attr5[2];
        PerDeviceOnce& a5 = attr5[p.mask ? 1 : 0];
        if (a5.needed()) {
            cudaError_t e = p.mask ? cudaFuncSetAttribute(istft_pk5_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)
                                   : cudaFuncSetAttribute(istft_pk5_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) return e;
            a5.mark();
        }
        if (p.mask) istft_pk5_kernel<true><<<(unsigned)(rows * p.segs), kSrThreads, smem5, stream>>>(p);
        else istft_pk5_kernel<false><<<(unsigned)(rows * p.segs), kSrThreads, smem5, stream>>>(p);
        count_launch();
        return cudaGetLastError();
    }
    if (ring == 2 && (long long)(p.n_frames + 8) * p.hop < 0x7fffffffLL && (long long)p.out_start + p.out_len < 0x7fffffffLL) {
        // the streaming kernel: any hop (the accumulator is position-addressed); positions are 32-bit inside a chunk
        p.ola_fast = 0;
        p.l2_prefetch = 0;
        // AL_IP_SPLIT=75 selects 7 spectrum-row + 5 mask-row slots (A/B of the split; default 8 + 4)
        static const int split75 = getenv("AL_IP_SPLIT") && atoi(getenv("AL_IP_SPLIT")) == 75;
        const size_t smem4 = split75 ? tk_launch_shape(p, n_chunks, n_sm, 7, 5) : tk_launch_shape(p, n_chunks, n_sm);
#define AL_TK_LAUNCH(KERNEL)                                                                                          \
    do {                                                                                                               \
        static PerDeviceOnce attr4;                                                                                    \
        if (attr4.needed()) {                                                                                          \
            cudaError_t e = cudaFuncSetAttribute(KERNEL, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);     \
            if (e != cudaSuccess) return e;                                                                            \
            attr4.mark();                                                                                              \
        }                                                                                                              \
        KERNEL<<<(unsigned)(rows * p.segs), kTkThreads, smem4, stream>>>(p);                                          \
    } while (0)
        if (split75) {
            if (p.mask) AL_TK_LAUNCH((istft_pk4_kernel<true, 7, 5>)); else AL_TK_LAUNCH((istft_pk4_kernel<false, 7, 5>));
        } else {
            if (p.mask) AL_TK_LAUNCH((istft_pk4_kernel<true>)); else AL_TK_LAUNCH((istft_pk4_kernel<false>));
        }
#undef AL_TK_LAUNCH
        count_launch();
        return cudaGetLastError();
    }
    if (ring && (kIpN + p.hop - 1) / p.hop <= kIpKMax) {
        static const int ola_fast3 = (getenv("AL_IP_OLAFAST") && atoi(getenv("AL_IP_OLAFAST")) == 0) ? 0 : 1;
        p.ola_fast = ola_fast3;
        p.l2_prefetch = 0;
        const size_t smem3 = rg_launch_shape(p, n_chunks, n_sm);
        if (smem3 <= (size_t)227 * 1024) {
            static PerDeviceOnce attr3[2];
            PerDeviceOnce& a3 = attr3[p.mask ? 1 : 0];
            if (a3.needed()) {
                cudaError_t e = p.mask ? cudaFuncSetAttribute(istft_pk3_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)
                                       : cudaFuncSetAttribute(istft_pk3_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
                if (e != cudaSuccess) return e;
                a3.mark();
            }
            if (p.mask) istft_pk3_kernel<true><<<(unsigned)(rows * p.segs), kRgThreads, smem3, stream>>>(p);
            else istft_pk3_kernel<false><<<(unsigned)(rows * p.segs), kRgThreads, smem3, stream>>>(p);
            count_launch();
            return cudaGetLastError();
        }
    }
    static const int W = (getenv("AL_IP_WARPS") && atoi(getenv("AL_IP_WARPS")) == 8) ? 8 : 4;
    static const int ola_fast = (getenv("AL_IP_OLAFAST") && atoi(getenv("AL_IP_OLAFAST")) == 0) ? 0 : 1;
    p.ola_fast = ola_fast;
    static const int l2pf = (getenv("AL_IP_L2PF") && atoi(getenv("AL_IP_L2PF")) == 0) ? 0 : 1;
    p.l2_prefetch = l2pf;
    const size_t smem = ip_launch_shape(p, n_chunks, n_sm, W);
    const size_t cap = 227 * 1024;
    if (smem > cap) return cudaErrorInvalidValue;
#define AL_IP_LAUNCH(MSK, WW, PP)                                                                                    \
    do {                                                                                                      \
        static PerDeviceOnce attr;                                                                             \
        if (attr.needed()) {                                                                                          \
            cudaError_t e = cudaFuncSetAttribute(istft_pk2_kernel<MSK, WW, PP>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                                 (int)cap);                                                   \
            if (e != cudaSuccess) return e;                                                                   \
            e = cudaFuncSetAttribute(istft_pk2_kernel<MSK, WW, PP>, cudaFuncAttributePreferredSharedMemoryCarveout,  \
                                     cudaSharedmemCarveoutMaxShared);                                         \
            if (e != cudaSuccess) return e;                                                                   \
            attr.mark();                                                                                      \
        }                                                                                                     \
        istft_pk2_kernel<MSK, WW, PP><<<(unsigned)(rows * p.segs), WW * 32, smem, stream>>>(p);                 \
    } while (0)
    static const int PRE = getenv("AL_IP_PRE") ? atoi(getenv("AL_IP_PRE")) : 0;   // measured (profiles/r01l): 0 -> 37.1 %, 2 -> 28.6 %, 3 -> 26.4 % of HBM
#define AL_IP_PICK(WW, PP) do { if (p.mask) AL_IP_LAUNCH(true, WW, PP); else AL_IP_LAUNCH(false, WW, PP); } while (0)
    if (W == 8) {
        if (PRE == 0) AL_IP_PICK(8, 0); else AL_IP_PICK(8, 2);
    } else {
        if (PRE == 0) AL_IP_PICK(4, 0); else if (PRE == 3) AL_IP_PICK(4, 3); else AL_IP_PICK(4, 2);
    }
#undef AL_IP_PICK
#undef AL_IP_LAUNCH
    count_launch();
    return cudaGetLastError();
}

}  // namespace al
