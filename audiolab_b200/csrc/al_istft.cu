// K2: fused (complex mask (.) spec | CaC -> complex) + zero freq-pad + complex-to-real iFFT +
// synthesis window + overlap-add over frames + / sum(window^2) + centre trim (+ chunk weight,
// + trim-and-concat placement).  Replaces torch.istft and its surroundings
// (reference: modules/rvc/infer/modules/uvr5/mdxnet.py:58-75, :178-183).
//
// HBM-bound.  Algorithmic bytes per frame: n_bins_in*8 (spectrum) [+ n_bins*8 mask] + hop*4 (wave).
//
// Determinism: every output sample is owned by exactly one thread of one CTA and is the
// left-to-right sum of its frames in ascending frame order (a carry buffer hands partial sums
// from one round of G frames to the next), so the result does not depend on the tiling, the
// batch composition or the number of GPUs.  The R-1 frames that precede a CTA's first owned
// sample are recomputed (halo) instead of exchanged through global atomics.
#include "al_kernels.h"

namespace al {

// [emul-begin]
constexpr int kOlaKMax = 6;   // frames covering one position in the register form of the overlap-add (n_fft <= 6 hop)

template <int D>
__global__ void __launch_bounds__(Cfg<D>::UW * 32)
istft_kernel(const IstftParams p) {
    constexpr int G = Cfg<D>::G, UW = Cfg<D>::UW, NT = UW * 32, N = D * 1024, HW = D / 2;
    AL_DYN_SMEM(unsigned char, smem_raw);
    float2* s_tw = reinterpret_cast<float2*>(smem_raw);     // [1024]
    float2* s_slot = s_tw + 1024;                            // [UW][kSlotF2]
    float* s_carry = reinterpret_cast<float*>(s_slot + UW * kSlotF2);   // [N - hop], updated in place

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int orow = blockIdx.x / p.segs, seg = blockIdx.x - orow * p.segs;   // orow = (chunk*stems + s)*channels + c
    const int ch = orow % p.channels;
    const int cs = orow / p.channels;
    const int stem = cs % p.stems, chunk = cs / p.stems;
    const long long srow = p.spec_has_stems ? orow : (long long)chunk * p.channels + ch;
    const int hop = p.hop;
    const int carry_len = N - hop;
    const int ola_k = (N + hop - 1) / hop;              // frames that can cover one position

    const SpecView sview{const_cast<float*>(p.spec), p.layout, p.n_frames_in, p.n_bins_in, p.channels};
    const SpecView mview{const_cast<float*>(p.mask), p.layout, p.n_frames_in, N / 2 + 1, p.channels};

    // owned untrimmed OLA positions [Pa, Pb)
    const long long Pa = (long long)p.out_start + (long long)seg * p.hops_per_cta * hop;
    const long long Pend = (long long)p.out_start + p.out_len;
    const long long Pb = min(Pa + (long long)p.hops_per_cta * hop, Pend);
    if (Pa >= Pb) return;
    int ta = (int)((Pa - N) / hop) + 1;                 // first frame touching Pa  (Pa >= N/2 > 0)
    if (Pa < N) ta = 0;
    ta = max(ta, 0);
    const int tb = min((int)((Pb - 1) / hop), p.n_frames_total - 1);   // last frame touching Pb-1
    // T-innermost planes (AL_LAYOUT_CAC): four consecutive frames of one bin are one aligned 16-byte load when the
    // rounds start on a multiple of 4 frames of the stored spectrogram.  The first round may then start up to 3
    // frames early (even below frame 0): those frames end before Pa or do not exist, so they add nothing to the
    // owned samples and the ascending-frame order of every emitted sample is unchanged.
    const bool vec4 = p.layout == 2 && !p.mask && (G % 4) == 0 && (p.n_frames_in & 3) == 0 &&
                      (reinterpret_cast<uintptr_t>(p.spec) & 15) == 0;
    if (vec4) ta -= (((ta - p.frame_pad) % 4) + 4) % 4;

    const long long place = p.dst_offsets ? p.dst_offsets[chunk] : p.dst_off0 + (long long)chunk * p.dst_off_step;
    float* __restrict__ dst = p.dst + ((long long)stem * p.channels + ch) * p.dst_ch_stride +
                              (long long)chunk * p.dst_chunk_stride + place;

    for (int i = tid; i < 1024; i += NT) s_tw[i] = p.tw[i];
    for (int i = tid; i < carry_len; i += NT) s_carry[i] = 0.f;

    // rounds run past the last frame until the carry has been flushed up to Pb
    const int t_last = (int)((Pb - 1) / hop);
    for (int tr = ta; tr <= t_last; tr += G) {
        __syncthreads();   // slots free (previous OLA finished), tables visible
        const int nf = max(0, min(G, tb - tr + 1));   // live frames in this round (CTA-uniform)
        if (nf > 0) {

        // ---- stage A: load (x mask), Hermitian extension, inverse radix-D -> X_r[kappa] -------------
        const bool t_fast = (p.layout == 1 || p.layout == 2);
        if (vec4) {
            // item = (kappa, 4 frames): 2 x D aligned 128-bit loads (re / im plane of each of the D bins), all in
            // flight before the first butterfly; then one radix-D butterfly per frame
            constexpr int GV = G / 4 > 0 ? G / 4 : 1;
            const long long plane = (long long)p.n_bins_in * p.n_frames_in;
            const float* __restrict__ rowp = p.spec + srow * 2 * plane;
            for (int it = tid; it < GV * 513; it += NT) {
                const int kappa = it / GV, gq = it - kappa * GV;
                const int t0 = tr + 4 * gq, ts0 = t0 - p.frame_pad;          // ts0 is a multiple of 4
                const bool any = t0 <= tb && ts0 >= 0 && ts0 < p.n_frames_in;
                float4 re4[D], im4[D];
#pragma unroll
                for (int q = 0; q < D; ++q) {
                    const int k = kappa + 1024 * q;
                    const int bin = (k <= N / 2) ? k : N - k;
                    re4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                    im4[q] = re4[q];
                    if (any && bin < p.n_bins_in && bin >= p.zero_low_bins) {
                        const float* __restrict__ src = rowp + (long long)bin * p.n_frames_in + ts0;
                        re4[q] = __ldg(reinterpret_cast<const float4*>(src));
                        im4[q] = __ldg(reinterpret_cast<const float4*>(src + plane));
                    }
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    float2 y[D];
#pragma unroll
                    for (int q = 0; q < D; ++q) {
                        const int k = kappa + 1024 * q;
                        const int bin = (k <= N / 2) ? k : N - k;
                        float2 v;
                        v.x = e == 0 ? re4[q].x : e == 1 ? re4[q].y : e == 2 ? re4[q].z : re4[q].w;
                        v.y = e == 0 ? im4[q].x : e == 1 ? im4[q].y : e == 2 ? im4[q].z : im4[q].w;
                        if (t0 + e > tb) v = make_float2(0.f, 0.f);          // like the scalar path: frames past tb are empty
                        if (bin == 0 || bin == N / 2) v.y = 0.f;             // C2R ignores Im of DC / Nyquist
                        if (k > N / 2) v.y = -v.y;
                        y[q] = v;
                    }
                    SmallDft<D, true>::run(y);
                    float2* xs = s_slot + ((4 * gq + e) * HW) * kSlotF2 + kappa;
                    xs[0] = y[0];
#pragma unroll
                    for (int r = 1; r < D; ++r)
                        xs[(r >> 1) * kSlotF2 + (r & 1) * kXHalf] = cmul_conj(y[r], __ldg(p.ctw + (r - 1) * 513 + kappa));
                }
            }
        } else {
        // two (frame, kappa) items per iteration: the 2 x D (x 2 planes, x 2 with a mask) loads of both are in
        // flight before the first butterfly
        auto sa_load = [&](int it, float2 (&y)[D], int& f, int& kappa) {
            if (t_fast) { kappa = it / G; f = it - kappa * G; }
            else        { f = it / 513;  kappa = it - f * 513; }
            const int t = tr + f;
            const int ts = t - p.frame_pad;
            const bool live = (t <= tb) && ts >= 0 && ts < p.n_frames_in;
#pragma unroll
            for (int q = 0; q < D; ++q) {
                const int k = kappa + 1024 * q;
                const int bin = (k <= N / 2) ? k : N - k;
                float2 v = make_float2(0.f, 0.f);
                if (live && bin < p.n_bins_in && bin >= p.zero_low_bins) {
                    v = spec_load(sview, srow, ts, bin);
                    if (p.mask) v = cmul(v, spec_load(mview, orow, ts, bin));
                }
                if (bin == 0 || bin == N / 2) v.y = 0.f;   // C2R ignores Im of DC / Nyquist
                if (k > N / 2) v.y = -v.y;
                y[q] = v;
            }
        };
        auto sa_finish = [&](float2 (&y)[D], int f, int kappa) {
            SmallDft<D, true>::run(y);
            float2* xs = s_slot + (f * HW) * kSlotF2 + kappa;
            xs[0] = y[0];
#pragma unroll
            for (int r = 1; r < D; ++r)
                xs[(r >> 1) * kSlotF2 + (r & 1) * kXHalf] = cmul_conj(y[r], __ldg(p.ctw + (r - 1) * 513 + kappa));
        };
        int it = tid;
        for (; it + NT < G * 513; it += 2 * NT) {
            float2 ya[D], yb[D];
            int fa, ka, fb, kb;
            sa_load(it, ya, fa, ka);
            sa_load(it + NT, yb, fb, kb);
            sa_finish(ya, fa, ka);
            sa_finish(yb, fb, kb);
        }
        if (it < G * 513) {
            float2 ya[D];
            int fa, ka;
            sa_load(it, ya, fa, ka);
            sa_finish(ya, fa, ka);
        }
        }  // scalar stage A
        __syncthreads();

        // ---- unit inverse FFT: warp = (frame f, pair w) ----------------------------------------------
        {
            const int f = warp / HW, w = warp - f * HW;
            float2* slot = s_slot + warp * kSlotF2;
            float re[32], im[32];
            // Z_w[kappa] = X_2w[kappa] + i X_2w+1[kappa]; kappa > 512 through Hermitian symmetry
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                const int kappa = 32 * r + lane;
                if (r < 16 || (r == 16 && lane == 0)) {
                    const float2 a = slot[kappa], b = slot[kXHalf + kappa];
                    re[r] = a.x - b.y;
                    im[r] = a.y + b.x;
                } else {
                    const float2 a = slot[1024 - kappa], b = slot[kXHalf + 1024 - kappa];
                    re[r] = a.x + b.y;
                    im[r] = b.x - a.y;
                }
            }
            __syncwarp();
            warp_fft1024<true>(re, im, slot, s_tw, lane);
            // z[n] = (x[D n + 2w], x[D n + 2w + 1]); apply the synthesis window, park the frame in the slot
            const float2* __restrict__ win2 = reinterpret_cast<const float2*>(p.window) + w + HW * lane;
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                const float2 wv = __ldg(win2 + HW * 32 * r);
                slot[32 * r + lane] = make_float2(re[r] * wv.x, im[r] * wv.y);
            }
        }
        __syncthreads();
        }  // nf > 0

        // ---- overlap-add: span [S, S + G*hop + carry_len), ascending frame order ---------------------
        // A thread owns offsets j (< hop) and walks the hop-blocks h of the span: position i = h hop + j
        // receives frame f at sample (h - f) hop + j, f = max(0, h - kj) .. min(nf - 1, h), with
        // kj = (N - 1 - j) / hop fixed per thread (no per-sample division by the run-time hop).  The
        // carry is updated in place: carry[(h - G) hop + j] is written after the same thread read it.
        const long long S = (long long)tr * hop;
        const float* frames = reinterpret_cast<const float*>(s_slot);
        const int emit = G * hop;
        const int span = emit + carry_len;
        // Interior round (CTA-uniform): every frame is live, every emitted position is owned and lands inside the
        // destination -- the register form without per-position predicates, 32-bit indexing from three bases.
        const long long rel = S - p.out_start;
        if (ola_k <= kOlaKMax && nf == G && S >= Pa && S + emit <= Pb && place + rel >= 0 && place + rel + emit <= p.dst_limit) {
            constexpr int NH = G + kOlaKMax - 1;
            const float* __restrict__ envp = p.inv_env + S;
            const float* __restrict__ wgp = p.weight ? p.weight + rel : nullptr;
            float* __restrict__ d = dst + rel;
            for (int j = tid; j < hop; j += NT) {
                const int kj = (ola_k - 1) * hop + j < N ? ola_k - 1 : ola_k - 2;
                float ev[G], wg[G];
#pragma unroll
                for (int h = 0; h < G; ++h) {
                    ev[h] = __ldg(envp + h * hop + j);
                    wg[h] = wgp ? __ldg(wgp + h * hop + j) : 1.f;
                }
                int idx[kOlaKMax];
#pragma unroll
                for (int k = 0; k < kOlaKMax; ++k) {
                    const int o = k * hop + j;
                    idx[k] = (k <= kj) ? ((((o % D) >> 1) * kSlotF2 + o / D) * 2 + (o & 1)) : -1;
                }
                float out[NH];
#pragma unroll
                for (int h = 0; h < NH; ++h) {
                    const int i = h * hop + j;
                    out[h] = (i < carry_len) ? s_carry[i] : 0.f;
                }
#pragma unroll
                for (int f = 0; f < G; ++f) {
                    const float* __restrict__ fr = frames + f * (HW * kSlotF2 * 2);
#pragma unroll
                    for (int k = 0; k < kOlaKMax; ++k)
                        if (idx[k] >= 0) out[f + k] += fr[idx[k]];
                }
#pragma unroll
                for (int h = 0; h < G; ++h) {
                    float v = out[h] * ev[h];
                    if (wgp) v *= wg[h];
                    d[h * hop + j] = v;
                }
#pragma unroll
                for (int h = G; h < NH; ++h) {
                    const int i = h * hop + j;
                    if (i < span) s_carry[i - emit] = out[h];
                }
            }
            continue;
        }
        for (int j = tid; j < hop; j += NT) {
            const int kj = (ola_k - 1) * hop + j < N ? ola_k - 1 : ola_k - 2;
            // 1 / envelope and chunk weight of the emitted positions: loads issued before their first use
            float ev[G], wg[G];
#pragma unroll
            for (int h = 0; h < G; ++h) {
                const long long P = S + h * hop + j;
                const bool mine = P >= Pa && P < Pb;
                ev[h] = mine ? __ldg(p.inv_env + P) : 0.f;
                wg[h] = (mine && p.weight) ? __ldg(p.weight + (P - p.out_start)) : 1.f;
            }
            if (ola_k <= kOlaKMax) {
                // Register form (CTA-uniform branch): the thread's G + K - 1 hop-blocks are accumulators; frames are
                // added in ascending order, frame f feeding blocks f .. f + kj from the same K slot offsets.
                constexpr int NH = G + kOlaKMax - 1;
                int idx[kOlaKMax];
#pragma unroll
                for (int k = 0; k < kOlaKMax; ++k) {
                    const int o = k * hop + j;                // sample of the frame that lands on block f + k
                    idx[k] = (k <= kj) ? ((((o % D) >> 1) * kSlotF2 + o / D) * 2 + (o & 1)) : -1;
                }
                float out[NH];
#pragma unroll
                for (int h = 0; h < NH; ++h) {
                    const int i = h * hop + j;
                    out[h] = (i < carry_len) ? s_carry[i] : 0.f;
                }
#pragma unroll
                for (int f = 0; f < G; ++f) {
                    if (f < nf) {
                        const float* __restrict__ fr = frames + f * (HW * kSlotF2 * 2);
#pragma unroll
                        for (int k = 0; k < kOlaKMax; ++k)
                            if (idx[k] >= 0) out[f + k] += fr[idx[k]];
                    }
                }
#pragma unroll
                for (int h = 0; h < G; ++h) {
                    const long long P = S + h * hop + j;
                    if (P >= Pa && P < Pb) {
                        const long long pp = P - p.out_start;
                        float v = out[h] * ev[h];
                        if (p.weight) v *= wg[h];
                        const long long q = place + pp;
                        if (q >= 0 && q < p.dst_limit) dst[pp] = v;
                    }
                }
#pragma unroll
                for (int h = G; h < NH; ++h) {
                    const int i = h * hop + j;
                    if (i < span) s_carry[i - emit] = out[h];
                }
                continue;
            }
#pragma unroll
            for (int h = 0; h < G; ++h) {                     // emitted hop-blocks
                const int i = h * hop + j;
                float acc = (i < carry_len) ? s_carry[i] : 0.f;
                const int f_lo = max(0, h - kj);
                const int f_hi = min(nf - 1, h);
                int o = (h - f_lo) * hop + j;                 // sample of frame f_lo, 0 <= o < N
                for (int f = f_lo; f <= f_hi; ++f, o -= hop) {
                    const int w = (o % D) >> 1, n = o / D, c = o & 1;
                    acc += frames[((f * HW + w) * kSlotF2 + n) * 2 + c];
                }
                const long long P = S + i;
                if (P >= Pa && P < Pb) {
                    const long long pp = P - p.out_start;
                    float v = acc * ev[h];
                    if (p.weight) v *= wg[h];
                    const long long q = place + pp;
                    if (q >= 0 && q < p.dst_limit) dst[pp] = v;
                }
            }
            for (int h = G, i = emit + j; i < span; ++h, i += hop) {   // hop-blocks that stay in the carry
                float acc = (i < carry_len) ? s_carry[i] : 0.f;
                const int f_lo = max(0, h - kj);
                const int f_hi = min(nf - 1, h);
                int o = (h - f_lo) * hop + j;
                for (int f = f_lo; f <= f_hi; ++f, o -= hop) {
                    const int w = (o % D) >> 1, n = o / D, c = o & 1;
                    acc += frames[((f * HW + w) * kSlotF2 + n) * 2 + c];
                }
                s_carry[i - emit] = acc;
            }
        }
    }
}

// launch shape of istft_kernel<D>: fills hops_per_cta / segs, returns the dynamic shared memory size
template <int D>
static size_t istft_tiling(IstftParams& p, int rows) {
    constexpr int UW = Cfg<D>::UW, N = D * 1024;
    const int total_hops = (p.out_len + p.hop - 1) / p.hop;
    // enough CTAs for ~4 waves over 148 SMs, but segments no shorter than 16 hops (halo <= ~30 %)
    int segs = (4 * 148 + rows - 1) / rows;
    int hpc = (total_hops + segs - 1) / segs;
    hpc = hpc > 16 ? hpc : 16;
    p.hops_per_cta = hpc;
    p.segs = (total_hops + hpc - 1) / hpc;
    return 1024 * sizeof(float2) + (size_t)UW * kSlotF2 * sizeof(float2) + (size_t)(N - p.hop) * sizeof(float);
}
// [emul-end]

template <int D>
static cudaError_t launch_istft_d(const IstftParams& p0, int n_chunks, cudaStream_t stream) {
    constexpr int UW = Cfg<D>::UW;
    IstftParams p = p0;
    const int rows = n_chunks * p.stems * p.channels;
    const size_t smem = istft_tiling<D>(p, rows);
    static PerDeviceOnce attr_set;
    if (attr_set.needed()) {
        cudaError_t e = cudaFuncSetAttribute(istft_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        attr_set.mark();
    }
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    istft_kernel<D><<<(unsigned)(rows * p.segs), UW * 32, smem, stream>>>(p);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_istft(const IstftParams& p, int n_fft, int n_chunks, cudaStream_t stream) {
    switch (n_fft) {
        case 2048: return launch_istft_d<2>(p, n_chunks, stream);
        case 4096: return launch_istft_d<4>(p, n_chunks, stream);
        case 6144: return launch_istft_d<6>(p, n_chunks, stream);
        default: return cudaErrorInvalidValue;
    }
}

// inv_env[P] = 1 / sum_t w[P - t*hop]^2 over frames t in [0, n_frames_total)   (0 where empty)
__global__ void env_kernel(const float* __restrict__ w, int n_fft, int hop, int n_frames_total,
                           float* __restrict__ inv_env, long long total) {
    const long long P = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (P >= total) return;
    long long t_lo = (P - n_fft) / hop + 1;
    if (P < n_fft) t_lo = 0;
    long long t_hi = P / hop;
    if (t_hi > n_frames_total - 1) t_hi = n_frames_total - 1;
    double acc = 0.0;
    for (long long t = t_lo; t <= t_hi; ++t) {
        const float v = w[P - t * hop];
        acc += (double)v * (double)v;
    }
    inv_env[P] = acc > 1e-11 ? (float)(1.0 / acc) : 0.f;
}

cudaError_t launch_env(const float* window_raw, int n_fft, int hop, int n_frames_total, float* inv_env,
                       cudaStream_t stream) {
    const long long total = (long long)(n_frames_total - 1) * hop + n_fft;
    env_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(window_raw, n_fft, hop, n_frames_total,
                                                                    inv_env, total);
    count_launch();
    return cudaGetLastError();
}

}  // namespace al
