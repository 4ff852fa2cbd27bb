// K1: fused pad-and-chunk + reflect-pad + framing + window + real-to-complex FFT + output layout.
// Replaces torch.stft and the layout shuffles around it
// (reference: modules/rvc/infer/modules/uvr5/mdxnet.py:41-56, :152-164).
//
// HBM-bound.  Algorithmic bytes per frame: hop*4 (wave read) + n_bins_out*8 (spectrum write).
// CTA = UW unit warps; per round it stages the input span of G consecutive frames ONCE into
// shared memory (coalesced, phase-split by D so unit reads are stride-1), runs the G*D/2 unit
// FFTs, and the radix-D combine writes the spectrum in the consumer's layout (frame-fastest
// lanes for the T-innermost layouts so every store fills whole 32-byte sectors).
#include "al_kernels.h"

namespace al {

// [emul-begin]
template <int D>
__global__ void __launch_bounds__(Cfg<D>::UW * 32)
stft_kernel(const StftParams p) {
    constexpr int G = Cfg<D>::G, UW = Cfg<D>::UW, NT = UW * 32, N = D * 1024, HW = D / 2;
    AL_DYN_SMEM(unsigned char, smem_raw);
    float2* s_tw = reinterpret_cast<float2*>(smem_raw);             // [1024]
    float2* s_slot = s_tw + 1024;                                    // [UW][kSlotF2]
    float* s_stage = reinterpret_cast<float*>(s_slot + UW * kSlotF2);  // [D][ps]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int row = blockIdx.x / p.tiles, tile = blockIdx.x - row * p.tiles;
    const int chunk = row / p.channels, ch = row - chunk * p.channels;
    const long long coff = p.chunk_offsets ? p.chunk_offsets[chunk] : p.off0 + (long long)chunk * p.off_step;
    const float* __restrict__ src = p.track + (long long)ch * p.ch_stride;
    const int ps = p.ps;
    const int span = (G - 1) * p.hop + N;
    const SpecView view{p.spec, p.layout, p.n_frames, p.n_bins_out, p.channels};

    for (int i = tid; i < 1024; i += NT) s_tw[i] = p.tw[i];

    for (int round = 0; round < p.rounds_per_cta; ++round) {
        const int t0 = (tile * p.rounds_per_cta + round) * G;
        if (t0 >= p.n_frames) break;   // uniform over the CTA
        __syncthreads();               // previous combine is done with slots / stage; s_tw visible

        // ---- stage the span: reflect about the chunk, zero outside the track ----------------------
        // NT is a multiple of D, so a thread's decimation phase i % D never changes and i / D advances by NT / D:
        // no division in the loop.  Interior spans (no reflection, inside the track) skip the edge logic.
        const long long s0 = (long long)t0 * p.hop - p.center;
        static_assert(NT % D == 0, "the staging loop relies on NT % D == 0");
        float* __restrict__ sdst = s_stage + (tid % D) * ps + tid / D;
        if (s0 >= 0 && s0 + span <= p.chunk_len && coff + s0 >= 0 && coff + s0 + span <= p.n_valid) {
            const float* __restrict__ g = src + coff + s0;
            int q = 0;
            for (int i = tid; i < span; i += NT, q += NT / D) sdst[q] = __ldg(g + i);
        } else {
            int q = 0;
            for (int i = tid; i < span; i += NT, q += NT / D) {
                long long j = s0 + i;
                if (j < 0) j = -j;
                if (j >= p.chunk_len) j = 2LL * (p.chunk_len - 1) - j;
                float v = 0.f;
                if (j >= 0 && j < p.chunk_len) {
                    const long long g = coff + j;
                    if (g >= 0 && g < p.n_valid) v = __ldg(src + g);
                }
                sdst[q] = v;
            }
        }
        __syncthreads();

        // ---- unit FFT: warp = (frame f, pair w) ----------------------------------------------------
        {
            const int f = warp / HW, w = warp - f * HW;
            float re[32], im[32];
            const int e0 = f * p.hop + 2 * w;   // span index of (n = 0, c = 0)
            const float* sa = s_stage + (e0 % D) * ps + e0 / D + lane;
            const float* sb = s_stage + ((e0 + 1) % D) * ps + (e0 + 1) / D + lane;
            const float2* __restrict__ win2 = reinterpret_cast<const float2*>(p.window) + w + HW * lane;
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                const float2 wv = __ldg(win2 + HW * 32 * r);
                re[r] = sa[32 * r] * wv.x;
                im[r] = sb[32 * r] * wv.y;
            }
            float2* slot = s_slot + warp * kSlotF2;
            warp_fft1024<false>(re, im, slot, s_tw, lane);
            // separate the two real spectra (kappa = 32 r + lane <= 512) with the lane-mirror trick
            const int ml = (32 - lane) & 31;
#pragma unroll
            for (int r = 0; r <= 16; ++r) {
                float pr = __shfl_sync(0xffffffffu, re[31 - r], ml);
                float pi = __shfl_sync(0xffffffffu, im[31 - r], ml);
                if (lane == 0) { pr = re[(32 - r) & 31]; pi = im[(32 - r) & 31]; }
                if (r < 16 || lane == 0) {
                    const int kappa = 32 * r + lane;
                    slot[kappa] = make_float2(0.5f * (re[r] + pr), 0.5f * (im[r] - pi));
                    slot[kXHalf + kappa] = make_float2(0.5f * (im[r] + pi), 0.5f * (pr - re[r]));
                }
            }
        }
        __syncthreads();

        // ---- radix-D combine + store ----------------------------------------------------------------
        const bool t_fast = (p.layout == 1 || p.layout == 2);
        // T-innermost planes (AL_LAYOUT_CAC): a thread combines 4 consecutive frames of one kappa and writes each of
        // its 2 x D plane rows as one aligned 16-byte store (t0 is a multiple of G, rows are 16-byte aligned)
        const bool vec4 = p.layout == 2 && (G % 4) == 0 && (p.n_frames & 3) == 0 &&
                          (reinterpret_cast<uintptr_t>(p.spec) & 15) == 0;
        if (vec4) {
            constexpr int GV = G / 4 > 0 ? G / 4 : 1;
            const long long plane = (long long)p.n_bins_out * p.n_frames;
            float* __restrict__ rowp = p.spec + (long long)row * 2 * plane;
            for (int it = tid; it < GV * 513; it += NT) {
                const int kappa = it / GV, gq = it - kappa * GV;
                const int t = t0 + 4 * gq;
                if (t >= p.n_frames) continue;                 // n_frames % 4 == 0: all four frames in or out
                float re4[D][4], im4[D][4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const float2* xs = s_slot + ((4 * gq + e) * HW) * kSlotF2 + kappa;
                    float2 y[D];
#pragma unroll
                    for (int r = 0; r < D; ++r) y[r] = xs[(r >> 1) * kSlotF2 + (r & 1) * kXHalf];
#pragma unroll
                    for (int r = 1; r < D; ++r) y[r] = cmul(y[r], __ldg(p.ctw + (r - 1) * 513 + kappa));
                    SmallDft<D, false>::run(y);
#pragma unroll
                    for (int q = 0; q < D; ++q) {
                        re4[q][e] = y[q].x;
                        im4[q][e] = y[q].y;
                    }
                }
#pragma unroll
                for (int q = 0; q < D; ++q) {
                    const int k = kappa + 1024 * q;
                    int bin;
                    float sgn = 1.f;
                    if (k <= N / 2) {
                        bin = k;
                    } else {
                        if (kappa == 0 || kappa == 512) continue;   // duplicates of directly produced bins
                        bin = N - k;
                        sgn = -1.f;
                    }
                    if (bin >= p.n_bins_out) continue;
                    float4 vr = make_float4(re4[q][0], re4[q][1], re4[q][2], re4[q][3]);
                    float4 vi = make_float4(sgn * im4[q][0], sgn * im4[q][1], sgn * im4[q][2], sgn * im4[q][3]);
                    if (bin < p.zero_low_bins) vr = vi = make_float4(0.f, 0.f, 0.f, 0.f);
                    float* dstp = rowp + (long long)bin * p.n_frames + t;
                    *reinterpret_cast<float4*>(dstp) = vr;
                    *reinterpret_cast<float4*>(dstp + plane) = vi;
                }
            }
            continue;
        }
        for (int it = tid; it < G * 513; it += NT) {
            int f, kappa;
            if (t_fast) { kappa = it / G; f = it - kappa * G; }
            else        { f = it / 513;  kappa = it - f * 513; }
            const int t = t0 + f;
            if (t >= p.n_frames) continue;
            const float2* xs = s_slot + (f * HW) * kSlotF2 + kappa;
            float2 y[D];
#pragma unroll
            for (int r = 0; r < D; ++r) y[r] = xs[(r >> 1) * kSlotF2 + (r & 1) * kXHalf];
#pragma unroll
            for (int r = 1; r < D; ++r) y[r] = cmul(y[r], __ldg(p.ctw + (r - 1) * 513 + kappa));
            SmallDft<D, false>::run(y);
#pragma unroll
            for (int q = 0; q < D; ++q) {
                const int k = kappa + 1024 * q;
                int bin;
                float2 val = y[q];
                if (k <= N / 2) {
                    bin = k;
                } else {
                    if (kappa == 0 || kappa == 512) continue;   // duplicates of directly produced bins
                    bin = N - k;
                    val.y = -val.y;
                }
                if (bin >= p.n_bins_out) continue;
                if (bin < p.zero_low_bins) val = make_float2(0.f, 0.f);
                spec_store(view, row, t, bin, val);
            }
        }
    }
}

// launch shape of stft_kernel<D>: fills ps / rounds_per_cta / tiles, returns the dynamic shared memory size
template <int D>
static size_t stft_tiling(StftParams& p) {
    constexpr int G = Cfg<D>::G, UW = Cfg<D>::UW, N = D * 1024;
    const int span = (G - 1) * p.hop + N;
    p.ps = ((span + D - 1) / D + 31) / 32 * 32 + 32 / D;
    p.rounds_per_cta = (D == 2) ? 1 : 2;
    const int frames_per_cta = G * p.rounds_per_cta;
    p.tiles = (p.n_frames + frames_per_cta - 1) / frames_per_cta;
    return 1024 * sizeof(float2) + (size_t)UW * kSlotF2 * sizeof(float2) + (size_t)D * p.ps * sizeof(float);
}
// [emul-end]

template <int D>
static cudaError_t launch_stft_d(const StftParams& p0, int rows, cudaStream_t stream) {
    constexpr int UW = Cfg<D>::UW;
    StftParams p = p0;
    const size_t smem = stft_tiling<D>(p);
    static PerDeviceOnce attr_set;
    if (attr_set.needed()) {
        cudaError_t e = cudaFuncSetAttribute(stft_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        attr_set.mark();
    }
    if (smem > 200 * 1024) return cudaErrorInvalidValue;
    stft_kernel<D><<<(unsigned)(rows * p.tiles), UW * 32, smem, stream>>>(p);
    count_launch();
    return cudaGetLastError();
}

cudaError_t launch_stft(const StftParams& p, int n_fft, int rows, cudaStream_t stream) {
    switch (n_fft) {
        case 2048: return launch_stft_d<2>(p, rows, stream);
        case 4096: return launch_stft_d<4>(p, rows, stream);
        case 6144: return launch_stft_d<6>(p, rows, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace al
