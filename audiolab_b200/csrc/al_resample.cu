// K3: polyphase FIR resampler with scipy.signal.resample_poly semantics (zero-phase, zero ends):
//   out[m] = sum_k taps[k] * x_up[m*down + half - k],  x_up[j*up] = x[j]
//          = sum_i taps[phi + up*i] * x[b - i],  c = m*down + half, phi = c mod up, b = c div up.
// Replaces librosa.load(sr=44100) / res_type "polyphase" (reference:
// modules/separator/stem_separator.py:865; modules/rvc/infer/lib/uvr5_pack/lib_v5/model_param_init.py:22).
//
// HBM-bound: 4*down/up B read + 4 B written per output sample.  A CTA produces a tile of outputs,
// stages the input span and the tap table in shared memory; taps are stored [i][phi] so the phase
// stride across lanes (down mod up) spreads over the banks.
#include "al_kernels.h"

namespace al {

constexpr int kResTile = 2048;   // outputs per CTA

__global__ void __launch_bounds__(256)
resample_kernel(const float* __restrict__ in, long long in_stride, float* __restrict__ out,
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

cudaError_t launch_resample(const float* in, long long in_stride, float* out, long long out_stride,
                            int rows, long long n_in, long long n_out, int up, int down,
                            const float* taps, int n_taps, cudaStream_t stream) {
    if (rows <= 0 || n_out <= 0) return cudaSuccess;
    const int tpp = (n_taps + up - 1) / up;
    const int span_cap = (int)(((long long)kResTile * down) / up + tpp + 4);
    const size_t smem = ((size_t)tpp * up + span_cap) * sizeof(float);
    if (smem > 160 * 1024) return cudaErrorInvalidValue;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    dim3 grid((unsigned)((n_out + kResTile - 1) / kResTile), (unsigned)rows);
    resample_kernel<<<grid, 256, smem, stream>>>(in, in_stride, out, out_stride, n_in, n_out, up, down, taps,
                                                 n_taps, tpp, span_cap);
    count_launch();
    return cudaGetLastError();
}

}  // namespace al
