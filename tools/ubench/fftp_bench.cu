// Microbenchmark: packed warp FFT (al_fftp.cuh) throughput vs resident warps per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../audiolab_b200/csrc -o fftp_bench fftp_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "al_fftp.cuh"
using namespace al;

template <int MODE>
__global__ void __launch_bounds__(384, 1) k(const float2* tw_g, float2* out, int iters, long long* cyc) {
    extern __shared__ __align__(16) unsigned char sm[];
    float2* s_tw = (float2*)sm;
    float2* s_scr = s_tw + 1024;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) s_tw[i] = tw_g[i];
    __syncthreads();
    float2 re[32], im[32];
#pragma unroll
    for (int r = 0; r < 32; ++r) { re[r] = make_float2(lane + r, lane - r); im[r] = make_float2(r * 0.5f, lane * 0.25f); }
    float2* scr = s_scr + warp * kScrF2;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) warp_fft1024p<false>(re, im, scr, s_tw, lane);
        if (MODE == 3) {   // single-pass transposition through a 16-byte-element scratch
            fft32p_fwd(re, im);
#pragma unroll
            for (int k1 = 1; k1 < 32; ++k1) pcmul<false>(re[k1], im[k1], s_tw[k1 * 32 + lane]);
            float4* wr = (float4*)(s_scr + warp * 2 * kScrF2) + lane * 33;
            const float4* rd = (const float4*)(s_scr + warp * 2 * kScrF2) + lane;
#pragma unroll
            for (int k1 = 0; k1 < 32; ++k1) wr[k1] = make_float4(re[k1].x, re[k1].y, im[k1].x, im[k1].y);
            __syncwarp();
#pragma unroll
            for (int n2 = 0; n2 < 32; ++n2) { const float4 v = rd[n2 * 33]; re[n2] = make_float2(v.x, v.y); im[n2] = make_float2(v.z, v.w); }
            __syncwarp();
            fft32p_fwd(re, im);
        }
        if (MODE == 1) { fft32p_fwd(re, im); fft32p_fwd(re, im); }          // FP only
        if (MODE == 2) {                                                     // transposition only
            float2* wr = scr + lane * 33; const float2* rd = scr + lane;
#pragma unroll
            for (int k1 = 0; k1 < 32; ++k1) wr[k1] = re[k1];
            __syncwarp();
#pragma unroll
            for (int n2 = 0; n2 < 32; ++n2) re[n2] = rd[n2 * 33];
            __syncwarp();
#pragma unroll
            for (int k1 = 0; k1 < 32; ++k1) wr[k1] = im[k1];
            __syncwarp();
#pragma unroll
            for (int n2 = 0; n2 < 32; ++n2) im[n2] = rd[n2 * 33];
            __syncwarp();
        }
    }
    long long t1 = clock64();
    float2 acc = make_float2(0, 0);
#pragma unroll
    for (int r = 0; r < 32; ++r) acc = padd(acc, padd(re[r], im[r]));
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int warps, float2* tw, float2* out, long long* cyc) {
    const int iters = 200;
    size_t smem = (1024 + (size_t)warps * kScrF2 * (MODE == 3 ? 2 : 1)) * sizeof(float2);
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    k<MODE><<<148, warps * 32, smem>>>(tw, out, 10, cyc);
    cudaDeviceSynchronize();
    k<MODE><<<148, warps * 32, smem>>>(tw, out, iters, cyc);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-14s warps/SM=%2d  cycles per iteration per warp = %8.1f   per-SM cycles per FFT(2ch) = %8.1f  (%s)\n", name, warps,
           (double)c / iters, (double)c / iters / warps, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    float2 *tw, *out; long long* cyc;
    cudaMalloc(&tw, 1024 * 8); cudaMemset(tw, 0, 1024 * 8);
    cudaMalloc(&out, 148 * 512 * 8); cudaMalloc(&cyc, 8);
    for (int w : {4, 8, 12}) run<0>("fft1024p", w, tw, out, cyc);
    for (int w : {4, 8, 11, 12}) run<3>("fft1024p 1pass", w, tw, out, cyc);
    return 0;
}
