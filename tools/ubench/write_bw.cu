// Microbenchmark: pure global-store bandwidth vs resident threads per SM (STG.128, fully coalesced).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void w(float4* out, long long n4, float v) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) out[i] = make_float4(v, v, v, v);
}
// each warp writes 33 consecutive 512 B rows per "frame" like the STFT kernel, frames strided across warps
__global__ void wf(float4* out, long long frames, float v) {
    const int lane = threadIdx.x & 31;
    const long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long f = wid; f < frames; f += nw) {
        float4* o = out + f * 1025;
#pragma unroll
        for (int r = 0; r < 32; ++r) o[32 * r + lane] = make_float4(v, v + r, v, v);
        if (lane == 0) o[1024] = make_float4(v, v, v, v);
    }
}
int main() {
    const long long bytes = 840ll << 20; float4* d; cudaMalloc(&d, bytes);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int cfgs[][2] = {{148, 384}, {148, 1024}, {296, 1024}, {148 * 8, 256}, {148 * 16, 128}};
    for (auto& c : cfgs) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0); w<<<c[0], c[1]>>>(d, bytes / 16, 1.f); cudaEventRecord(e1); cudaDeviceSynchronize();
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep) printf("linear  grid %5d x %4d : %7.1f GB/s\n", c[0], c[1], bytes / ms / 1e6);
        }
        const long long frames = bytes / (1025 * 16);
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0); wf<<<c[0], c[1]>>>(d, frames, 1.f); cudaEventRecord(e1); cudaDeviceSynchronize();
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep) printf("frames  grid %5d x %4d : %7.1f GB/s\n", c[0], c[1], frames * 1025 * 16 / ms / 1e6);
        }
    }
    return 0;
}
