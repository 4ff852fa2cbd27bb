// Microbenchmark: FFMA vs packed FFMA2 (fma.rn.f32x2) / FADD vs FADD2 issue throughput on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu && ./f32x2
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float s) {
    float2 a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f - i);
    const float2 m = make_float2(s, s * 0.5f), c = make_float2(s * 0.25f, -s);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) { a[i].x = fmaf(a[i].x, m.x, c.x); a[i].y = fmaf(a[i].y, m.y, c.y); }
            if (MODE == 1) { a[i] = __ffma2_rn(a[i], m, c); }
            if (MODE == 2) { a[i].x = a[i].x + c.x; a[i].y = a[i].y + c.y; }
            if (MODE == 3) { a[i] = __fadd2_rn(a[i], c); }
            if (MODE == 4) { a[i].x = a[i].x * m.x; a[i].y = a[i].y * m.y; }
            if (MODE == 5) { a[i] = __fmul2_rn(a[i], m); }
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += a[i].x + a[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char* name, float* d) {
    const int iters = 20000, blocks = 148 * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(d, 100, 1.0001f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(d, iters, 1.0001f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)blocks * 256 * iters * 16;   // scalar fp32 ops (fma counted once)
    printf("%-8s %8.3f ms  %8.2f Tlane-op/s  (%.1f lane-ops/clk/SM at 1.965 GHz)\n", name, ms, ops / ms / 1e9,
           ops / (ms * 1e-3) / 148 / 1.965e9);
}

int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("FFMA", d); run<1>("FFMA2", d); run<2>("FADD", d); run<3>("FADD2", d); run<4>("FMUL", d); run<5>("FMUL2", d);
    return 0;
}
