// Per-phase cycle breakdown of the packed STFT kernel (consumer warps), BASELINE cfg2 shape.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DAL_PK_PROF -I../../audiolab_b200/csrc -o stft_pk_phases stft_pk_phases.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
namespace al { void count_launch() {} }
#include "al_stft_pk.cu"

int main(int argc, char** argv) {
    const int stagger = argc > 1 ? atoi(argv[1]) : 0;
    const int n_chunks = 16, chunk = 352800, T = 801, hop = 441;
    const long long n = (long long)n_chunks * chunk;
    float *track, *spec, *win; float2 *tw, *ctw; long long* prof;
    cudaMalloc(&track, 2 * n * 4); cudaMemset(track, 0, 2 * n * 4);
    cudaMalloc(&spec, (size_t)n_chunks * T * 1025 * 16);
    std::vector<float> hw(2048); for (int i = 0; i < 2048; ++i) hw[i] = 0.5f - 0.5f * cosf(6.2831853f * i / 2048);
    std::vector<float2> htw(1024), hctw(544, make_float2(0, 0));
    for (int k1 = 0; k1 < 32; ++k1) for (int n2 = 0; n2 < 32; ++n2) { double a = -6.283185307179586 * k1 * n2 / 1024; htw[k1 * 32 + n2] = make_float2(cos(a), sin(a)); }
    for (int k = 0; k <= 512; ++k) { double a = -6.283185307179586 * k / 2048; hctw[k] = make_float2(0.5 * cos(a), 0.5 * sin(a)); }
    cudaMalloc(&win, 2048 * 4); cudaMemcpy(win, hw.data(), 2048 * 4, cudaMemcpyHostToDevice);
    cudaMalloc(&tw, 1024 * 8); cudaMemcpy(tw, htw.data(), 1024 * 8, cudaMemcpyHostToDevice);
    cudaMalloc(&ctw, 544 * 8); cudaMemcpy(ctw, hctw.data(), 544 * 8, cudaMemcpyHostToDevice);
    const int nprof = 148 * 16 * 5;
    cudaMalloc(&prof, nprof * 8); cudaMemset(prof, 0, nprof * 8);
    al::StftPkParams p{};
    p.track = track; p.n_valid = n; p.ch_stride = n; p.chunk_offsets = nullptr; p.off0 = 0; p.off_step = chunk;
    p.n_chunks = n_chunks; p.chunk_len = chunk; p.center = 1024; p.hop = hop; p.n_frames = T; p.window = win; p.tw = tw;
    p.ctw_half = ctw; p.spec = spec; p.layout = 3; p.n_bins_out = 1025; p.zero_low_bins = 0; p.aligned = 1; p.prof = prof;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        cudaError_t e = al::launch_stft_pk(p, 0);
        cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("launch %d: %s  %.1f us  -> %.0f GB/s\n", rep, cudaGetErrorString(e), ms * 1e3, 255340800.0 / ms / 1e6);
    }
    std::vector<long long> h(nprof); cudaMemcpy(h.data(), prof, nprof * 8, cudaMemcpyDeviceToHost);
    double s[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < nprof; ++i) s[i % 5] += h[i];
    const double frames = 16.0 * 801;
    const char* names[5] = {"wait full", "load+window", "fft1024", "combine+store", "loop overhead"};
    double tot = 0; for (int k = 0; k < 5; ++k) tot += s[k];
    for (int k = 0; k < 5; ++k) printf("%-14s %8.0f cycles per frame per warp (%4.1f%%)\n", names[k], s[k] / frames, 100 * s[k] / tot);
    printf("total %8.0f cycles per frame per warp;  x %d warps -> per-SM cycles per frame = %.0f\n", tot / frames, al::kPkWarps, tot / frames / al::kPkWarps);
    return 0;
}
