// Minimal host emulation of the CUDA execution model for the index-logic kernels (no warp intrinsics):
// one std::thread per CUDA thread, CTAs run one after another, __syncthreads() = std::barrier.
// Test infrastructure only (tests/test_kernel_emulation.py); never part of the product.
#pragma once
#include <algorithm>
#include <barrier>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

struct dim3 { unsigned x = 1, y = 1, z = 1; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct alignas(16) float4 { float x, y, z, w; };
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static thread_local dim3 threadIdx, blockIdx;
static dim3 blockDim, gridDim;
#define __global__
#define __device__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
template <class T> static inline T __ldg(const T* p) { return *p; }
using std::max;
using std::min;
static std::barrier<>* g_bar = nullptr;
static inline void __syncthreads() { g_bar->arrive_and_wait(); }
alignas(16) static unsigned char g_smem[228 * 1024];
#define AL_DYN_SMEM(T, name) T* name = reinterpret_cast<T*>(g_smem)
static inline void al_cp_async16(void* d, const void* s) { std::memcpy(d, s, 16); }
static inline void al_cp_async_commit() {}
template <int N> static inline void al_cp_async_wait() {}

template <class F>
static void emul_launch(dim3 grid, dim3 block, F f) {
    gridDim = grid;
    blockDim = block;
    for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx) {
            std::barrier<> bar(block.x);
            g_bar = &bar;
            std::vector<std::thread> th;
            for (unsigned t = 0; t < block.x; ++t)
                th.emplace_back([=, &f] {
                    threadIdx = dim3(t);
                    blockIdx = dim3(bx, by);
                    f();
                });
            for (auto& x : th) x.join();
        }
}
