// Minimal host emulation of the CUDA execution model (CTA barrier, warp barrier, warp shuffle; no tensor ops):
// one std::thread per CUDA thread, CTAs run one after another, __syncthreads() = std::barrier.
// Test infrastructure only (tests/test_kernel_emulation.py); never part of the product.
#pragma once
#include <algorithm>
#include <atomic>
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
struct alignas(8) float2 { float x, y; };
static inline float4 make_float4(float a, float b, float c, float d) { return float4{a, b, c, d}; }
static inline float2 make_float2(float a, float b) { return float2{a, b}; }
// packed fp32 pair intrinsics (FADD2 / FMUL2 / FFMA2 on sm_100a): element-wise on the host
static inline float2 __fadd2_rn(float2 a, float2 b) { return float2{a.x + b.x, a.y + b.y}; }
static inline float2 __fmul2_rn(float2 a, float2 b) { return float2{a.x * b.x, a.y * b.y}; }
static inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return float2{std::fma(a.x, b.x, c.x), std::fma(a.y, b.y, c.y)}; }
// bfloat16 (round-to-nearest-even conversions, like cuda_bf16.h)
struct __nv_bfloat16 { unsigned short v; };
struct __nv_bfloat162 { __nv_bfloat16 x, y; };
static inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
static inline float __bfloat162float(__nv_bfloat16 h) { return __uint_as_float((unsigned)h.v << 16); }
static inline __nv_bfloat16 __float2bfloat16_rn(float f) {
    unsigned u = __float_as_uint(f);
    if ((u & 0x7FFFFFFFu) > 0x7F800000u) return __nv_bfloat16{(unsigned short)0x7FFF};   // NaN
    u += 0x7FFFu + ((u >> 16) & 1u);
    return __nv_bfloat16{(unsigned short)(u >> 16)};
}
static inline unsigned short __bfloat16_as_ushort(__nv_bfloat16 h) { return h.v; }
static inline float2 __bfloat1622float2(__nv_bfloat162 h) { return float2{__bfloat162float(h.x), __bfloat162float(h.y)}; }
static inline __nv_bfloat162 __floats2bfloat162_rn(float a, float b) { return __nv_bfloat162{__float2bfloat16_rn(a), __float2bfloat16_rn(b)}; }
#define __expf(x) std::exp((float)(x))      // glibc declares __expf itself
struct alignas(16) uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
typedef int cudaError_t;
typedef void* cudaStream_t;
static thread_local dim3 threadIdx, blockIdx;
static dim3 blockDim, gridDim;
#define __global__
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __device__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
template <class T> static inline T __ldg(const T* p) { return *p; }
using std::max;
using std::min;
static std::barrier<>* g_bar = nullptr;
static inline void __syncthreads() { g_bar->arrive_and_wait(); }
// warp-level primitives: one barrier and one 32-word exchange buffer per warp of the running CTA
static std::vector<std::barrier<>*> g_warp_bar;
static unsigned g_shfl[64][32];
static inline void __syncwarp(unsigned = 0xffffffffu) { g_warp_bar[threadIdx.x >> 5]->arrive_and_wait(); }
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int lane_mask);
template <class T>
static inline T __shfl_sync(unsigned, T v, int src) {
    static_assert(sizeof(T) == 4, "4-byte shuffles only");
    const unsigned w = threadIdx.x >> 5, l = threadIdx.x & 31;
    std::memcpy(&g_shfl[w][l], &v, 4);
    g_warp_bar[w]->arrive_and_wait();
    T r;
    std::memcpy(&r, &g_shfl[w][src & 31], 4);
    g_warp_bar[w]->arrive_and_wait();
    return r;
}
alignas(16) static unsigned char g_smem[228 * 1024];
#define AL_DYN_SMEM(T, name) T* name = reinterpret_cast<T*>(g_smem)
static inline void al_cp_async16(void* d, const void* s) { std::memcpy(d, s, 16); }
static inline void al_cp_async_commit() {}
template <int N> static inline void al_cp_async_wait() {}

template <class T>
static inline T __shfl_xor_sync(unsigned m, T v, int lane_mask) { return __shfl_sync(m, v, (int)((threadIdx.x & 31) ^ lane_mask)); }

// mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 as a warp collective: every lane publishes its fragments, then
// computes its four outputs from the assembled 16x16 A and 16x8 B (fragment layouts of the PTX ISA; CuTe states the same
// in cute/atom/mma_traits_sm80.hpp: A (m = lane/4 + 8 v1, k = 2 (lane%4) + v0 + 8 v2), B (n = lane/4, k = 2 (lane%4) + v0 + 8 v1),
// C (m = lane/4 + 8 v1, n = 2 (lane%4) + v0)).
static unsigned g_mma_a[64][32][4], g_mma_b[64][32][2];
static inline void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    const unsigned w = threadIdx.x >> 5, l = threadIdx.x & 31;
    for (int i = 0; i < 4; ++i) g_mma_a[w][l][i] = a[i];
    for (int i = 0; i < 2; ++i) g_mma_b[w][l][i] = b[i];
    g_warp_bar[w]->arrive_and_wait();
    auto bf = [](unsigned u, int half) { return __uint_as_float(((u >> (16 * half)) & 0xFFFFu) << 16); };
    for (int i = 0; i < 4; ++i) {
        const int row = (int)(l >> 2) + 8 * (i >> 1), col = 2 * (int)(l & 3) + (i & 1);
        float acc = d[i];
        for (int k = 0; k < 16; ++k) {
            const int la = (row & 7) * 4 + (k & 7) / 2, ra = (row >= 8 ? 1 : 0) + (k >= 8 ? 2 : 0);
            const int lb = col * 4 + (k & 7) / 2, rb = k >= 8 ? 1 : 0;
            acc += bf(g_mma_a[w][la][ra], k & 1) * bf(g_mma_b[w][lb][rb], k & 1);
        }
        d[i] = acc;
    }
    g_warp_bar[w]->arrive_and_wait();
}

template <class F>
static void emul_launch(dim3 grid, dim3 block, F f) {
    gridDim = grid;
    blockDim = block;
    for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx) {
            std::barrier<> bar(block.x);
            g_bar = &bar;
            for (auto* b : g_warp_bar) delete b;
            g_warp_bar.clear();
            for (unsigned w = 0; w * 32 < block.x; ++w)
                g_warp_bar.push_back(new std::barrier<>(std::min(32u, block.x - w * 32)));
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
