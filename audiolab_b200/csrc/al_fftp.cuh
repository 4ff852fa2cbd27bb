// Channel-packed warp FFT building blocks for the stereo fast-path STFT / iSTFT kernels (sm_100a).
//
// Every value is a float2 = (left, right) sample of the SAME FFT element, so each FADD2 / FMUL2 /
// FFMA2 advances both channels' transforms and twiddles / window values are scalar broadcast
// operands.  One warp owns one 1024-point complex FFT per channel (element 32*r + lane in register r):
// two in-register radix-32 butterflies around one shared-memory transposition, done as two half
// passes (re, then im) through an 8.25 KB per-warp scratch.
//
// Also here: the mbarrier helpers of the producer / consumer stage ring.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "fft32p_gen.cuh"

namespace al {

#ifndef AL_CPU_EMUL
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
#endif

constexpr int kScrF2 = 33 * 32;   // float2 per warp scratch: [32][33] transposition tile

// (re, im) *= (w.x + i w.y), or its conjugate when CONJ
template <bool CONJ>
__device__ __forceinline__ void pcmul(float2& re, float2& im, float2 w) {
    const float s = CONJ ? -w.y : w.y;
    const float2 r = pfma(im, -s, pscale(re, w.x));
    const float2 i = pfma(im, w.x, pscale(re, s));
    re = r;
    im = i;
}

__device__ __forceinline__ float2 shfl2(float2 v, int src) {
    return make_float2(__shfl_sync(0xffffffffu, v.x, src), __shfl_sync(0xffffffffu, v.y, src));
}

// in : re/im[r] = z[32*r + lane];  out: re/im[r] = Z[32*r + lane],
// Z[k] = sum_n z[n] exp(-+2 pi i n k / 1024) (unnormalised).  tw[k1*32 + n2] = exp(-2 pi i k1 n2 / 1024).
template <bool INV>
__device__ __forceinline__ void warp_fft1024p(float2 (&re)[32], float2 (&im)[32], float2* scr,
                                              const float2* tw, int lane) {
    if (INV) fft32p_inv(re, im); else fft32p_fwd(re, im);
#pragma unroll
    for (int k1 = 1; k1 < 32; ++k1) pcmul<INV>(re[k1], im[k1], tw[k1 * 32 + lane]);
    float2* wr = scr + lane * 33;
    const float2* rd = scr + lane;
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
    if (INV) fft32p_inv(re, im); else fft32p_fwd(re, im);
}

// Same transform with a single transposition pass through a 16-byte-element scratch
// (float4 = (re.L, re.R, im.L, im.R), [32][33] float4 = 16.5 KB per warp): half the LDS/STS
// instructions and two warp barriers instead of four.
constexpr int kScrF4 = 33 * 32;
template <bool INV>
__device__ __forceinline__ void warp_fft1024p_wide(float2 (&re)[32], float2 (&im)[32], float4* scr,
                                                   const float2* tw, int lane) {
    if (INV) fft32p_inv(re, im); else fft32p_fwd(re, im);
    float4* wr = scr + lane * 33;
    const float4* rd = scr + lane;
    wr[0] = make_float4(re[0].x, re[0].y, im[0].x, im[0].y);
#pragma unroll
    for (int k1 = 1; k1 < 32; ++k1) {
        pcmul<INV>(re[k1], im[k1], tw[k1 * 32 + lane]);
        wr[k1] = make_float4(re[k1].x, re[k1].y, im[k1].x, im[k1].y);
    }
    __syncwarp();
#pragma unroll
    for (int n2 = 0; n2 < 32; ++n2) {
        const float4 v = rd[n2 * 33];
        re[n2] = make_float2(v.x, v.y);
        im[n2] = make_float2(v.z, v.w);
    }
    __syncwarp();
    if (INV) fft32p_inv(re, im); else fft32p_fwd(re, im);
}

// The same transform with ONE copy of the 32-point butterfly in the instruction stream (a two-trip loop around it):
// half the code of the form above.  For kernels whose warps drift against each other, where the code every warp streams
// through has to fit the 32 KB L1.5 instruction cache (profiles/r02zf_*: 34 % of the stall samples of a kernel with a
// 37 KB consumer loop were instruction fetches).
template <bool INV>
__device__ __forceinline__ void warp_fft1024p_wide_rolled(float2 (&re)[32], float2 (&im)[32], float4* scr,
                                                          const float2* tw, int lane, uint64_t* release = nullptr);
__device__ __forceinline__ void mbar_arrive(uint64_t* bar);
__device__ __forceinline__ void fence_async_smem();
// release (optional): an mbarrier that takes one arrival as soon as the scratch has been read back (the scratch is a ring
// slot that goes back to its producer half way through the transform)
template <bool INV>
__device__ __forceinline__ void warp_fft1024p_wide_rolled(float2 (&re)[32], float2 (&im)[32], float4* scr,
                                                          const float2* tw, int lane, uint64_t* release) {
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
        if (INV) fft32p_inv(re, im); else fft32p_fwd(re, im);
        if (pass == 0) {
            float4* wr = scr + lane * 33;
            const float4* rd = scr + lane;
            wr[0] = make_float4(re[0].x, re[0].y, im[0].x, im[0].y);
#pragma unroll
            for (int k1 = 1; k1 < 32; ++k1) {
                pcmul<INV>(re[k1], im[k1], tw[k1 * 32 + lane]);
                wr[k1] = make_float4(re[k1].x, re[k1].y, im[k1].x, im[k1].y);
            }
            __syncwarp();
#pragma unroll
            for (int n2 = 0; n2 < 32; ++n2) {
                const float4 v = rd[n2 * 33];
                re[n2] = make_float2(v.x, v.y);
                im[n2] = make_float2(v.z, v.w);
            }
            if (release) fence_async_smem();            // generic-proxy accesses of the scratch before the next bulk copy into it
            __syncwarp();
            if (release && lane == 0) mbar_arrive(release);
        }
    }
}

#ifndef AL_CPU_EMUL   // (tools/cpu_emul emulates the barrier / shuffle kernels only)
// ---- bulk asynchronous shared -> global store (TMA engine, no tensor map) ------------------------
__device__ __forceinline__ void bulk_store_s2g(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)),
                 "r"(bytes)
                 : "memory");
}
// bulk asynchronous global -> shared copy whose bytes are counted on an mbarrier (16-byte aligned, size a multiple of 16)
__device__ __forceinline__ void bulk_load_g2s(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sdst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- mbarrier helpers (shared::cta) ----------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// one arrival that also announces `bytes` of bulk copies (issued BEFORE this call: the emulation copies at issue time)
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!ok);
}
// wait whose suspension is bounded by `ns` nanoseconds per try (a waiter on a serial hand-off chain must notice the phase
// flip at once; 0 = the plain form)
__device__ __forceinline__ void mbar_wait_hint(uint64_t* bar, uint32_t parity, uint32_t ns) {
    if (ns == 0) { mbar_wait(bar, parity); return; }
    const uint32_t addr = smem_u32(bar);
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity), "r"(ns)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ void mbar_arrive_relaxed(uint64_t* bar) {
    asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
#else   // host emulation (tools/cpu_emul): the same protocol on std::atomic_ref, bulk copies done at issue time
__device__ __forceinline__ void bulk_store_s2g(void* gdst, const void* ssrc, uint32_t bytes) { std::memcpy(gdst, ssrc, bytes); }
__device__ __forceinline__ void bulk_load_g2s(void* sdst, const void* gsrc, uint32_t bytes, uint64_t*) { std::memcpy(sdst, gsrc, bytes); }
__device__ __forceinline__ void bulk_commit() {}
__device__ __forceinline__ void bulk_wait_read_all() {}
__device__ __forceinline__ void bulk_wait_all() {}
__device__ __forceinline__ void fence_async_smem() {}
// word = expected arrivals << 32 | pending arrivals << 1 | phase parity
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    std::atomic_ref<uint64_t>(*bar).store(((uint64_t)count << 32) | ((uint64_t)count << 1));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    std::atomic_ref<uint64_t> a(*bar);
    uint64_t v = a.load();
    for (;;) {
        const uint64_t count = v >> 32, pending = (v & 0xFFFFFFFFull) >> 1, phase = v & 1;
        const uint64_t nv = pending > 1 ? ((count << 32) | ((pending - 1) << 1) | phase)
                                        : ((count << 32) | (count << 1) | (phase ^ 1));
        if (a.compare_exchange_weak(v, nv)) return;
    }
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t) { mbar_arrive(bar); }
__device__ __forceinline__ void mbar_arrive_relaxed(uint64_t* bar) { mbar_arrive(bar); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    std::atomic_ref<uint64_t> a(*bar);
    while ((a.load() & 1) == parity) std::this_thread::yield();
}
__device__ __forceinline__ void mbar_wait_hint(uint64_t* bar, uint32_t parity, uint32_t) { mbar_wait(bar, parity); }
#endif  // AL_CPU_EMUL

}  // namespace al
