// sm_100a building blocks of the tensor-core kernels (al_gemm.cu, al_fattn.cu): mbarrier, TMA tensor copies
// (cp.async.bulk.tensor), tensor memory (tcgen05.alloc / ld), tcgen05.mma with shared-memory descriptors, and
// the proxy / thread-sync fences that tie them together.  Thin wrappers of single PTX instructions, nothing else.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace al {
namespace tc {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ int lane_id() { return (int)(threadIdx.x & 31); }

// ---- mbarrier (shared::cta, addressed by 32-bit shared addresses) -------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Remote arrive on the barrier at the same offset in CTA `cta` of the cluster.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 r;\n\t"
        "mapa.shared::cluster.u32 r, %0, %1;\n\t"
        "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [r];\n\t}" ::"r"(bar), "r"(cta)
        : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// A kernel whose pipeline protocol is broken would spin here for ever and take the GPU with it; the watchdog turns
// that into a trap (the launch fails with an error) after ~2 s.  It costs nothing on the path where waits succeed.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}
// Acquire at cluster scope (the barrier is also arrived on by the peer CTA / by multicast commits).
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}

// ---- TMA: tiled tensor copies through a CUtensorMap -----------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// Pull the box at these coordinates into L2 (no shared memory, no completion to wait for).
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* m, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];" ::"l"(reinterpret_cast<uint64_t>(m)), "r"(c0),
                 "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(src), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(src), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
// generic-proxy writes to shared memory -> visible to the async proxy (TMA stores, tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- tensor memory ---------------------------------------------------------------------------------------------
// Whole-warp instructions.  `cols` is a power of two in [32, 512]; the base address lands in *dst (shared memory).
__device__ __forceinline__ void tmem_alloc(uint32_t dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 32 lanes x 32 columns of 32-bit: thread i of the warp receives lane (base + i), register j = column (base + j).
// A warp can only touch the lane quadrant 32 * (warp_id % 4) .. + 31.
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}

// ---- tcgen05.mma (kind::f16: bf16 x bf16 -> fp32 in tensor memory), issued by ONE thread -----------------------
// Shared-memory matrix descriptor of a K-major operand tile stored as rows of 128 bytes (64 bf16) in the
// 128-byte swizzle pattern TMA writes (CU_TENSOR_MAP_SWIZZLE_128B): 8-row groups 1024 bytes apart.
//   bits [0,14)  start address >> 4        bits [16,30) leading byte offset >> 4 (unused for swizzled K-major)
//   bits [32,46) stride byte offset >> 4   bits [46,48) descriptor version (1 on sm_100)
//   bits [61,64) layout: 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_byte_addr) {
    return (uint64_t)((smem_byte_addr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) |
           ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}
// Instruction descriptor: D fp32 (bits [4,6) = 1), A and B bf16 (bits [7,10) = [10,13) = 1), both K-major
// (bits 15, 16 = 0), N >> 3 in bits [17,23), M >> 4 in bits [24,29).
__device__ __forceinline__ uint32_t umma_idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
// same with the A / B format selectable: 0 = IEEE half (f16), 1 = bfloat16 -- both kind::f16, same rate
__device__ __forceinline__ uint32_t umma_idesc_16bit(int m, int n, bool f16) {
    const uint32_t fmt = f16 ? 0u : 1u;
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// All tcgen05.mma issued so far by this thread -> one arrive on the mbarrier when they have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---- CTA pairs (cta_group::2): one tcgen05.mma spans two SMs (M = 256: 128 rows per CTA), each CTA holds its own A tile
// and HALF of the B tile, so a CTA moves 2/3 of the operand bytes per flop of the one-CTA form ---------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// The shared::cluster address of the same shared-memory offset in CTA 0 of the pair (bit 24 = CTA rank inside the pair)
__device__ __forceinline__ uint32_t in_leader(uint32_t addr) { return addr & 0xFEFFFFFFu; }
__device__ __forceinline__ void tmem_alloc2(uint32_t dst, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// TMA load into THIS CTA's shared memory whose bytes are counted on the barrier at the same offset in the pair's CTA 0
__device__ __forceinline__ void tma_load_3d_pair(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(in_leader(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void umma_bf16_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                  uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
// all tcgen05.mma of the pair issued so far -> one arrive on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((unsigned short)3)
                 : "memory");
}
// arrive on the barrier at this offset in CTA 0 of the pair (from either CTA)
__device__ __forceinline__ void mbar_arrive_leader(uint32_t bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(in_leader(bar)) : "memory");
}

// ---- small numerics shared by the epilogues ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// IEEE half pair; saturates to +-65504 instead of producing inf (half has 5 exponent bits)
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
template <bool F16>
__device__ __forceinline__ uint32_t pack16(float lo, float hi) { return F16 ? pack_f16(lo, hi) : pack_bf16(lo, hi); }
__device__ __forceinline__ float fast_ex2(float x) {
    float r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_rcp(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
// Exact-form GELU x * Phi(x) with erf(|z|) = 1 - 2^(t P(t)), t = min(|z|, 4): |erf error| < 3.1e-7 in fp32
// (degree-5 weighted minimax fit, tools/fit_erf.py), far inside the bf16 rounding of the stored result.
__device__ __forceinline__ float gelu_erf(float x) {
    const float ax = fabsf(x);
    const float t = fminf(ax * 0.70710678118654752440f, 4.0f);
    float p = 1.4203174214344472e-4f;
    p = fmaf(p, t, -3.6642320919781923e-3f);
    p = fmaf(p, t, 3.089611791074276e-2f);
    p = fmaf(p, t, -1.496993899345398e-1f);
    p = fmaf(p, t, -9.181655049324036e-1f);
    p = fmaf(p, t, -1.6279250383377075f);
    const float erf_abs = 1.0f - fast_ex2(p * t);
    return 0.5f * fmaf(ax, erf_abs, x);
}
// Two elements at a time on the packed fp32 pipe (FFMA2 / FMUL2), 13 issue slots per pair.  Same function and the same
// roundings as gelu_erf: u = sat(|z| / 4) replaces min(|z|, 4) (one saturating multiply instead of multiply + min), the
// polynomial runs in u with coefficients scaled by powers of 4 (exact), and relu(x) is 0.5 x + 0.5 |x| (exact):
// x Phi(x) = 0.5 x + 0.5 |x| - 0.5 |x| 2^(t P(t)).
__device__ __forceinline__ float2 gelu_erf2(float2 x) {
    float2 u;
    u.x = __saturatef(fabsf(x.x) * 0.17677669529663688110f);      // |x| / sqrt(2) / 4, clamped to 1
    u.y = __saturatef(fabsf(x.y) * 0.17677669529663688110f);
    constexpr float b5 = 1.4203174214344472e-4f * 4096.f, b4 = -3.6642320919781923e-3f * 1024.f,
                    b3 = 3.089611791074276e-2f * 256.f, b2 = -1.496993899345398e-1f * 64.f,
                    b1 = -9.181655049324036e-1f * 16.f, b0 = -1.6279250383377075f * 4.f;
    float2 p = __ffma2_rn(make_float2(b5, b5), u, make_float2(b4, b4));
    p = __ffma2_rn(p, u, make_float2(b3, b3));
    p = __ffma2_rn(p, u, make_float2(b2, b2));
    p = __ffma2_rn(p, u, make_float2(b1, b1));
    p = __ffma2_rn(p, u, make_float2(b0, b0));
    const float2 pt = __fmul2_rn(p, u);
    const float2 e = make_float2(fast_ex2(pt.x), fast_ex2(pt.y));
    const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
    const float2 nh = __fmul2_rn(ax, make_float2(-0.5f, -0.5f));
    const float2 relu = __ffma2_rn(x, make_float2(0.5f, 0.5f), make_float2(-nh.x, -nh.y));
    return __ffma2_rn(nh, e, relu);
}
// tanh(x) = 1 - 2 / (1 + e^(2x)), two MUFU ops, ~1e-7 absolute
__device__ __forceinline__ float tanh_fast(float x) {
    const float e = fast_ex2(x * 2.8853900817779268f);
    return 1.0f - 2.0f * fast_rcp(1.0f + e);
}

__device__ __forceinline__ float sigmoid_fast(float x) { return fast_rcp(1.0f + fast_ex2(x * -1.4426950408889634f)); }

}  // namespace tc
}  // namespace al
