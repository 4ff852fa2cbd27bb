// Time-axis attention of the RoFormer mask network on the 5th-generation tensor cores (sm_100a):
//   o[b, t, f, h, :] = sigmoid(gate[b, t, f, h]) * sum_t' softmax_t'(q[b, t, f, h, :] . k[b, t', f, h, :] * scale) v[b, t', f, h, :]
// upstream Attention.forward inside the time transformer (driven by MDXCSeparator.demix behind
// modules/separator/stem_separator.py:281).  Replaces F.scaled_dot_product_attention (cuDNN) + the separate gate pass.
//
// q, k, v, o: [B * T * F, H * 64] 16-bit, token (b, t, f) in row (b T + t) F + f -- the token-major layout of the residual
// stream.  A sequence (b, f, h) is the T rows with row stride F: the TMA tensor map views the buffers as
// [B][T][F * H][64], so a 128 x 64 tile of one sequence is ONE box and no transposition copy exists on either side.
//
// One persistent CTA per SM, three warpgroups, warp-specialised; a work unit is (sequence, pair of 128-row query tiles):
//   warp 8       producer : TMA loads of the two Q tiles and of the K / V tiles (128 x 64, 128-byte swizzle) into a ring
//   warps 9, 10  MMA issuers of softmax group A / B (one thread each, blocking waits in a fixed order):
//                           S_L(j+1) = Q_L K_{j+1}^T (128 x 128 x 64) as soon as group L has taken S_L(j) out of tensor
//                           memory, O_L (+)= P_L(j) V_j (128 x 64 x 128, P read from TENSOR MEMORY) when P_L(j) is ready
//                           (warps 8..11 hand their registers to the softmax groups: setmaxnreg 56 / 224)
//   warps 0..3   softmax group A (query tile 2p), warps 4..7 group B (tile 2p + 1): thread = query row.  The row of S
//                           comes out of tensor memory once (128 registers), p = 2^(s c - m) with a LAZY reference m: it
//                           only moves when the row maximum grows by more than 2^8 (then O is rescaled in tensor memory),
//                           P goes back to tensor memory as packed 16-bit pairs (tcgen05.st), O accumulates in tensor memory.
// While group A runs its exponentials the tensor core works for group B and vice versa; the kernel is paced by the
// MUFU.EX2 rate (16 / clk / SM), not by the tensor pipe (d = 64).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <mutex>

#include "al_kernels.h"
#include "al_tc.cuh"

namespace al {
namespace fa {

using namespace al::tc;

constexpr int kD = 64;            // head dimension
constexpr int kBM = 128;          // query rows per tile = tensor memory lanes
constexpr int kBN = 128;          // keys per tile
constexpr int kStages = 5;        // K / V ring
constexpr int kThreads = 384;    // softmax groups A, B + the producer / MMA warpgroup
constexpr int kTile = kBM * kD * 2;           // 16 KB: one 128 x 64 16-bit tile
constexpr int kTurn = 3;                      // named barriers 3, 4: whose turn it is to run exponentials
constexpr float kLazy = 8.0f;                 // log2 of the growth of the row maximum that triggers a rescale

struct Maps {
    CUtensorMap q, k, v, o;
};

struct Args {
    int T, F, H, B;
    int n_qt, n_kv;                // query tiles, key tiles per sequence
    long long n_seq;               // B * F * H sequences
    long long n_units;             // ceil(n_seq / 2) * n_qt: the 2 n_qt query tiles of two sequences, taken two at a time
    float scale_log2;              // scale * log2(e)
    const void* gates;             // [rows, gate_ld] 16-bit or NULL
    long long gate_ld;
};

struct Smem {
    static constexpr int kQ = 0;                                   // 2 x 16 KB
    static constexpr int kKV = 2 * kTile;                          // kStages x (K 16 KB + V 16 KB)
    static constexpr int kO = kKV + kStages * 2 * kTile;           // 2 x 16 KB: O tiles on their way out (TMA store)
    static constexpr int kBar = kO + 2 * kTile;
    // barriers: q_full[2] q_empty[2] s_full[2] s_empty[2] p_full[2] o_full[2] k_full[ST] v_full[ST] kv_empty[ST]
    static constexpr int kNumBars = 12 + 3 * kStages;
    static constexpr int kTotal = kBar + kNumBars * 8 + 16;
    static constexpr int kDynamic = kTotal + 1024;
};
static_assert(Smem::kDynamic <= 232448, "shared memory budget");

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(m)),
                 "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
// non-blocking probe of an mbarrier phase (the MMA warp polls several barriers)
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void tmem_ld_32x8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_st_32x8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
                 "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]: the A operand (128 x 16, 16-bit) sits in tensor memory, lane = row, one 32-bit column
// = two consecutive K elements (low half first)
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void named_bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// instruction descriptor: D fp32, A / B 16-bit (f16 = 0, bf16 = 1), A K-major, B K-major or MN-major (bit 16)
__device__ __forceinline__ uint32_t idesc(int m, int n, bool f16, bool b_mn_major) {
    const uint32_t fmt = f16 ? 0u : 1u;
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((b_mn_major ? 1u : 0u) << 16) | ((uint32_t)(n >> 3) << 17) |
           ((uint32_t)(m >> 4) << 24);
}

// 2^x for x <= 8 on the FMA pipe (every 4th pair of a row: the MUFU is the scarce unit): x = n + r with n = round(x) taken
// from the low mantissa bits of x + 1.5 * 2^23, 2^r by a degree-4 polynomial on [-0.5, 0.5] (relative error 3.6e-6, far
// below the 16-bit rounding of P), n added to the exponent field.
__device__ __forceinline__ float2 exp2_fma2(float2 x) {
    x.x = fmaxf(x.x, -126.f);
    x.y = fmaxf(x.y, -126.f);
    const float2 magic = make_float2(12582912.f, 12582912.f);
    const float2 t = __fadd2_rn(x, magic);
    const float2 n = __fadd2_rn(t, make_float2(-12582912.f, -12582912.f));
    const float2 r = __ffma2_rn(n, make_float2(-1.f, -1.f), x);
    float2 p = __ffma2_rn(make_float2(0.00966636836528778f, 0.00966636836528778f), r,
                          make_float2(0.055921975523233414f, 0.055921975523233414f));
    p = __ffma2_rn(p, r, make_float2(0.2402234971523285f, 0.2402234971523285f));
    p = __ffma2_rn(p, r, make_float2(0.6931210160255432f, 0.6931210160255432f));
    p = __ffma2_rn(p, r, make_float2(1.0f, 1.0f));
    return make_float2(__int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23)),
                       __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23)));
}

template <bool F16>
__global__ void __launch_bounds__(kThreads, 1)
time_attn_kernel(const __grid_constant__ Maps tm, const Args g) {
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t raw = smem_addr(smem_dyn);
    const uint32_t base = (raw + 1023u) & ~1023u;
    unsigned char* base_ptr = smem_dyn + (base - raw);
    const uint32_t bars = base + Smem::kBar;
    auto q_full = [&](int l) { return bars + 8u * l; };
    auto q_empty = [&](int l) { return bars + 8u * (2 + l); };
    auto s_full = [&](int l) { return bars + 8u * (4 + l); };
    auto s_empty = [&](int l) { return bars + 8u * (6 + l); };
    auto p_full = [&](int l) { return bars + 8u * (8 + l); };
    auto o_full = [&](int l) { return bars + 8u * (10 + l); };
    auto k_full = [&](int s) { return bars + 8u * (12 + s); };
    auto v_full = [&](int s) { return bars + 8u * (12 + kStages + s); };
    auto kv_empty = [&](int s) { return bars + 8u * (12 + 2 * kStages + s); };
    const uint32_t tmem_slot = bars + 8u * Smem::kNumBars;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + Smem::kBar + 8 * Smem::kNumBars);

    const int warp = (int)(threadIdx.x >> 5);
    const int lane = lane_id();

    if (warp == 8) {
        if (lane == 0) {
            tma_prefetch_desc(&tm.q);
            tma_prefetch_desc(&tm.k);
            tma_prefetch_desc(&tm.v);
            tma_prefetch_desc(&tm.o);
        }
        tmem_alloc(tmem_slot, 512);
        tmem_relinquish();
    } else if (warp == 9 && lane == 0) {
        for (int l = 0; l < 2; ++l) {
            mbar_init(q_full(l), 1);
            mbar_init(q_empty(l), 1);
            mbar_init(s_full(l), 1);
            mbar_init(s_empty(l), 4);
            mbar_init(p_full(l), 4);
            mbar_init(o_full(l), 1);
        }
        for (int s = 0; s < kStages; ++s) {
            mbar_init(k_full(s), 1);
            mbar_init(v_full(s), 1);
            mbar_init(kv_empty(s), 2);            // one arrival per MMA issuer
        }
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;
    // tensor memory columns: S_A 0..127, S_B 128..255, O_A 256..319, O_B 320..383, P_A 384..447, P_B 448..511 (16-bit pairs)
    auto s_col = [&](int l) { return (uint32_t)(l * kBN); };
    auto o_col = [&](int l) { return (uint32_t)(2 * kBN + l * kD); };
    auto p_col = [&](int l) { return (uint32_t)(2 * kBN + 2 * kD + l * (kBN / 2)); };

    // Work units.  Two consecutive sequences (b, f, h) form a group; the 2 n_qt query tiles of the group are taken two at a
    // time, tile 2 r + l of the list going to softmax group l of unit r.  With an odd number of tiles per sequence (801
    // frames = 6 full tiles + 33 rows) one unit of the group is MIXED: its two query tiles belong to different sequences and
    // the K / V ring carries the tiles of both, alternating -- instead of one unit per sequence that keeps a whole softmax
    // group idle.  The unit index inside the group is rotated by the CTA's iteration count so that every CTA sees the same mix.
    const uint32_t U = (uint32_t)g.n_qt;
    const bool rotate = gridDim.x % U == 0;
    struct Lane { uint32_t hs; int qt; bool active; };
    auto decode = [&](uint32_t u, int l) {
        const uint32_t grp = u / U;
        uint32_t r = u - grp * U;
        if (rotate) r = (r + u / gridDim.x) % U;
        const uint32_t k = 2 * r + (uint32_t)l;
        const uint32_t second = k >= U ? 1u : 0u;
        Lane x;
        x.hs = 2 * grp + second;
        x.qt = (int)(k - second * U);
        x.active = (long long)x.hs < g.n_seq;
        return x;
    };
    const int n_kv = g.n_kv;
    const int last_cols = g.T - (n_kv - 1) * kBN;                 // valid keys of the last key tile
    const int last_n16 = (last_cols + 15) & ~15;
    const int FH = g.F * g.H;

    if (warp >= 8) {
        // the producer / MMA warpgroup hands its registers to the softmax groups (168 * 384 = 224 * 256 + 56 * 128)
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    }
    if (warp == 8) {
        // ================================= TMA producer =================================
        if (lane == 0) {
            uint32_t kvc = 0;
            uint32_t uc[2] = {0, 0};
            for (uint32_t u = blockIdx.x; u < (uint32_t)g.n_units; u += gridDim.x) {
                Lane ln[2] = {decode(u, 0), decode(u, 1)};
                int bb[2], jj[2];
                for (int l = 0; l < 2; ++l) {
                    bb[l] = (int)(ln[l].hs / (uint32_t)FH);
                    jj[l] = (int)(ln[l].hs - (uint32_t)bb[l] * (uint32_t)FH);
                    if (!ln[l].active) continue;
                    mbar_wait(q_empty(l), (uc[l] & 1u) ^ 1u);
                    mbar_arrive_expect_tx(q_full(l), (uint32_t)kTile);
                    tma_load_4d(base + Smem::kQ + (uint32_t)l * kTile, &tm.q, q_full(l), 0, jj[l], ln[l].qt * kBM, bb[l]);
                    ++uc[l];
                }
                const bool mixed = ln[0].active && ln[1].active && ln[0].hs != ln[1].hs;
                const int first = ln[0].active ? 0 : 1;
                for (int jt = 0; jt < n_kv; ++jt) {
                    for (int l = first; l < (mixed ? 2 : first + 1); ++l, ++kvc) {       // mixed: K / V of both sequences, alternating
                        const int s = (int)(kvc % kStages);
                        const uint32_t ph = (kvc / kStages) & 1u;
                        mbar_wait(kv_empty(s), ph ^ 1u);
                        const uint32_t kb = base + Smem::kKV + (uint32_t)s * 2 * kTile;
                        mbar_arrive_expect_tx(k_full(s), (uint32_t)kTile);
                        tma_load_4d(kb, &tm.k, k_full(s), 0, jj[l], jt * kBN, bb[l]);
                        mbar_arrive_expect_tx(v_full(s), (uint32_t)kTile);
                        tma_load_4d(kb + kTile, &tm.v, v_full(s), 0, jj[l], jt * kBN, bb[l]);
                    }
                }
            }
        }
    } else if (warp == 9 || warp == 10) {
        // ================================= MMA issuer of group l =================================
        if (lane == 0) {
            const int l = warp - 9;
            uint32_t kvc = 0;                  // ring position of key tile 0 of the current unit
            uint32_t n_q = 0, n_se = 0, n_pf = 0, n_s = 0;     // completions consumed: q_full, s_empty, p_full; S products issued
            const uint64_t q_desc = umma_desc_sw128(base + Smem::kQ + (uint32_t)l * kTile);
            const uint32_t id_pv = idesc(kBM, kD, F16, true);
            const uint32_t s_tmem = tmem_base + s_col(l), o_tmem = tmem_base + o_col(l), p_tmem = tmem_base + p_col(l);
            auto issue_s = [&](uint32_t kv, int jt) {
                const int st = (int)(kv % kStages);
                mbar_wait(k_full(st), (kv / kStages) & 1u);
                if (n_s != 0) {                                   // S_l is free once the group has read the previous product
                    mbar_wait(s_empty(l), n_se & 1u);
                    ++n_se;
                }
                tc_fence_after();
                const uint32_t id = idesc(kBM, jt == n_kv - 1 ? last_n16 : kBN, F16, false);
                const uint64_t b_desc = umma_desc_sw128(base + Smem::kKV + (uint32_t)st * 2 * kTile);
#pragma unroll
                for (int k = 0; k < kD / 16; ++k)
                    umma_bf16_ss(s_tmem, q_desc + (uint64_t)(2 * k), b_desc + (uint64_t)(2 * k), id, (uint32_t)(k != 0));
                umma_commit(s_full(l));
                ++n_s;
            };
            // a ring slot this issuer has no use for (the other group's sequence, or no query tile at all): hand it back
            auto pass = [&](uint32_t kv) {
                const int st = (int)(kv % kStages);
                mbar_wait(k_full(st), (kv / kStages) & 1u);
                mbar_wait(v_full(st), (kv / kStages) & 1u);
                mbar_arrive(kv_empty(st));
            };
            for (uint32_t u = blockIdx.x; u < (uint32_t)g.n_units; u += gridDim.x) {
                const Lane me = decode(u, l), other = decode(u, l ^ 1);
                const bool mixed = me.active && other.active && me.hs != other.hs;
                const uint32_t n_pos = (uint32_t)(mixed ? 2 * n_kv : n_kv);          // ring slots of the unit
                if (!me.active) {
                    for (uint32_t q = 0; q < n_pos; ++q) pass(kvc + q);
                    kvc += n_pos;
                    continue;
                }
                // own slots: jt (shared K / V) or 2 jt + l (mixed unit); the slots in between belong to the other sequence
                auto own = [&](int jt) { return kvc + (uint32_t)(mixed ? 2 * jt + l : jt); };
                uint32_t next_pass = kvc + (uint32_t)(1 - l);
                auto pass_below = [&](uint32_t kv) {               // slots below a slot that is already full: no waiting
                    if (!mixed) return;
                    for (; next_pass < kv; next_pass += 2) pass(next_pass);
                };
                mbar_wait(q_full(l), n_q & 1u);
                ++n_q;
                issue_s(own(0), 0);
                pass_below(own(0));
                for (int jt = 0; jt < n_kv; ++jt) {
                    if (jt + 1 < n_kv) {
                        issue_s(own(jt + 1), jt + 1);
                        pass_below(own(jt + 1));
                    } else {
                        umma_commit(q_empty(l));                  // Q tile free once the last S has read it
                    }
                    const uint32_t kv = own(jt);
                    const int st = (int)(kv % kStages);
                    mbar_wait(v_full(st), (kv / kStages) & 1u);
                    mbar_wait(p_full(l), n_pf & 1u);              // P_l(jt) is in tensor memory, O_l is ours
                    ++n_pf;
                    tc_fence_after();
                    const int nk16 = (jt == n_kv - 1 ? last_n16 : kBN) / 16;
                    const uint64_t v_desc = umma_desc_sw128(base + Smem::kKV + (uint32_t)st * 2 * kTile + kTile);
#pragma unroll
                    for (int kk = 0; kk < kBN / 16; ++kk) {
                        // A = P: 16 keys = 8 columns of packed pairs; B = V: MN-major, 16 keys = 2 KB of the swizzled tile
                        if (kk < nk16)
                            umma_f16_ts(o_tmem, p_tmem + (uint32_t)(kk * 8), v_desc + (uint64_t)(kk * 128), id_pv,
                                        (uint32_t)((jt | kk) != 0));
                    }
                    umma_commit(o_full(l));
                    umma_commit(kv_empty(st));
                }
                pass_below(kvc + n_pos);
                kvc += n_pos;
            }
        }
    } else if (warp < 8) {
        // ================================= softmax groups =================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        const int l = warp >> 2;                          // group A / B
        const int qd = warp & 3;                          // tensor memory lane quadrant of this warp
        const int row = qd * 32 + lane;                   // query row inside the tile
        const int gtid = (int)threadIdx.x - l * 128;      // 0..127 inside the group
        const uint32_t lane_taddr = tmem_base + ((uint32_t)(qd * 32) << 16);
        const uint32_t obuf = base + Smem::kO + (uint32_t)l * kTile;
        const uint32_t orow = obuf + (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
        uint32_t n_s = 0, n_o = 0;                        // completions of s_full / o_full consumed so far
        if (l == 1) named_bar_arrive(kTurn, 256);         // group A takes the first turn
        for (uint32_t u = blockIdx.x; u < (uint32_t)g.n_units; u += gridDim.x) {
            const Lane me = decode(u, l);
            if (!me.active) {
                // no query tile for this group in the unit: keep the turn-taking of the exponential phases going
                for (int jt = 0; jt < n_kv; ++jt) {
                    named_bar_sync(kTurn + l, 256);
                    named_bar_arrive(kTurn + (l ^ 1), 256);
                }
                continue;
            }
            const int qt = me.qt;
            const int b = (int)(me.hs / (uint32_t)FH), j = (int)(me.hs - (uint32_t)b * (uint32_t)FH);
            const int t = qt * kBM + row;
            const bool warp_valid = qt * kBM + qd * 32 < g.T;     // any valid query row in this warp
            uint16_t gate_raw = 0;                                // read now, used in the unit's epilogue
            if (g.gates != nullptr && t < g.T) {
                const int f = j / g.H, h = j - f * g.H;
                const long long grow = ((long long)b * g.T + t) * g.F + f;
                gate_raw = __ldg(reinterpret_cast<const uint16_t*>(g.gates) + grow * g.gate_ld + h);
            }
            float m_ref = 0.f, lsum = 0.f;
            for (int jt = 0; jt < n_kv; ++jt) {
                const int ncols = jt == n_kv - 1 ? last_cols : kBN;
                const int n16 = (ncols + 15) & ~15;
                uint32_t s[4][32];
                mbar_wait(s_full(l), n_s & 1u);
                ++n_s;
                tc_fence_after();
                if (warp_valid) {
                    const uint32_t ta = lane_taddr + s_col(l);
                    tmem_ld_32x32(ta, s[0]);
                    if (n16 > 32) tmem_ld_32x32(ta + 32, s[1]);
                    if (n16 > 64) tmem_ld_32x32(ta + 64, s[2]);
                    if (n16 > 96) tmem_ld_32x32(ta + 96, s[3]);
                    tmem_wait_ld();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(s_empty(l));          // S_l may be overwritten by the next product
                bool o_waited = false;
                if (warp_valid) {
                    // ---- row maximum over the valid keys (columns >= ncols of the last tile are zero-filled keys)
                    if (ncols < kBN) {
#pragma unroll
                        for (int c = 0; c < 4; ++c)
#pragma unroll
                            for (int i = 0; i < 32; ++i)
                                if (c * 32 + i >= ncols) s[c][i] = 0xff800000u;          // -inf
                    }
                    // eight independent chains (3-input maxima), chunks beyond the valid keys never loaded: skipped
                    float m8[8];
#pragma unroll
                    for (int k = 0; k < 8; ++k) m8[k] = fmaxf(__uint_as_float(s[0][k]), __uint_as_float(s[0][8 + k]));
#pragma unroll
                    for (int k = 0; k < 8; ++k) m8[k] = fmaxf(m8[k], fmaxf(__uint_as_float(s[0][16 + k]), __uint_as_float(s[0][24 + k])));
#pragma unroll
                    for (int c = 1; c < 4; ++c) {
                        if (c * 32 < n16) {
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                m8[k] = fmaxf(m8[k], fmaxf(__uint_as_float(s[c][k]), __uint_as_float(s[c][8 + k])));
                                m8[k] = fmaxf(m8[k], fmaxf(__uint_as_float(s[c][16 + k]), __uint_as_float(s[c][24 + k])));
                            }
                        }
                    }
                    float mx = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
                    mx *= g.scale_log2;
                    if (jt == 0) {
                        m_ref = mx;
                    } else {
                        const bool need = mx > m_ref + kLazy;
                        if (__any_sync(0xffffffffu, need)) {
                            // the reference moves: rescale the running sum and the accumulator in tensor memory
                            const float alpha = need ? fast_ex2(m_ref - mx) : 1.f;
                            if (need) m_ref = mx;
                            lsum *= alpha;
                            mbar_wait(o_full(l), n_o & 1u);      // P V (jt - 1) has finished writing O_l
                            ++n_o;
                            o_waited = true;
                            tc_fence_after();
                            for (int c = 0; c < kD / 8; ++c) {            // rare path: 8 columns at a time, few registers
                                uint32_t o[8];
                                tmem_ld_32x8(lane_taddr + o_col(l) + (uint32_t)(c * 8), o);
                                tmem_wait_ld();
#pragma unroll
                                for (int i = 0; i < 8; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                                tmem_st_32x8(lane_taddr + o_col(l) + (uint32_t)(c * 8), o);
                            }
                            tmem_wait_st();
                        }
                    }
                }
                // ---- p = 2^(s c - m): the MUFU phase.  The two groups take turns (named barriers kTurn, kTurn + 1), so that
                // one group's exponentials run at the full MUFU rate while the other group loads / reduces / stores
                named_bar_sync(kTurn + l, 256);
                if (warp_valid) {
                    const float2 sc2 = make_float2(g.scale_log2, g.scale_log2);
                    const float2 nm2 = make_float2(-m_ref, -m_ref);
                    float2 acc2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        if (c * 32 < n16) {
                            float2 x[16];
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                x[i] = __ffma2_rn(make_float2(__uint_as_float(s[c][2 * i]), __uint_as_float(s[c][2 * i + 1])), sc2, nm2);
#pragma unroll
                            for (int i = 0; i < 16; ++i)
                                x[i] = (i & 3) == 3 ? exp2_fma2(x[i]) : make_float2(fast_ex2(x[i].x), fast_ex2(x[i].y));
#pragma unroll
                            for (int i = 0; i < 16; ++i) {
                                acc2[i & 1] = __fadd2_rn(acc2[i & 1], x[i]);
                                s[c][i] = pack16<F16>(x[i].x, x[i].y);
                            }
                        }
                    }
                    lsum += (acc2[0].x + acc2[1].x) + (acc2[0].y + acc2[1].y);
                }
                named_bar_arrive(kTurn + (l ^ 1), 256);
                if (jt > 0 && !o_waited) {
                    mbar_wait(o_full(l), n_o & 1u);              // P V (jt - 1) has read P_l: its columns are free
                    ++n_o;
                    tc_fence_after();
                }
                if (warp_valid) {
                    // P_l: key pair i of chunk c -> column 16 c + i (what the A operand of the second product expects)
                    const uint32_t tp = lane_taddr + p_col(l);
                    tmem_st_32x16(tp, s[0]);
                    if (n16 > 32) tmem_st_32x16(tp + 16, s[1]);
                    if (n16 > 64) tmem_st_32x16(tp + 32, s[2]);
                    if (n16 > 96) tmem_st_32x16(tp + 48, s[3]);
                    tmem_wait_st();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(p_full(l));
            }
            // ---- epilogue of the unit: O / l * sigmoid(gate) -> 16-bit tile -> TMA store
            if (gtid == 0) bulk_wait_read<0>();                  // the previous unit's store has read the staging tile
            named_bar_sync(1 + l, 128);
            mbar_wait(o_full(l), n_o & 1u);
            ++n_o;
            tc_fence_after();
            if (warp_valid) {
                float gate = 1.f;
                if (g.gates != nullptr) {
                    gate = F16 ? __half2float(__ushort_as_half(gate_raw)) : __bfloat162float(__ushort_as_bfloat16(gate_raw));
                    gate = sigmoid_fast(gate);
                }
                const float inv = gate / lsum;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t o[32];
                    tmem_ld_32x32(lane_taddr + o_col(l) + (uint32_t)(c * 32), o);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int c16 = c * 4 + i;               // 16-byte chunk of the 128-byte row, 128-byte swizzle
                        st_shared_v4(orow + (uint32_t)((c16 ^ (row & 7)) << 4),
                                     pack16<F16>(__uint_as_float(o[8 * i]) * inv, __uint_as_float(o[8 * i + 1]) * inv),
                                     pack16<F16>(__uint_as_float(o[8 * i + 2]) * inv, __uint_as_float(o[8 * i + 3]) * inv),
                                     pack16<F16>(__uint_as_float(o[8 * i + 4]) * inv, __uint_as_float(o[8 * i + 5]) * inv),
                                     pack16<F16>(__uint_as_float(o[8 * i + 6]) * inv, __uint_as_float(o[8 * i + 7]) * inv));
                    }
                }
                fence_proxy_async();
            }
            tc_fence_before();
            named_bar_sync(1 + l, 128);
            if (gtid == 0) {
                tma_store_4d(&tm.o, obuf, 0, j, qt * kBM, b);
                bulk_commit();
            }
        }
        if (gtid == 0) bulk_wait<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    });
    return fn;
}

// [B][T][F * H][64] view of a token-major [B * T * F, H * 64] buffer; box = one 128 x 64 tile of one (b, f, h) sequence
static bool make_map(CUtensorMap* m, const void* ptr, bool fp16, long long B, long long T, long long FH) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[4] = {(cuuint64_t)kD, (cuuint64_t)FH, (cuuint64_t)T, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)(kD * 2), (cuuint64_t)(FH * kD * 2), (cuuint64_t)(T * FH * kD * 2)};
    cuuint32_t box[4] = {(cuuint32_t)kD, 1u, (cuuint32_t)kBM, 1u};
    cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    return fn(m, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims,
              strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

}  // namespace fa

const char* launch_time_attention(const void* q, const void* k, const void* v, void* o, const void* gates, long long gate_ld,
                                  long long n_batch, int seq_len, int inner, int heads, int dim_head, float scale, int fp16,
                                  cudaStream_t stream, cudaError_t* cuda_err) {
    using namespace fa;
    *cuda_err = cudaSuccess;
    if (dim_head != kD) return "dim_head must be 64";
    if (n_batch <= 0 || seq_len <= 0 || inner <= 0 || heads <= 0) return "bad sizes";
    if ((reinterpret_cast<uintptr_t>(q) | reinterpret_cast<uintptr_t>(k) | reinterpret_cast<uintptr_t>(v) |
         reinterpret_cast<uintptr_t>(o)) & 15)
        return "q, k, v, o must be 16-byte aligned";
    const long long FH = (long long)inner * heads;
    if (FH > 0x7fffffffll || n_batch > 0x7fffffffll || n_batch * FH * ((seq_len + kBM - 1) / kBM) > 0x7fffffffll)
        return "too many sequences";
    Maps tm;
    if (!make_map(&tm.q, q, fp16 != 0, n_batch, seq_len, FH) || !make_map(&tm.k, k, fp16 != 0, n_batch, seq_len, FH) ||
        !make_map(&tm.v, v, fp16 != 0, n_batch, seq_len, FH) || !make_map(&tm.o, o, fp16 != 0, n_batch, seq_len, FH))
        return "cuTensorMapEncodeTiled failed";
    Args g{};
    g.T = seq_len; g.F = inner; g.H = heads; g.B = (int)n_batch;
    g.n_qt = (seq_len + kBM - 1) / kBM;
    g.n_kv = (seq_len + kBN - 1) / kBN;
    g.n_seq = n_batch * FH;
    g.n_units = ((g.n_seq + 1) / 2) * g.n_qt;
    g.scale_log2 = scale * 1.4426950408889634f;
    g.gates = gates;
    g.gate_ld = gates != nullptr ? (gate_ld > 0 ? gate_ld : heads) : 0;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) { *cuda_err = e; return "cudaGetDevice failed"; }
    static std::mutex mu;
    static bool attr[64][2] = {};
    static int n_sm[64] = {};
    {
        std::lock_guard<std::mutex> lk(mu);
        if (n_sm[dev & 63] == 0) {
            cudaDeviceGetAttribute(&n_sm[dev & 63], cudaDevAttrMultiProcessorCount, dev);
            if (n_sm[dev & 63] <= 0) n_sm[dev & 63] = 148;
        }
        if (!attr[dev & 63][fp16 ? 1 : 0]) {
            e = fp16 ? cudaFuncSetAttribute(time_attn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem::kDynamic)
                     : cudaFuncSetAttribute(time_attn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem::kDynamic);
            if (e != cudaSuccess) { *cuda_err = e; return "cudaFuncSetAttribute failed"; }
            attr[dev & 63][fp16 ? 1 : 0] = true;
        }
    }
    const long long grid = g.n_units < n_sm[dev & 63] ? g.n_units : n_sm[dev & 63];
    if (fp16) time_attn_kernel<true><<<(unsigned)grid, kThreads, Smem::kDynamic, stream>>>(tm, g);
    else time_attn_kernel<false><<<(unsigned)grid, kThreads, Smem::kDynamic, stream>>>(tm, g);
    count_launch();
    *cuda_err = cudaGetLastError();
    return *cuda_err == cudaSuccess ? nullptr : "launch failed";
}

}  // namespace al
