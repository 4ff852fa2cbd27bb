// K4: the dense contractions of the RoFormer mask network on the 5th-generation tensor cores (sm_100a).
//
//   out[g][m, n] = epilogue( sum_k A[g][m, k] * W[g][n, k] )        A, W bf16 K-major (nn.Linear layout), fp32 accumulate
//
// One persistent CTA per SM, warp-specialised (the roles never meet on a CTA-wide barrier after start-up):
//   warp 0     producer : TMA (cp.async.bulk.tensor) loads of 128 x 64 A tiles and BN x 64 W tiles, 128-byte swizzle,
//                         into a STAGES-deep shared-memory ring (mbarrier full / empty)
//   warp 1     MMA      : one thread issues tcgen05.mma (128 x BN x 16, bf16 -> fp32) from shared-memory descriptors
//                         into one of TWO accumulators in tensor memory; tcgen05.commit releases ring slots and
//                         publishes the finished accumulator
//   warps 2..9 epilogue : tcgen05.ld of their 32-lane quadrant (two warps per quadrant split the columns), the fused
//                         epilogue in registers on the packed fp32 pipe, swizzled staging in shared memory and TMA
//                         stores; accumulator t+1 is being computed while t is drained
//
// Fused epilogues (what the reference runs as separate PyTorch kernels between two GEMMs):
//   EPI_BF16 : v = acc * rowscale[m] + bias[n]  -> optional rotary embedding on column pairs -> optional GELU / tanh
//              -> bf16.  rowscale = sqrt(d) / max(||x_m||, eps) from the partial sums of squares the residual epilogue
//              left behind: RMSNorm(x) W^T = diag(rowscale) x (gamma (.) W)^T, so the normalisation pass disappears
//              (gamma is folded into W once on the host).  The output columns may be routed to up to 4 tensors
//              (q / k / v / gates).
//   EPI_GLU  : W's rows interleaved (a_i, b_i): out[m, i] = (acc_2i + bias_2i) * sigmoid(acc_2i+1 + bias_2i+1), fp32 -- the
//              second Linear + GLU of the mask estimator, written straight into the mask tensor.
//   EPI_RES  : x32 += acc + bias (fp32 residual stream, read and written in place through TMA), xb = bf16(x32) (the
//              next GEMM's A operand), ss[m][n_tile][half] = partial sums of squares of the new row (the next rowscale).
//
// Reference call site accelerated: modules/separator/stem_separator.py:281 (separator.separate -> model forward under
// autocast, :106); upstream module tree restated in nets/roformer.py.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdlib.h>

#include <mutex>

#include "al_gemm.h"
#include "al_kernels.h"
#include "al_tc.cuh"

namespace al {
namespace tc {

constexpr int BM = 128;          // rows of A per tile = tensor memory lanes
constexpr int BK = 64;           // bf16 elements per 128-byte swizzled row
constexpr int UK = 16;           // K of one tcgen05.mma kind::f16
constexpr int kResSlots = 2;     // EPI_RES: ring of 32-column fp32 chunks per epilogue warp
// Epilogue warps EW: 8 = two per tensor-memory lane quadrant.  16 (four per quadrant, 3 ring stages to pay for their
// staging buffers) was measured 12-19 % SLOWER on the K = 512 shapes (profiles/r02l_gemm_epilogue_width.log): these shapes
// are paced by the depth of the operand ring, not by the epilogue arithmetic.
constexpr int threads_of(int ew) { return 64 + 32 * ew; }     // + TMA producer warp + MMA issuer warp

struct Tmaps {
    CUtensorMap a, b;
    CUtensorMap o[4];   // EPI_BF16: outputs (bf16, box 64 x 32, 128-byte swizzle)
                        // EPI_RES : o[0] = x32 (fp32, box 32 x 32, 128-byte swizzle), o[1] = xb (bf16, box 32 x 32, 64-byte
                        //           swizzle), o[2] = x32 again with a 64 x 128 box for the L2 prefetch
};

template <int BN, int STAGES, int EPI, int EW, int CG = 1>
struct SmemLayout {
    static constexpr int kA = BM * BK * 2;                    // 16 KB
    static constexpr int kB = BN / CG * BK * 2;               // a CTA of a pair holds half of the W tile
    static constexpr int kStage = kA + kB;
    static constexpr int kEpiPerWarp = EPI == EPI_RES ? (kResSlots * 4096 + 2048) : 4096;
    static constexpr int kEpiOff = STAGES * kStage;
    static constexpr int kBarOff = kEpiOff + EW * kEpiPerWarp;
    static constexpr int kNumBars = 2 * STAGES + 4 + EW * kResSlots;
    static constexpr int kTotal = kBarOff + kNumBars * 8 + 16;
    static constexpr int kDynamic = kTotal + 1024;            // slack for the manual 1024-byte alignment
};

// CG = 2: CTA pairs (cluster of two, tcgen05 cta_group::2).  The pair computes a 256 x BN tile -- CTA r owns rows 128 r ..
// 128 r + 127 of it in its own tensor memory and loads its own A tile and rows (BN / 2) r .. of the W tile; the MMA is issued
// by CTA 0 alone, TMA bytes of both CTAs are counted on CTA 0's full barrier, ring slots and finished accumulators are
// published to both CTAs by multicast commits, and the epilogue warps of both CTAs release an accumulator on CTA 0's barrier.
// A CTA then moves 32 KB instead of 48 KB of operands per 128 x 256 x 64 block and the ring is 6 deep instead of 4: the
// K = 512 shapes of the network are paced by operand traffic from L2 (28 GB per to_qkv launch = the LTS throughput cap).
// VAR = what the EPI_BF16 epilogue computes, fixed at compile time for the shapes of the network (one kernel with every
// ingredient behind run-time flags is 44 KB of SASS and 159 registers): VAR_ANY = run-time flags (any combination),
// VAR_ROT = row scale + bias + rotary (to_qkv + to_gates), VAR_GELU = row scale + bias + GELU (FeedForward Linear 1).
enum { VAR_ANY = 0, VAR_ROT = 1, VAR_GELU = 2 };
template <int BN, int STAGES, int EPI, bool F16, int EW, int CG, int VAR = VAR_ANY>
__global__ void __launch_bounds__(threads_of(EW), 1)
gemm_bf16_kernel(const __grid_constant__ Tmaps tm, const GemmArgs g) {
    constexpr bool kMayRot = VAR == VAR_ANY || VAR == VAR_ROT;
    constexpr bool kFixed = VAR != VAR_ANY;
    using L = SmemLayout<BN, STAGES, EPI, EW, CG>;
    constexpr int kEpiWarps = EW;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t raw = smem_addr(smem_dyn);
    const uint32_t base = (raw + 1023u) & ~1023u;            // 128-byte swizzle atoms are 1024-byte aligned
    unsigned char* base_ptr = smem_dyn + (base - raw);
    const uint32_t bars = base + L::kBarOff;
    auto full_bar = [&](int s) { return bars + 8u * s; };
    auto empty_bar = [&](int s) { return bars + 8u * (STAGES + s); };
    auto tfull_bar = [&](int a) { return bars + 8u * (2 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bars + 8u * (2 * STAGES + 2 + a); };
    auto res_bar = [&](int w, int s) { return bars + 8u * (2 * STAGES + 4 + w * kResSlots + s); };
    const uint32_t tmem_slot = bars + 8u * L::kNumBars;
    volatile uint32_t* tmem_slot_ptr = reinterpret_cast<volatile uint32_t*>(base_ptr + L::kBarOff + 8 * L::kNumBars);

    const int warp = (int)(threadIdx.x >> 5);
    const int lane = lane_id();
    constexpr uint32_t kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;   // two accumulators; BN is a power of two >= 16

    if (warp == 0) {
        if (lane == 0) {
            tma_prefetch_desc(&tm.a);
            tma_prefetch_desc(&tm.b);
            tma_prefetch_desc(&tm.o[0]);
            if (EPI == EPI_RES) tma_prefetch_desc(&tm.o[1]);
        }
        if constexpr (CG == 2) {
            tmem_alloc2(tmem_slot, kTmemCols);
            tmem_relinquish2();
        } else {
            tmem_alloc(tmem_slot, kTmemCols);
            tmem_relinquish();
        }
    } else if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), kEpiWarps * CG);            // pairs: the epilogue warps of both CTAs arrive on CTA 0's
        }
        for (int w = 0; w < kEpiWarps; ++w)
            for (int s = 0; s < kResSlots; ++s) mbar_init(res_bar(w, s), 1);
        fence_mbar_init();
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (CG == 2) cluster_sync_all();                   // the peer's barriers exist before anything arrives on them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    // tiles: (group, row block, column block); a CTA pair walks PAIRS of row blocks, CTA r taking row block 2 pair + r
    const int m_units = CG == 2 ? (g.m_tiles + 1) / 2 : g.m_tiles;
    const int tiles_per_group = m_units * g.n_tiles;
    const int total_tiles = tiles_per_group * g.groups;
    const int k_blocks = (g.K + BK - 1) / BK;
    const int tile0 = CG == 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int tile_step = CG == 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    auto row_block = [&](int m_unit) { return CG == 2 ? 2 * m_unit + (int)rank : m_unit; };

    if (warp == 0) {
        // ================================= TMA producer =================================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = tile0; tile < total_tiles; tile += tile_step) {
                const int grp = tile / tiles_per_group;
                const int rem = tile - grp * tiles_per_group;
                const int m_unit = rem / g.n_tiles, n_blk = rem - m_unit * g.n_tiles;
                const int m_blk = row_block(m_unit);
                int n_cur = g.N - n_blk * BN;
                n_cur = n_cur >= BN ? BN : ((n_cur + 15) & ~15);
                if (EPI == EPI_RES && g.accumulate != 0 && g.pf_x != 0) {
                    // the fp32 residual tile this accumulator will be added to: into L2 now, so that the epilogue's
                    // small ring of TMA loads sees L2 latency, not HBM latency
#pragma unroll
                    for (int c = 0; c < BN / 64; ++c) tma_prefetch_3d(&tm.o[2], n_blk * BN + c * 64, m_blk * BM, grp);
                }
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1u);
                    const uint32_t sa = base + (uint32_t)stage * L::kStage;
                    if constexpr (CG == 2) {
                        // both CTAs' bytes land on CTA 0's barrier, which expects them all
                        if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), (uint32_t)(2 * L::kStage));
                        tma_load_3d_pair(sa, &tm.a, full_bar(stage), kb * BK, m_blk * BM, grp);
                        // CTA r supplies rows [r n / 2, (r + 1) n / 2) of the n-row B operand (n < BN in a ragged last tile)
                        tma_load_3d_pair(sa + L::kA, &tm.b, full_bar(stage), kb * BK, n_blk * BN + (int)rank * (n_cur / 2), grp);
                    } else {
                        mbar_arrive_expect_tx(full_bar(stage), (uint32_t)L::kStage);
                        tma_load_3d(sa, &tm.a, full_bar(stage), kb * BK, m_blk * BM, grp);
                        tma_load_3d(sa + L::kA, &tm.b, full_bar(stage), kb * BK, n_blk * BN, grp);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================= MMA issuer =================================
        if (lane == 0 && rank == 0) {                             // pairs: CTA 0 issues for both CTAs
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = tile0; tile < total_tiles; tile += tile_step, ++it) {
                const int rem = tile % tiles_per_group;
                const int n_blk = rem % g.n_tiles;
                int n_cur = g.N - n_blk * BN;
                n_cur = n_cur >= BN ? BN : ((n_cur + 15) & ~15);
                const uint32_t idesc = umma_idesc_16bit(BM * CG, n_cur, F16);
                const int acc = it & 1;
                const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
                // epilogue has drained this accumulator (pairs: in both CTAs, the peer's warps arrive remotely)
                if constexpr (CG == 2) mbar_wait_cluster(tempty_bar(acc), acc_phase ^ 1u);
                else mbar_wait(tempty_bar(acc), acc_phase ^ 1u);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const uint32_t sa = base + (uint32_t)stage * L::kStage;
                    const uint64_t a_desc = umma_desc_sw128(sa);
                    const uint64_t b_desc = umma_desc_sw128(sa + L::kA);
#pragma unroll
                    for (int k = 0; k < BK / UK; ++k) {
                        // advancing K inside the 128-byte swizzle atom = advancing the start address by 32 bytes
                        if constexpr (CG == 2)
                            umma_bf16_ss_pair(d_tmem, a_desc + (uint64_t)(k * ((UK * 2) >> 4)), b_desc + (uint64_t)(k * ((UK * 2) >> 4)),
                                              idesc, (uint32_t)((kb | k) != 0));
                        else
                            umma_bf16_ss(d_tmem, a_desc + (uint64_t)(k * ((UK * 2) >> 4)), b_desc + (uint64_t)(k * ((UK * 2) >> 4)),
                                         idesc, (uint32_t)((kb | k) != 0));
                    }
                    // ring slot free once these MMAs have read it (pairs: in both CTAs)
                    if constexpr (CG == 2) umma_commit_pair(empty_bar(stage));
                    else umma_commit(empty_bar(stage));
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                if constexpr (CG == 2) umma_commit_pair(tfull_bar(acc));         // accumulator complete
                else umma_commit(tfull_bar(acc));
            }
        }
    } else {
        // ================================= epilogue warps =================================
        // 8 warps: warp w may touch tensor memory lanes 32 (w % 4) .. + 31 only, so two warps share a lane quadrant
        // and split the accumulator's columns (half 0 / half 1).  Two warps per scheduler keep the issue slots busy
        // while one of them waits on tcgen05.ld / shared memory / the TMA store.
        const int ew = warp - 2;                 // 0..EW-1: private staging buffers / barriers
        const int q = warp & 3;                  // tensor memory lane quadrant
        const int half = ew >> 2;                // which column block of the N tile (EW / 4 blocks)
        const uint32_t ebuf = base + L::kEpiOff + (uint32_t)ew * L::kEpiPerWarp;
        unsigned char* ebuf_ptr = base_ptr + L::kEpiOff + ew * L::kEpiPerWarp;
        const uint32_t lane_taddr = tmem_base + ((uint32_t)(q * 32) << 16);

        if constexpr (EPI != EPI_RES) {
            constexpr int kColBlocks = EW / 4;
            constexpr int kWarpCols = BN / kColBlocks >= 64 ? BN / kColBlocks : 64;   // columns per warp, in 64-column steps
            const int wcol0 = half * kWarpCols;                         // first column of this warp inside the tile
            // The tile sequence of this CTA (tile0, tile0 + tile_step, ...) is decoded incrementally (group, row unit, column
            // block): the per-tile prologue of the eight epilogue warps was a fifth of their time (integer divisions with long
            // dependent chains, profiles/r02z_*).
            const int dn = tile_step % g.n_tiles, dm = tile_step / g.n_tiles;
            int grp = tile0 / tiles_per_group;
            int m_unit = (tile0 - grp * tiles_per_group) / g.n_tiles;
            int n_blk = tile0 - grp * tiles_per_group - m_unit * g.n_tiles;
            auto advance = [&](int& gr, int& mu, int& nb) {
                nb += dn;
                mu += dm;
                if (nb >= g.n_tiles) { nb -= g.n_tiles; ++mu; }
                while (mu >= m_units) { mu -= m_units; ++gr; }
            };
            // partial sums of squares of this lane's row (the previous residual epilogue left them): the RAW values are fetched
            // one tile ahead and only summed when the tile comes up, so that their latency hides behind a tile of arithmetic
            const bool ss_vec = (reinterpret_cast<uintptr_t>(g.row_ss) & 15) == 0;
            auto load_ss = [&](bool valid, int gr, int mu) -> float4 {
                float4 r4 = make_float4(1.f, 0.f, 0.f, 0.f);
                if (g.row_ss == nullptr || !valid) return r4;
                const int row_ = row_block(mu) * BM + q * 32 + lane;
                if (row_ >= g.M) return r4;
                const float* sp = g.row_ss + ((long long)gr * g.side_gs + (long long)row_ * g.side_rs) * g.ss_parts;
                if (g.ss_parts == 2 && ss_vec) {
                    const float2 t2 = __ldg(reinterpret_cast<const float2*>(sp));
                    r4 = make_float4(t2.x, t2.y, 0.f, 0.f);
                } else if (g.ss_parts == 4 && ss_vec) {
                    r4 = __ldg(reinterpret_cast<const float4*>(sp));
                } else {
                    float ss = 0.f;
                    for (int p = 0; p < g.ss_parts; ++p) ss += __ldg(sp + p);
                    r4 = make_float4(ss, 0.f, 0.f, 0.f);
                }
                return r4;
            };
            float4 ss_cur = load_ss(tile0 < total_tiles, grp, m_unit);
            int it = 0;
            for (int tile = tile0; tile < total_tiles; tile += tile_step, ++it) {
                const int m_blk = row_block(m_unit);
                const int acc = it & 1;
                const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
                const int row0 = m_blk * BM + q * 32;
                const int row = row0 + lane;
                const int cur_grp = grp, cur_n = n_blk;
                advance(grp, m_unit, n_blk);                                  // (grp, m_unit, n_blk) now name the NEXT tile
                const float4 ss_next = load_ss(tile + tile_step < total_tiles, grp, m_unit);
                float rs = 1.f;
                if (g.row_ss != nullptr && row < g.M)
                    rs = g.ss_scale / fmaxf(sqrtf((ss_cur.x + ss_cur.y) + (ss_cur.z + ss_cur.w)), g.ss_eps);
                ss_cur = ss_next;
                const float2 rs2 = make_float2(rs, rs);
                const int n0 = cur_n * BN;
                const int n_cols = min(BN, g.N - n0);
                int pos = 0;
                const bool rot_tile = kMayRot && g.cos_sin != nullptr && n0 + wcol0 < g.rot_cols;
                if (rot_tile) {
                    pos = (int)(((uint32_t)row / (uint32_t)g.pos_div) % (uint32_t)g.pos_mod);
                    // this lane's (cos, sin) row (256 bytes) into L1 while the main loop of the tile still runs
                    const char* cr = reinterpret_cast<const char*>(g.cos_sin) + (long long)pos * 256;
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(cr));
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(cr + 128));
                }
                if (g.bias != nullptr && lane < kWarpCols / 32 && n0 + wcol0 + lane * 32 < g.N)
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(g.bias + (long long)cur_grp * g.N + n0 + wcol0 + lane * 32));
                mbar_wait(tfull_bar(acc), acc_phase);
                tc_fence_after();
                int steps = 0;
                if (wcol0 < BN && wcol0 < n_cols) steps = (min(n_cols - wcol0, kWarpCols) + 63) >> 6;
                if (steps == 0) {                                   // nothing of this tile is ours (ragged / narrow tile)
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) { if constexpr (CG == 2) mbar_arrive_leader(tempty_bar(acc)); else mbar_arrive(tempty_bar(acc)); }
                }
                for (int st = 0; st < steps; ++st) {
                    uint32_t r[2][32];
                    const int tcol = wcol0 + st * 64;
                    const uint32_t ta = lane_taddr + (uint32_t)(acc * BN + tcol);
                    tmem_ld_32x32(ta, r[0]);
                    tmem_ld_32x32(ta + 32, r[1]);
                    tmem_wait_ld();
                    if (st == steps - 1) {                      // our part of the accumulator is in registers
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) { if constexpr (CG == 2) mbar_arrive_leader(tempty_bar(acc)); else mbar_arrive(tempty_bar(acc)); }
                    }
                    const int col0 = n0 + tcol;
                    const bool rot = kMayRot && g.cos_sin != nullptr && col0 < g.rot_cols;
                    // the staging buffer is still being read by the TMA store of the previous step
                    if (lane == 0) bulk_wait_read<0>();
                    __syncwarp();
                    unsigned char* sb = ebuf_ptr + lane * 128;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        float2 v[16];
                        if (kFixed || g.bias != nullptr) {
                            // all eight loads go out together: the index is clamped instead of predicated (a predicated load is
                            // a branch per load, each waiting for the one before); columns >= N are clipped by the TMA store
                            const float* bg = g.bias + (long long)cur_grp * g.N;
                            float4 bq[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                bq[j] = __ldg(reinterpret_cast<const float4*>(bg + min(col0 + h * 32 + j * 4, g.N - 4)));
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float4 b = bq[j];
                                v[2 * j] = __ffma2_rn(make_float2(__uint_as_float(r[h][4 * j]), __uint_as_float(r[h][4 * j + 1])), rs2,
                                                      make_float2(b.x, b.y));
                                v[2 * j + 1] = __ffma2_rn(make_float2(__uint_as_float(r[h][4 * j + 2]), __uint_as_float(r[h][4 * j + 3])),
                                                          rs2, make_float2(b.z, b.w));
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                v[j] = __fmul2_rn(make_float2(__uint_as_float(r[h][2 * j]), __uint_as_float(r[h][2 * j + 1])), rs2);
                        }
                        if (rot) {
                            // pair (2i, 2i+1) of a 64-wide head turns by pos * freq_i; a 64-column step is one head
                            const float4* cp = reinterpret_cast<const float4*>(g.cos_sin) + ((long long)pos * 32 + h * 16) / 2;
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const float4 c = __ldg(cp + j);          // (cos, sin) of pairs 2j, 2j+1 of this half
                                const float2 a = v[2 * j], b = v[2 * j + 1];
                                v[2 * j] = make_float2(a.x * c.x - a.y * c.y, a.y * c.x + a.x * c.y);
                                v[2 * j + 1] = make_float2(b.x * c.z - b.y * c.w, b.y * c.z + b.x * c.w);
                            }
                        }
                        if (VAR == VAR_GELU || (!kFixed && g.act == ACT_GELU)) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = gelu_erf2(v[j]);
                        } else if (!kFixed && g.act == ACT_TANH) {
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = make_float2(tanh_fast(v[j].x), tanh_fast(v[j].y));
                        }
                        if constexpr (EPI == EPI_GLU) {
                            // W's rows were interleaved on the host: column 2i = a_i, 2i + 1 = b_i; out_i = a_i sigmoid(b_i), fp32
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float4 u;
                                u.x = v[4 * j].x * sigmoid_fast(v[4 * j].y);
                                u.y = v[4 * j + 1].x * sigmoid_fast(v[4 * j + 1].y);
                                u.z = v[4 * j + 2].x * sigmoid_fast(v[4 * j + 2].y);
                                u.w = v[4 * j + 3].x * sigmoid_fast(v[4 * j + 3].y);
                                const int chunk = h * 4 + j;             // 16 outputs of this half = 4 chunks of the 128-byte row
                                *reinterpret_cast<float4*>(sb + ((chunk ^ (lane & 7)) << 4)) = u;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                uint4 u;
                                u.x = pack16<F16>(v[4 * j].x, v[4 * j].y);
                                u.y = pack16<F16>(v[4 * j + 1].x, v[4 * j + 1].y);
                                u.z = pack16<F16>(v[4 * j + 2].x, v[4 * j + 2].y);
                                u.w = pack16<F16>(v[4 * j + 3].x, v[4 * j + 3].y);
                                const int chunk = h * 4 + j;             // 16-byte chunk of the 128-byte row
                                *reinterpret_cast<uint4*>(sb + ((chunk ^ (lane & 7)) << 4)) = u;
                            }
                        }
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (EPI == EPI_GLU) {
                            tma_store_3d(&tm.o[0], ebuf, col0 >> 1, row0, cur_grp);
                        } else {
                            const int oi = col0 / g.out_split;
                            tma_store_3d(&tm.o[oi], ebuf, col0 - oi * g.out_split, row0, cur_grp);
                        }
                        bulk_commit();
                    }
                }
            }
            if (lane == 0) bulk_wait<0>();
        } else {
            // ---------------- EPI_RES: x32 += acc + bias ; xb = bf16(x32) ; ss partials ----------------
            static_assert(EPI != EPI_RES || EW == 8, "the residual epilogue splits the tile over two warps per quadrant");
            // Each warp streams the 32-column chunks of its half of the tile: the fp32 residual chunk arrives by TMA
            // (prefetched into L2 by the producer warp when the tile's main loop started) into a 2-slot ring, is
            // updated in place and leaves by TMA together with its bf16 image.
            constexpr int CH = BN / 64;                                 // chunks per tile per warp
            const int n_my = tile0 < total_tiles ? (total_tiles - tile0 + tile_step - 1) / tile_step : 0;
            const long long n_chunks = (long long)n_my * CH;
            auto coords = [&](long long gc, int& grp, int& row0, int& col0, int& tcol) {
                const int it = (int)(gc / CH), c = (int)(gc - (long long)it * CH);
                const int tile = tile0 + it * tile_step;
                grp = tile / tiles_per_group;
                const int rem = tile - grp * tiles_per_group;
                const int m_unit = rem / g.n_tiles, n_blk = rem - m_unit * g.n_tiles;
                row0 = row_block(m_unit) * BM + q * 32;
                tcol = half * (BN / 2) + c * 32;
                col0 = n_blk * BN + tcol;
            };
            auto issue_load = [&](long long gc) {
                int grp, row0, col0, tcol;
                coords(gc, grp, row0, col0, tcol);
                const int slot = (int)(gc & (kResSlots - 1));
                mbar_arrive_expect_tx(res_bar(ew, slot), 4096u);
                tma_load_3d(ebuf + (uint32_t)slot * 4096, &tm.o[0], res_bar(ew, slot), col0, row0, grp);
            };
            const bool acc_in = g.accumulate != 0;               // 0: x32 = acc + bias (the stream starts here), no residual read
            if (lane == 0 && n_chunks > 0 && acc_in) issue_load(0);
            float ss = 0.f;
            for (long long gc = 0; gc < n_chunks; ++gc) {
                const int it = (int)(gc / CH), c = (int)(gc - (long long)it * CH);
                const int acc = it & 1;
                const uint32_t acc_phase = (uint32_t)(it >> 1) & 1u;
                int grp, row0, col0, tcol;
                coords(gc, grp, row0, col0, tcol);
                if (c == 0) {
                    mbar_wait(tfull_bar(acc), acc_phase);
                    tc_fence_after();
                    ss = 0.f;
                }
                uint32_t r[32];
                tmem_ld_32x32(lane_taddr + (uint32_t)(acc * BN + tcol), r);
                const int slot = (int)(gc & (kResSlots - 1));
                if (lane == 0) {
                    bulk_wait_read<0>();                      // chunk gc-1's stores have read their buffers
                    if (gc + 1 < n_chunks && acc_in) issue_load(gc + 1);
                }
                __syncwarp();
                if (acc_in) mbar_wait(res_bar(ew, slot), (uint32_t)(gc / kResSlots) & 1u);
                tmem_wait_ld();
                if (c == CH - 1) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) { if constexpr (CG == 2) mbar_arrive_leader(tempty_bar(acc)); else mbar_arrive(tempty_bar(acc)); }
                }
                unsigned char* rb = ebuf_ptr + slot * 4096 + lane * 128;
                unsigned char* xb = ebuf_ptr + kResSlots * 4096 + lane * 64;
                const float* bias = g.bias != nullptr ? g.bias + (long long)grp * g.N + col0 : nullptr;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4* p = reinterpret_cast<float4*>(rb + ((j ^ (lane & 7)) << 4));
                    float4 x = acc_in ? *p : make_float4(0.f, 0.f, 0.f, 0.f);
                    x.x += __uint_as_float(r[4 * j]);
                    x.y += __uint_as_float(r[4 * j + 1]);
                    x.z += __uint_as_float(r[4 * j + 2]);
                    x.w += __uint_as_float(r[4 * j + 3]);
                    if (bias != nullptr) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(bias) + j);
                        x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w;
                    }
                    ss = fmaf(x.x, x.x, ss);
                    ss = fmaf(x.y, x.y, ss);
                    ss = fmaf(x.z, x.z, ss);
                    ss = fmaf(x.w, x.w, ss);
                    *p = x;
                    r[4 * j] = __float_as_uint(x.x);
                    r[4 * j + 1] = __float_as_uint(x.y);
                    r[4 * j + 2] = __float_as_uint(x.z);
                    r[4 * j + 3] = __float_as_uint(x.w);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint4 u;
                    u.x = pack16<F16>(__uint_as_float(r[8 * j]), __uint_as_float(r[8 * j + 1]));
                    u.y = pack16<F16>(__uint_as_float(r[8 * j + 2]), __uint_as_float(r[8 * j + 3]));
                    u.z = pack16<F16>(__uint_as_float(r[8 * j + 4]), __uint_as_float(r[8 * j + 5]));
                    u.w = pack16<F16>(__uint_as_float(r[8 * j + 6]), __uint_as_float(r[8 * j + 7]));
                    *reinterpret_cast<uint4*>(xb + ((j ^ ((lane >> 1) & 3)) << 4)) = u;   // 64-byte swizzle
                }
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) {
                    tma_store_3d(&tm.o[0], ebuf + (uint32_t)slot * 4096, col0, row0, grp);
                    tma_store_3d(&tm.o[1], ebuf + (uint32_t)(kResSlots * 4096), col0, row0, grp);
                    bulk_commit();
                }
                if (c == CH - 1 && g.ss_out != nullptr) {
                    // the two warps of a lane quadrant each hold half of the slab's sum: [row][n_tile][half]
                    const long long row = (long long)row0 + lane;
                    if (row < g.M)
                        g.ss_out[(((long long)grp * g.side_gs + row * g.side_rs) * g.n_tiles + (col0 / BN)) * 2 + half] = ss;
                }
            }
            if (lane == 0) bulk_wait<0>();
        }
    }
    tc_fence_before();
    __syncthreads();
    if constexpr (CG == 2) {
        cluster_sync_all();          // neither CTA leaves (or frees tensor memory) while the pair still works on its behalf
        if (warp == 0) tmem_dealloc2(tmem_base, kTmemCols);
    } else {
        if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ------------------------------------------------------------------------------------------------------------------
// host side
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

// rank-3 map of a [groups][rows][cols] tensor (cols innermost), box = box_cols x box_rows x 1
static bool make_map(CUtensorMap* m, const void* ptr, CUtensorMapDataType dt, int elem, long long cols, long long rows,
                     long long groups, long long ld, long long group_stride, int box_cols, int box_rows,
                     CUtensorMapSwizzle sw) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    cuuint64_t dims[3] = {(cuuint64_t)cols, (cuuint64_t)rows, (cuuint64_t)groups};
    cuuint64_t strides[2] = {(cuuint64_t)(ld * elem), (cuuint64_t)((groups > 1 ? group_stride : ld * rows) * elem)};
    cuuint32_t box[3] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows, 1u};
    cuuint32_t estr[3] = {1u, 1u, 1u};
    return fn(m, dt, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct DevInfo {
    int n_sm = 0;
    bool attr[32] = {};
};
static DevInfo& dev_info(int dev) {
    static DevInfo info[64];
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    DevInfo& d = info[dev & 63];
    if (d.n_sm == 0) {
        cudaDeviceGetAttribute(&d.n_sm, cudaDevAttrMultiProcessorCount, dev);
        if (d.n_sm <= 0) d.n_sm = 148;
    }
    return d;
}

template <int BN, int STAGES, int EPI, bool F16, int EW, int CG, int VAR = VAR_ANY>
static cudaError_t launch_cfg1(const Tmaps& tm, const GemmArgs& g, int slot, cudaStream_t stream) {
    using L = SmemLayout<BN, STAGES, EPI, EW, CG>;
    static_assert(L::kDynamic <= 232448, "shared memory budget");
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    DevInfo& d = dev_info(dev);
    auto kernel = gemm_bf16_kernel<BN, STAGES, EPI, F16, EW, CG, VAR>;
    if (!d.attr[slot]) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kDynamic);
        if (e != cudaSuccess) return e;
        d.attr[slot] = true;
    }
    const int units = (CG == 2 ? (g.m_tiles + 1) / 2 : g.m_tiles) * g.n_tiles * g.groups;
    int grid = units * CG < d.n_sm ? units * CG : (d.n_sm / CG) * CG;
    if (g.max_ctas > 0 && grid > g.max_ctas) grid = (g.max_ctas / CG) * CG > 0 ? (g.max_ctas / CG) * CG : CG;
    if constexpr (CG == 2) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3((unsigned)threads_of(EW));
        cfg.dynamicSmemBytes = L::kDynamic;
        cfg.stream = stream;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 2;
        at[0].val.clusterDim.y = 1;
        at[0].val.clusterDim.z = 1;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        e = cudaLaunchKernelEx(&cfg, kernel, tm, g);
        count_launch();
        return e != cudaSuccess ? e : cudaGetLastError();
    } else {
        kernel<<<grid, threads_of(EW), L::kDynamic, stream>>>(tm, g);
        count_launch();
        return cudaGetLastError();
    }
}

template <int BN, int STAGES, int EPI, int EW = 8, int CG = 1, int VAR = VAR_ANY>
static cudaError_t launch_cfg(const Tmaps& tm, const GemmArgs& g, int slot, cudaStream_t stream) {
    return g.fp16 ? launch_cfg1<BN, STAGES, EPI, true, EW, CG, VAR>(tm, g, slot + 16, stream)
                  : launch_cfg1<BN, STAGES, EPI, false, EW, CG, VAR>(tm, g, slot, stream);
}

// AL_GEMM_PAIRS=1 selects CTA pairs (cta_group::2) for the 256-wide EPI_BF16 / EPI_GLU tiles.  Measured
// (profiles/r02y_gemm_cta_pairs.log): the plain GEMM gains 8-12 % (FF1 shape 2.71 -> 2.53 ms: less operand traffic from L2,
// 6-deep ring), but with the fused epilogues the kernel is paced by the epilogue warps and the pair adds cross-CTA hand-offs:
// to_qkv 2.12 -> 2.25 ms, contract line 161.2 -> 158.7 audio-s/s.  Off by default.
static bool cta_pairs() {
    static const bool on = [] {
        const char* e = getenv("AL_GEMM_PAIRS");
        return e != nullptr && e[0] == '1';
    }();
    return on;
}

}  // namespace tc

const char* launch_gemm_bf16(const GemmCall& c, cudaStream_t stream, cudaError_t* cuda_err) {
    using namespace tc;
    *cuda_err = cudaSuccess;
    if (c.M <= 0 || c.N <= 0 || c.K <= 0 || c.groups <= 0) return "bad sizes";
    if ((c.K & 7) != 0 || (c.lda & 7) != 0 || (c.ldw & 7) != 0) return "K, lda, ldw must be multiples of 8 (16-byte rows)";
    if ((reinterpret_cast<uintptr_t>(c.A) | reinterpret_cast<uintptr_t>(c.W)) & 15) return "A, W must be 16-byte aligned";
    const int BN = c.epi == EPI_RES ? (c.N % 256 == 0 ? 256 : 128) : (c.N > 128 ? 256 : (c.N > 64 ? 128 : 64));
    GemmArgs g{};
    g.M = (int)c.M; g.N = c.N; g.K = c.K; g.groups = c.groups;
    g.m_tiles = (int)((c.M + BM - 1) / BM);
    g.n_tiles = (c.N + BN - 1) / BN;
    g.bias = c.bias;
    g.row_ss = c.row_ss; g.ss_parts = c.ss_parts; g.ss_scale = c.ss_scale; g.ss_eps = c.ss_eps;
    g.cos_sin = c.cos_sin; g.pos_div = c.pos_div > 0 ? c.pos_div : 1; g.pos_mod = c.pos_mod > 0 ? c.pos_mod : 1;
    g.rot_cols = c.rot_cols; g.act = c.act;
    g.ss_out = c.ss_out;
    g.max_ctas = c.max_ctas;
    g.side_rs = c.side_row_stride > 0 ? c.side_row_stride : 1;
    g.side_gs = c.side_row_stride > 0 ? c.side_group_stride : c.M;
    g.accumulate = c.no_accumulate ? 0 : 1;
    g.fp16 = c.operand_fp16 ? 1 : 0;
    {
        // AL_GEMM_PFX=1: L2 prefetch of the fp32 residual tile when its main loop starts.  Measured SLOWER
        // (profiles/r02m_gemm_residual_prefetch_variants.log: to_out 1.86 ms with, 1.59 ms without), so off by default.
        static const int pfx = [] { const char* e = getenv("AL_GEMM_PFX"); return e ? atoi(e) : 0; }();
        g.pf_x = pfx;
    }
    const CUtensorMapDataType dt16 = c.operand_fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
    if ((long long)g.m_tiles * g.n_tiles * g.groups > 0x7fffffffll) return "too many tiles";
    Tmaps tm;
    if (!make_map(&tm.a, c.A, dt16, 2, c.K, c.M, c.groups, c.lda, c.a_group_stride, BK, BM,
                  CU_TENSOR_MAP_SWIZZLE_128B))
        return "cuTensorMapEncodeTiled(A) failed";
    // CTA pairs (each CTA loads half of the W tile).  Residual epilogue: on for long reductions -- FF Linear 2 (K = 2048) is
    // paced by operand traffic from L2 (32 GB per launch) and gains 6 % (3.00 -> 2.81 ms), to_out (K = 512) is paced by the
    // residual stream in HBM and loses 12 % to the pair's hand-offs (profiles/r02y_gemm_cta_pairs.log); AL_GEMM_RES_PAIRS=0
    // switches them off.  EPI_BF16 / EPI_GLU: opt-in, see cta_pairs().
    static const bool res_pairs = [] { const char* e = getenv("AL_GEMM_RES_PAIRS"); return !(e != nullptr && e[0] == '0'); }();
    const bool pairs = BN == 256 && (c.epi == EPI_RES ? (res_pairs && c.K >= 1024) : cta_pairs());
    if (!make_map(&tm.b, c.W, dt16, 2, c.K, c.N, c.groups, c.ldw, c.w_group_stride, BK, pairs ? BN / 2 : BN,
                  CU_TENSOR_MAP_SWIZZLE_128B))
        return "cuTensorMapEncodeTiled(W) failed";
    if (c.epi == EPI_RES) {
        if (c.N % 128 != 0) return "residual epilogue needs N to be a multiple of 128";
        if (!c.x32 || !c.xb) return "residual epilogue needs x32 and xb";
        if ((c.ldx & 3) != 0 || (c.ldxb & 7) != 0) return "ldx / ldxb must give 16-byte rows";
        if (!make_map(&tm.o[0], c.x32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, c.N, c.M, c.groups, c.ldx, c.x_group_stride, 32, 32,
                      CU_TENSOR_MAP_SWIZZLE_128B))
            return "cuTensorMapEncodeTiled(x32) failed";
        if (!make_map(&tm.o[1], c.xb, dt16, 2, c.N, c.M, c.groups, c.ldxb, c.xb_group_stride, 32, 32,
                      CU_TENSOR_MAP_SWIZZLE_64B))
            return "cuTensorMapEncodeTiled(xb) failed";
        if (!make_map(&tm.o[2], c.x32, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, c.N, c.M, c.groups, c.ldx, c.x_group_stride, 64, BM,
                      CU_TENSOR_MAP_SWIZZLE_NONE))
            return "cuTensorMapEncodeTiled(x32 prefetch) failed";
        tm.o[3] = tm.o[1];
        g.out_split = c.N;
        if (BN == 256 && pairs) *cuda_err = launch_cfg<256, 4, EPI_RES, 8, 2>(tm, g, 12, stream);
        else if (BN == 256) *cuda_err = launch_cfg<256, 3, EPI_RES>(tm, g, 0, stream);
        else *cuda_err = launch_cfg<128, 4, EPI_RES>(tm, g, 4, stream);
        return *cuda_err == cudaSuccess ? nullptr : "launch failed";
    }
    if (c.epi == EPI_GLU) {
        if ((c.N & 15) != 0) return "GLU epilogue needs N to be a multiple of 16 (interleaved (a, b) rows)";
        if (!c.out[0] || (reinterpret_cast<uintptr_t>(c.out[0]) & 15) != 0 || (c.ldo[0] & 3) != 0)
            return "GLU epilogue needs a 16-byte aligned fp32 output";
        const int out_cols = c.out_split > 0 ? c.out_split : c.N / 2;     // columns that exist in the output (<= N / 2)
        if (out_cols > c.N / 2) return "GLU epilogue: out_split (valid output columns) exceeds N / 2";
        if (!make_map(&tm.o[0], c.out[0], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, out_cols, c.M, c.groups, c.ldo[0], c.o_group_stride[0],
                      32, 32, CU_TENSOR_MAP_SWIZZLE_128B))
            return "cuTensorMapEncodeTiled(glu out) failed";
        tm.o[1] = tm.o[2] = tm.o[3] = tm.o[0];
        g.out_split = c.N;
        if (BN == 256 && pairs) *cuda_err = launch_cfg<256, 6, EPI_GLU, 8, 2>(tm, g, 8, stream);
        else if (BN == 256) *cuda_err = launch_cfg<256, 4, EPI_GLU>(tm, g, 5, stream);
        else if (BN == 128) *cuda_err = launch_cfg<128, 6, EPI_GLU>(tm, g, 6, stream);
        else *cuda_err = launch_cfg<64, 8, EPI_GLU>(tm, g, 7, stream);
        return *cuda_err == cudaSuccess ? nullptr : "launch failed";
    }
    // EPI_BF16
    const int split = c.out_split > 0 ? c.out_split : c.N;
    if (split % 64 != 0 && split != c.N) return "out_split must be a multiple of 64";
    if (split != c.N && split % BN != 0) return "out_split must be a multiple of the N tile";
    const int n_out = (c.N + split - 1) / split;
    if (n_out > 4) return "at most 4 output tensors";
    if ((c.N & 7) != 0) return "N must be a multiple of 8";
    if (c.cos_sin != nullptr && (c.rot_cols % 64 != 0)) return "rot_cols must be a multiple of 64";
    g.out_split = split;
    for (int i = 0; i < 4; ++i) {
        const void* o = c.out[i < n_out ? i : 0];
        const int src = i < n_out ? i : 0;
        if (!o) return "missing output tensor";
        if ((reinterpret_cast<uintptr_t>(o) & 15) != 0 || (c.ldo[src] & 7) != 0) return "outputs need 16-byte aligned rows";
        const int cols = (src == n_out - 1) ? c.N - split * (n_out - 1) : split;
        if (!make_map(&tm.o[i], o, dt16, 2, cols, c.M, c.groups, c.ldo[src],
                      c.o_group_stride[src], 64, 32, CU_TENSOR_MAP_SWIZZLE_128B))
            return "cuTensorMapEncodeTiled(out) failed";
    }
    // the two shapes that carry the step get their own instantiation (AL_GEMM_VARIANTS=0: the general kernel for everything)
    static const bool variants = [] { const char* e = getenv("AL_GEMM_VARIANTS"); return !(e != nullptr && e[0] == '0'); }();
    const bool v_rot = variants && c.bias && c.row_ss && c.cos_sin && c.act == ACT_NONE;
    const bool v_gelu = variants && c.bias && c.row_ss && !c.cos_sin && c.act == ACT_GELU;
    if (BN == 256 && pairs && v_rot) *cuda_err = launch_cfg<256, 6, EPI_BF16, 8, 2, VAR_ROT>(tm, g, 13, stream);
    else if (BN == 256 && pairs && v_gelu) *cuda_err = launch_cfg<256, 6, EPI_BF16, 8, 2, VAR_GELU>(tm, g, 14, stream);
    else if (BN == 256 && pairs) *cuda_err = launch_cfg<256, 6, EPI_BF16, 8, 2>(tm, g, 9, stream);
    else if (BN == 256 && v_rot) *cuda_err = launch_cfg<256, 4, EPI_BF16, 8, 1, VAR_ROT>(tm, g, 10, stream);
    else if (BN == 256 && v_gelu) *cuda_err = launch_cfg<256, 4, EPI_BF16, 8, 1, VAR_GELU>(tm, g, 11, stream);
    else if (BN == 256) *cuda_err = launch_cfg<256, 4, EPI_BF16>(tm, g, 1, stream);
    else if (BN == 128) *cuda_err = launch_cfg<128, 6, EPI_BF16>(tm, g, 2, stream);
    else *cuda_err = launch_cfg<64, 8, EPI_BF16>(tm, g, 3, stream);
    return *cuda_err == cudaSuccess ? nullptr : "launch failed";
}

}  // namespace al
