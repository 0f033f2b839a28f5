// K1 (v2): CTA-pair tensor-core distance contraction + fused BMU argmin + fused per-BMU accumulate.
//
// Same contract as bmu_tc.cuh (score = x . w' + bias minimised over neurons, 3xTF32 split, the
// (n, K) score matrix lives only in TMEM) re-tiled for the two-SM tensor-core mode of sm_100a:
//
//   * thread-block cluster of 2 CTAs on one TPC, tcgen05.mma.cta_group::2, UMMA tile M=256 N=256
//     K=8: each CTA owns 128 sample rows (its own TMEM accumulators, 2 x 256 columns) and stages
//     only HALF of every W' tile (128 neurons), so the codebook traffic L2->SMEM and the SMEM reads
//     of the B operand are halved with respect to the one-CTA kernel;
//   * operand rings in 192 KB of shared memory, SWIZZLE_128B.  Streaming (D > 32): three 64 KB stages
//     (X chunk hi | lo, W'hi half, W'lo half).  Resident (D <= 32, one feature block): two A buffers (the X
//     tile of a row tile is loaded and converted ONCE and reused by all its neuron tiles) + four B slots;
//   * W' tiles arrive by TMA with the .cta_group::2 form: both CTAs' loads complete on the
//     leader CTA's mbarrier; X chunks complete on a local mbarrier that the converter waits on;
//   * tcgen05.commit ... .multicast::cluster releases the stage / publishes the accumulator in
//     both CTAs at once;
//   * four extra "scatter" warps per CTA take the finished BMUs of a 128-row tile and issue the
//     S[bmu] += x reductions (red.global.add.v4.f32), so the ~1 element/clk/SM RED throughput
//     overlaps with the MMA and the argmin epilogue instead of stalling them;
//   * the bias is folded into the contraction when the last 32-feature block has three spare columns
//     (W'hi[:, d..d+2] = the TF32 pieces of bias_k, prepared by codebook_split_kernel; the converter sets
//     X[:, d..d+2] = 1), so the epilogue only compares; 8-column MMA steps that hold nothing but the TMA
//     zero fill are not issued (D = 16: 7 MMAs per tile instead of 12);
//   * the producer and MMA warps run their loops with all 32 lanes and issue under elect.sync, which keeps
//     every descriptor in uniform registers (a `lane == 0` branch costs an R2UR waterfall per instruction).
//
// Warp roles per CTA (576 threads): 0 TMA producer | 1 MMA issuer (leader CTA only) + TMEM alloc |
// 2-5 converter (X -> TF32 hi/lo) | 6-13 epilogue (TMEM -> argmin; two warps per TMEM lane quarter,
// each taking half of the 256 columns, so one hides the other's TMEM/L1 latency) | 14-17 scatter.
#pragma once
#include <cuda.h>
#include "common.cuh"
#include "bmu_tc.cuh"

namespace somb200 {
namespace tc2 {

using tc::FusedAcc;
using tc::smem_u32;

constexpr int BM = 128;        // sample rows per CTA (256 per pair)
constexpr int BN = 256;        // neurons per accumulator tile (UMMA N)
constexpr int BNH = 128;       // neurons staged by each CTA
constexpr int BK = 32, STAGES = 3, UMMA_K = 8;
constexpr int A_BYTES = BM * BK * 4;     // 16 KB
constexpr int BH_BYTES = BNH * BK * 4;   // 16 KB
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * BH_BYTES;   // 64 KB
constexpr int NUM_THREADS = 576;
constexpr int CONV_WARP0 = 2, EPI_WARP0 = 6, SCAT_WARP0 = 14;
constexpr int EPI_THREADS = 256;
constexpr int MAXS = 4;                  // ring slots per barrier family (A ring: 3 or 2, B ring: 3 or 4)
constexpr int NUM_BARS = 5 * MAXS + 8;
// stages | per-warp bias slices | merge buffers [2][BM] (float + int) | bmu hand-off [2][BM] | barriers | tmem slot | slack
constexpr int EPI_STAGE_BYTES = 8 * 128 * 4;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_STAGE_BYTES + tc::SCAT_STAGE_BYTES + 2 * BM * 8 + 2 * BM * 4 + NUM_BARS * 8 + 64 + 1024;

// ---- cluster / 2-SM PTX wrappers -------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_cta(uint32_t saddr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
// arrive on a barrier that may live in the peer CTA (shared::cluster address).  Default semantics, as
// CUTLASS's ClusterBarrier::arrive(cta_id): an explicit .release.cluster compiles to MEMBAR.ALL.GPU.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" :: "r"(cluster_addr) : "memory");
}
// Wait on a barrier of THIS CTA that threads of the peer CTA (and multicast tcgen05.commit) also arrive on.  The
// default CTA-scope acquire is what the data needs: everything the waiter consumes afterwards goes through the
// async proxy (tensor core reading SMEM the peer fenced with fence.proxy.async) or TMEM (tcgen05 fences).  A
// cluster-scope acquire here compiles to CCTL.IVALL -- a full L1 invalidate of the SM, three times per tile.
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity), "r"(tc::kSuspendHintNs) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait_cluster(bar, parity)) return;
    uint32_t polls = 0;
    long long t0 = 0;
    while (!mbar_try_wait_cluster(bar, parity)) {
        if ((++polls & 1023u) == 0) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            if (now - t0 > tc::kTrapCycles) {
                printf("som_b200(tc2): mbarrier timeout (block %d thread %d bar 0x%x parity %u)\n",
                       (int)blockIdx.x, (int)threadIdx.x, bar, parity);
                __trap();
            }
        }
    }
}
// TMA load whose completion bytes go to the LEADER CTA's mbarrier (peer bit cleared)
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar & 0xFEFFFFFFu) : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t slot_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(slot_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(bar), "h"((uint16_t)3) : "memory");
}

// UMMA tile of the pair: M = 256 (two CTAs x 128 lanes), N = 256
constexpr uint32_t kIdesc2 = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);

// RES (one 32-feature block, D <= 32): the converted X tile of a row tile stays RESIDENT across its neuron tiles
// (two A buffers, loaded and converted once per row tile; four B slots) instead of travelling with every stage.
template <bool RES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
bmu_tc2_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_whi,
               const __grid_constant__ CUtensorMap map_wlo, const float *__restrict__ bias,
               const unsigned int *__restrict__ gstat, int64_t n, int num_pair_tiles, int num_n_tiles, int num_k_blocks,
               int32_t *__restrict__ bmu_out, float *__restrict__ best_out, const FusedAcc acc) {
    pdl_wait(); pdl_trigger();        // programmatic dependent launch (common.cuh)
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *smem = smem_raw + (smem_base - smem_u32(smem_raw));

    float    *epi_stage = reinterpret_cast<float *>(smem + STAGES * STAGE_BYTES);                    // [8 warps][128]
    uint8_t  *scat_stage = smem + STAGES * STAGE_BYTES + EPI_STAGE_BYTES;            // [4 warps][SCAT_NBUF][128] int64
    uint8_t  *tail   = scat_stage + tc::SCAT_STAGE_BYTES;
    float    *mrg_v  = reinterpret_cast<float *>(tail);                       // [2][BM]
    int      *mrg_i  = reinterpret_cast<int *>(tail + 2 * BM * 4);            // [2][BM]
    int      *bmu_s  = reinterpret_cast<int *>(tail + 2 * BM * 8);            // [2][BM]
    uint64_t *bars   = reinterpret_cast<uint64_t *>(tail + 2 * BM * 8 + 2 * BM * 4);
    const uint32_t bar0 = smem_u32(bars);
    auto xfull_bar  = [&](int s) { return bar0 + 8u * s; };                     // local: X chunk landed
    auto bfull_bar  = [&](int s) { return bar0 + 8u * (MAXS + s); };            // leader: both W' halves landed
    auto ready_bar  = [&](int s) { return bar0 + 8u * (2 * MAXS + s); };        // leader: both converters done
    auto aempty_bar = [&](int s) { return bar0 + 8u * (3 * MAXS + s); };        // local: A slot consumed (multicast commit)
    auto bempty_bar = [&](int s) { return bar0 + 8u * (4 * MAXS + s); };        // local: B slot consumed (multicast commit)
    auto tfull_bar  = [&](int a) { return bar0 + 8u * (5 * MAXS + a); };        // local: accumulator complete (multicast commit)
    auto tempty_bar = [&](int a) { return bar0 + 8u * (5 * MAXS + 2 + a); };    // leader: both epilogues drained it
    auto bfullq_bar = [&](int b) { return bar0 + 8u * (5 * MAXS + 4 + b); };    // local: BMUs of a tile published
    auto bemptyq_bar = [&](int b) { return bar0 + 8u * (5 * MAXS + 6 + b); };   // local: scatter warps done with them
    // operand rings.  streaming: three 64 KB stages [A hi | A lo | B hi | B lo], A and B advance together;
    // resident: two A buffers [A hi | A lo] (one per row tile) followed by four B slots [B hi | B lo]
    constexpr int NA = RES ? 2 : STAGES, NB = RES ? 4 : STAGES;
    auto a_off = [&](int s) { return (uint32_t)(RES ? s * 2 * A_BYTES : s * STAGE_BYTES); };
    auto b_off = [&](int s) { return (uint32_t)(RES ? 2 * 2 * A_BYTES + s * 2 * BH_BYTES : s * STAGE_BYTES + 2 * A_BYTES); };
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(bars + NUM_BARS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    const bool fused = acc.S != nullptr;
    // bias folded into the contraction by prepare_codebook (three spare feature columns of the last K block hold
    // the TF32 pieces of the bias on the W'hi side, ones on the X side): the epilogue adds nothing
    const bool fold = gstat[3] != 0u;
    const int probe = (acc.dbg == 9 && blockIdx.x == 0) ? 1 : 0;     // timeline probe (tools/bmu_probe.py)

    if (threadIdx.x == 0) {
        for (int s = 0; s < NA; ++s) { tc::mbar_init(xfull_bar(s), 1); tc::mbar_init(ready_bar(s), 2 * 4); tc::mbar_init(aempty_bar(s), 1); }
        for (int s = 0; s < NB; ++s) { tc::mbar_init(bfull_bar(s), 1); tc::mbar_init(bempty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) {
            tc::mbar_init(tfull_bar(a), 1); tc::mbar_init(tempty_bar(a), 2 * EPI_THREADS / 32);
            tc::mbar_init(bfullq_bar(a), 4); tc::mbar_init(bemptyq_bar(a), 4);
        }
        tc::fence_barrier_init();
        tc::tma_prefetch_desc(&map_x); tc::tma_prefetch_desc(&map_whi); tc::tma_prefetch_desc(&map_wlo);
    }
    if (warp == 1) tmem_alloc_2sm(smem_u32((const void *)tmem_slot), 512);
    tc::tc_fence_before();
    cluster_sync();                      // barriers of BOTH CTAs initialised before anyone signals the peer
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer (each CTA loads its own rows and its half of W') ===========
        {
            uint32_t it = 0, tile_it = 0;
            for (int pt = pair; pt < num_pair_tiles; pt += num_pairs, ++tile_it)
                for (int nt = 0; nt < num_n_tiles; ++nt)
                    for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
                        const uint32_t ia = RES ? tile_it : it;                  // A ring position
                        const bool a_step = !RES || nt == 0;                     // this iteration brings a new X chunk
                        const int sa = ia % NA, sb = it % NB;
                        if (a_step) tc::mbar_wait(aempty_bar(sa), ((ia / NA) & 1) ^ 1);
                        tc::mbar_wait(bempty_bar(sb), ((it / NB) & 1) ^ 1);
                        if (tc::elect_one()) {
                            if (a_step) {
                                tc::mbar_expect_tx(xfull_bar(sa), A_BYTES);
                                tc::tma_load_2d(smem_base + a_off(sa), &map_x, kb * BK, pt * (2 * BM) + (int)rank * BM, xfull_bar(sa));
                            }
                            if (leader) tc::mbar_expect_tx(bfull_bar(sb), 4 * BH_BYTES);  // hi+lo halves of both CTAs
                            tma_load_2d_2sm(smem_base + b_off(sb),            &map_whi, kb * BK, nt * BN + (int)rank * BNH, bfull_bar(sb));
                            tma_load_2d_2sm(smem_base + b_off(sb) + BH_BYTES, &map_wlo, kb * BK, nt * BN + (int)rank * BNH, bfull_bar(sb));
                        }
                        __syncwarp();
                    }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer: one thread of the LEADER CTA drives both tensor cores ========
        if (leader) {
            uint32_t it = 0, acc_it = 0, tile_it = 0;
            for (int pt = pair; pt < num_pair_tiles; pt += num_pairs, ++tile_it)
                for (int nt = 0; nt < num_n_tiles; ++nt, ++acc_it) {
                    const int a = acc_it & 1; const uint32_t aph = (acc_it >> 1) & 1;
                    tc::dbg_stamp(probe, 0, acc_it);
                    const uint32_t tmem_d = tmem_base + (uint32_t)(a * BN);
                    for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
                        const uint32_t ia = RES ? tile_it : it;
                        const bool a_step = !RES || nt == 0;
                        const int sa = ia % NA, sb = it % NB;
                        mbar_wait_cluster(bfull_bar(sb), (it / NB) & 1);          // operands first: they are ready long before
                        if (a_step) mbar_wait_cluster(ready_bar(sa), (ia / NA) & 1);   // the accumulator is
                        if (kb == 0) {
                            mbar_wait_cluster(tempty_bar(a), aph ^ 1);  // both CTAs' epilogues drained this accumulator
                            tc::dbg_stamp(probe, 1, acc_it);
                        }
                        tc::tc_fence_after();
                        if (kb == 0) tc::dbg_stamp(probe, 2, acc_it);
                        const uint32_t sta = smem_base + a_off(sa), stb = smem_base + b_off(sb);
                        const uint64_t a_hi = tc::make_smem_desc(sta), a_lo = tc::make_smem_desc(sta + A_BYTES);
                        const uint64_t b_hi = tc::make_smem_desc(stb), b_lo = tc::make_smem_desc(stb + BH_BYTES);
                        // only the 8-column steps that hold real features are issued (the rest of the block is the
                        // TMA zero fill); the folded bias columns d..d+2 live in the hi*hi term alone
                        const int dl = acc.d - kb * BK;
                        const int kk_x = dl >= BK ? BK / UMMA_K : (dl + UMMA_K - 1) / UMMA_K;
                        const int kk_h = (fold && kb == num_k_blocks - 1) ? (dl + 3 + UMMA_K - 1) / UMMA_K : kk_x;
                        if (tc::elect_one()) {
#pragma unroll
                        for (int kk = 0; kk < BK / UMMA_K; ++kk) {
                            const uint64_t off = (uint64_t)((kk * UMMA_K * 4) >> 4);
                            if (kk < kk_x) {
                                umma_tf32_2sm(tmem_d, a_lo + off, b_hi + off, kIdesc2, (kb | kk) != 0);
                                umma_tf32_2sm(tmem_d, a_hi + off, b_lo + off, kIdesc2, 1);
                            }
                            if (kk < kk_h) umma_tf32_2sm(tmem_d, a_hi + off, b_hi + off, kIdesc2, 1);
                        }
                        umma_commit_2sm(bempty_bar(sb));
                        if (!RES || nt == num_n_tiles - 1) umma_commit_2sm(aempty_bar(sa));
                        if (kb == num_k_blocks - 1) {
                            umma_commit_2sm(tfull_bar(a));
                            tc::dbg_stamp(probe, 3, acc_it);
                        }
                        }
                        __syncwarp();
                    }
                }
        }
    } else if (warp < EPI_WARP0) {
        // ===================== converter: split the local X chunk into TF32 hi / lo =====================
        const int t = threadIdx.x - CONV_WARP0 * 32;
        // Folded bias: this thread's float4s (e = t + 128 i) all sit at 16-byte chunk (t & 7) of rows with the same
        // r & 7, i.e. at the SAME feature columns under the 128-byte swizzle: decide once which lanes hold d..d+2.
        const int o = acc.d - ((num_k_blocks - 1) * BK + ((((t & 7) ^ ((t >> 3) & 7))) << 2));
        const bool f0 = o <= 0 && o >= -2, f1 = o <= 1 && o >= -1, f2 = o <= 2 && o >= 0, f3 = o <= 3 && o >= 1;
        const bool any_f = fold && (f0 || f1 || f2 || f3);
        uint32_t it = 0;            // A ring position: one item per row tile (resident) or per (row tile, neuron tile, k block)
        for (int pt = pair; pt < num_pair_tiles; pt += num_pairs)
            for (int nt = 0; nt < (RES ? 1 : num_n_tiles); ++nt)
                for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
                    const int s = it % NA; const uint32_t ph = (it / NA) & 1;
                    const bool ones = any_f && kb == num_k_blocks - 1;
                    tc::mbar_wait(xfull_bar(s), ph);
                    float4 *ahi = reinterpret_cast<float4 *>(smem + a_off(s));
                    float4 *alo = reinterpret_cast<float4 *>(smem + a_off(s) + A_BYTES);
#pragma unroll
                    for (int i = 0; i < A_BYTES / 16 / 128; ++i) {
                        const int e = t + 128 * i;
                        float4 v = ahi[e];
                        if (ones) {          // zero-filled features d..d+2 become 1 (hi = 1, lo = 0 after the split)
                            v.x = f0 ? 1.f : v.x; v.y = f1 ? 1.f : v.y; v.z = f2 ? 1.f : v.z; v.w = f3 ? 1.f : v.w;
                        }
                        float4 h, l;
                        h.x = tc::tf32_rna_dev(v.x); h.y = tc::tf32_rna_dev(v.y); h.z = tc::tf32_rna_dev(v.z); h.w = tc::tf32_rna_dev(v.w);
                        l.x = tc::tf32_rna_dev(v.x - h.x); l.y = tc::tf32_rna_dev(v.y - h.y);
                        l.z = tc::tf32_rna_dev(v.z - h.z); l.w = tc::tf32_rna_dev(v.w - h.w);
                        ahi[e] = h; alo[e] = l;
                    }
                    tc::fence_proxy_async();
                    __syncwarp();                                           // one arrival per warp: every arrival
                    if (lane == 0) mbar_arrive_cluster(map_to_cta(ready_bar(s), 0));   // wakes the sleeping waiter
                }
    } else if (warp < SCAT_WARP0) {
        // ===================== epilogue: TMEM -> registers -> running argmin =====================
        // warp w: TMEM lane quarter q = w % 4 (hardware rule), column half h = (w - EPI_WARP0) / 4
        const int q = warp & 3;
        const int h = (warp - EPI_WARP0) >> 2;
        const int row_in_tile = q * 32 + lane;
        uint32_t acc_it = 0, tile_it = 0;
        for (int pt = pair; pt < num_pair_tiles; pt += num_pairs, ++tile_it) {
            tc::RunMin rm; rm.reset();
            // this warp's 128 bias values of the current neuron tile live in its private shared-memory
            // slice; the next tile's are prefetched into registers while this one is drained
            float *bs = epi_stage + (warp - EPI_WARP0) * 128;
            float4 nb = __ldg(reinterpret_cast<const float4 *>(bias + h * (BN / 2)) + lane);
            for (int nt = 0; nt < num_n_tiles; ++nt, ++acc_it) {
                const int a = acc_it & 1; const uint32_t aph = (acc_it >> 1) & 1;
                const int col0 = nt * BN + h * (BN / 2);
                __syncwarp();
                reinterpret_cast<float4 *>(bs)[lane] = nb;
                __syncwarp();
                nb = __ldg(reinterpret_cast<const float4 *>(bias + (nt + 1 < num_n_tiles ? nt + 1 : 0) * BN + h * (BN / 2)) + lane);
                if (warp == EPI_WARP0 && lane == 0) tc::dbg_stamp(probe, 4, acc_it);
                tc::mbar_wait(tfull_bar(a), aph);
                tc::tc_fence_after();
                if (warp == EPI_WARP0 && lane == 0) tc::dbg_stamp(probe, 5, acc_it);
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BN + h * (BN / 2));
#pragma unroll 1
                for (int c = 0; c < BN / 2 / 32 - 1; ++c) {
                    uint32_t v[32];
                    tc::tmem_ld32(taddr + c * 32, v);
                    tc::tmem_ld_wait_dep(v);
                    if (fold) rm.chunk_nobias(v, col0 + c * 32);
                    else      rm.chunk(v, bs + c * 32, col0 + c * 32);
                }
                {   // last chunk: once it is in registers the accumulator is free -- release it BEFORE the compares,
                    // so the MMAs of the tile after next start a chunk's worth of work earlier
                    constexpr int c = BN / 2 / 32 - 1;
                    uint32_t v[32];
                    tc::tmem_ld32(taddr + c * 32, v);
                    tc::tmem_ld_wait_dep(v);
                    tc::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(map_to_cta(tempty_bar(a), 0));
                    if (fold) rm.chunk_nobias(v, col0 + c * 32);
                    else      rm.chunk(v, bs + c * 32, col0 + c * 32);
                }
                if (warp == EPI_WARP0 && lane == 0) tc::dbg_stamp(probe, 6, acc_it);
            }
            float best; int bidx;
            rm.result(best, bidx);
            // merge the two column halves of each row (double-buffered hand-off through shared memory)
            const int mb = (tile_it & 1) * BM;
            if (h == 1) { mrg_v[mb + row_in_tile] = best; mrg_i[mb + row_in_tile] = bidx; }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (h == 0) {
                argmin_merge(best, bidx, mrg_v[mb + row_in_tile], mrg_i[mb + row_in_tile]);
                const int64_t row = (int64_t)pt * (2 * BM) + (int64_t)rank * BM + row_in_tile;
                if (row < n) {
                    if (bmu_out) bmu_out[row] = bidx;
                    if (best_out) best_out[row] = best;
                }
                if (fused) {          // hand the tile's BMUs to the scatter warps (double buffered)
                    const int b = tile_it & 1; const uint32_t bph = (tile_it >> 1) & 1;
                    tc::mbar_wait(bemptyq_bar(b), bph ^ 1);
                    bmu_s[b * BM + row_in_tile] = (row < n) ? bidx : -1;
                    __syncwarp();
                    if (lane == 0) tc::mbar_arrive(bfullq_bar(b));
                }
            }
        }
    } else {
        // ===================== scatter: S[bmu[r], :] += X[r, :], cnt[bmu[r]] += 1 =====================
        if (fused) {
            const int wq = warp - SCAT_WARP0;
            long long *stage = reinterpret_cast<long long *>(scat_stage) + wq * (tc::SCAT_NBUF * ACC_PIECE);
            uint32_t tile_it = 0, bulk_it = 0;
            const FusedAcc racc = acc.for_cta(blockIdx.x >> 1);          // this pair's replica of the accumulator
            for (int pt = pair; pt < num_pair_tiles; pt += num_pairs, ++tile_it) {
                const int b = tile_it & 1; const uint32_t bph = (tile_it >> 1) & 1;
                tc::mbar_wait_relaxed(bfullq_bar(b), bph, 400);
                const int64_t row0 = (int64_t)pt * (2 * BM) + (int64_t)rank * BM;
                scatter_rows_exact<tc::SCAT_NBUF>(racc, bmu_s + b * BM, row0, BM, wq, 4, lane, stage, bulk_it);
                __syncwarp();
                if (lane == 0) tc::mbar_arrive(bemptyq_bar(b));
            }
            bulk_wait_all();          // every bulk reduction of this thread has been performed
        }
    }

    __syncwarp();
    tc::tc_fence_before();
    cluster_sync();                      // the peer may still be signalling this CTA's barriers / reading its SMEM
    if (warp == 1) { tc::tc_fence_after(); tmem_dealloc_2sm(tmem_base, 512); }

}

inline int launch_bmu_tc2(const float *X, int64_t n, int d, int64_t ldx, int k, const WsLayout &L, uint8_t *ws,
                          int32_t *bmu, float *best, const AccTarget &T, int sm_count, cudaStream_t st) {
    SOM_REQUIRE(L.k_pad < (1 << 24), SOM_E_SHAPE,
                "tensor-core BMU kernels track the winning neuron as an exact fp32 integer: at most 2^24 neurons (k=%d)", k);
    SOM_REQUIRE(tc::shape_ok(X, n, d, ldx), SOM_E_SHAPE,
                "tensor-core BMU kernel needs ldx %% 4 == 0 and a 16-byte aligned X for TMA (d=%d ldx=%lld)", d, (long long)ldx);
    CUtensorMap mx, mhi, mlo;
    int rc;
    if ((rc = tc::make_map_2d(&mx, X, (uint64_t)d, (uint64_t)n, (uint64_t)ldx * 4, BK, BM))) return rc;
    if ((rc = tc::make_map_2d(&mhi, ws + L.whi_off, (uint64_t)L.d_pad, (uint64_t)L.k_pad, (uint64_t)L.d_pad * 4, BK, BNH))) return rc;
    if ((rc = tc::make_map_2d(&mlo, ws + L.wlo_off, (uint64_t)L.d_pad, (uint64_t)L.k_pad, (uint64_t)L.d_pad * 4, BK, BNH))) return rc;
    static bool attr_set[64] = {};
    if (tc::first_launch_on_device(attr_set)) {
        SOM_CUDA(cudaFuncSetAttribute(bmu_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        SOM_CUDA(cudaFuncSetAttribute(bmu_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    }
    const int num_pair_tiles = (int)ceil_div(n, 2 * BM);
    const int num_n_tiles = L.k_pad / BN;
    const int num_k_blocks = L.d_pad / BK;
    int pairs = sm_count / 2;
    if (pairs > num_pair_tiles) pairs = num_pair_tiles;
    if (pairs < 1) pairs = 1;
    FusedAcc acc;
    acc.X = X; acc.ldx = ldx; acc.d = d; acc.k = k; acc.S = T.S; acc.cnt = T.cnt; acc.qscale = T.qscale;
    acc.lds = acc_ld(d);
    acc.reps = T.reps; acc.rep_words = T.rep_words;
    acc.vec = (d % 4 == 0) ? 1 : 0;              // X rows are 16-byte aligned here (tc::shape_ok)
    static const int dbg = tc::env_int("SOM_B200_DBG");
    acc.dbg = dbg;
    static const int no_res = tc::env_int("SOM_B200_TC2_STREAM");     // A/B measurements: force the streaming variant
    const float *bias_p = reinterpret_cast<const float *>(ws + L.bias_off);
    const unsigned int *gstat_p = reinterpret_cast<const unsigned int *>(ws + L.gstat_off);
    const cudaError_t e = (num_k_blocks == 1 && !no_res)
        ? launch_pdl(bmu_tc2_kernel<true>, dim3(2 * pairs), dim3(NUM_THREADS), SMEM_BYTES, st, mx, mhi, mlo, bias_p, gstat_p, n,
                     num_pair_tiles, num_n_tiles, num_k_blocks, bmu, best, acc)
        : launch_pdl(bmu_tc2_kernel<false>, dim3(2 * pairs), dim3(NUM_THREADS), SMEM_BYTES, st, mx, mhi, mlo, bias_p, gstat_p, n,
                     num_pair_tiles, num_n_tiles, num_k_blocks, bmu, best, acc);
    return check_cuda(e, "bmu_tc2_kernel launch");
}

}  // namespace tc2
}  // namespace somb200
