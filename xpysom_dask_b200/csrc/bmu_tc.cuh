// K1: tensor-core distance contraction with a fused BMU argmin (sm_100a).
//
// Replaces xp.dot(x, w.T) + w_sq.T + argmin of the reference
// (distances.py:22-23 / 55-59 and xpysom.py:416) for the Euclidean and cosine
// activation distances.  score[r, k] = x_r . w'_k + bias_k is minimised over k,
// with w' = -2 w, bias = |w|^2 (euclidean) or w' = -w/|w|, bias = 0 (cosine),
// both prepared once per epoch by prepare_codebook_kernel.
//
// fp32 accuracy from TF32 tensor cores: every fp32 operand is split into
// hi = rna_tf32(v) and lo = rna_tf32(v - hi) and three MMAs accumulate
// lo*hi + hi*lo + hi*hi into the same fp32 TMEM accumulator (the lo*lo term,
// ~2^-22 relative, is dropped).  W is split by the prepare kernel; X is split
// inside this kernel, in shared memory, by a converter warpgroup, so X is read
// from HBM exactly once and never re-written.
//
// Pipeline (one CTA per SM, persistent over 128-row tiles of X):
//   warp 0      TMA producer   X chunk [128 x 32] + W'hi/W'lo chunks [256 x 32], SWIZZLE_128B
//   warp 1      MMA issuer     tcgen05.mma.kind::tf32, M=128 N=256 K=8, accumulators in TMEM (2 x 256 cols)
//   warps 2-5   converter      X chunk -> hi (in place) and lo, generic->async proxy fence
//   warps 6-9   epilogue       tcgen05.ld 32 columns at a time, + bias, running (min, argmin) per row;
//                              then (fused K3) S[bmu] += x for the tile's 128 rows, re-read from L2
// The (n, K) score matrix lives only in TMEM.
//
// Fused accumulate (acc.S != nullptr): once a 128-row tile has its BMUs, the epilogue warps add the
// tile's rows into the (K, D) accumulator with red.global.add.v4.f32 — the rows were just streamed
// through L2 by TMA, so HBM sees X once per epoch — and bump exact int32 counts; the last CTA to
// finish folds the counts into the fp32 c vector (fp32 increments of 1 are lost above 2^24).
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace somb200 {
namespace tc {

constexpr int BM = 128, BN = 256, BK = 32, STAGES = 2, UMMA_K = 8;
constexpr int A_BYTES = BM * BK * 4;   // 16 KB
constexpr int B_BYTES = BN * BK * 4;   // 32 KB
constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;   // A_hi | A_lo | B_hi | B_lo = 96 KB
constexpr int NUM_THREADS = 320;
constexpr int CONV_WARP0 = 2, EPI_WARP0 = 6;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 2 * BN * 4 /*bias tiles*/ + 256 /*barriers*/ + 1024 /*align slack*/;

// ---- PTX wrappers ------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes (prompt
// wake-up) or the hint expires, instead of burning issue slots that the epilogue warps need
constexpr uint32_t kSuspendHintNs = 20000;
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity), "r"(kSuspendHintNs) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must trap (-> CUDA error on the host), never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {   // ~2 s at 2 GHz
            printf("som_b200: mbarrier timeout (block %d thread %d bar 0x%x parity %u)\n",
                   (int)blockIdx.x, (int)threadIdx.x, bar, parity);
            __trap();
        }
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" :: "l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(slot_smem), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "setp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                   "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                   "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait_dep16(uint32_t (&v)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
                 :: "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// wait::ld that also "touches" the destination registers, so the compiler cannot schedule their
// consumers above the wait
__device__ __forceinline__ void tmem_ld_wait_dep(uint32_t (&v)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]),
                   "+r"(v[8]), "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]),
                   "+r"(v[16]), "+r"(v[17]), "+r"(v[18]), "+r"(v[19]), "+r"(v[20]), "+r"(v[21]), "+r"(v[22]), "+r"(v[23]),
                   "+r"(v[24]), "+r"(v[25]), "+r"(v[26]), "+r"(v[27]), "+r"(v[28]), "+r"(v[29]), "+r"(v[30]), "+r"(v[31])
                 :: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 format):
//   [0,14)  start address >> 4      [16,30) leading byte offset >> 4 (ignored for swizzled K-major; 1)
//   [32,46) stride byte offset >> 4 (8 rows x 128 B = 1024 B between 8-row groups)
//   [46,48) version = 1             [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// instruction descriptor: D=f32 (bit4), A=B=tf32 (2 at bits 7 and 10), K-major both, N>>3 at 17, M>>4 at 24
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

// Running argmin of the epilogue.  A single (best, index) pair would make every one of the K compares
// wait for the previous one (FSETP -> FSEL latency per element, ~2x the MMA time of a D=64 tile);
// EPI_ACC independent pairs over interleaved columns keep the ALU pipe busy instead.  Each pair sees
// its columns in increasing order with a strict <, the final lexicographic merge restores numpy's
// first-minimum rule.
constexpr int EPI_ACC = 8;
struct RunMin {
    float v[EPI_ACC];
    int   i[EPI_ACC];
    __device__ __forceinline__ void reset() {
#pragma unroll
        for (int a = 0; a < EPI_ACC; ++a) { v[a] = INFINITY; i[a] = 0x7fffffff; }
    }
    // 32 accumulator columns (TMEM registers) + their bias (shared or global memory, 16-byte aligned)
    __device__ __forceinline__ void chunk(const uint32_t (&acc)[32], const float *bias32, int colbase) {
        const float4 *b4 = reinterpret_cast<const float4 *>(bias32);
        float4 bq[8];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) bq[j4] = b4[j4];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
            const float4 b = bq[j4];
            const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int j = j4 * 4 + e;
                const float sc = __uint_as_float(acc[j]) + bb[e];
                const int a = j % EPI_ACC;
                if (sc < v[a]) { v[a] = sc; i[a] = colbase + j; }
            }
        }
    }
    __device__ __forceinline__ void result(float &best, int &bidx) const {
        best = v[0]; bidx = i[0];
#pragma unroll
        for (int a = 1; a < EPI_ACC; ++a) argmin_merge(best, bidx, v[a], i[a]);
        if (bidx == 0x7fffffff) bidx = 0;      // every score was +inf / NaN: numpy's argmin would say 0
    }
};

// Drain one accumulator tile (NCHUNK x 32 columns) into the running argmin.  TMEM loads are double
// buffered in registers: the load of chunk c+1 is in flight while chunk c is compared.
template <int NCHUNK>
__device__ __forceinline__ void drain_accumulator(RunMin &rm, uint32_t taddr, const float *bs, int colbase) {
    uint32_t va[32], vb[32];
    tmem_ld32(taddr, va);
#pragma unroll 1
    for (int c = 0; c < NCHUNK; c += 2) {
        tmem_ld_wait_dep(va);
        tmem_ld32(taddr + (c + 1) * 32, vb);
        rm.chunk(va, bs + c * 32, colbase + c * 32);
        tmem_ld_wait_dep(vb);
        if (c + 2 < NCHUNK) tmem_ld32(taddr + (c + 2) * 32, va);
        rm.chunk(vb, bs + (c + 1) * 32, colbase + (c + 1) * 32);
    }
}

__device__ __forceinline__ float tf32_rna_dev(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

// Timeline probe for experiments (SOM_B200_DBG=9): the leader CTA of pair 0 stamps clock64() at the
// hand-off points of its first tiles.  record(slot, tile) writes g_dbg[tile * 8 + slot].
__device__ long long g_dbg[8 * 256];
__device__ __forceinline__ void dbg_stamp(int on, int slot, uint32_t tile) {
    if (on && tile < 256) g_dbg[tile * 8 + slot] = clock64();
}

// arguments of the fused accumulate; S == nullptr turns it off
struct FusedAcc {
    const float *X;      // samples (row stride ldx), the same matrix map_x describes
    int64_t ldx;
    int d, k;
    float *S, *c;        // (K, D) sums and (K) counts, accumulated into
    int *cnt;            // k ints, zero on entry, zero again on exit
    unsigned int *done;  // ticket counter, zero on entry and exit
    int vec;             // rows are 16-byte aligned and d % 4 == 0: 128-bit path
    int dbg;             // experiments only (SOM_B200_DBG); 0 in production
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
bmu_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_whi,
              const __grid_constant__ CUtensorMap map_wlo, const float *__restrict__ bias,
              int64_t n, int num_m_tiles, int num_n_tiles, int num_k_blocks,
              int32_t *__restrict__ bmu_out, float *__restrict__ best_out, const FusedAcc acc) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;   // SWIZZLE_128B needs 1024-byte alignment
    uint8_t *smem = smem_raw + (smem_base - smem_u32(smem_raw));

    float    *bias_s = reinterpret_cast<float *>(smem + STAGES * STAGE_BYTES);          // [2][BN]
    uint64_t *bars   = reinterpret_cast<uint64_t *>(smem + STAGES * STAGE_BYTES + 2 * BN * 4);
    // barrier slots: full[S], ready[S], empty[S], tmem_full[2], tmem_empty[2], then the TMEM base slot
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar  = [&](int s) { return bar0 + 8u * s; };
    auto ready_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
    auto empty_bar = [&](int s) { return bar0 + 8u * (2 * STAGES + s); };
    auto tfull_bar = [&](int a) { return bar0 + 8u * (3 * STAGES + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (3 * STAGES + 2 + a); };
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(bars + 3 * STAGES + 4);
    __shared__ int bmu_s[BM];          // BMUs of the current tile, shared among the epilogue warps
    __shared__ unsigned int last_cta;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(full_bar(s), 1); mbar_init(ready_bar(s), 128); mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(tfull_bar(a), 1); mbar_init(tempty_bar(a), 128); }
        fence_barrier_init();
        tma_prefetch_desc(&map_x); tma_prefetch_desc(&map_whi); tma_prefetch_desc(&map_wlo);
    }
    if (warp == 1) tmem_alloc(smem_u32((const void *)tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            uint32_t it = 0;
            for (int mt = blockIdx.x; mt < num_m_tiles; mt += gridDim.x)
                for (int nt = 0; nt < num_n_tiles; ++nt)
                    for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
                        const int s = it % STAGES; const uint32_t ph = (it / STAGES) & 1;
                        mbar_wait(empty_bar(s), ph ^ 1);
                        const uint32_t st = smem_base + s * STAGE_BYTES;
                        mbar_expect_tx(full_bar(s), A_BYTES + 2 * B_BYTES);
                        tma_load_2d(st,                         &map_x,   kb * BK, mt * BM, full_bar(s));
                        tma_load_2d(st + 2 * A_BYTES,           &map_whi, kb * BK, nt * BN, full_bar(s));
                        tma_load_2d(st + 2 * A_BYTES + B_BYTES, &map_wlo, kb * BK, nt * BN, full_bar(s));
                    }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            uint32_t it = 0, acc_it = 0;
            for (int mt = blockIdx.x; mt < num_m_tiles; mt += gridDim.x)
                for (int nt = 0; nt < num_n_tiles; ++nt, ++acc_it) {
                    const int a = acc_it & 1; const uint32_t aph = (acc_it >> 1) & 1;
                    mbar_wait(tempty_bar(a), aph ^ 1);      // epilogue has drained this accumulator
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + (uint32_t)(a * BN);
                    for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
                        const int s = it % STAGES; const uint32_t ph = (it / STAGES) & 1;
                        mbar_wait(full_bar(s), ph);          // W' tiles landed (async proxy)
                        mbar_wait(ready_bar(s), ph);         // X hi/lo written by the converter
                        tc_fence_after();
                        const uint32_t st = smem_base + s * STAGE_BYTES;
                        const uint64_t a_hi = make_smem_desc(st), a_lo = make_smem_desc(st + A_BYTES);
                        const uint64_t b_hi = make_smem_desc(st + 2 * A_BYTES), b_lo = make_smem_desc(st + 2 * A_BYTES + B_BYTES);
#pragma unroll
                        for (int kk = 0; kk < BK / UMMA_K; ++kk) {
                            const uint64_t off = (uint64_t)((kk * UMMA_K * 4) >> 4);   // +32 B along K inside the swizzle atom
                            umma_tf32(tmem_d, a_lo + off, b_hi + off, kIdesc, (kb | kk) != 0);
                            umma_tf32(tmem_d, a_hi + off, b_lo + off, kIdesc, 1);
                            umma_tf32(tmem_d, a_hi + off, b_hi + off, kIdesc, 1);
                        }
                        umma_commit(empty_bar(s));           // stage reusable once these MMAs retire
                    }
                    umma_commit(tfull_bar(a));               // accumulator complete -> epilogue
                }
        }
    } else if (warp < EPI_WARP0) {
        // ===================== converter: split X into TF32 hi / lo =====================
        const int t = threadIdx.x - CONV_WARP0 * 32;   // 0..127
        uint32_t it = 0;
        for (int mt = blockIdx.x; mt < num_m_tiles; mt += gridDim.x)
            for (int nt = 0; nt < num_n_tiles; ++nt)
                for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
                    const int s = it % STAGES; const uint32_t ph = (it / STAGES) & 1;
                    mbar_wait(full_bar(s), ph);
                    float4 *ahi = reinterpret_cast<float4 *>(smem + s * STAGE_BYTES);
                    float4 *alo = reinterpret_cast<float4 *>(smem + s * STAGE_BYTES + A_BYTES);
#pragma unroll
                    for (int i = 0; i < A_BYTES / 16 / 128; ++i) {
                        const int e = t + 128 * i;          // position-preserving, so the swizzle is irrelevant here
                        const float4 v = ahi[e];
                        float4 h, l;
                        h.x = tf32_rna_dev(v.x); h.y = tf32_rna_dev(v.y); h.z = tf32_rna_dev(v.z); h.w = tf32_rna_dev(v.w);
                        l.x = tf32_rna_dev(v.x - h.x); l.y = tf32_rna_dev(v.y - h.y);
                        l.z = tf32_rna_dev(v.z - h.z); l.w = tf32_rna_dev(v.w - h.w);
                        ahi[e] = h; alo[e] = l;
                    }
                    fence_proxy_async();                     // generic-proxy writes -> visible to the tensor core
                    mbar_arrive(ready_bar(s));
                }
    } else {
        // ===================== epilogue: TMEM -> registers -> running argmin =====================
        const int q = warp & 3;                              // TMEM lane quarter this warp may access
        const int e = threadIdx.x - EPI_WARP0 * 32;          // 0..127
        const int row_in_tile = q * 32 + lane;
        uint32_t acc_it = 0;
        for (int mt = blockIdx.x; mt < num_m_tiles; mt += gridDim.x) {
            RunMin rm; rm.reset();
            for (int nt = 0; nt < num_n_tiles; ++nt, ++acc_it) {
                const int a = acc_it & 1; const uint32_t aph = (acc_it >> 1) & 1;
                float *bs = bias_s + a * BN;
                bs[e] = __ldg(bias + (int64_t)nt * BN + e);
                bs[e + 128] = __ldg(bias + (int64_t)nt * BN + e + 128);
                asm volatile("bar.sync 1, 128;" ::: "memory");   // bias tile visible to the 4 epilogue warps
                mbar_wait(tfull_bar(a), aph);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BN);
                drain_accumulator<BN / 32>(rm, taddr, bs, nt * BN);
                tc_fence_before();
                mbar_arrive(tempty_bar(a));
            }
            float best; int bidx;
            rm.result(best, bidx);
            const int64_t row = (int64_t)mt * BM + row_in_tile;
            if (row < n) {
                if (bmu_out) bmu_out[row] = bidx;
                if (best_out) best_out[row] = best;
            }
            if (acc.S != nullptr) {
                // ---- fused K3: S[bmu[r], :] += X[r, :] for the rows of this tile -----------------
                bmu_s[row_in_tile] = (row < n) ? bidx : -1;
                if (row < n) atomicAdd(acc.cnt + bidx, 1);
                asm volatile("bar.sync 2, 128;" ::: "memory");
                const int64_t row0 = (int64_t)mt * BM;
                const int wq = warp - EPI_WARP0;                 // this warp takes rows wq, wq+4, ...
                if (acc.vec) {
                    const int d4 = acc.d >> 2;
                    // 32 lanes cover `rows_per_pass` rows x d4 float4 at a time (d4 < 32), or one row in strips
                    if (d4 <= 32) {
                        const int lanes_per_row = d4 <= 1 ? 1 : d4 <= 2 ? 2 : d4 <= 4 ? 4 : d4 <= 8 ? 8 : d4 <= 16 ? 16 : 32;
                        const int rows_per_pass = 32 / lanes_per_row;
                        const int sub = lane / lanes_per_row, c4 = lane % lanes_per_row;
                        for (int r = wq * rows_per_pass + sub; r < BM; r += 4 * rows_per_pass) {
                            const int b = bmu_s[r];
                            if (b >= 0 && c4 < d4) {
                                const float4 v = __ldg(reinterpret_cast<const float4 *>(acc.X + (row0 + r) * acc.ldx) + c4);
                                red_add_v4(acc.S + (int64_t)b * acc.d + c4 * 4, v);
                            }
                        }
                    } else {
                        for (int r = wq; r < BM; r += 4) {
                            const int b = bmu_s[r];
                            if (b < 0) continue;
                            const float4 *xr = reinterpret_cast<const float4 *>(acc.X + (row0 + r) * acc.ldx);
                            float *sr = acc.S + (int64_t)b * acc.d;
                            for (int c4 = lane; c4 < d4; c4 += 32) red_add_v4(sr + c4 * 4, __ldg(xr + c4));
                        }
                    }
                } else {
                    for (int r = wq; r < BM; r += 4) {
                        const int b = bmu_s[r];
                        if (b < 0) continue;
                        for (int cc = lane; cc < acc.d; cc += 32)
                            atomicAdd(acc.S + (int64_t)b * acc.d + cc, __ldg(acc.X + (row0 + r) * acc.ldx + cc));
                    }
                }
                asm volatile("bar.sync 2, 128;" ::: "memory");   // bmu_s is rewritten by the next tile
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }

    if (acc.S != nullptr) {
        // the last CTA to get here folds the exact integer counts into the fp32 vector c
        if (threadIdx.x == 0) {
            __threadfence();
            last_cta = (atomicAdd(acc.done, 1u) == gridDim.x - 1) ? 1u : 0u;
        }
        __syncthreads();
        if (last_cta) {
            __threadfence();
            for (int i = threadIdx.x; i < acc.k; i += blockDim.x) {
                const int v = atomicExch(acc.cnt + i, 0);
                if (v) atomicAdd(acc.c + i, (float)v);
            }
            if (threadIdx.x == 0) *acc.done = 0u;
        }
    }
}

// ---- host side -----------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

inline int make_map_2d(CUtensorMap *m, const void *base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                       uint32_t box_inner, uint32_t box_outer) {
    EncodeTiledFn enc = get_encode_fn();
    SOM_REQUIRE(enc != nullptr, SOM_E_NODEVICE, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {row_stride_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SOM_REQUIRE(r == CUDA_SUCCESS, SOM_E_SHAPE, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 0;
}

inline bool shape_ok(const float *X, int64_t n, int d, int64_t ldx) {
    return n > 0 && d >= 1 && (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0) &&
           n < ((int64_t)1 << 31) - BM;
}

inline int launch_bmu_tc(const float *X, int64_t n, int d, int64_t ldx, int k, const WsLayout &L, uint8_t *ws,
                         int32_t *bmu, float *best, float *S, float *c, int sm_count, cudaStream_t st) {
    SOM_REQUIRE(shape_ok(X, n, d, ldx), SOM_E_SHAPE,
                "tensor-core BMU kernel needs ldx %% 4 == 0 and a 16-byte aligned X for TMA (d=%d ldx=%lld)", d, (long long)ldx);
    CUtensorMap mx, mhi, mlo;
    int rc;
    if ((rc = make_map_2d(&mx, X, (uint64_t)d, (uint64_t)n, (uint64_t)ldx * 4, BK, BM))) return rc;
    if ((rc = make_map_2d(&mhi, ws + L.whi_off, (uint64_t)L.d_pad, (uint64_t)L.k_pad, (uint64_t)L.d_pad * 4, BK, BN))) return rc;
    if ((rc = make_map_2d(&mlo, ws + L.wlo_off, (uint64_t)L.d_pad, (uint64_t)L.k_pad, (uint64_t)L.d_pad * 4, BK, BN))) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        SOM_CUDA(cudaFuncSetAttribute(bmu_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
        attr_set = true;
    }
    const int num_m_tiles = (int)ceil_div(n, BM);
    const int num_n_tiles = L.k_pad / BN;
    const int num_k_blocks = L.d_pad / BK;
    const int grid = num_m_tiles < sm_count ? num_m_tiles : sm_count;
    FusedAcc acc;
    acc.X = X; acc.ldx = ldx; acc.d = d; acc.k = k; acc.S = S; acc.c = c;
    acc.cnt = reinterpret_cast<int *>(ws + L.cnt_off);
    acc.done = reinterpret_cast<unsigned int *>(ws + L.done_off);
    acc.vec = (d % 4 == 0) && S != nullptr && ((reinterpret_cast<uintptr_t>(S) & 15) == 0);
    acc.dbg = 0;
    bmu_tc_kernel<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(mx, mhi, mlo, reinterpret_cast<const float *>(ws + L.bias_off),
                                                         n, num_m_tiles, num_n_tiles, num_k_blocks, bmu, best, acc);
    return check_cuda(cudaGetLastError(), "bmu_tc_kernel launch");
}

}  // namespace tc
}  // namespace somb200
