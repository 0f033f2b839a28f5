// Shared building blocks of the tensor-core BMU kernels (bmu_tc2.cuh: kind::tf32, bmu_tc3.cuh:
// kind::f16): PTX wrappers for mbarrier / TMA / tcgen05 / TMEM, the shared-memory matrix descriptor,
// the running-argmin epilogue state, the fused-accumulate arguments and the TMA tensor-map helpers.
//
// Contract of those kernels.  They replace xp.dot(x, w.T) + w_sq.T + argmin of the reference
// (distances.py:22-23 / 55-59 and xpysom.py:416) for the Euclidean and cosine activation distances:
// score[r, k] = x_r . w'_k + bias_k is minimised over k, with w' = -2 w, bias = |w|^2 (euclidean) or
// w' = -w/|w|, bias = 0 (cosine), both prepared once per epoch (misc.cuh).  fp32 accuracy on
// 11-bit-significand tensor-core inputs comes from splitting every fp32 operand into hi + lo and
// accumulating lo*hi + hi*lo + hi*hi in the fp32 TMEM accumulator (the lo*lo term, ~2^-22
// relative, is dropped).  The (n, K) score matrix lives only in TMEM.
#pragma once
#include <cuda.h>
#include "common.cuh"

namespace somb200 {
namespace tc {

constexpr int BM = 128;     // sample rows per CTA (TMEM lanes)

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
constexpr long long kTrapCycles = 8000000000LL;     // ~4 s at 2 GHz
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity), "r"(kSuspendHintNs) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must trap (-> CUDA error on the host), never hang the GPU.  The clock is looked at
// only every 1024 failed polls: a sleeping waiter is woken by EVERY arrival on its barrier, and each wake-up costs
// issue slots that the epilogue warps need.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    uint32_t polls = 0;
    long long t0 = 0;
    for (;;) {
        if (mbar_try_wait(bar, parity) || mbar_try_wait(bar, parity) || mbar_try_wait(bar, parity) ||
            mbar_try_wait(bar, parity)) return;
        if ((++polls & 255u) == 0) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            if (now - t0 > kTrapCycles) {      // a protocol bug must trap quickly; profilers slow kernels down, hence seconds
                printf("som_b200: mbarrier timeout (block %d thread %d bar 0x%x parity %u)\n",
                       (int)blockIdx.x, (int)threadIdx.x, bar, parity);
                __trap();
            }
        }
    }
}
// Wait of a warp that is OFF the critical path (scatter warps, TMA producers running stages ahead): the suspend
// hint of try_wait is not honoured for long (ncu: one poll every ~50 cycles), so back off explicitly and leave
// the issue slots to the epilogue warps.
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity, uint32_t sleep_ns) {
    if (mbar_try_wait(bar, parity)) return;
    uint32_t polls = 0;
    long long t0 = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(sleep_ns);
        if ((++polls & 1023u) == 0) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            if (now - t0 > kTrapCycles) {
                printf("som_b200: mbarrier timeout (block %d thread %d bar 0x%x parity %u)\n",
                       (int)blockIdx.x, (int)threadIdx.x, bar, parity);
                __trap();
            }
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
// (instruction descriptors: D format at bit 4, A / B formats at bits 7 / 10, N >> 3 at bit 17, M >> 4 at
// bit 24 — kIdesc2 in bmu_tc2.cuh, kIdescF16 in bmu_tc3.cuh)

// Running argmin of the epilogue.  A single (best, index) pair would make every one of the K compares
// wait for the previous one (FSETP -> FSEL latency per element, ~2x the MMA time of a D=64 tile);
// EPI_ACC independent pairs over interleaved columns keep the ALU pipe busy instead.  Each pair sees
// its columns in increasing order with a strict <, the final lexicographic merge restores numpy's
// first-minimum rule.
constexpr int EPI_ACC = 4;
struct RunMin {
    float v[EPI_ACC];
    float fi[EPI_ACC];      // winning column kept as a float (exact: k_pad < 2^24, checked by the launchers)
    __device__ __forceinline__ void reset() {
#pragma unroll
        for (int a = 0; a < EPI_ACC; ++a) { v[a] = INFINITY; fi[a] = -1.f; }
    }
    // One candidate.  The kernels are bound by the ALU pipe (ncu: 80 % busy with the FMA pipe at 20 %), and a
    // compare + two selects per score would all be ALU work.  Only the compare stays there: the index is written
    // by a predicated FADD (base + immediate) and the value by a predicated FADD of -0.0 (an exact copy for every
    // input, -0.0 and denormals included), both on the FMA pipe.
    __device__ __forceinline__ void upd(int a, float sc, float basef, float jf) {
        asm("{\n\t.reg .pred p;\n\t"
            "setp.lt.f32 p, %2, %0;\n\t"
            "@p add.f32 %1, %3, %4;\n\t"
            "@p add.f32 %0, %2, 0f80000000;\n\t}"
            : "+f"(v[a]), "+f"(fi[a]) : "f"(sc), "f"(basef), "f"(jf));
    }
    // 32 accumulator columns (TMEM registers) + their bias (shared or global memory, 16-byte aligned)
    __device__ __forceinline__ void chunk(const uint32_t (&acc)[32], const float *bias32, int colbase) {
        const float4 *b4 = reinterpret_cast<const float4 *>(bias32);
        const float basef = (float)colbase;
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
                upd(j % EPI_ACC, __uint_as_float(acc[j]) + bb[e], basef, (float)j);
            }
        }
    }
    // bias already inside the accumulator (folded into the contraction): compare the raw accumulator
    __device__ __forceinline__ void chunk_nobias(const uint32_t (&acc)[32], int colbase) {
        const float basef = (float)colbase;
#pragma unroll
        for (int j = 0; j < 32; ++j) upd(j % EPI_ACC, __uint_as_float(acc[j]), basef, (float)j);
    }
    __device__ __forceinline__ void result(float &best, int &bidx) const {
        best = v[0]; bidx = fi[0] < 0.f ? 0x7fffffff : (int)fi[0];
#pragma unroll
        for (int a = 1; a < EPI_ACC; ++a) argmin_merge(best, bidx, v[a], fi[a] < 0.f ? 0x7fffffff : (int)fi[a]);
        if (bidx == 0x7fffffff) bidx = 0;      // every score was +inf / NaN: numpy's argmin would say 0
    }
};

// Experiment knobs (SOM_B200_DBG: timeline probe of tools/bmu_probe.py; SOM_B200_TBN: neuron-tile width of the
// resident fp16 kernel; ...) exist only in a library built with -DSOM_B200_EXPERIMENTS (python -m
// xpysom_dask_b200.build --experiments); the production library never looks at the environment.
inline int env_int(const char *name) {
#ifdef SOM_B200_EXPERIMENTS
    const char *e = getenv(name);
    return e ? atoi(e) : 0;
#else
    (void)name;
    return 0;
#endif
}
// cudaFuncSetAttribute is per device: remember which devices have the large dynamic-shared-memory opt-in.
inline bool first_launch_on_device(bool (&done)[64]) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return true;
    if (done[dev]) return false;
    done[dev] = true;
    return true;
}

// One lane of a converged warp.  The whole warp runs the producer / MMA loops (so every address stays in uniform
// registers -- no per-instruction R2UR waterfall) and only the async instructions sit under the election.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
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

// arguments of the fused accumulate (common.cuh: ExactAcc; S == nullptr turns it off) plus the experiment probe
typedef ExactAcc FusedAcc;
constexpr int SCAT_NBUF = 4;                                        // staging buffers per scatter warp
constexpr int SCAT_STAGE_BYTES = 4 * SCAT_NBUF * ACC_PIECE * 8;     // four scatter warps

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
                       uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeTiledFn enc = get_encode_fn();
    SOM_REQUIRE(enc != nullptr, SOM_E_NODEVICE, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {row_stride_bytes};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void *>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SOM_REQUIRE(r == CUDA_SUCCESS, SOM_E_SHAPE, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return 0;
}

inline bool shape_ok(const float *X, int64_t n, int d, int64_t ldx) {
    return n > 0 && d >= 1 && (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0) &&
           n < ((int64_t)1 << 31) - BM;
}

}  // namespace tc
}  // namespace somb200
