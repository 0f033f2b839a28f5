// Shared helpers for the B200 batch-SOM kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <math.h>
#include "../../include/som_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libsom_b200 is written for sm_100a (B200) only"
#endif

namespace somb200 {

constexpr int kKPad = 256;   // neuron padding of the prepared codebook (one UMMA N tile)
constexpr int kDPad = 32;    // feature padding: one 128-byte swizzle row of fp32

__host__ __device__ inline int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }
__host__ __device__ inline int64_t ceil_div(int64_t v, int64_t m) { return (v + m - 1) / m; }

// Workspace carved out of the caller's buffer by prepare_codebook and read by the BMU kernels.
struct WsLayout {
    int    k_pad, d_pad;
    size_t aux_off;   // k_pad floats: |w|^2 (euclidean) / 1/|w| (cosine, SIMT) / bias (tensor-core path)
    size_t bias_off;  // k_pad floats: additive bias of the tensor-core epilogue (+inf on padding)
    size_t whi_off;   // k_pad x d_pad floats: TF32 "hi" part of the scaled codebook
    size_t wlo_off;   // k_pad x d_pad floats: TF32 "lo" part
    int    d_pad64;   // feature padding of the fp16 operand copies: one 128-byte swizzle row of halves
    size_t w16hi_off; // k_pad x d_pad64 halves: fp16 "hi" part of the scaled codebook times 2^b_k
    size_t w16lo_off; // k_pad x d_pad64 halves: fp16 "lo" part
    size_t wsinv_off; // k_pad floats: 2^-b_k, the inverse of the per-neuron power-of-two scale
    size_t wfold_off; // k_pad x 8 floats, stored as the shared-memory IMAGE of the fold operand of the fp16 kernel: per
                      // 128 neurons one 4 KB block in the no-swizzle K-major core-matrix layout (bmu_tc3.cuh)
    size_t cnt_off;   // k_pad int32: spare
    size_t done_off;  // the grid barrier of epoch_tail_kernel lives at +256 (arrivals) and +320 (generation); zeroed by
                      // prepare_codebook
    size_t gstat_off; // codebook statistics of the current prepare: [0] bits of max_k amax_k, [1] ~bits of the
                      // smallest non-zero amax_k (both via atomicMax, zeroed by prepare), [2] uniform-scale flag
    size_t amax_off;  // k_pad floats: amax_k = max_c |w'_k[c]|
    size_t total;
};

__host__ inline WsLayout ws_layout(int k, int d) {
    WsLayout L;
    L.k_pad = (int)round_up(k, kKPad);
    L.d_pad = (int)round_up(d, kDPad);
    size_t off = 0;
    L.aux_off = off;  off += round_up((size_t)L.k_pad * 4, 1024);
    L.bias_off = off; off += round_up((size_t)L.k_pad * 4, 1024);
    L.whi_off = off;  off += round_up((size_t)L.k_pad * L.d_pad * 4, 1024);
    L.wlo_off = off;  off += round_up((size_t)L.k_pad * L.d_pad * 4, 1024);
    L.d_pad64 = (int)round_up(d, 64);
    L.w16hi_off = off; off += round_up((size_t)L.k_pad * L.d_pad64 * 2, 1024);
    L.w16lo_off = off; off += round_up((size_t)L.k_pad * L.d_pad64 * 2, 1024);
    L.wsinv_off = off; off += round_up((size_t)L.k_pad * 4, 1024);
    L.wfold_off = off; off += round_up((size_t)L.k_pad * 32, 1024);
    L.cnt_off = off;  off += round_up((size_t)L.k_pad * 4, 1024);
    L.done_off = off; off += 1024;
    L.gstat_off = off; off += 1024;
    L.amax_off = off; off += round_up((size_t)L.k_pad * 4, 1024);
    L.total = off;
    return L;
}

// ---- error plumbing (host) -------------------------------------------------
void set_error(const char *fmt, ...);
int  check_cuda(cudaError_t e, const char *what);

#define SOM_CUDA(call)                                                 \
    do {                                                               \
        int _rc = ::somb200::check_cuda((call), #call);                \
        if (_rc) return _rc;                                           \
    } while (0)

#define SOM_REQUIRE(cond, code, ...)                                   \
    do {                                                               \
        if (!(cond)) { ::somb200::set_error(__VA_ARGS__); return (code); } \
    } while (0)

// ---- device helpers -------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// 16-byte vector reduction into global memory (sm_90+): one L2 atomic transaction for 4 floats.
__device__ __forceinline__ void red_add_v4(float *addr, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// ---- exact, order-independent accumulation of the per-BMU sums ---------------------------------------------
// S[b, c] = sum over the samples whose BMU is b of x[c] is kept as a 64-bit FIXED-POINT integer: every element is
// converted as rint(x * 2^q_c) with one power-of-two scale per feature column (q_c = 62 - (e_c + 1) - ceil(log2 n_total),
// e_c the exponent of the column's largest magnitude, so n_total samples can never overflow 63 bits), and integer
// additions are associative: the sum does not depend on the order in which the GPU's atomics arrive, on the tiling,
// or on how the samples are sharded over GPUs -- two runs are bit-identical and N ranks give the bits of one rank.
// An element keeps all 24 bits of its significand as long as it is within 2^14 of its column's maximum; smaller
// ones are rounded at 2^-38 of the column maximum (for n_total = 2^24), far below fp32 resolution of the sum.
// Rows travel to the accumulator as TMA bulk reductions (cp.reduce.async.bulk ... .add.u64) of up to 128 columns
// staged as int64 in shared memory: measured on B200 (tools/micro/red_bw.cu) 0.90 elements/clk/SM at D = 64 against
// 1.16 for red.global.add.v4.f32 and 0.60 for scalar 64-bit atomics.
constexpr int ACC_PIECE = 128;            // columns per bulk reduction (1 KB of int64)

__host__ __device__ inline int acc_ld(int d) { return (d + 1) & ~1; }   // row stride of S in int64: 16-byte rows

__device__ __forceinline__ void bulk_reduce_add_u64(unsigned long long *dst, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.u64 [%0], [%1], %2;"
                 :: "l"(dst), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void red_add_u64(unsigned long long *addr, long long v) {
    asm volatile("red.global.add.u64 [%0], %1;" :: "l"(addr), "l"(v) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// what a scatter warp needs: the samples, the column scales and the accumulator
struct ExactAcc {
    const float *X;                 // samples (row stride ldx)
    int64_t ldx;
    int d, k;
    const float *qscale;            // d floats: 2^q_c
    unsigned long long *S;          // (k, lds) fixed-point sums, two's complement; nullptr turns the accumulate off
    unsigned long long *cnt;        // (k) counts
    int lds;
    int vec;                        // rows of X are 16-byte aligned and d % 4 == 0: 128-bit loads
    int dbg;                        // experiments builds only (timeline probes); 0 in production
    int reps;                       // replicas of [S | cnt], rep_words apart: CTA i adds into replica i % reps (see acc_replicas)
    size_t rep_words;
    // this CTA's replica
    __device__ __forceinline__ ExactAcc for_cta(unsigned cta) const {
        ExactAcc a = *this;
        if (reps > 1 && S != nullptr) { const size_t off = (size_t)(cta % (unsigned)reps) * rep_words; a.S += off; a.cnt += off; }
        return a;
    }
};
#ifdef SOM_B200_EXPERIMENTS
__device__ long long g_scat_dbg[8 * 64];
#define SCAT_STAMP(slot) do { if (A.dbg == 8 && blockIdx.x == 0 && first == 0 && lane == 0 && sidx < 64) g_scat_dbg[sidx * 8 + (slot)] = clock64(); } while (0)
#else
#define SCAT_STAMP(slot) do { } while (0)
#endif

// host side: where a launch accumulates (S == nullptr: BMU search only)
struct AccTarget {
    unsigned long long *S = nullptr, *cnt = nullptr;
    const float *qscale = nullptr;
    int reps = 1;
    size_t rep_words = 0;
};

// Replicas of a small accumulator.  Samples that share a BMU add into the same rows of S, and reductions on one L2 line
// serialise: with most rows on a few "hot" neurons (blob data on a young map) the bulk reductions ran at half their
// rate (tools/micro/red_bw.cu, 90 % of the rows on 16 neurons: 0.352 ms against 0.242 ms at config 2; with 4 replicas
// 0.242 ms again).  So accumulators of up to 4 MB are kept 4 times, CTAs spread over the copies, and the finalize
// phase sums them -- integer sums, so the result is the same bits whatever the number of replicas.
__host__ __device__ inline size_t acc_words_one(int k, int d) { return ((size_t)k * acc_ld(d) + (size_t)k + 1) & ~(size_t)1; }   // 16-byte replicas
__host__ inline int acc_replicas(int k, int d) { return acc_words_one(k, d) * 8 <= ((size_t)4 << 20) ? 4 : 1; }

// One warp scatters rows of a tile whose BMUs sit in shared memory (bm[r] < 0: no such row).  The warp takes the
// rows r = first, first + stride, ... below `rows`; `stage` is THIS warp's staging area of NBUF * ACC_PIECE int64
// (128-byte aligned); `it` counts the bulk groups this warp has committed so far (buffer rotation).
// Rows of up to 128 columns: a group of g = pow2 >= d/4 lanes owns a row, 32/g rows per pass share one staging
// buffer, every row is ONE bulk reduction.  Longer rows: the whole warp walks a row in 128-column pieces.
template <int NBUF>
__device__ __forceinline__ void scatter_rows_exact(const ExactAcc &A, const int *bm, int64_t row0, int rows, int first,
                                                   int stride, int lane, long long *stage, uint32_t &it) {
    const int d = A.d, lds = A.lds;
    auto buffer = [&](uint32_t i) -> long long * { return stage + (i % NBUF) * ACC_PIECE; };
    if (d <= ACC_PIECE) {
        // One pass = 32 / g rows staged in one buffer and sent as 32 / g bulk reductions; NBUF passes form a batch: all
        // their global loads are issued first (the X rows come from L2, ~600 cycles away, and a scatter warp has
        // nothing else to overlap them with), then the buffers -- read out by the previous batch's reductions by now --
        // are filled, fenced ONCE (the generic -> async proxy fence and the waits cost ~1000 cycles per round, which is
        // what a batch amortises: one pass per round measured 0.49 ms at config 2, four 0.37 ms) and sent.
        const int d4 = (d + 3) >> 2;
        const int g = d4 <= 1 ? 1 : d4 <= 2 ? 2 : d4 <= 4 ? 4 : d4 <= 8 ? 8 : d4 <= 16 ? 16 : 32;
        const int G = 32 / g, sub = lane / g, c = (lane % g) * 4;
        const bool col_ok = c < lds;
        float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
        if (col_ok) {
            if (A.vec) q = __ldg(reinterpret_cast<const float4 *>(A.qscale + c));
            else { q.x = __ldg(A.qscale + c); if (c + 1 < d) q.y = __ldg(A.qscale + c + 1);
                   if (c + 2 < d) q.z = __ldg(A.qscale + c + 2); if (c + 3 < d) q.w = __ldg(A.qscale + c + 3); }
        }
        // (written without branches: a scatter warp runs alone on its scheduler slot, so every dependent
        // branch / address chain is exposed latency -- the first version spent ~900 cycles per phase and batch)
        const float *xbase = A.X + row0 * A.ldx + (col_ok ? c : 0);                 // row r of the tile: xbase + r * ldx
        const int ldx32 = (int)A.ldx;
        auto load_batch = [&](int r0, int (&bb)[NBUF], float4 (&v)[NBUF]) {
            int rr[NBUF];
#pragma unroll
            for (int u = 0; u < NBUF; ++u) {
                const int r = r0 + u * stride * G + sub;
                const bool in = r < rows;
                rr[u] = in ? r : 0;
                bb[u] = bm[rr[u]];
                bb[u] = in ? bb[u] : -1;
            }
#pragma unroll
            for (int u = 0; u < NBUF; ++u) {
                // rows that do not exist (bb < 0: past the end of the samples) read row 0 of the tile instead: never sent
                const float *xr = xbase + (size_t)((bb[u] >= 0 ? rr[u] : 0) * ldx32);
                if (A.vec) v[u] = __ldcg(reinterpret_cast<const float4 *>(xr));
                else { v[u].x = __ldcg(xr); v[u].y = c + 1 < d ? __ldcg(xr + 1) : 0.f;
                       v[u].z = c + 2 < d ? __ldcg(xr + 2) : 0.f; v[u].w = c + 3 < d ? __ldcg(xr + 3) : 0.f; }
            }
        };
        const int step = stride * G * NBUF;
        int bb[NBUF], bbn[NBUF];
        float4 v[NBUF], vn[NBUF];
        int r0 = first * G;
        load_batch(r0, bb, v);
        long long *slotp[NBUF];
        uint32_t slota[NBUF];
#pragma unroll
        for (int u = 0; u < NBUF; ++u) {
            slotp[u] = buffer(u) + sub * (4 * g);
            slota[u] = (uint32_t)__cvta_generic_to_shared(slotp[u]);
        }
        for (; r0 < rows; r0 += step) {                                              // same trip count for every lane
            const uint32_t sidx = it; (void)sidx;
            SCAT_STAMP(0);
            load_batch(r0 + step < rows ? r0 + step : r0, bbn, vn);    // the next batch's rows are in flight while this one is sent
            SCAT_STAMP(1);
            bulk_wait_read<0>();                 // the previous batch's bulk reductions have read the staging buffers
            __syncwarp();
            SCAT_STAMP(2);
#pragma unroll
            for (int u = 0; u < NBUF; ++u) {
                if (col_ok) {                    // (rows that do not exist are converted too: their buffer is never sent)
                    longlong2 lo, hi;
                    lo.x = __float2ll_rn(v[u].x * q.x); lo.y = __float2ll_rn(v[u].y * q.y);
                    hi.x = __float2ll_rn(v[u].z * q.z); hi.y = __float2ll_rn(v[u].w * q.w);
                    *reinterpret_cast<longlong2 *>(slotp[u] + c) = lo;                  // c + 1 < lds: lds is even
                    if (c + 2 < lds) *reinterpret_cast<longlong2 *>(slotp[u] + c + 2) = hi;
                }
                if (c == 0 && bb[u] >= 0) atomicAdd(A.cnt + bb[u], 1ull);
            }
            SCAT_STAMP(3);
            fence_proxy_async_smem();
            __syncwarp();
            SCAT_STAMP(4);
#pragma unroll
            for (int u = 0; u < NBUF; ++u)
                if (bb[u] >= 0 && c == 0)
                    bulk_reduce_add_u64(A.S + (size_t)bb[u] * (size_t)lds, slota[u], (uint32_t)lds * 8u);
            bulk_commit();
            SCAT_STAMP(5);
            ++it;
            const bool more = r0 + step < rows;
#pragma unroll
            for (int u = 0; u < NBUF; ++u) { bb[u] = more ? bbn[u] : -1; v[u] = vn[u]; }
        }
    } else {
        for (int r = first; r < rows; r += stride) {
            const int bb = bm[r];
            if (bb < 0) continue;                                   // warp-uniform
            const float *xr = A.X + (row0 + r) * A.ldx;
            if (lane == 0) atomicAdd(A.cnt + bb, 1ull);
            for (int p0 = 0; p0 < lds; p0 += ACC_PIECE, ++it) {
                long long *buf = buffer(it);
                bulk_wait_read<NBUF - 1>();
                __syncwarp();
                const int c = p0 + lane * 4;
                if (c < lds) {
                    float v[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
                    if (A.vec) {
                        const float4 a = __ldg(reinterpret_cast<const float4 *>(xr + c));
                        const float4 s = __ldg(reinterpret_cast<const float4 *>(A.qscale + c));
                        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; q[0] = s.x; q[1] = s.y; q[2] = s.z; q[3] = s.w;
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) if (c + j < d) { v[j] = __ldg(xr + c + j); q[j] = __ldg(A.qscale + c + j); }
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) if (c + j < lds) buf[lane * 4 + j] = __float2ll_rn(v[j] * q[j]);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    const int cols = lds - p0 < ACC_PIECE ? lds - p0 : ACC_PIECE;
                    bulk_reduce_add_u64(A.S + (int64_t)bb * lds + p0, (uint32_t)__cvta_generic_to_shared(buf), (uint32_t)cols * 8u);
                }
                bulk_commit();
            }
        }
    }
}

// power-of-two factor that brings a row whose largest magnitude is amax into [2^14, 2^15): the fp16
// hi/lo split then keeps 22 significant bits of every element that matters (exact scaling, undone
// exactly in the epilogue).  Zero / non-finite rows are left alone.
__device__ __forceinline__ float pow2_scale_for(float amax) {
    if (!(amax > 0.f) || !isfinite(amax)) return 1.f;
    int a = 14 - ilogbf(amax);
    a = a < -100 ? -100 : (a > 100 ? 100 : a);
    return ldexpf(1.f, a);
}

// lexicographic (value, index) minimum: the first minimum wins, as numpy's argmin.
__device__ __forceinline__ void argmin_merge(float &v, int &i, float ov, int oi) {
    if (ov < v || (ov == v && oi < i)) { v = ov; i = oi; }
}


// ---- programmatic dependent launch -------------------------------------------------------------------
// The epoch is a serial chain of short kernels around the BMU kernel (stats -> split -> BMU -> tables -> apply ->
// merge); launched with the programmatic-stream-serialization attribute, each one is made resident while its
// predecessor is still running and blocks in pdl_wait() until that grid has completed and flushed, which hides
// the launch latency between them (~6 us per boundary at config 2).  Every kernel launched through launch_pdl
// MUST call pdl_wait() before its first global-memory access.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

}  // namespace somb200
