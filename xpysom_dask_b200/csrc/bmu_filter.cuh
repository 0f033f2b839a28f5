// K1 (filter + refine): the BMU search for long rows (D >= 256) in ONE tensor-core pass instead of three.
//
// bmu_tc3.cuh gets fp32-accurate scores out of 11-bit tensor-core inputs by running three MMAs per K step
// (lo*hi + hi*lo + hi*hi); at config 4 (K = 10^4, D = 784) that kernel sits at the 3-pass ceiling of the tensor pipe.
// Here the contraction is run ONCE, on the fp16 "hi" parts only, and what it produces per (sample, neuron) is not the
// score but an INTERVAL that is guaranteed to contain it:
//     s^_k = x_hi . w'_hi,k + bias_k,      |s_k - s^_k| <= E_k = A_r nwh_k + B_r nwl_k
// (A_r, B_r: residual norms of the sample's split, nwh_k, nwl_k: of the neuron's; Cauchy-Schwarz on the three dropped
// terms plus the roundings listed below).  A neuron can only be the BMU if its lower bound does not exceed the
// smallest upper bound of the row, so the epilogue keeps, per row, the running minimum U of the upper bounds and a
// short list of (neuron, lower bound) with lower bound <= U.  A second kernel re-scores the surviving candidates
// DIRECTLY (fp32 sum of (x - w)^2 from the caller's arrays, no cancellation) and takes the first minimum: at least as
// accurate as the three-pass kernel, at a third of the tensor work.
//
// Everything is CENTRED: the Euclidean BMU is translation invariant, so samples and codebook are shifted by the
// same vector mu (column means of the first rows of the upload); the bound is proportional to |x - mu| |w - mu|, and on a
// young, smooth map -- all neurons close to the data mean -- that is what makes the candidate lists short
// (profiles/r2_filter_candidates_c4.txt: median 6 candidates of 10^4 after centring, 430 without).
//
// What the bound E_k covers, in units of the unscaled score (x = sample, w' = -2 (w - mu), ^ = after fp16 split):
//   x_lo . w'_hi + x_hi . w'_lo + x_lo . w'_lo      <= nxl nwh + (nxh + nxl) nwl
//   fp32 accumulation of D products in TMEM          <= 2^-14 nxh nwh               (D <= 1024)
//   fl(x - mu), fl(w - mu) (the centring itself)      <= 2^-23 (|x| + |mu|) nwh,  2^-22 (|w| + |mu|) nxh
//   fp32 roundings of s^ itself                        <= 2^-22 bias_k
//   => A_r = 1.01 (nxl + 2^-14 nxh + 2^-23 (|x| + |mu|)),  B_r = 1.01 (nxh + nxl),  nwl_k includes 2^-22 (|w_k| + |mu|)
//
// Pipeline per CTA pair (same 2-SM structure as bmu_tc3.cuh, no converter: the fp16 sample copy is made once per
// upload): warp 0 TMA producer (A = fp16 sample tile 128 x 64, B = this CTA's half of the fp16 codebook tile, one
// 32 KB stage, 6 stages) | warp 1 MMA issuer (4 MMAs per 64-feature block) | warps 4-11 epilogue (two per TMEM lane
// quarter, 128 columns each).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include "common.cuh"
#include "bmu_tc.cuh"
#include "bmu_tc2.cuh"
#include "bmu_tc3.cuh"

namespace somb200 {
namespace flt {

using tc::smem_u32;
using namespace tc2;   // cluster / 2-SM wrappers

constexpr int BM = 128, TBN = 256, TBNH = 128, BK = 64, UMMA_K = 16, NACC = 2;
constexpr int NST = 6;                                   // pipeline stages of [A 16 KB | B 16 KB]
constexpr int TILE_BYTES = BM * BK * 2;                  // 16 KB
constexpr int STAGE_BYTES = 2 * TILE_BYTES;
constexpr int NUM_THREADS = 384;
constexpr int EPI_WARP0 = 4, NEPI_WARPS = 8;
constexpr int EPI_STAGE_FLOATS = 4 * 128;                // per epilogue warp: bias | sinv | nwh | nwl of its 128 columns
constexpr int NUM_BARS = 2 * NST + 2 * NACC;
constexpr int SMEM_BYTES = NST * STAGE_BYTES + NEPI_WARPS * EPI_STAGE_FLOATS * 4 + 2 * BM * 4 + NUM_BARS * 8 + 64 + 1024;
static_assert(SMEM_BYTES <= 227 * 1024, "shared memory of the filter kernel");

constexpr int CAPH = 128;                                // candidate entries per row and column half
constexpr int CAND_PER_ROW = 2 * CAPH;
constexpr int REF_NBUF = 2;                               // staging buffers per refine warp for the fused accumulate
constexpr int MU_ROWS = 65536;                           // rows the centre is estimated from

// where the sample-side and codebook-side buffers of the filter live in the caller's workspace
struct FilterLayout {
    int d_pad8, d_pad64, k_pad;
    size_t mu_off;        // d_pad8 floats: the centre; then d_pad8 doubles of column sums + 1 double |mu|
    size_t xh_off;        // n x d_pad8 halves: fp16 hi part of the centred, row-scaled samples
    size_t rstat_off;     // n x float4: (2^-a_r, A_r, B_r, |x - mu|^2)
    size_t cand_off;      // n x CAND_PER_ROW int2: (neuron, bits of the lower bound)
    size_t meta_off;      // n x 2 int2: per column half (entries | overflow mark, bits of U)
    size_t seed_off;      // n floats: upper bound of each row's minimum before the sweep (filter_seed_kernel)
    size_t wh_off;        // k_pad x d_pad64 halves: fp16 hi part of the centred, scaled codebook -2 (w - mu) 2^b_k
    size_t wstat_off;     // 4 x k_pad floats: bias | 2^-b_k | nwh | nwl
    size_t ovf_off;       // int[0]: rows that overflowed their lists in this launch; uint64 at +8: candidates re-scored
    size_t total;
};
__host__ inline FilterLayout filter_layout(int64_t n, int k, int d) {
    FilterLayout L;
    L.d_pad8 = (int)round_up(d, 8); L.d_pad64 = (int)round_up(d, 64); L.k_pad = (int)round_up(k, kKPad);
    size_t off = 0;
    L.mu_off = off;    off += round_up((size_t)L.d_pad8 * 4 + (size_t)L.d_pad8 * 8 + 64, 1024);
    L.xh_off = off;    off += round_up((size_t)n * L.d_pad8 * 2, 1024);
    L.rstat_off = off; off += round_up((size_t)n * 16, 1024);
    L.cand_off = off;  off += round_up((size_t)n * CAND_PER_ROW * 8, 1024);
    L.meta_off = off;  off += round_up((size_t)n * 16, 1024);
    L.seed_off = off;  off += round_up((size_t)n * 4, 1024);
    L.wh_off = off;    off += round_up((size_t)L.k_pad * L.d_pad64 * 2, 1024);
    L.wstat_off = off; off += round_up((size_t)4 * L.k_pad * 4, 1024);
    L.ovf_off = off;   off += 1024;
    L.total = off;
    return L;
}

// the filter pays off where the three-pass kernel is bound by the tensor pipe: long rows, large maps, Euclidean
__host__ inline bool filter_eligible(const float *X, int64_t n, int d, int64_t ldx, int k, int dist_kind) {
    return dist_kind == SOM_DIST_EUCLIDEAN && d >= 256 && d <= 1024 && d % 4 == 0 && ldx % 4 == 0 && k >= 1024 &&
           n >= 4096 && ((reinterpret_cast<uintptr_t>(X) & 15) == 0) && n < ((int64_t)1 << 31) - BM;
}

// ---- once per upload: centre, fp16 copy, per-row statistics -------------------------------------------------
__global__ void column_sum_kernel(const float *__restrict__ X, int64_t rows, int d, int64_t ldx, double *__restrict__ sums) {
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
        double s = 0.0;
        for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) s += (double)__ldg(X + r * ldx + c);
        atomicAdd(sums + c, s);
    }
}
__global__ void centre_finish_kernel(const double *__restrict__ sums, double rows, int d, int d_pad8, float *__restrict__ mu,
                                     double *__restrict__ mu_norm) {
    __shared__ double part[256];
    double s = 0.0;
    for (int c = threadIdx.x; c < d_pad8; c += blockDim.x) {
        const float m = c < d ? (float)(sums[c] / rows) : 0.f;
        mu[c] = m;
        s += (double)m * m;
    }
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < (int)blockDim.x; ++i) t += part[i];
        *mu_norm = sqrt(t);
    }
}

// one warp per row: xc = x - mu, scaled by 2^a (largest magnitude into [2^14, 2^15)), hi = fp16(xc 2^a)
__global__ void __launch_bounds__(256)
filter_samples_kernel(const float *__restrict__ X, int64_t n, int d, int64_t ldx, const float *__restrict__ mu,
                      const double *__restrict__ mu_norm, int d_pad8, __half *__restrict__ Xh, float4 *__restrict__ rstat) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int d4 = d >> 2;
    constexpr int MAXV = 8;                        // d <= 1024: at most 8 float4 per lane
    const float mun = (float)*mu_norm;
    for (int64_t r = warp; r < n; r += nwarps) {
        const float4 *xr = reinterpret_cast<const float4 *>(X + r * ldx);
        float4 v[MAXV];
        float amax = 0.f; double xsq = 0.0;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int c4 = lane + 32 * i;
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c4 < d4) {
                const float4 x = __ldg(xr + c4), m = __ldg(reinterpret_cast<const float4 *>(mu) + c4);
                xsq += (double)x.x * x.x + (double)x.y * x.y + (double)x.z * x.z + (double)x.w * x.w;
                v[i] = make_float4(x.x - m.x, x.y - m.y, x.z - m.z, x.w - m.w);
                amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v[i].x), fabsf(v[i].y)), fmaxf(fabsf(v[i].z), fabsf(v[i].w))));
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        const float sc = pow2_scale_for(amax);
        double h2 = 0.0, l2 = 0.0, c2 = 0.0;
        __half *dst = Xh + r * d_pad8;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int c4 = lane + 32 * i;
            if (c4 < d4) {
                const float x0 = v[i].x * sc, x1 = v[i].y * sc, x2 = v[i].z * sc, x3 = v[i].w * sc;
                c2 += (double)x0 * x0 + (double)x1 * x1 + (double)x2 * x2 + (double)x3 * x3;
                const __half2 h01 = __floats2half2_rn(x0, x1), h23 = __floats2half2_rn(x2, x3);
                const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                h2 += (double)f01.x * f01.x + (double)f01.y * f01.y + (double)f23.x * f23.x + (double)f23.y * f23.y;
                const float e0 = x0 - f01.x, e1 = x1 - f01.y, e2 = x2 - f23.x, e3 = x3 - f23.y;   // exact residuals
                l2 += (double)e0 * e0 + (double)e1 * e1 + (double)e2 * e2 + (double)e3 * e3;
                uint2 hv;
                hv.x = *reinterpret_cast<const uint32_t *>(&h01); hv.y = *reinterpret_cast<const uint32_t *>(&h23);
                *reinterpret_cast<uint2 *>(dst + c4 * 4) = hv;
            }
        }
        h2 = warp_sum(h2); l2 = warp_sum(l2); xsq = warp_sum(xsq); c2 = warp_sum(c2);
        if (lane == 0) {
            const double inv = 1.0 / (double)sc;
            const double nxh = sqrt(h2) * inv, nxl = sqrt(l2) * inv;
            const double A = 1.01 * (nxl + ldexp(nxh, -14) + ldexp(sqrt(xsq) + (double)mun, -23));
            const double B = 1.01 * (nxh + nxl);
            // round the bounds UP when they become floats
            // .w: |fl(x - mu)|^2, the row constant the filter's scores leave out (for the seed of the upper bound)
            rstat[r] = make_float4((float)inv, __double2float_ru(A), __double2float_ru(B), (float)(c2 * inv * inv));
        }
        // padding columns d .. d_pad8 of the fp16 row stay whatever they are: the tensor map's inner extent is d
    }
}

// ---- once per epoch: centred fp16 codebook copy and per-neuron statistics ------------------------------------
// one warp per (padded) neuron: w' = -2 (w - mu), scaled by its own power of two 2^b_k
__global__ void __launch_bounds__(256)
filter_codebook_kernel(const float *W, int k, int d, int k_pad, int d_pad64, const float *__restrict__ mu,
                       const double *__restrict__ mu_norm, __half *__restrict__ Wh, float *__restrict__ wstat) {
    pdl_wait(); pdl_trigger();
    const int lane = threadIdx.x & 31;
    const int row = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (row >= k_pad) return;
    float *bias = wstat, *sinv = wstat + k_pad, *nwh = wstat + 2 * (size_t)k_pad, *nwl = wstat + 3 * (size_t)k_pad;
    __half *dst = Wh + (size_t)row * d_pad64;
    if (row >= k) {                                  // padding neuron: zero operand, +inf bias
        for (int c = lane; c < d_pad64; c += 32) dst[c] = __float2half_rn(0.f);
        if (lane == 0) { bias[row] = INFINITY; sinv[row] = 1.f; nwh[row] = 0.f; nwl[row] = 0.f; }
        return;
    }
    const float *wr = W + (size_t)row * d;
    float amax = 0.f; double wc2 = 0.0, w2 = 0.0;
    for (int c = lane; c < d; c += 32) {
        const float w = wr[c], wc = w - mu[c];
        amax = fmaxf(amax, fabsf(2.f * wc));
        wc2 += (double)wc * wc; w2 += (double)w * w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    const float sc = pow2_scale_for(amax);
    double h2 = 0.0, l2 = 0.0;
    for (int c = lane; c < d_pad64; c += 32) {
        float v = 0.f;
        if (c < d) v = -2.f * (wr[c] - mu[c]) * sc;
        const __half h = __float2half_rn(v);
        const float hf = __half2float(h), lo = v - hf;
        h2 += (double)hf * hf; l2 += (double)lo * lo;
        dst[c] = h;
    }
    h2 = warp_sum(h2); l2 = warp_sum(l2); wc2 = warp_sum(wc2); w2 = warp_sum(w2);
    if (lane == 0) {
        const double inv = 1.0 / (double)sc;
        bias[row] = (float)wc2;                      // |w - mu|^2 (its fp32 rounding is inside the 1 % slack of A_r, B_r)
        sinv[row] = (float)inv;
        nwh[row] = __double2float_ru(sqrt(h2) * inv);
        nwl[row] = __double2float_ru(sqrt(l2) * inv + ldexp(sqrt(w2) + *mu_norm, -22) + ldexp(sqrt(wc2), -22));
    }
}

// ---- the one-pass contraction with interval epilogue ---------------------------------------------------------
// list full: keep what can still win under the current bound U, then append (cold path, kept out of line)
__device__ __noinline__ int compact_and_push(int2 *list, int cnt, int col, float lb, float U) {
    if (cnt > CAPH) return cnt;                                   // already overflowed
    int m = 0;
    for (int i = 0; i < CAPH; ++i) {
        const int2 e = list[i];
        if (__int_as_float(e.y) <= U) list[m++] = e;
    }
    // a list that stays nearly full would be compacted again at almost every push: give the row up instead (the caller
    // re-does it with the three-pass kernel)
    if (m <= (CAPH * 3) / 4) { list[m++] = make_int2(col, __float_as_int(lb)); return m; }
    return CAPH + 1;
}
__device__ __forceinline__ void push_candidate(int2 *list, int &cnt, int col, float lb, float U) {
    if (cnt < CAPH) { list[cnt++] = make_int2(col, __float_as_int(lb)); return; }
    cnt = compact_and_push(list, cnt, col, lb, U);
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
bmu_filter_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                  const float *__restrict__ wstat, const float4 *__restrict__ rstat, int64_t n, int k, int k_pad,
                  int num_pair_tiles, int num_n_tiles, int num_k_blocks, int d, const float *__restrict__ seed,
                  int2 *__restrict__ cand, int2 *__restrict__ meta) {
    pdl_wait(); pdl_trigger();
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *smem = smem_raw + (smem_base - smem_u32(smem_raw));
    float *epi_stage = reinterpret_cast<float *>(smem + NST * STAGE_BYTES);
    float *ushare = epi_stage + NEPI_WARPS * EPI_STAGE_FLOATS;                      // [2][BM]: each half's running U
    uint64_t *bars = reinterpret_cast<uint64_t *>(ushare + 2 * BM);
    const uint32_t bar0 = smem_u32(bars);
    auto full_bar  = [&](int s) { return bar0 + 8u * s; };                           // leader: A and B of both CTAs landed
    auto empty_bar = [&](int s) { return bar0 + 8u * (NST + s); };                   // local: stage consumed
    auto tfull_bar  = [&](int a) { return bar0 + 8u * (2 * NST + a); };
    auto tempty_bar = [&](int a) { return bar0 + 8u * (2 * NST + NACC + a); };
    volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(bars + NUM_BARS);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int pair = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    const int n_last = (int)round_up(k - (num_n_tiles - 1) * TBN, 16);               // narrow last tile (bmu_tc3.cuh)

    if (threadIdx.x == 0) {
        for (int s = 0; s < NST; ++s) { tc::mbar_init(full_bar(s), 1); tc::mbar_init(empty_bar(s), 1); }
        for (int a = 0; a < NACC; ++a) { tc::mbar_init(tfull_bar(a), 1); tc::mbar_init(tempty_bar(a), 2 * NEPI_WARPS); }
        tc::fence_barrier_init();
        tc::tma_prefetch_desc(&map_x); tc::tma_prefetch_desc(&map_w);
    }
    if (warp == 1) tmem_alloc_2sm(smem_u32((const void *)tmem_slot), 512);
    tc::tc_fence_before();
    cluster_sync();
    tc::tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== producer: A (fp16 samples) and this CTA's half of B (fp16 codebook) ===============
        uint32_t it = 0;
        for (int pt = pair; pt < num_pair_tiles; pt += num_pairs) {
            const int row0 = pt * (2 * BM) + (int)rank * BM;
            for (int nt = 0; nt < num_n_tiles; ++nt) {
                const int nrow = nt * TBN + (int)rank * ((nt == num_n_tiles - 1 ? n_last : TBN) / 2);
                for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
                    const int s = it % NST; const uint32_t ph = (it / NST) & 1;
                    tc::mbar_wait(empty_bar(s), ph ^ 1);
                    const uint32_t st = smem_base + s * STAGE_BYTES;
                    if (tc::elect_one()) {
                        if (leader) tc::mbar_expect_tx(full_bar(s), 4 * TILE_BYTES);      // bytes of both CTAs
                        tma_load_2d_2sm(st, &map_x, kb * BK, row0, full_bar(s));
                        tma_load_2d_2sm(st + TILE_BYTES, &map_w, kb * BK, nrow, full_bar(s));
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA): one pass, 4 MMAs per 64-feature block ======================
        if (leader) {
            uint32_t it = 0, acc_it = 0;
            for (int pt = pair; pt < num_pair_tiles; pt += num_pairs)
                for (int nt = 0; nt < num_n_tiles; ++nt, ++acc_it) {
                    const int a = acc_it % NACC; const uint32_t aph = (acc_it / NACC) & 1;
                    const uint32_t tmem_d = tmem_base + (uint32_t)(a * TBN);
                    const uint32_t idesc = tc3::idesc_f16(nt == num_n_tiles - 1 ? n_last : TBN);
                    for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
                        const int s = it % NST; const uint32_t ph = (it / NST) & 1;
                        mbar_wait_cluster(full_bar(s), ph);
                        if (kb == 0) mbar_wait_cluster(tempty_bar(a), aph ^ 1);
                        tc::tc_fence_after();
                        const uint32_t st = smem_base + s * STAGE_BYTES;
                        const uint64_t adesc = tc::make_smem_desc(st), bdesc = tc::make_smem_desc(st + TILE_BYTES);
                        const int dl = d - kb * BK;
                        const int kk_n = dl >= BK ? BK / UMMA_K : (dl + UMMA_K - 1) / UMMA_K;
                        if (tc::elect_one()) {
#pragma unroll
                            for (int kk = 0; kk < BK / UMMA_K; ++kk) {
                                if (kk >= kk_n) break;
                                const uint64_t off = (uint64_t)((kk * UMMA_K * 2) >> 4);
                                tc3::umma_f16_2sm(tmem_d, adesc + off, bdesc + off, idesc, (kb | kk) != 0);
                            }
                            umma_commit_2sm(empty_bar(s));
                            if (kb == num_k_blocks - 1) umma_commit_2sm(tfull_bar(a));
                        }
                        __syncwarp();
                    }
                }
        }
    } else if (warp >= EPI_WARP0) {
        // ===================== epilogue: intervals, running upper bound, candidate lists ===========================
        const int q = warp & 3, h = (warp - EPI_WARP0) >> 2;
        const int row_in_tile = q * 32 + lane;
        float *wst = epi_stage + (warp - EPI_WARP0) * EPI_STAGE_FLOATS;              // bias | sinv | nwh | nwl, 128 each
        const float *gb = wstat, *gs = wstat + k_pad, *gh = wstat + 2 * (size_t)k_pad, *gl = wstat + 3 * (size_t)k_pad;
        uint32_t acc_it = 0;
        float4 nx[4];                                                                 // next tile's slice, one float4 per array
        {
            const int c = h * (TBN / 2) + lane * 4;
            nx[0] = __ldg(reinterpret_cast<const float4 *>(gb + c)); nx[1] = __ldg(reinterpret_cast<const float4 *>(gs + c));
            nx[2] = __ldg(reinterpret_cast<const float4 *>(gh + c)); nx[3] = __ldg(reinterpret_cast<const float4 *>(gl + c));
        }
        for (int pt = pair; pt < num_pair_tiles; pt += num_pairs) {
            const int64_t row = (int64_t)pt * (2 * BM) + (int64_t)rank * BM + row_in_tile;
            const bool live = row < n;
            float fr = 0.f, Ar = 0.f, Br = 0.f;
            if (live) { const float4 rs = __ldg(rstat + row); fr = rs.x; Ar = rs.y; Br = rs.z; }
            float U = live ? __ldg(seed + row) : -INFINITY;                          // rows past the end collect nothing
            int cnt = 0;
            int2 *list = cand + (live ? row : 0) * CAND_PER_ROW + h * CAPH;
            // the two warps of a lane quarter exchange their running bounds through `ushare`: nobody may still be reading
            // the previous row tile's values when the new ones are written, nor read before they are written
            asm volatile("bar.sync %0, 64;" :: "r"(2 + q) : "memory");
            ushare[h * BM + row_in_tile] = U;
            asm volatile("bar.sync %0, 64;" :: "r"(2 + q) : "memory");
            for (int nt = 0; nt < num_n_tiles; ++nt, ++acc_it) {
                const int a = acc_it % NACC; const uint32_t aph = (acc_it / NACC) & 1;
                const int col0 = nt * TBN + h * (TBN / 2);
                __syncwarp();
#pragma unroll
                for (int j = 0; j < 4; ++j) reinterpret_cast<float4 *>(wst + j * 128)[lane] = nx[j];
                __syncwarp();
                {
                    const int c = (nt + 1 < num_n_tiles ? nt + 1 : 0) * TBN + h * (TBN / 2) + lane * 4;
                    nx[0] = __ldg(reinterpret_cast<const float4 *>(gb + c)); nx[1] = __ldg(reinterpret_cast<const float4 *>(gs + c));
                    nx[2] = __ldg(reinterpret_cast<const float4 *>(gh + c)); nx[3] = __ldg(reinterpret_cast<const float4 *>(gl + c));
                }
                // the other column half's bound is as good as mine (stale values are still valid upper bounds)
                U = fminf(U, ushare[(h ^ 1) * BM + row_in_tile]);
                tc::mbar_wait(tfull_bar(a), aph);
                tc::tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * TBN + h * (TBN / 2));
                int nch = 4;
                { const int left = k - col0; nch = left <= 0 ? 0 : ((left + 31) >> 5 < 4 ? (left + 31) >> 5 : 4); }
#pragma unroll 1
                for (int c = 0; c < nch; ++c) {
                    uint32_t v[32];
                    tc::tmem_ld32(taddr + c * 32, v);
                    tc::tmem_ld_wait_dep(v);
                    const float4 *b4 = reinterpret_cast<const float4 *>(wst + c * 32);
                    const float4 *s4 = reinterpret_cast<const float4 *>(wst + 128 + c * 32);
                    const float4 *h4 = reinterpret_cast<const float4 *>(wst + 256 + c * 32);
                    const float4 *l4 = reinterpret_cast<const float4 *>(wst + 384 + c * 32);
                    // straight-line part: every score becomes its lower bound (kept in v[]), U takes the upper bounds
#pragma unroll
                    for (int j4 = 0; j4 < 8; ++j4) {
                        const float4 b = b4[j4], si = s4[j4], nh = h4[j4], nl = l4[j4];
                        const float bb[4] = {b.x, b.y, b.z, b.w}, ss[4] = {si.x, si.y, si.z, si.w};
                        const float hh[4] = {nh.x, nh.y, nh.z, nh.w}, ll[4] = {nl.x, nl.y, nl.z, nl.w};
                        float ub[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const int j = j4 * 4 + e;
                            const float sc = fmaf(__uint_as_float(v[j]), fr * ss[e], bb[e]);
                            const float er = fmaf(Ar, hh[e], fmaf(Br, ll[e], bb[e] * 2.4e-7f));   // + 2^-22 |bias|: fp32 roundings of s
                            v[j] = __float_as_uint(sc - er);
                            ub[e] = sc + er;
                        }
                        asm("min.f32 %0, %0, %1, %2;" : "+f"(U) : "f"(ub[0]), "f"(ub[1]));
                        asm("min.f32 %0, %0, %1, %2;" : "+f"(U) : "f"(ub[2]), "f"(ub[3]));
                    }
                    // candidates: lower bound <= the bound AFTER this chunk (tighter than the running one, still valid).
                    // Groups of 8 columns; a group is only looked at when some lane of the warp has a candidate in it.
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        float m;
                        asm("min.f32 %0, %1, %2, %3;" : "=f"(m) : "f"(__uint_as_float(v[g * 8])), "f"(__uint_as_float(v[g * 8 + 1])), "f"(__uint_as_float(v[g * 8 + 2])));
                        asm("min.f32 %0, %0, %1, %2;" : "+f"(m) : "f"(__uint_as_float(v[g * 8 + 3])), "f"(__uint_as_float(v[g * 8 + 4])));
                        asm("min.f32 %0, %0, %1, %2;" : "+f"(m) : "f"(__uint_as_float(v[g * 8 + 5])), "f"(__uint_as_float(v[g * 8 + 6])));
                        m = fminf(m, __uint_as_float(v[g * 8 + 7]));
                        if (__any_sync(0xffffffffu, m <= U)) {
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                const float lb = __uint_as_float(v[g * 8 + e]);
                                if (lb <= U) push_candidate(list, cnt, col0 + c * 32 + g * 8 + e, lb, U);
                            }
                        }
                    }
                }
                tc::tc_fence_before();
                ushare[h * BM + row_in_tile] = U;
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(map_to_cta(tempty_bar(a), 0));
            }
            if (live) meta[row * 2 + h] = make_int2(cnt, __float_as_int(U));
        }
    }

    __syncwarp();
    tc::tc_fence_before();
    cluster_sync();
    if (warp == 1) { tc::tc_fence_after(); tmem_dealloc_2sm(tmem_base, 512); }
}

// ---- seed: an upper bound of every row's minimum before the sweep --------------------------------------------------
// Without one the running bound U starts at +inf and the first neurons of a sweep are all "candidates" (record
// minima): 50 - 100 list entries per row that later compactions throw away.  The BMU of the previous epoch is almost
// as good as this epoch's: its squared distance under the CURRENT codebook, minus the row constant |x - mu|^2 the
// filter's scores leave out, bounds the minimum from above (any neuron's score does), and the sweep then only lists
// neurons that can really win.  2^-16 of slack covers the fp32 evaluation and the centring roundings.
__global__ void __launch_bounds__(256)
filter_seed_kernel(const float *__restrict__ X, int64_t n, int d, int64_t ldx, const float *W, int k,
                   const int32_t *__restrict__ bmu_prev, const float4 *__restrict__ rstat, float *__restrict__ seed) {
    pdl_wait(); pdl_trigger();
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int d4 = d >> 2;
    for (int64_t r = warp; r < n; r += nwarps) {
        const int b = bmu_prev[r];
        if (b < 0 || b >= k) { if (lane == 0) seed[r] = INFINITY; continue; }
        const float4 *xr = reinterpret_cast<const float4 *>(X + r * ldx);
        const float4 *wr = reinterpret_cast<const float4 *>(W + (size_t)b * d);
        float s0 = 0.f, s1 = 0.f;
        for (int c4 = lane; c4 < d4; c4 += 32) {
            const float4 x = __ldg(xr + c4), w = wr[c4];
            const float a = x.x - w.x, bb = x.y - w.y, c = x.z - w.z, e = x.w - w.w;
            s0 = fmaf(a, a, fmaf(bb, bb, s0)); s1 = fmaf(c, c, fmaf(e, e, s1));
        }
        const float d2 = warp_sum(s0 + s1);
        if (lane == 0) {
            const float xc2 = rstat[r].w;
            seed[r] = (d2 - xc2) + 1.53e-5f * (d2 + xc2);
        }
    }
}

// ---- refine: the surviving candidates re-scored directly, first minimum ----------------------------------------------
// |x - w|^2 as a plain fp32 sum of squared differences (four accumulators per lane, butterfly over the warp): relative
// error ~1e-6 of the DISTANCE, far inside the stated epsilon (which is relative to |x|^2 + |d|) and more accurate than
// the reference's own |w|^2 - 2 x.w in fp32.  (An fp64 version was 8x slower: fp32 -> fp64 conversions run at 16 per
// clock and SM.)
// one warp per row.  A row whose list overflowed gets bmu = -1 and is counted in ovf[0] (the caller re-does those rows
// with the three-pass kernel); ovf[2..3] (one uint64) = candidates re-scored in this launch (the caller's policy input)
__global__ void __launch_bounds__(256, 2)
bmu_refine_kernel(const float *__restrict__ X, int64_t n, int d, int64_t ldx, const float *W, int k,
                  const int2 *__restrict__ cand, const int2 *__restrict__ meta, int32_t *__restrict__ bmu, int *__restrict__ ovf,
                  const ExactAcc A0) {
    pdl_wait(); pdl_trigger();
    // fused exact accumulate (A0.S != nullptr): the warp that has just found a row's BMU sends the row to the accumulator
    // itself (scatter_rows_exact, common.cuh) -- the row is hot in L1/L2 and the L2 reductions overlap the other warps'
    // re-scoring instead of running as a separate pass (2.3 ms per 1M rows at D = 784)
    __shared__ __align__(128) long long stage[8][REF_NBUF][ACC_PIECE];
    __shared__ int bm_s[8];
    const ExactAcc A = A0.for_cta(blockIdx.x);
    const int wib = threadIdx.x >> 5;
    uint32_t bulk_it = 0;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int d4 = d >> 2;
    constexpr int MAXV = 8;
    int n_ovf = 0, n_eval = 0;
    for (int64_t r = warp; r < n; r += nwarps) {
        const float4 *xr = reinterpret_cast<const float4 *>(X + r * ldx);
        float4 x[MAXV];
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int c4 = lane + 32 * i;
            x[i] = c4 < d4 ? __ldg(xr + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        auto load_w = [&](int kk, float4 (&w)[MAXV]) {
            const float4 *wr = reinterpret_cast<const float4 *>(W + (size_t)kk * d);
#pragma unroll
            for (int i = 0; i < MAXV; ++i) {
                const int c4 = lane + 32 * i;
                w[i] = c4 < d4 ? wr[c4] : x[i];                 // (difference 0 past the end)
            }
        };
        auto dist2 = [&](const float4 (&w)[MAXV]) -> float {
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int i = 0; i < MAXV; ++i) {
                const float a = x[i].x - w[i].x, b = x[i].y - w[i].y, c = x[i].z - w[i].z, e = x[i].w - w[i].w;
                s0 = fmaf(a, a, s0); s1 = fmaf(b, b, s1); s2 = fmaf(c, c, s2); s3 = fmaf(e, e, s3);
            }
            return warp_sum((s0 + s1) + (s2 + s3));    // (xor butterfly: every lane holds the same bits)
        };
        const int2 m0 = meta[r * 2], m1 = meta[r * 2 + 1];
        const float U = fminf(__int_as_float(m0.y), __int_as_float(m1.y));
        float best = INFINITY; int bidx = 0x7fffffff;
        if (m0.x > CAPH || m1.x > CAPH) {
            n_ovf += 1;                                // the caller re-does this row with the three-pass kernel
            if (lane == 0) bmu[r] = -1;
            continue;
        }
        // The surviving entries of both halves, software-pipelined: the next candidate's codebook row is in flight while
        // this one is reduced (the re-scoring is bound by L2 latency / bandwidth: 4 D bytes per candidate).
        const int2 *list0 = cand + r * CAND_PER_ROW;
        const int total = m0.x + m1.x;
        auto next_live = [&](int i) -> int {           // first entry >= i that survives the final bound; its neuron in nk
            for (; i < total; ++i) {
                const int2 e = i < m0.x ? list0[i] : list0[CAPH + (i - m0.x)];
                if (__int_as_float(e.y) <= U && e.x < k) return i;
            }
            return total;
        };
        auto unit_of = [&](int i) -> int { return (i < m0.x ? list0[i] : list0[CAPH + (i - m0.x)]).x; };
        int cur = next_live(0);
        float4 wa[MAXV], wb[MAXV];
        if (cur < total) load_w(unit_of(cur), wa);
        while (cur < total) {
            const int ku = unit_of(cur);
            const int nxt = next_live(cur + 1);
            if (nxt < total) load_w(unit_of(nxt), wb);
            const float v = dist2(wa);
            n_eval += 1;
            if (v < best || (v == best && ku < bidx)) { best = v; bidx = ku; }
#pragma unroll
            for (int i = 0; i < MAXV; ++i) wa[i] = wb[i];
            cur = nxt;
        }
        if (bidx == 0x7fffffff) bidx = 0;
        if (lane == 0) bmu[r] = bidx;
        if (A.S != nullptr) {
            __syncwarp();
            if (lane == 0) bm_s[wib] = bidx;
            __syncwarp();
            scatter_rows_exact<REF_NBUF>(A, bm_s + wib, r, 1, 0, 1, lane, &stage[wib][0][0], bulk_it);
        }
    }
    bulk_wait_all();
    if (lane == 0 && (n_ovf | n_eval)) {
        if (n_ovf) atomicAdd(ovf, n_ovf);
        atomicAdd(reinterpret_cast<unsigned long long *>(ovf + 2), (unsigned long long)n_eval);
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------
inline int filter_prepare_samples(const float *X, int64_t n, int d, int64_t ldx, uint8_t *fws, int sm_count, cudaStream_t st) {
    const FilterLayout L = filter_layout(n, 1, d);             // (the sample side does not depend on k)
    float *mu = reinterpret_cast<float *>(fws + L.mu_off);
    double *sums = reinterpret_cast<double *>(fws + L.mu_off + (size_t)L.d_pad8 * 4);
    double *mu_norm = sums + L.d_pad8;
    SOM_CUDA(cudaMemsetAsync(sums, 0, (size_t)(L.d_pad8 + 1) * 8, st));
    const int64_t rows = n < MU_ROWS ? n : MU_ROWS;
    column_sum_kernel<<<(unsigned)(rows < 512 ? rows : 512), 256, 0, st>>>(X, rows, d, ldx, sums);
    centre_finish_kernel<<<1, 256, 0, st>>>(sums, (double)rows, d, L.d_pad8, mu, mu_norm);
    int64_t blocks = ceil_div(n, 8);
    if (blocks > (int64_t)sm_count * 16) blocks = (int64_t)sm_count * 16;
    filter_samples_kernel<<<(unsigned)blocks, 256, 0, st>>>(X, n, d, ldx, mu, mu_norm, L.d_pad8,
                                                          reinterpret_cast<__half *>(fws + L.xh_off),
                                                          reinterpret_cast<float4 *>(fws + L.rstat_off));
    return check_cuda(cudaGetLastError(), "filter_prepare_samples launch");
}

// BMUs of all rows into bmu[]; ovf (device int, zeroed here) counts the rows that fell back to the full scan
// bmu[] on entry: the BMUs of the previous epoch (or -1), used to seed the bounds; on return: this epoch's
inline int launch_bmu_filter(const float *X, int64_t n, int d, int64_t ldx, const float *W, int k, uint8_t *fws, int32_t *bmu,
                             const AccTarget &T, int sm_count, cudaStream_t st) {
    const FilterLayout L = filter_layout(n, k, d);
    const FilterLayout Ls = filter_layout(n, 1, d);
    // the sample side was laid out without knowing k: its offsets must not depend on it
    SOM_REQUIRE(L.xh_off == Ls.xh_off && L.rstat_off == Ls.rstat_off && L.cand_off == Ls.cand_off && L.seed_off == Ls.seed_off, SOM_E_BADARG, "filter layout");
    const float *mu = reinterpret_cast<const float *>(fws + L.mu_off);
    const double *mu_norm = reinterpret_cast<const double *>(fws + L.mu_off + (size_t)L.d_pad8 * 4) + L.d_pad8;
    __half *Wh = reinterpret_cast<__half *>(fws + L.wh_off);
    float *wstat = reinterpret_cast<float *>(fws + L.wstat_off);
    int *ovf = reinterpret_cast<int *>(fws + L.ovf_off);
    SOM_CUDA(cudaMemsetAsync(ovf, 0, 16, st));
    SOM_CUDA(launch_pdl(filter_codebook_kernel, dim3((unsigned)ceil_div(L.k_pad, 8)), dim3(256), 0, st, W, k, d, L.k_pad,
                        L.d_pad64, mu, mu_norm, Wh, wstat));
    CUtensorMap mx, mw;
    int rc;
    if ((rc = tc3::make_map_2d_f16(&mx, fws + L.xh_off, (uint64_t)d, (uint64_t)n, (uint64_t)L.d_pad8 * 2, BK, BM))) return rc;
    if ((rc = tc3::make_map_2d_f16(&mw, Wh, (uint64_t)L.d_pad64, (uint64_t)L.k_pad, (uint64_t)L.d_pad64 * 2, BK, TBNH))) return rc;
    static bool attr_set[64] = {};
    if (tc::first_launch_on_device(attr_set))
        SOM_CUDA(cudaFuncSetAttribute(bmu_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    int64_t blocks = ceil_div(n, 8);
    if (blocks > (int64_t)sm_count * 8) blocks = (int64_t)sm_count * 8;
    float *seed = reinterpret_cast<float *>(fws + L.seed_off);
    SOM_CUDA(launch_pdl(filter_seed_kernel, dim3((unsigned)blocks), dim3(256), 0, st, X, n, d, ldx, W, k, (const int32_t *)bmu,
                        reinterpret_cast<const float4 *>(fws + L.rstat_off), seed));
    const int num_pair_tiles = (int)ceil_div(n, 2 * BM);
    int pairs = sm_count / 2;
    if (pairs > num_pair_tiles) pairs = num_pair_tiles;
    if (pairs < 1) pairs = 1;
    SOM_CUDA(launch_pdl(bmu_filter_kernel, dim3(2 * pairs), dim3(NUM_THREADS), (size_t)SMEM_BYTES, st, mx, mw, (const float *)wstat,
                        reinterpret_cast<const float4 *>(fws + L.rstat_off), n, k, L.k_pad, num_pair_tiles, L.k_pad / TBN,
                        L.d_pad64 / BK, d, (const float *)seed, reinterpret_cast<int2 *>(fws + L.cand_off),
                        reinterpret_cast<int2 *>(fws + L.meta_off)));
    ExactAcc A;
    A.X = X; A.ldx = ldx; A.d = d; A.k = k; A.qscale = T.qscale; A.S = T.S; A.cnt = T.cnt; A.lds = acc_ld(d); A.dbg = 0;
    A.vec = 1; A.reps = T.reps; A.rep_words = T.rep_words;
    SOM_CUDA(launch_pdl(bmu_refine_kernel, dim3((unsigned)blocks), dim3(256), 0, st, X, n, d, ldx, W, k,
                        reinterpret_cast<const int2 *>(fws + L.cand_off), reinterpret_cast<const int2 *>(fws + L.meta_off), bmu, ovf, A));
    return 0;
}

}  // namespace flt
}  // namespace somb200
