// K3: per-BMU sample sums, exact and order-independent.  S[bmu[r], :] += X[r, :],  cnt[bmu[r]] += 1.
//
// This is the sample side of the reference's second GEMM g^T X and of sum(g) (xpysom.py:436-440): because
// h(bmu, k) depends on the sample only through its BMU, g^T X == H^T S -- the "segmented
// reduction" of north_star item (4).  The sums are 64-bit fixed-point integers (common.cuh: ExactAcc), so the
// result does not depend on the order of the atomics, on the tiling or on the sharding; the fp32 values the
// neighbourhood apply reads are produced once per epoch by accum_finalize_kernel (one rounding per element).
//
// Kernels here:
//   column_absmax (inside row_scale_kernel, misc.cuh)   per-column largest magnitude of the samples, once per upload
//   accum_scales_kernel       q_c from the column maxima and the total sample count: qscale = 2^q_c, qinv = 2^-q_c
//   accumulate_kernel         the scatter for the SIMT BMU path (the tensor-core kernels scatter from their own warps)
//   accum_finalize_kernel     int64 -> fp32 S, c (and clears the integers for the next epoch)
//   accum_fold_kernel         int64 part -> running fp64 sums (several parts with different scales: chunked uploads,
//                             streamed out-of-core blocks), accum_finalize_f64_kernel rounds those once
#pragma once
#include "common.cuh"
#include "peer.cuh"

namespace somb200 {

constexpr int ACC_THREADS = 128;      // four scatter warps per CTA, as in the tensor-core kernels
constexpr int ACC_NBUF = 4;

// qscale[c] = 2^q_c, qinv[c] = 2^-q_c with q_c = 62 - (e_c + 1) - ceil(log2 n_total), e_c = ilogb(colmax[c]).
// A zero / non-finite column gets q = 0.  q is clamped to what fp32 can hold as a power of two.
__global__ void accum_scales_kernel(const float *__restrict__ colmax, int d, int d_pad, double n_total,
                                    float *__restrict__ qscale, float *__restrict__ qinv) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= d_pad) return;
    int q = 0;
    if (c < d) {
        const float m = colmax[c];
        if (m > 0.f && isfinite(m)) {
            int lg = 0;
            while (ldexp(1.0, lg) < n_total) ++lg;
            q = 62 - (ilogbf(m) + 1) - lg;
            q = q > 120 ? 120 : (q < -120 ? -120 : q);
        }
    }
    qscale[c] = c < d ? ldexpf(1.f, q) : 0.f;
    qinv[c] = c < d ? ldexpf(1.f, -q) : 0.f;
}

__global__ void __launch_bounds__(ACC_THREADS)
accumulate_kernel(const ExactAcc A0, const int32_t *__restrict__ bmu, int64_t n, int64_t rows_per_cta) {
    const ExactAcc A = A0.for_cta(blockIdx.x);
    __shared__ __align__(128) long long stage[4][ACC_NBUF][ACC_PIECE];
    __shared__ int bm[128];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t r1 = r0 + rows_per_cta < n ? r0 + rows_per_cta : n;
    uint32_t it = 0;
    for (int64_t t0 = r0; t0 < r1; t0 += 128) {
        const int rows = (int)(r1 - t0 < 128 ? r1 - t0 : 128);
        __syncthreads();
        if ((int)threadIdx.x < 128) bm[threadIdx.x] = (int)threadIdx.x < rows ? __ldg(bmu + t0 + threadIdx.x) : -1;
        __syncthreads();
        scatter_rows_exact<ACC_NBUF>(A, bm, t0, rows, warp, 4, lane, &stage[warp][0][0], it);
    }
    bulk_wait_all();
}

inline int launch_accumulate(const float *X, int64_t n, int d, int64_t ldx, const int32_t *bmu, int k,
                             const AccTarget &T, int sm_count, cudaStream_t st) {
    if (n <= 0) return 0;
    ExactAcc A;
    A.X = X; A.ldx = ldx; A.d = d; A.k = k; A.qscale = T.qscale; A.S = T.S; A.cnt = T.cnt; A.lds = acc_ld(d); A.dbg = 0;
    A.reps = T.reps; A.rep_words = T.rep_words;
    A.vec = ((d % 4 == 0) && (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0)) ? 1 : 0;
    // a multiple of the SM count; each CTA takes a contiguous slab of whole 128-row tiles so its loads are sequential
    int64_t ctas = (int64_t)sm_count * 8;
    int64_t rows_per_cta = round_up(ceil_div(n, ctas), 128);
    ctas = ceil_div(n, rows_per_cta);
    accumulate_kernel<<<(unsigned)ctas, ACC_THREADS, 0, st>>>(A, bmu, n, rows_per_cta);
    return check_cuda(cudaGetLastError(), "accumulate_kernel launch");
}

// int64 sums -> the fp32 S (K, D) and c (K) the neighbourhood apply reads (peer.cuh: accum_finalize_elements); the
// integers are cleared for the next epoch when `clear` is set.  With a PeerView the sums run over the accumulators of
// all ranks (the sharded path's exchange step, fused in here).
__global__ void accum_finalize_kernel(unsigned long long *__restrict__ Si, int reps, size_t rep_words,
                                      const float *__restrict__ qinv, int k, int d, int lds, float *__restrict__ S,
                                      float *__restrict__ c, int clear, const PeerView V) {
    pdl_wait(); pdl_trigger();
    accum_finalize_elements(Si, reps, rep_words, qinv, k, d, lds, S, c, clear, V,
                            (int64_t)blockIdx.x * blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x);
}

// one part of a multi-part epoch (its own column scales): Sd += double(S_int) * 2^-q, cd += count; integers cleared
__global__ void accum_fold_kernel(unsigned long long *__restrict__ Si, int reps, size_t rep_words, const float *__restrict__ qinv,
                                  int k, int d, int lds, double *__restrict__ Sd, double *__restrict__ cd) {
    const int64_t tot = (int64_t)k * lds, all = tot + k;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < all; e += (int64_t)gridDim.x * blockDim.x) {
        unsigned long long u = 0ull;
        for (int r = 0; r < reps; ++r) {
            const unsigned long long w = Si[(size_t)r * rep_words + e];
            u += w;
            if (w) Si[(size_t)r * rep_words + e] = 0ull;
        }
        if (!u) continue;
        if (e < tot) {
            const int row = (int)(e / lds), col = (int)(e % lds);
            if (col < d) Sd[(int64_t)row * d + col] += (double)(long long)u * (double)qinv[col];
        } else {
            cd[e - tot] += (double)u;
        }
    }
}

// the replicas of a local accumulator folded into ONE copy and cleared: dst == Si: replicas 1.. added into replica 0
// (before an all-reduce: one copy travels); dst != Si: all replicas summed into dst (a peer-memory accumulator, zero
// before) -- the sharded path on hot data: scatter into local replicas, fold into the mailbox, exchange
__global__ void accum_fold_replicas_kernel(unsigned long long *__restrict__ Si, int reps, size_t rep_words, int64_t all,
                                           unsigned long long *__restrict__ dst) {
    pdl_wait(); pdl_trigger();
    const int r0 = dst == Si ? 1 : 0;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < all; e += (int64_t)gridDim.x * blockDim.x) {
        unsigned long long u = 0ull;
        for (int r = r0; r < reps; ++r) {
            const unsigned long long w = Si[(size_t)r * rep_words + e];
            if (w) { u += w; Si[(size_t)r * rep_words + e] = 0ull; }
        }
        if (u) dst[e] += u;
    }
}

// running fp64 sums -> fp32 S, c; the doubles are cleared
__global__ void accum_finalize_f64_kernel(double *__restrict__ Sd, double *__restrict__ cd, int k, int d,
                                          float *__restrict__ S, float *__restrict__ c) {
    const int64_t tot = (int64_t)k * d;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (int64_t)gridDim.x * blockDim.x) {
        S[e] = (float)Sd[e];
        Sd[e] = 0.0;
    }
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < k; e += gridDim.x * blockDim.x) { c[e] = (float)cd[e]; cd[e] = 0.0; }
}

inline int grid_for(int64_t items, int sm_count) {
    int64_t b = ceil_div(items, 256);
    if (b > (int64_t)sm_count * 8) b = (int64_t)sm_count * 8;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace somb200
