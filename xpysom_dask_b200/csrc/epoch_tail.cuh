// Everything of an epoch that is not the BMU search, in ONE cooperative launch (small maps).
//
// Between two BMU kernels the reference's epoch needs: the neighbourhood tables of this epoch's sigma, the apply
// num = eta H^T S / den = eta H^T c (xpysom.py:434-441 through the H^T S identity), the merge W <- num / den
// (xpysom.py:446-455), the per-neuron statistics and operand copies of the NEW codebook for the next BMU search
// (xpysom.py:529-539) and clean accumulators.  As separate launches that is ten stream operations of 2-14 us
// each with a launch gap between every pair -- 70 us of a 420 us epoch at config 2.  Here the phases run back
// to back in one persistent grid, separated by grid-wide barriers:
//
//   phase 0  factor tables (fp64); the exact int64 per-BMU sums rounded once to the fp32 S, c the apply reads, and
//            cleared for the next epoch (accumulate.cuh); statistics cleared
//   phase 1  apply, the same 64x64 tiles as neigh_apply_kernel<4, 4>, a linear tile index per CTA; when the
//            reduction over BMUs is sliced every slice writes its own partial block (no atomics)
//   phase 2  one warp per neuron: partial blocks summed in slice order, merge, statistics of the new row
//   phase 3  one warp per neuron: TF32 and fp16 operand copies (needs the codebook-wide statistics)
//
// Measured (B200, bench.py --steps 10, same box): 0.426 vs 0.438 ms per epoch at config 2; maps that take the 128x128
// apply tiles were 2.5 % slower this way and keep the separate launches (som_api.cu decides).
// The grid barrier is a sense-reversing counter in the workspace; the kernel is launched cooperatively so that
// all CTAs are co-resident (the launch fails otherwise, it cannot deadlock), and the wait is bounded.
#pragma once
#include "common.cuh"
#include "peer.cuh"
#include "neigh.cuh"
#include "misc.cuh"

namespace somb200 {

struct TailArgs {
    NeighParams P;
    double sigma, dd;
    float *S, *c, *num, *den, *W;
    unsigned long long *Si;             // exact sums, then counts (nullptr: S and c already hold this epoch's fp32 values)
    int reps; size_t rep_words;         // replicas of a local accumulator (common.cuh: acc_replicas)
    PeerView peer;                      // world > 0: the sums run over all ranks' accumulators (peer.cuh)
    const float *qinv;
    int lds;
    float *partials;                    // [slices][K*D + K] when slices > 1
    int k, d, dist_kind, k_pad;
    int slices, b_per_slice, tiles_m, tiles_n;
    float *aux, *bias, *amax;
    unsigned int *gstat;
    SplitOut split;
    int do_split;
    unsigned int *bar;          // [0] arrivals, [16] generation (separate 64-byte lines)
};

__device__ __forceinline__ void grid_barrier(unsigned int *bar) {
    __syncthreads();
    if (threadIdx.x == 0) {
        volatile unsigned int *gen = bar + 16;
        const unsigned int g = *gen;                 // read BEFORE arriving
        __threadfence();
        if (atomicAdd(bar, 1u) == gridDim.x - 1) {
            *bar = 0u;                               // everyone has arrived: reset, then release
            __threadfence();
            atomicAdd(bar + 16, 1u);
        } else {
            const long long t0 = clock64();
            while (*gen == g) {
                if (clock64() - t0 > 4000000000LL) {
                    printf("som_b200: grid barrier timeout (block %d)\n", (int)blockIdx.x);
                    __trap();
                }
            }
        }
        __threadfence();
    }
    __syncthreads();
}

template <int RM, int RN>
__global__ void __launch_bounds__(NB_THREADS)
epoch_tail_kernel(TailArgs A) {
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nthr = gridDim.x * blockDim.x;
    const int K = A.k, D = A.d;
    // ---- phase 0 ----
    neigh_tables_fill(A.P.gx, A.P.gy, A.P.kind, A.P.compact, A.P.shifted, A.sigma, A.dd, const_cast<float *>(A.P.tx),
                      const_cast<float *>(A.P.ty), const_cast<float *>(A.P.mx), const_cast<float *>(A.P.my), tid, nthr);
    if (A.Si != nullptr)
        accum_finalize_elements(A.Si, A.reps, A.rep_words, A.qinv, K, D, A.lds, A.S, A.c, 1, A.peer, tid, nthr);
    if (tid < 4) A.gstat[tid] = 0u;
    grid_barrier(A.bar);
    // ---- phase 1: apply ----
    {
        const int ntiles = A.tiles_m * A.tiles_n * A.slices;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
            const int bz = t / (A.tiles_m * A.tiles_n), r = t % (A.tiles_m * A.tiles_n);
            const int by = r / A.tiles_m, bx = r % A.tiles_m;
            const int b_begin = bz * A.b_per_slice;
            const int b_end = min(K, b_begin + A.b_per_slice);
            float *num = A.num, *den = A.den;
            if (A.slices > 1) { num = A.partials + (size_t)bz * ((size_t)K * D + K); den = num + (size_t)K * D; }
            neigh_apply_tile<RM, RN>(A.P, A.P.inv_d, A.P.two_over_d, A.P.eta, A.S, A.c, num, den, bx * 16 * RM, by * 16 * RN,
                                     b_begin, b_end, by == 0);
        }
    }
    grid_barrier(A.bar);
    // ---- phase 2: merge + statistics of the new codebook + clean accumulators ----
    const int lane = threadIdx.x & 31, gwarp = tid >> 5, nwarps = nthr >> 5;
    for (int row = gwarp; row < A.k_pad; row += nwarps) {
        const bool real = row < K;
        double s = 0.0;
        float m = 0.f;
        if (real) {
            const size_t block = (size_t)K * D + K;
            float dn;
            if (A.slices > 1) {
                dn = 0.f;
                for (int z = 0; z < A.slices; ++z) dn += A.partials[(size_t)z * block + (size_t)K * D + row];
            } else {
                dn = A.den[row];
            }
            for (int cc = lane; cc < D; cc += 32) {
                const int64_t e = (int64_t)row * D + cc;
                float v = A.W[e];
                if (dn != 0.f) {
                    float nm;
                    if (A.slices > 1) {
                        nm = 0.f;
                        for (int z = 0; z < A.slices; ++z) nm += A.partials[(size_t)z * block + e];     // slice order
                    } else {
                        nm = A.num[e];
                    }
                    v = __fdiv_rn(nm, dn); A.W[e] = v;                              // xpysom.py:451-455
                }
                s += (double)v * v; m = fmaxf(m, fabsf(v));
            }
        }
        codebook_stats_row(row, real, s, m, lane, A.dist_kind, A.aux, A.bias, A.amax, A.gstat);
    }
    if (!A.do_split) return;
    grid_barrier(A.bar);
    // ---- phase 3: operand copies for the next BMU search ----
    for (int row = gwarp; row < A.k_pad; row += nwarps)
        codebook_split_row(A.W, row, row < K, lane, D, A.dist_kind, A.split, A.aux, A.bias, A.amax, A.gstat);
}

}  // namespace somb200
