// K3: per-BMU sample sums.  S[bmu[r], :] += X[r, :],  c[bmu[r]] += 1.
//
// This is the sample side of the reference's second GEMM g^T X and of sum(g)
// (xpysom.py:436-440): because h(bmu, k) depends on the sample only through its
// BMU, g^T X == H^T S.  HBM-bound streaming pass: 4*D bytes
// per sample read once, one 16-byte vector reduction per 4 features into the
// L2-resident (K, D) accumulator.  Counts go through a per-CTA shared-memory
// integer histogram so that fp32 increments are never lost above 2^24.
#pragma once
#include "common.cuh"

namespace somb200 {

constexpr int ACC_THREADS = 256;
constexpr int ACC_HIST_MAX_K = 12288;   // 48 KB of int32 bins in shared memory

template <bool VEC, bool HIST>
__global__ void __launch_bounds__(ACC_THREADS)
accumulate_kernel(const float *__restrict__ X, int64_t n, int d, int64_t ldx,
                  const int32_t *__restrict__ bmu, int k,
                  float *__restrict__ S, float *__restrict__ c, int64_t rows_per_cta) {
    extern __shared__ int hist[];
    if (HIST) {
        for (int i = threadIdx.x; i < k; i += blockDim.x) hist[i] = 0;
        __syncthreads();
    }
    const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta;
    const int64_t r1 = r0 + rows_per_cta < n ? r0 + rows_per_cta : n;
    if (VEC) {
        const int d4 = d >> 2;
        const int64_t items = (r1 - r0) * d4;
        for (int64_t it = threadIdx.x; it < items; it += blockDim.x) {
            const int64_t r = r0 + it / d4;
            const int c4 = (int)(it % d4);
            const int b = __ldg(bmu + r);
            const float4 v = __ldcs(reinterpret_cast<const float4 *>(X + r * ldx) + c4);  // streaming: read once
            red_add_v4(S + (int64_t)b * d + c4 * 4, v);
            if (c4 == 0) {
                if (HIST) atomicAdd(&hist[b], 1);
                else      atomicAdd(c + b, 1.0f);
            }
        }
    } else {
        const int64_t items = (r1 - r0) * d;
        for (int64_t it = threadIdx.x; it < items; it += blockDim.x) {
            const int64_t r = r0 + it / d;
            const int col = (int)(it % d);
            const int b = __ldg(bmu + r);
            atomicAdd(S + (int64_t)b * d + col, __ldcs(X + r * ldx + col));
            if (col == 0) {
                if (HIST) atomicAdd(&hist[b], 1);
                else      atomicAdd(c + b, 1.0f);
            }
        }
    }
    if (HIST) {
        __syncthreads();
        for (int i = threadIdx.x; i < k; i += blockDim.x) {
            const int h = hist[i];
            if (h) atomicAdd(c + i, (float)h);
        }
    }
}

inline int launch_accumulate(const float *X, int64_t n, int d, int64_t ldx, const int32_t *bmu, int k,
                             float *S, float *c, int sm_count, cudaStream_t st) {
    if (n <= 0) return 0;
    const bool vec = (d % 4 == 0) && (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(S) & 15) == 0);
    const bool hist = k <= ACC_HIST_MAX_K;
    // a multiple of the SM count; each CTA takes a contiguous slab of rows so its loads are sequential
    int64_t ctas = (int64_t)sm_count * 8;
    int64_t rows_per_cta = ceil_div(n, ctas);
    if (rows_per_cta < 32) rows_per_cta = 32;
    ctas = ceil_div(n, rows_per_cta);
    const size_t smem = hist ? (size_t)k * sizeof(int) : 0;
#define SOM_LAUNCH_ACC(V, H)                                                                  \
    do {                                                                                      \
        if (smem > 48 * 1024)                                                                 \
            cudaFuncSetAttribute(accumulate_kernel<V, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        accumulate_kernel<V, H><<<(unsigned)ctas, ACC_THREADS, smem, st>>>(X, n, d, ldx, bmu, k, S, c, rows_per_cta); \
    } while (0)
    if (vec && hist) SOM_LAUNCH_ACC(true, true);
    else if (vec)    SOM_LAUNCH_ACC(true, false);
    else if (hist)   SOM_LAUNCH_ACC(false, true);
    else             SOM_LAUNCH_ACC(false, false);
#undef SOM_LAUNCH_ACC
    return check_cuda(cudaGetLastError(), "accumulate_kernel launch");
}

}  // namespace somb200
