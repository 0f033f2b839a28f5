// K2: shared-memory tiled SIMT distance + fused first-minimum argmin.
//
// Replaces, per block of samples, the reference's distance matrix + argmin
// (xpysom.py:410-417 with distances.py:11-23 / 45-59 / 61-75 / 137-158) without
// ever materialising the (n, K) matrix.  Plain fp32 FMA chains, so this is both
// the kernel for the distances that are not contractions (Manhattan,
// Chebyshev, norm-p) and the exact-fp32 cross-check of the tensor-core kernel.
//
// Tiling: one CTA owns TM=128 sample rows and walks all K neurons in tiles of
// TN=128; features are streamed through shared memory DK=16 at a time; each of
// the 256 threads keeps an 8x8 register tile.  Roofline: SIMT fp32, 2 ops per
// (row, neuron, feature) triple; X is read once per 128 rows from HBM, the
// codebook from L2.
#pragma once
#include "common.cuh"

namespace somb200 {

constexpr int ST_TM = 128, ST_TN = 128, ST_DK = 16, ST_THREADS = 256;
constexpr int ST_LD = ST_TM + 4;  // padded leading dimension of the transposed smem tiles

template <int DIST>
__device__ __forceinline__ void simt_accum(float &acc, float x, float w, float p) {
    if (DIST == SOM_DIST_EUCLIDEAN || DIST == SOM_DIST_COSINE) {
        acc = fmaf(x, w, acc);
    } else if (DIST == SOM_DIST_MANHATTAN) {
        acc += fabsf(x - w);
    } else if (DIST == SOM_DIST_CHEBYSHEV) {
        acc = fmaxf(acc, fabsf(x - w));
    } else {  // NORM_P: p is 2, 3, 4 (fast paths) or anything else through powf
        float t = fabsf(x - w);
        if (p == 2.f)      acc = fmaf(t, t, acc);
        else if (p == 3.f) acc = fmaf(t * t, t, acc);
        else if (p == 4.f) { float t2 = t * t; acc = fmaf(t2, t2, acc); }
        else               acc += powf(t, p);
    }
}

// score that is minimised; aux = |w|^2 (euclidean) or 1/|w| (cosine, 0 for a zero neuron)
template <int DIST>
__device__ __forceinline__ float simt_score(float acc, float aux) {
    if (DIST == SOM_DIST_EUCLIDEAN) return fmaf(-2.f, acc, aux);   // distances.py:23
    if (DIST == SOM_DIST_COSINE)    return -(acc * aux);           // argmin(1 - sim) == argmin(-x.w/|w|)
    return acc;
}

template <int DIST>
__global__ void __launch_bounds__(ST_THREADS, 2)
bmu_simt_kernel(const float *__restrict__ X, int64_t n, int d, int64_t ldx,
                const float *__restrict__ W, int k, const float *__restrict__ aux, float p,
                int32_t *__restrict__ bmu_out, float *__restrict__ best_out, int vec_ok) {
    __shared__ __align__(16) float Xs[ST_DK][ST_LD];
    __shared__ __align__(16) float Ws[ST_DK][ST_LD];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int lrow = tid >> 1, lcol = (tid & 1) * 8;  // loader mapping: 2 threads per tile row, 8 floats each

    for (int64_t row0 = (int64_t)blockIdx.x * ST_TM; row0 < n; row0 += (int64_t)gridDim.x * ST_TM) {
        float best[8];
        int   bidx[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { best[i] = INFINITY; bidx[i] = 0x7fffffff; }

        for (int n0 = 0; n0 < k; n0 += ST_TN) {
            float acc[8][8];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

            for (int d0 = 0; d0 < d; d0 += ST_DK) {
                // ---- stage X[row0:row0+128, d0:d0+16] and W[n0:n0+128, d0:d0+16], transposed
                float xv[8], wv[8];
                const int64_t gr = row0 + lrow;
                const int     gn = n0 + lrow;
                const int     gc = d0 + lcol;
                if (vec_ok && gc + 8 <= d) {
                    if (gr < n) {
                        const float4 *px = reinterpret_cast<const float4 *>(X + gr * ldx + gc);
                        float4 a = __ldg(px), b = __ldg(px + 1);
                        xv[0] = a.x; xv[1] = a.y; xv[2] = a.z; xv[3] = a.w;
                        xv[4] = b.x; xv[5] = b.y; xv[6] = b.z; xv[7] = b.w;
                    } else {
#pragma unroll
                        for (int e = 0; e < 8; ++e) xv[e] = 0.f;
                    }
                    if (gn < k) {
                        const float4 *pw = reinterpret_cast<const float4 *>(W + (int64_t)gn * d + gc);
                        float4 a = __ldg(pw), b = __ldg(pw + 1);
                        wv[0] = a.x; wv[1] = a.y; wv[2] = a.z; wv[3] = a.w;
                        wv[4] = b.x; wv[5] = b.y; wv[6] = b.z; wv[7] = b.w;
                    } else {
#pragma unroll
                        for (int e = 0; e < 8; ++e) wv[e] = 0.f;
                    }
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        xv[e] = (gr < n && gc + e < d) ? __ldg(X + gr * ldx + gc + e) : 0.f;
                        wv[e] = (gn < k && gc + e < d) ? __ldg(W + (int64_t)gn * d + gc + e) : 0.f;
                    }
                }
                __syncthreads();  // previous chunk fully consumed
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    Xs[lcol + e][lrow] = xv[e];
                    Ws[lcol + e][lrow] = wv[e];
                }
                __syncthreads();

                // zero padding of the feature tail is neutral for every distance here
                // (x = w = 0 adds |0-0|^p = 0, max(.,0), 0*0).
#pragma unroll
                for (int kk = 0; kk < ST_DK; ++kk) {
                    float a[8], b[8];
                    const float4 a0 = *reinterpret_cast<const float4 *>(&Xs[kk][ty * 8]);
                    const float4 a1 = *reinterpret_cast<const float4 *>(&Xs[kk][ty * 8 + 4]);
                    const float4 b0 = *reinterpret_cast<const float4 *>(&Ws[kk][tx * 4]);
                    const float4 b1 = *reinterpret_cast<const float4 *>(&Ws[kk][64 + tx * 4]);
                    a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w;
                    a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
                    b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w;
                    b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int j = 0; j < 8; ++j) simt_accum<DIST>(acc[i][j], a[i], b[j], p);
                }
            }

            // ---- fused argmin over this tile's columns, increasing index order, strict <
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int col = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
                if (col < k) {
                    const float ax = (DIST == SOM_DIST_EUCLIDEAN || DIST == SOM_DIST_COSINE) ? __ldg(aux + col) : 0.f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float s = simt_score<DIST>(acc[i][j], ax);
                        if (s < best[i]) { best[i] = s; bidx[i] = col; }
                    }
                }
            }
        }

        // ---- merge the 16 threads that share a row (lanes tx=0..15 of a half warp)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float v = best[i];
            int   ix = bidx[i];
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                const float ov = __shfl_xor_sync(0xffffffffu, v, o);
                const int   oi = __shfl_xor_sync(0xffffffffu, ix, o);
                argmin_merge(v, ix, ov, oi);
            }
            const int64_t gr = row0 + ty * 8 + i;
            if (tx == 0 && gr < n) {
                // a row whose every score is NaN/inf keeps index 0, like numpy's argmin on equal values
                bmu_out[gr] = (ix == 0x7fffffff) ? 0 : ix;
                if (best_out) best_out[gr] = v;
            }
        }
    }
}

// Distance MATRIX (n, K), for the callers that really want it (activate, distance_from_weights,
// topographic_error: xpysom.py:323-354, 647-671, 709-746).  Same tiling as the BMU kernel; `mode` selects the
// value written:  0 = the activation distance as the reference defines it (euclidean: partial
// -2 x.w + |w|^2, distances.py:23; cosine: 1 - nan_to_num(x.w / sqrt(|x|^2 |w|^2)), distances.py:55-59;
// manhattan / chebyshev / norm_p sums), 1 = Euclidean distance sqrt(max(0, |x|^2 - 2 x.w + |w|^2)) with
// nan_to_num (distances.py:33-43), whatever DIST, 2 = the full squared Euclidean distance
// (-2 x.w + |w|^2) + |x|^2 of 'euclidean_no_opt' (distances.py:25-31).
template <int DIST>
__global__ void __launch_bounds__(ST_THREADS, 2)
dist_matrix_kernel(const float *__restrict__ X, int64_t n, int d, int64_t ldx, const float *__restrict__ W, int k,
                   const float *__restrict__ wsq, float p, int mode, float *__restrict__ out) {
    __shared__ __align__(16) float Xs[ST_DK][ST_LD];
    __shared__ __align__(16) float Ws[ST_DK][ST_LD];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int lrow = tid >> 1, lcol = (tid & 1) * 8;
    const int64_t row0 = (int64_t)blockIdx.x * ST_TM;
    const int n0 = blockIdx.y * ST_TN;
    float acc[8][8], xs[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { xs[i] = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f; }
    for (int d0 = 0; d0 < d; d0 += ST_DK) {
        float xv[8], wv[8];
        const int64_t gr = row0 + lrow;
        const int gn = n0 + lrow, gc = d0 + lcol;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            xv[e] = (gr < n && gc + e < d) ? __ldg(X + gr * ldx + gc + e) : 0.f;
            wv[e] = (gn < k && gc + e < d) ? __ldg(W + (int64_t)gn * d + gc + e) : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < 8; ++e) { Xs[lcol + e][lrow] = xv[e]; Ws[lcol + e][lrow] = wv[e]; }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < ST_DK; ++kk) {
            float a[8], b[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = Xs[kk][ty * 8 + i];
#pragma unroll
            for (int j = 0; j < 8; ++j) b[j] = Ws[kk][j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4)];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                xs[i] = fmaf(a[i], a[i], xs[i]);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (mode != 0) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
                    else simt_accum<DIST>(acc[i][j], a[i], b[j], p);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int col = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
        if (col >= k) continue;
        const float ws = __ldg(wsq + col);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int64_t gr = row0 + ty * 8 + i;
            if (gr >= n) continue;
            float v;
            if (mode == 1) {
                v = sqrtf((fmaf(-2.f, acc[i][j], ws)) + xs[i]);             // (-2 x.w + |w|^2) + |x|^2, then sqrt
                if (isnan(v)) v = 0.f;                                       // negative round-off -> NaN -> 0
            } else if (mode == 2) {
                v = fmaf(-2.f, acc[i][j], ws) + xs[i];
            } else if (DIST == SOM_DIST_EUCLIDEAN) {
                v = fmaf(-2.f, acc[i][j], ws);
            } else if (DIST == SOM_DIST_COSINE) {
                float q = acc[i][j] / sqrtf(xs[i] * ws);
                if (isnan(q)) q = 0.f;
                else if (isinf(q)) q = q > 0.f ? 3.4028234664e38f : -3.4028234664e38f;
                v = 1.f - q;
            } else {
                v = acc[i][j];
            }
            out[gr * (int64_t)k + col] = v;
        }
    }
}


// Best and second-best matching unit per row on the Euclidean distance sqrt(|x|^2 - 2 x.w + |w|^2) with nan_to_num
// (distances.py:33-43), fused: what topographic_error needs (xpysom.py:709-746 argsorts the full (n, K) matrix and
// keeps two columns).  Same tiling as bmu_simt_kernel; each thread keeps (d1, i1, d2, i2) for its 8 rows and the 16
// threads of a row merge by shuffles.  Ties are ordered by index (numpy's argsort leaves them unspecified).
__device__ __forceinline__ void top2_insert(float &d1, int &i1, float &d2, int &i2, float v, int idx) {
    if (v < d1 || (v == d1 && idx < i1)) { d2 = d1; i2 = i1; d1 = v; i1 = idx; }
    else if (v < d2 || (v == d2 && idx < i2)) { d2 = v; i2 = idx; }
}

__global__ void __launch_bounds__(ST_THREADS, 1)
top2_simt_kernel(const float *__restrict__ X, int64_t n, int d, int64_t ldx, const float *__restrict__ W, int k,
                 const float *__restrict__ wsq, int32_t *__restrict__ out2) {
    __shared__ __align__(16) float Xs[ST_DK][ST_LD];
    __shared__ __align__(16) float Ws[ST_DK][ST_LD];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int lrow = tid >> 1, lcol = (tid & 1) * 8;
    for (int64_t row0 = (int64_t)blockIdx.x * ST_TM; row0 < n; row0 += (int64_t)gridDim.x * ST_TM) {
        float d1[8], d2[8];
        int i1[8], i2[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { d1[i] = d2[i] = INFINITY; i1[i] = i2[i] = 0x7fffffff; }
        for (int n0 = 0; n0 < k; n0 += ST_TN) {
            float acc[8][8], xs[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) { xs[i] = 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = 0.f; }
            for (int d0 = 0; d0 < d; d0 += ST_DK) {
                float xv[8], wv[8];
                const int64_t gr = row0 + lrow;
                const int gn = n0 + lrow, gc = d0 + lcol;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    xv[e] = (gr < n && gc + e < d) ? __ldg(X + gr * ldx + gc + e) : 0.f;
                    wv[e] = (gn < k && gc + e < d) ? __ldg(W + (int64_t)gn * d + gc + e) : 0.f;
                }
                __syncthreads();
#pragma unroll
                for (int e = 0; e < 8; ++e) { Xs[lcol + e][lrow] = xv[e]; Ws[lcol + e][lrow] = wv[e]; }
                __syncthreads();
#pragma unroll
                for (int kk = 0; kk < ST_DK; ++kk) {
                    float a[8], b[8];
                    const float4 a0 = *reinterpret_cast<const float4 *>(&Xs[kk][ty * 8]);
                    const float4 a1 = *reinterpret_cast<const float4 *>(&Xs[kk][ty * 8 + 4]);
                    const float4 b0 = *reinterpret_cast<const float4 *>(&Ws[kk][tx * 4]);
                    const float4 b1 = *reinterpret_cast<const float4 *>(&Ws[kk][64 + tx * 4]);
                    a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
                    b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        xs[i] = fmaf(a[i], a[i], xs[i]);
#pragma unroll
                        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int col = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
                if (col < k) {
                    const float ws = __ldg(wsq + col);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float v = sqrtf(fmaf(-2.f, acc[i][j], ws) + xs[i]);
                        if (isnan(v)) v = 0.f;
                        top2_insert(d1[i], i1[i], d2[i], i2[i], v, col);
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                const float od1 = __shfl_xor_sync(0xffffffffu, d1[i], o), od2 = __shfl_xor_sync(0xffffffffu, d2[i], o);
                const int oi1 = __shfl_xor_sync(0xffffffffu, i1[i], o), oi2 = __shfl_xor_sync(0xffffffffu, i2[i], o);
                top2_insert(d1[i], i1[i], d2[i], i2[i], od1, oi1);
                top2_insert(d1[i], i1[i], d2[i], i2[i], od2, oi2);
            }
            const int64_t gr = row0 + ty * 8 + i;
            if (tx == 0 && gr < n) {
                out2[2 * gr] = i1[i] == 0x7fffffff ? 0 : i1[i];
                out2[2 * gr + 1] = i2[i] == 0x7fffffff ? (k > 1 ? 1 : 0) : i2[i];
            }
        }
    }
}

inline int launch_top2_simt(const float *X, int64_t n, int d, int64_t ldx, const float *W, int k, const float *wsq,
                            int32_t *out2, int sm_count, cudaStream_t st) {
    int64_t tiles = ceil_div(n, ST_TM);
    int grid = (int)(tiles < (int64_t)sm_count * 8 ? tiles : (int64_t)sm_count * 8);
    if (grid < 1) grid = 1;
    top2_simt_kernel<<<grid, ST_THREADS, 0, st>>>(X, n, d, ldx, W, k, wsq, out2);
    return check_cuda(cudaGetLastError(), "top2_simt_kernel launch");
}

inline int launch_dist_matrix(const float *X, int64_t n, int d, int64_t ldx, const float *W, int k, int dist_kind,
                              float p, int mode, const float *wsq, float *out, cudaStream_t st) {
    dim3 grid((unsigned)ceil_div(n, ST_TM), (unsigned)ceil_div(k, ST_TN));
#define SOM_LAUNCH_DM(DK) dist_matrix_kernel<DK><<<grid, ST_THREADS, 0, st>>>(X, n, d, ldx, W, k, wsq, p, mode, out)
    switch (dist_kind) {
        case SOM_DIST_EUCLIDEAN: SOM_LAUNCH_DM(SOM_DIST_EUCLIDEAN); break;
        case SOM_DIST_COSINE:    SOM_LAUNCH_DM(SOM_DIST_COSINE); break;
        case SOM_DIST_MANHATTAN: SOM_LAUNCH_DM(SOM_DIST_MANHATTAN); break;
        case SOM_DIST_CHEBYSHEV: SOM_LAUNCH_DM(SOM_DIST_CHEBYSHEV); break;
        case SOM_DIST_NORM_P:    SOM_LAUNCH_DM(SOM_DIST_NORM_P); break;
        default: set_error("unknown distance kind %d", dist_kind); return SOM_E_BADARG;
    }
#undef SOM_LAUNCH_DM
    return check_cuda(cudaGetLastError(), "dist_matrix_kernel launch");
}

inline int launch_bmu_simt(const float *X, int64_t n, int d, int64_t ldx, const float *W, int k,
                           int dist_kind, float p, const float *aux, int32_t *bmu, float *best,
                           int sm_count, cudaStream_t st) {
    const int vec_ok = (d % 4 == 0) && (ldx % 4 == 0) &&
                       ((reinterpret_cast<uintptr_t>(X) & 15) == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
    int64_t tiles = ceil_div(n, ST_TM);
    int grid = (int)(tiles < (int64_t)sm_count * 8 ? tiles : (int64_t)sm_count * 8);
    if (grid < 1) grid = 1;
#define SOM_LAUNCH_SIMT(DK)                                                                         \
    bmu_simt_kernel<DK><<<grid, ST_THREADS, 0, st>>>(X, n, d, ldx, W, k, aux, p, bmu, best, vec_ok)
    switch (dist_kind) {
        case SOM_DIST_EUCLIDEAN: SOM_LAUNCH_SIMT(SOM_DIST_EUCLIDEAN); break;
        case SOM_DIST_COSINE:    SOM_LAUNCH_SIMT(SOM_DIST_COSINE); break;
        case SOM_DIST_MANHATTAN: SOM_LAUNCH_SIMT(SOM_DIST_MANHATTAN); break;
        case SOM_DIST_CHEBYSHEV: SOM_LAUNCH_SIMT(SOM_DIST_CHEBYSHEV); break;
        case SOM_DIST_NORM_P:    SOM_LAUNCH_SIMT(SOM_DIST_NORM_P); break;
        default: set_error("unknown distance kind %d", dist_kind); return SOM_E_BADARG;
    }
#undef SOM_LAUNCH_SIMT
    return check_cuda(cudaGetLastError(), "bmu_simt_kernel launch");
}

}  // namespace somb200
