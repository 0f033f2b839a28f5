// Small kernels around the epoch: codebook preparation (Q), merge (M),
// quantization error and the U-matrix.
#pragma once
#include <cuda_fp16.h>
#include "common.cuh"

namespace somb200 {

// round-to-nearest TF32 (10 explicit mantissa bits); result is an fp32 bit pattern with 13 zero low bits
__device__ __forceinline__ float tf32_rna(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

// Q, pass 1: one warp per (padded) neuron.
//   aux[k]  = |w_k|^2 (euclidean; xpysom.py:529-537)  or 1/|w_k| (cosine; 0 for a zero neuron,
//             which makes its similarity 0 like nan_to_num in distances.py:57)
//   bias[k] = additive term of the tensor-core epilogue: |w|^2 / 0, +inf on padding neurons
//   amax[k] = largest magnitude of the SCALED row w'_k (-2 w for euclidean: exact; -w/|w| for cosine);
//   gstat   = running max / min-non-zero of amax over the codebook (decides the fp16 scaling mode)
__device__ __forceinline__ float codebook_row_scale(int dist_kind, double s) {
    if (dist_kind == SOM_DIST_EUCLIDEAN) return -2.f;
    if (dist_kind == SOM_DIST_COSINE) return s > 0.0 ? -(float)(1.0 / sqrt(s)) : 0.f;
    return 0.f;
}

// one warp: statistics of neuron `row` from the values its lanes hold (lane l: columns l, l + 32, ...),
// s = partial sum of squares (fp64), m = partial max |w|
__device__ __forceinline__ void codebook_stats_row(int row, bool real, double s, float m, int lane, int dist_kind,
                                                   float *__restrict__ aux, float *__restrict__ bias,
                                                   float *__restrict__ amax, unsigned int *__restrict__ gstat) {
    s = warp_sum(s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane != 0) return;
    const float wsq = (float)s;
    float a = 0.f, b = INFINITY;
    if (real) {
        if (dist_kind == SOM_DIST_EUCLIDEAN) { a = wsq; b = wsq; }
        else if (dist_kind == SOM_DIST_COSINE) { a = -codebook_row_scale(dist_kind, s); b = 0.f; }
        else { a = wsq; b = 0.f; }
    }
    aux[row] = a; bias[row] = b;
    const float am = real ? m * fabsf(codebook_row_scale(dist_kind, s)) : 0.f;
    amax[row] = am;
    if (am > 0.f && isfinite(am)) {
        atomicMax(gstat + 0, __float_as_uint(am));
        atomicMax(gstat + 1, ~__float_as_uint(am));
    }
}

__global__ void codebook_stats_kernel(const float *__restrict__ W, int k, int d, int dist_kind, int k_pad,
                                      float *__restrict__ aux, float *__restrict__ bias, float *__restrict__ amax,
                                      unsigned int *__restrict__ gstat) {
    pdl_wait(); pdl_trigger();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= k_pad) return;
    const bool real = warp < k;
    double s = 0.0;
    float m = 0.f;
    if (real)
        for (int c = lane; c < d; c += 32) { const float v = W[(int64_t)warp * d + c]; s += (double)v * v; m = fmaxf(m, fabsf(v)); }
    codebook_stats_row(warp, real, s, m, lane, dist_kind, aux, bias, amax, gstat);
}

// Q, pass 2: operand copies of the scaled codebook for the tensor-core kernels.
//   whi/wlo      TF32 split of w'_k (zero on padding)
//   w16hi/w16lo  fp16 split of w'_k * 2^b_k, wsinv[k] = 2^-b_k.  When every non-zero neuron's amax is within
//                2^12 of the largest one, ONE power of two serves the whole codebook (b_k = b, gstat[2] = 1) and
//                the epilogue needs a single fused multiply-add per score; otherwise each neuron gets its own
//                b_k (exact, undone per column in the epilogue) so that tiny neurons keep their 22 bits.
struct SplitOut {
    float *whi, *wlo;            // (k_pad, d_pad) TF32 hi / lo
    __half *w16hi, *w16lo;       // (k_pad, d_pad64) fp16 hi / lo; d <= 16: w16hi rows are PACKED [hi | hi | lo | 0] x 16
    float *wsinv;                // (k_pad)
    float *wfold;                // fold operand image of the fp16 kernel: (bias_k 2^b_k) as three TF32 pieces per neuron
    int d_pad, d_pad64;
    int allow_fold;              // 0: keep the bias in the epilogue (A/B measurements)
};

// one warp: the operand copies of neuron `row`.  (No __restrict__ here on purpose: epoch_tail_kernel writes W, aux,
// bias and amax earlier in the SAME launch, so these loads must not go through the non-coherent path.)
__device__ __forceinline__ void codebook_split_row(const float *W, int row, bool real, int lane, int d, int dist_kind,
                                                   const SplitOut &O, const float *aux, const float *bias, const float *amax,
                                                   unsigned int *gstat) {
    float scale = 0.f;
    if (real) scale = dist_kind == SOM_DIST_EUCLIDEAN ? -2.f : -aux[row];     // aux = 1/|w| for cosine
    const float gmax = __uint_as_float(gstat[0]);
    const unsigned int nmin = gstat[1];
    const float gmin = nmin ? __uint_as_float(~nmin) : gmax;
    const bool uniform = !(gmax > 0.f) || gmin >= gmax * (1.f / 4096.f);
    // TF32 copies: when the last 32-feature block has three spare columns, the epilogue bias is FOLDED into the
    // contraction: columns d, d+1, d+2 of W'hi carry the three TF32 pieces of bias_k (33 mantissa bits, more than
    // the fp32 accumulator keeps) and the kernel sets the matching X columns to 1; gstat[3] = 1 tells it so.
    const bool fold = O.allow_fold && O.d_pad - d >= 3;
    if (row == 0 && lane == 0) { gstat[2] = uniform ? 1u : 0u; gstat[3] = fold ? 1u : 0u; }
    const float bk = bias[row];                        // |w|^2, 0 (cosine) or +inf (padding neuron)
    float b0 = tf32_rna(bk), b1 = 0.f, b2 = 0.f;
    if (isfinite(bk)) { b1 = tf32_rna(bk - b0); b2 = tf32_rna((bk - b0) - b1); }
    for (int c = lane; c < O.d_pad; c += 32) {
        float v = 0.f;
        if (real && c < d) v = W[(int64_t)row * d + c] * scale;
        float hi = tf32_rna(v), lo = tf32_rna(v - hi);
        if (fold && c >= d && c < d + 3) { hi = c == d ? b0 : (c == d + 1 ? b1 : b2); lo = 0.f; }
        O.whi[(int64_t)row * O.d_pad + c] = hi;
        O.wlo[(int64_t)row * O.d_pad + c] = lo;
    }
    const float ps = pow2_scale_for(uniform ? gmax : amax[row]);
    if (lane == 0) {
        O.wsinv[row] = 1.f / ps;             // exact: ps is a power of two
        // fold operand of the fp16 kernel (one extra kind::tf32 MMA adds 2^a_r * (2^b_k bias_k) to the accumulator, so its
        // epilogue only compares): three TF32 pieces of bias_k * 2^b_k in the first 16-byte K chunk of the neuron's row
        // of the no-swizzle K-major image [128-neuron block][K chunk 0 | K chunk 1][8-row group] (bmu_tc3.cuh)
        const float bs = bk * ps;                                      // +inf stays +inf on padding neurons
        float f0 = tf32_rna(bs), f1 = 0.f, f2 = 0.f;
        if (isfinite(bs)) { f1 = tf32_rna(bs - f0); f2 = tf32_rna((bs - f0) - f1); } else f0 = bs;
        float *blk = O.wfold + (size_t)(row >> 7) * 1024 + (size_t)((row & 127) >> 3) * 64 + (size_t)(row & 7) * 4;
        *reinterpret_cast<float4 *>(blk) = make_float4(f0, f1, f2, 0.f);            // K chunk 0
        *reinterpret_cast<float4 *>(blk + 32) = make_float4(0.f, 0.f, 0.f, 0.f);   // K chunk 1 (+128 bytes)
    }
    if (d <= 16) {
        // packed operand of the fp16 kernel for short rows: ONE 64-column tile carries the three products of the
        // split as three 16-column K steps, x^ = [hi | lo | hi | 0] against w^ = [hi | hi | lo | 0]
        const int c = lane & 15;
        float v = 0.f;
        if (real && c < d) v = W[(int64_t)row * d + c] * scale * ps;
        const __half hi = __float2half_rn(v), lo = __float2half_rn(v - __half2float(hi));
        __half *dst = O.w16hi + (int64_t)row * O.d_pad64;
        if (lane < 16) { dst[c] = hi; dst[16 + c] = hi; }
        else           { dst[32 + c] = lo; dst[48 + c] = __float2half_rn(0.f); }
    } else {
        for (int c = lane; c < O.d_pad64; c += 32) {
            float v = 0.f;
            if (real && c < d) v = W[(int64_t)row * d + c] * scale * ps;
            const __half hi = __float2half_rn(v);
            O.w16hi[(int64_t)row * O.d_pad64 + c] = hi;
            O.w16lo[(int64_t)row * O.d_pad64 + c] = __float2half_rn(v - __half2float(hi));
        }
    }
}

__global__ void codebook_split_kernel(const float *__restrict__ W, int k, int d, int dist_kind, int k_pad, SplitOut O,
                                      const float *__restrict__ aux, const float *__restrict__ bias,
                                      const float *__restrict__ amax, unsigned int *__restrict__ gstat) {
    pdl_wait(); pdl_trigger();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= k_pad) return;
    codebook_split_row(W, warp, warp < k, lane, d, dist_kind, O, aux, bias, amax, gstat);
}

// |w_k|^2 per neuron (fp64 accumulation, rounded once), one warp per row
__global__ void row_sq_kernel(const float *__restrict__ W, int k, int d, float *__restrict__ out) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= k) return;
    double s = 0.0;
    for (int c = lane; c < d; c += 32) { const double v = W[(int64_t)warp * d + c]; s += v * v; }
    s = warp_sum(s);
    if (lane == 0) out[warp] = (float)s;
}

// per-row power-of-two scale of the samples for the fp16-split kernel: xscale[r] = 2^a_r with
// max_c |x[r,c]| * 2^a_r in [2^14, 2^15).  One HBM pass, done once per upload.  VEC: 16-byte aligned rows,
// a group of LPR lanes (power of two >= d/4, at most 32) owns a row and 128-bit loads keep many rows in
// flight per warp; otherwise one warp per row with scalar loads.
template <bool VEC>
__global__ void __launch_bounds__(256)
row_scale_kernel(const float *__restrict__ X, int64_t n, int d, int64_t ldx, float *__restrict__ xscale, int lpr) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    if (VEC) {
        const int d4 = d >> 2, rpw = 32 / lpr, sub = lane / lpr, c0 = lane % lpr;
        for (int64_t r0 = warp * rpw; r0 < n; r0 += warps * rpw) {
            const int64_t r = r0 + sub;
            float amax = 0.f;
            if (r < n)
                for (int c4 = c0; c4 < d4; c4 += lpr) {
                    const float4 v = __ldg(reinterpret_cast<const float4 *>(X + r * ldx) + c4);
                    amax = fmaxf(amax, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));
                }
            for (int o = lpr >> 1; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
            if (c0 == 0 && r < n) xscale[r] = pow2_scale_for(amax);
        }
    } else {
        for (int64_t r = warp; r < n; r += warps) {
            float amax = 0.f;
            for (int c = lane; c < d; c += 32) amax = fmaxf(amax, fabsf(__ldg(X + r * ldx + c)));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
            if (lane == 0) xscale[r] = pow2_scale_for(amax);
        }
    }
}

// per-column largest magnitude of the samples (the scales of the exact accumulation, common.cuh: ExactAcc):
// colmax_bits[c] = max(colmax_bits[c], bits(|x[r, c]|)) over the rows -- non-negative floats order like their bit
// patterns, so the running maximum is an integer atomicMax.  One HBM pass, once per upload.  L threads (a power of
// two) cover the columns (float4 groups when VEC), 256 / L rows are in flight per CTA.
template <bool VEC>
__global__ void __launch_bounds__(256)
column_absmax_kernel(const float *__restrict__ X, int64_t n, int d, int64_t ldx, unsigned int *__restrict__ colmax_bits, int L) {
    __shared__ unsigned int cm[1024];
    const int t = threadIdx.x, R = 256 / L, sub = t / L, cl = t % L;
    const int units = VEC ? (d >> 2) : d;                   // column units: float4 groups or single columns
    for (int u0 = 0; u0 < units; u0 += L) {
        const int u = u0 + cl;
        float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f;
        if (u < units)
            for (int64_t r = (int64_t)blockIdx.x * R + sub; r < n; r += (int64_t)gridDim.x * R) {
                if (VEC) {
                    const float4 v = __ldg(reinterpret_cast<const float4 *>(X + r * ldx) + u);
                    m0 = fmaxf(m0, fabsf(v.x)); m1 = fmaxf(m1, fabsf(v.y)); m2 = fmaxf(m2, fabsf(v.z)); m3 = fmaxf(m3, fabsf(v.w));
                    // fmaxf drops NaNs: keep them visible (a NaN column gets scale 1 and poisons its sums, as in numpy)
                    if (v.x != v.x) m0 = v.x; if (v.y != v.y) m1 = v.y; if (v.z != v.z) m2 = v.z; if (v.w != v.w) m3 = v.w;
                } else {
                    const float v = __ldg(X + r * ldx + u);
                    m0 = fmaxf(m0, fabsf(v));
                    if (v != v) m0 = v;
                }
            }
        const int per = VEC ? 4 : 1;
        for (int i = t; i < L * per; i += 256) cm[i] = 0u;
        __syncthreads();
        if (u < units) {
            atomicMax(&cm[cl * per], __float_as_uint(fabsf(m0)));
            if (VEC) {
                atomicMax(&cm[cl * 4 + 1], __float_as_uint(fabsf(m1)));
                atomicMax(&cm[cl * 4 + 2], __float_as_uint(fabsf(m2)));
                atomicMax(&cm[cl * 4 + 3], __float_as_uint(fabsf(m3)));
            }
        }
        __syncthreads();
        for (int i = t; i < L * per; i += 256) {
            const int col = u0 * per + i;
            if (col < d && cm[i]) atomicMax(colmax_bits + col, cm[i]);
        }
        __syncthreads();
    }
}

// M: W <- den != 0 ? num/den : W      (xpysom.py:451-455)
__global__ void merge_kernel(float *__restrict__ W, const float *__restrict__ num, const float *__restrict__ den,
                             int k, int d) {
    pdl_wait(); pdl_trigger();
    const int64_t tot = (int64_t)k * d;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < tot; e += (int64_t)gridDim.x * blockDim.x) {
        const float dn = __ldg(den + e / d);
        if (dn != 0.f) W[e] = __fdiv_rn(num[e], dn);
    }
}

// quantization support: one warp per row.  q[r] = W[bmu[r]];  err[r] = ||x_r - W[bmu[r]]||_2
// (xpysom.py:632-645, 699-705: the quantization error subtracts the code vector and takes the norm).
__global__ void quantize_kernel(const float *__restrict__ X, int64_t n, int d, int64_t ldx,
                                const float *__restrict__ W, const int32_t *__restrict__ bmu,
                                float *__restrict__ q, float *__restrict__ err) {
    const int lane = threadIdx.x & 31;
    const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += warps) {
        const int b = __ldg(bmu + r);
        float s = 0.f;
        for (int c = lane; c < d; c += 32) {
            const float w = __ldg(W + (int64_t)b * d + c);
            if (q) q[r * d + c] = w;
            const float t = __ldg(X + r * ldx + c) - w;     // data -= quantization (fp32, :703)
            s = fmaf(t, t, s);
        }
        s = warp_sum(s);
        if (err && lane == 0) err[r] = sqrtf(s);
    }
}

// distance_map (xpysom.py:788-817): one warp per neuron, sum of ||w - w_neighbour|| over the
// 8 (rectangular) or 6 (hexagonal, offsets depend on the parity of j) grid neighbours.
__global__ void distance_map_kernel(const float *__restrict__ W, int gx, int gy, int d, int topology,
                                    float *__restrict__ um) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= gx * gy) return;
    const int x = warp / gy, y = warp % gy;
    const int rect_i[8] = {0, -1, -1, -1, 0, 1, 1, 1}, rect_j[8] = {-1, -1, 0, 1, 1, 1, 0, -1};
    const int hex_i_even[6] = {0, 1, 0, -1, -1, -1}, hex_i_odd[6] = {1, 1, 1, 0, -1, 0};
    const int hex_j[6] = {1, 0, -1, -1, 0, 1};
    const int nn = topology == SOM_TOPO_HEXAGONAL ? 6 : 8;
    double tot = 0.0;
    for (int e = 0; e < nn; ++e) {
        int di, dj;
        if (topology == SOM_TOPO_HEXAGONAL) { di = (y % 2 == 0) ? hex_i_even[e] : hex_i_odd[e]; dj = hex_j[e]; }
        else { di = rect_i[e]; dj = rect_j[e]; }
        const int xi = x + di, yj = y + dj;
        if (xi < 0 || xi >= gx || yj < 0 || yj >= gy) continue;
        double s = 0.0;
        for (int c = lane; c < d; c += 32) {
            const double t = (double)W[(int64_t)warp * d + c] - (double)W[((int64_t)xi * gy + yj) * d + c];
            s += t * t;
        }
        s = warp_sum(s);
        tot += sqrt(s);
    }
    if (lane == 0) um[warp] = (float)tot;
}

}  // namespace somb200
