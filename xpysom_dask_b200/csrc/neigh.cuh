// K4: neighbourhood apply.  num = eta * H^T S,  den = eta * H^T c  with
// H[b, k] = h(bmu = b, neuron = k) evaluated on the fly (never a K x K matrix
// in memory: 400 MB at K = 10^4).
//
// Replaces neighborhoods.py:14-130 (gaussian / mexican_hat / bubble / triangle
// on rectangular and hexagonal maps, with compact_support) together with the
// neighbourhood side of XPySom._update (xpysom.py:434-441).
//
// Every neighbourhood of the reference is built from per-axis factors, so a
// tiny fp64 pre-kernel tabulates them once per epoch:
//   TX[q][bi][i], TY[bj][j]     q = hexagonal row-parity shift index
// and the apply kernel turns two L1-resident table reads into h:
//   product form  h = TX * TY                 (gaussian, bubble, triangle)
//   mexican hat   h = exp(-p/d) (1 - 2p/d),  p = PX * mask + PY
// Work: 2 K^2 D flops, SIMT fp32, tiled 64 neurons x 64 features per CTA.
#pragma once
#include "common.cuh"

namespace somb200 {

struct NeighParams {
    int    gx, gy, d;
    int    topology, kind, compact;
    int    shifted;       // 1: real (hexagonal) coordinates are used -> 3 parity planes
    int    mex_rect_quirk;  // rectangular mexican_hat + compact: the y-window is applied at index i
    float  inv_d;         // 1/d,   d = 2 std_coeff^2 sigma^2
    float  two_over_d;    // 2/d
    float  eta;
    // optional device-side schedule (CUDA-graph replay): sigma = sched[2e], eta = sched[2e+1], e = *epoch
    const double *sched;
    const int    *epoch;
    double std_coeff;
    const float *tx, *ty;     // product factors or squared offsets
    const float *mx, *my;     // compact-support windows (mexican hat only)
};

__device__ __forceinline__ void neigh_resolve(const NeighParams &P, float &inv_d, float &two_over_d, float &eta) {
    if (P.sched == nullptr) { inv_d = P.inv_d; two_over_d = P.two_over_d; eta = P.eta; return; }
    const int e = *P.epoch;
    const double sigma = P.sched[2 * e];
    const double dd = 2.0 * P.std_coeff * P.std_coeff * sigma * sigma;
    inv_d = (float)(1.0 / dd); two_over_d = (float)(2.0 / dd); eta = (float)P.sched[2 * e + 1];
}

// hexagonal rule of xpysom.py:201-206: rows (gy-1-j) even are shifted by -0.5
__host__ __device__ inline int hex_shift(int j, int gy) { return ((gy - 1 - j) & 1) == 0 ? 1 : 0; }

// one thread per table entry (grid-stride over `nthr` threads); all arithmetic in fp64, stored as fp32
__device__ __forceinline__ void neigh_tables_fill(int gx, int gy, int kind, int compact, int shifted, double sigma, double dd,
                                                  float *tx, float *ty, float *mx, float *my, int tid, int nthr) {
    const int nxe = 3 * gx * gx, nye = gy * gy;
    for (int e = tid; e < nxe + nye; e += nthr) {
        const bool isx = e < nxe;
        int q = 1, c, nn;
        if (isx) { q = e / (gx * gx); c = (e / gx) % gx; nn = e % gx; }
        else     { int f = e - nxe; c = f / gy; nn = f % gy; }
        // offset along the axis in the coordinates the reference uses for this function
        const double shift = (isx && shifted) ? 0.5 * (double)(q - 1) : 0.0;
        const double delta = (double)(nn - c) + shift;       // n - c  (neighborhoods.py:26,48)
        const bool   inwin = (delta > -sigma) && (delta < sigma);  // strict window (:30, :108)
        const float  sq = (float)(delta * delta);            // power(..., 2, dtype=float32), exact for .5 grids
        float val = 0.f, msk = 1.f;
        if (kind == SOM_NEIGH_GAUSSIAN) {
            double v = exp(-(double)sq / dd);
            if (compact && !inwin) v = 0.0;
            val = (float)v;
        } else if (kind == SOM_NEIGH_BUBBLE) {
            val = inwin ? 1.f : 0.f;
        } else if (kind == SOM_NEIGH_TRIANGLE) {
            double v = sigma - fabs(delta);                   // (-|c - n|) + sigma  (:121)
            if (v < 0.0) v = 0.0;
            if (compact && !inwin) v = 0.0;
            val = (float)v;
        } else {  // mexican hat: tabulate the squared offset and the window
            val = sq;
            msk = inwin ? 1.f : 0.f;
        }
        if (isx) { tx[e] = val; mx[e] = msk; }
        else     { ty[e - nxe] = val; my[e - nxe] = msk; }
    }
}

__global__ void neigh_tables_kernel(int gx, int gy, int kind, int compact, int shifted,
                                    double sigma, double dd, float *tx, float *ty, float *mx, float *my,
                                    const double *sched, const int *epoch, double std_coeff) {
    pdl_wait(); pdl_trigger();
    if (sched != nullptr) {          // schedule read on the device (graph replay)
        sigma = sched[2 * (*epoch)];
        dd = 2.0 * std_coeff * std_coeff * sigma * sigma;
    }
    neigh_tables_fill(gx, gy, kind, compact, shifted, sigma, dd, tx, ty, mx, my, blockIdx.x * blockDim.x + threadIdx.x,
                      gridDim.x * blockDim.x);
}

__device__ __forceinline__ float neigh_eval(const NeighParams &P, float inv_d, float two_over_d, int bi, int bj, int i, int j) {
    const int q = P.shifted ? (hex_shift(bj, P.gy) - hex_shift(j, P.gy) + 1) : 1;
    const int xe = (q * P.gx + bi) * P.gx + i;
    const int ye = bj * P.gy + j;
    // (plain loads, not __ldg: epoch_tail_kernel fills the tables earlier in the same launch)
    const float fx = P.tx[xe], fy = P.ty[ye];
    if (P.kind != SOM_NEIGH_MEXICAN_HAT) return fx * fy;
    float px = fx;
    if (P.compact) {
        // neighborhoods.py:69-71 / 91-93: px is multiplied by both windows, py by none.
        const float wy = P.mex_rect_quirk ? P.my[bj * P.gy + i] : P.my[ye];
        px *= P.mx[xe] * wy;
    }
    const float p = px + fy;
    return expf(-p * inv_d) * (1.f - two_over_d * p);
}

// Tile = (16 RM) neurons x (16 RN) features per CTA of 256 threads, RM x RN outputs per thread, 16 BMUs per step.
// <4, 4> (64 x 64) keeps small maps / few features busy; <8, 8> (128 x 128) has 4x the FMAs per H evaluation and
// per shared-memory byte, which is what the non-separable maps with many features need (config 5: hexagonal
// mexican hat, K = 2500, D = 128: 2 K^2 D = 1.6 GFLOP per epoch).
constexpr int NB_K = 16, NB_THREADS = 256;

// One output tile: neurons [k0, k0 + 16 RM) x features [n0, n0 + 16 RN), BMUs [b_begin, b_end), written with plain
// stores to num / den -- when the reduction over BMUs is sliced, the caller passes the slice's own partial buffers
// and the slices are summed afterwards in slice order (no atomics: the result does not depend on scheduling).
// `with_den`: this tile also produces den (one tile per neuron range does).
template <int RM, int RN>
__device__ __forceinline__ void neigh_apply_tile(const NeighParams &P, float inv_d, float two_over_d, float eta,
                                                 const float *S, const float *c, float *num, float *den,
                                                 int k0, int n0, int b_begin, int b_end, bool with_den) {
    constexpr int TM = 16 * RM, TN = 16 * RN;
    __shared__ __align__(16) float Hs[NB_K][TM + 4];
    __shared__ __align__(16) float Ss[NB_K][TN + 4];
    __shared__ float cs[NB_K];
    const int K = P.gx * P.gy, D = P.d;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;

    // this thread generates H entries for b = b0 + (tid >> 4), k = k0 + (tid & 15) * RM + {0..RM-1}
    const int hb = tid >> 4, hk = (tid & 15) * RM, hn = (tid & 15) * RN;
    int ki[RM], kj[RM];
#pragma unroll
    for (int e = 0; e < RM; ++e) {
        int kk = k0 + hk + e; if (kk >= K) kk = K - 1;
        ki[e] = kk / P.gy; kj[e] = kk % P.gy;
    }
    // and loads S entries for b = b0 + (tid >> 4), cols n0 + (tid & 15) * RN + {0..RN-1}
    float acc[RM][RN];
#pragma unroll
    for (int a = 0; a < RM; ++a)
#pragma unroll
        for (int b = 0; b < RN; ++b) acc[a][b] = 0.f;
    float dacc[RM];
#pragma unroll
    for (int a = 0; a < RM; ++a) dacc[a] = 0.f;

    for (int b0 = b_begin; b0 < b_end; b0 += NB_K) {
        const int b = b0 + hb;
        float hv[RM], sv[RN];
#pragma unroll
        for (int e = 0; e < RM; ++e) hv[e] = 0.f;
#pragma unroll
        for (int e = 0; e < RN; ++e) sv[e] = 0.f;
        float cb = 0.f;
        if (b < b_end) {
            cb = c[b];         // (plain loads: epoch_tail_kernel clears S and c later in the same launch)
            if (cb != 0.f) {   // an empty BMU contributes nothing (S[b] = 0 as well)
                const int bi = b / P.gy, bj = b % P.gy;
#pragma unroll
                for (int e = 0; e < RM; ++e) hv[e] = neigh_eval(P, inv_d, two_over_d, bi, bj, ki[e], kj[e]);
#pragma unroll
                for (int e = 0; e < RN; ++e) {
                    const int col = n0 + hn + e;
                    sv[e] = col < D ? S[(int64_t)b * D + col] : 0.f;
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int e = 0; e < RM; ++e) Hs[hb][hk + e] = hv[e];
#pragma unroll
        for (int e = 0; e < RN; ++e) Ss[hb][hn + e] = sv[e];
        if (hk == 0) cs[hb] = cb;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < NB_K; ++kk) {
            float h[RM], s[RN];
#pragma unroll
            for (int a = 0; a < RM; a += 4) {
                const float4 v = *reinterpret_cast<const float4 *>(&Hs[kk][ty * RM + a]);
                h[a] = v.x; h[a + 1] = v.y; h[a + 2] = v.z; h[a + 3] = v.w;
            }
#pragma unroll
            for (int bb = 0; bb < RN; bb += 4) {
                const float4 v = *reinterpret_cast<const float4 *>(&Ss[kk][tx * RN + bb]);
                s[bb] = v.x; s[bb + 1] = v.y; s[bb + 2] = v.z; s[bb + 3] = v.w;
            }
#pragma unroll
            for (int a = 0; a < RM; ++a)
#pragma unroll
                for (int bb = 0; bb < RN; ++bb) acc[a][bb] = fmaf(h[a], s[bb], acc[a][bb]);
            if (tx == 0) {
                const float cv = cs[kk];
#pragma unroll
                for (int a = 0; a < RM; ++a) dacc[a] = fmaf(h[a], cv, dacc[a]);
            }
        }
    }
#pragma unroll
    for (int a = 0; a < RM; ++a) {
        const int kk = k0 + ty * RM + a;
        if (kk >= K) continue;
#pragma unroll
        for (int bb = 0; bb < RN; ++bb) {
            const int col = n0 + tx * RN + bb;
            if (col < D) num[(int64_t)kk * D + col] = acc[a][bb] * eta;       // g = h * eta (xpysom.py:434)
        }
        if (tx == 0 && with_den) den[kk] = dacc[a] * eta;
    }
    __syncthreads();       // the shared tiles may be reused by the caller's next tile
}

template <int RM, int RN>
__global__ void __launch_bounds__(NB_THREADS)
neigh_apply_kernel(NeighParams P, const float *__restrict__ S, const float *__restrict__ c,
                   float *__restrict__ num, float *__restrict__ den, float *__restrict__ partials, int b_per_slice) {
    pdl_wait(); pdl_trigger();
    float inv_d, two_over_d, eta;
    neigh_resolve(P, inv_d, two_over_d, eta);
    // gridDim.z slices the reduction over BMUs so that small maps still fill the GPU; slice z writes its own
    // partial [num | den] block, slice_reduce_kernel sums them in slice order
    const int K = P.gx * P.gy;
    const int b_begin = blockIdx.z * b_per_slice;
    const int b_end = min(K, b_begin + b_per_slice);
    if (gridDim.z != 1) {
        num = partials + (size_t)blockIdx.z * ((size_t)K * P.d + K);
        den = num + (size_t)K * P.d;
    }
    neigh_apply_tile<RM, RN>(P, inv_d, two_over_d, eta, S, c, num, den, blockIdx.x * 16 * RM, blockIdx.y * 16 * RN, b_begin, b_end,
                             blockIdx.y == 0);
}

// [num | den] = sum over the slices, in slice order (fixed association: bit-reproducible)
__global__ void slice_reduce_kernel(const float *__restrict__ partials, int slices, int K, int D,
                                    float *__restrict__ num, float *__restrict__ den) {
    pdl_wait(); pdl_trigger();
    const size_t block = (size_t)K * D + K;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < block; e += (size_t)gridDim.x * blockDim.x) {
        float a = 0.f;
        for (int z = 0; z < slices; ++z) a += partials[(size_t)z * block + e];
        if (e < (size_t)K * D) num[e] = a; else den[e - (size_t)K * D] = a;
    }
}

// ---- separable path -------------------------------------------------------------------------------
// On a rectangular map the gaussian, bubble and triangle neighbourhoods are products of one factor
// per axis, h((bi,bj),(i,j)) = TX[bi][i] * TY[bj][j] (neighborhoods.py:33,112,130), so
//   num[i,j,:] = sum_bi TX[bi][i] * ( sum_bj TY[bj][j] * S[bi,bj,:] )
// is two small axis contractions, 2 K (gx+gy) D flops instead of 2 K^2 D (50x fewer at 100x100).
// out[a, j, c] = scale * sum_b M[b * ldm + j] * in[a, b, c]   for a < A, j < J, c < C  (in: (A, B, C))
constexpr int AX_THREADS = 128, AX_J = 16;
__global__ void __launch_bounds__(AX_THREADS)
axis_contract_kernel(const float *__restrict__ in, const float *__restrict__ M, int ldm, int A, int B, int J, int64_t C,
                     float scale, const double *__restrict__ sched, const int *__restrict__ epoch, float *__restrict__ out) {
    pdl_wait(); pdl_trigger();
    if (sched != nullptr) scale = (float)sched[2 * (*epoch) + 1];      // eta of the current epoch (graph replay)
    extern __shared__ float Ms[];                         // [B][AX_J] slice of M for this block's j range
    const int a = blockIdx.z, j0 = blockIdx.y * AX_J;
    const int64_t c = (int64_t)blockIdx.x * AX_THREADS + threadIdx.x;
    for (int e = threadIdx.x; e < B * AX_J; e += AX_THREADS) {
        const int b = e / AX_J, jj = e % AX_J;
        Ms[e] = (j0 + jj < J) ? __ldg(M + (int64_t)b * ldm + j0 + jj) : 0.f;
    }
    __syncthreads();
    if (c >= C) return;
    float acc[AX_J];
#pragma unroll
    for (int jj = 0; jj < AX_J; ++jj) acc[jj] = 0.f;
    const float *src = in + (int64_t)a * B * C + c;
    for (int b = 0; b < B; ++b) {
        const float x = __ldg(src + (int64_t)b * C);
        const float4 *m4 = reinterpret_cast<const float4 *>(Ms + b * AX_J);
#pragma unroll
        for (int q4 = 0; q4 < AX_J / 4; ++q4) {
            const float4 m = m4[q4];
            acc[q4 * 4 + 0] = fmaf(m.x, x, acc[q4 * 4 + 0]);
            acc[q4 * 4 + 1] = fmaf(m.y, x, acc[q4 * 4 + 1]);
            acc[q4 * 4 + 2] = fmaf(m.z, x, acc[q4 * 4 + 2]);
            acc[q4 * 4 + 3] = fmaf(m.w, x, acc[q4 * 4 + 3]);
        }
    }
    float *dst = out + (int64_t)a * J * C + c;
#pragma unroll
    for (int jj = 0; jj < AX_J; ++jj)
        if (j0 + jj < J) dst[(int64_t)(j0 + jj) * C] = acc[jj] * scale;
}

inline int launch_axis_contract(const float *in, const float *M, int ldm, int A, int B, int J, int64_t C, float scale,
                                const double *sched, const int *epoch, float *out, cudaStream_t st) {
    dim3 grid((unsigned)ceil_div(C, AX_THREADS), (unsigned)ceil_div(J, AX_J), (unsigned)A);
    const size_t smem = (size_t)B * AX_J * sizeof(float);
    return check_cuda(launch_pdl(axis_contract_kernel, grid, dim3(AX_THREADS), smem, st, in, M, ldm, A, B, J, C, scale, sched,
                                 epoch, out), "axis_contract_kernel launch");
}

// scratch (floats) after the factor tables: the (gx, gy, D) and (gx, gy) intermediates of the separable path
inline size_t neigh_separable_floats(int gx, int gy, int d) { return (size_t)gx * gy * ((size_t)d + 1); }

inline bool neigh_is_separable(int topology, int kind, int gx, int gy) {
    return topology == SOM_TOPO_RECTANGULAR && kind != SOM_NEIGH_MEXICAN_HAT && gx <= 512 && gy <= 512 &&
           (int64_t)gx * gy >= 4096;            // smaller maps: the direct kernel is one launch and already tiny
}

inline size_t neigh_table_floats(int gx, int gy) {
    return (size_t)2 * ((size_t)3 * gx * gx + (size_t)gy * gy) + 64;
}

// How the direct kernel slices the reduction over BMUs on a device with `sm_count` SMs (>= 2 CTAs per SM, at least
// 64 BMUs per slice) and the scratch its per-slice partial [num | den] blocks need.
struct SliceGeom { int tm, tn, slices, b_per_slice; };
inline SliceGeom neigh_slice_geom(int K, int d, int sm_count) {
    SliceGeom g;
    const bool big = d > 64 && K >= 512;                         // 128 x 128 tiles, 8 x 8 per thread
    g.tm = big ? 128 : 64; g.tn = big ? 128 : 64;
    const int gxy = (int)(ceil_div(K, g.tm) * ceil_div(d, g.tn));
    int slices = (2 * sm_count) / gxy;                            // ONE wave of 2 CTAs per SM (rounding up left a 4-CTA second wave at config 5)
    const int max_slices = (int)ceil_div(K, 4 * 16);
    if (slices > max_slices) slices = max_slices;
    if (slices < 1) slices = 1;
    g.b_per_slice = (int)round_up(ceil_div(K, slices), 16);
    g.slices = (int)ceil_div(K, g.b_per_slice);
    return g;
}
constexpr int kSmCountBound = 192;        // sizing bound for the partial blocks (B200: 148 SMs)
inline size_t neigh_partial_floats(int gx, int gy, int d) {
    const int K = gx * gy;
    const SliceGeom g = neigh_slice_geom(K, d, kSmCountBound);
    return g.slices > 1 ? (size_t)g.slices * ((size_t)K * d + K) : 0;
}

// host: everything of NeighParams that does not depend on a device-side schedule; the factor tables live in `tables`
inline void neigh_params(NeighParams &P, int gx, int gy, int d, int topology, int kind, double sigma, double eta,
                         double std_coeff, int compact, float *tables) {
    P.sched = nullptr; P.epoch = nullptr; P.std_coeff = std_coeff;
    P.gx = gx; P.gy = gy; P.d = d; P.topology = topology; P.kind = kind; P.compact = compact ? 1 : 0;
    // bubble and triangle use integer grid indices on both topologies (xpysom.py:266-269, 277-278)
    P.shifted = (topology == SOM_TOPO_HEXAGONAL && (kind == SOM_NEIGH_GAUSSIAN || kind == SOM_NEIGH_MEXICAN_HAT)) ? 1 : 0;
    P.mex_rect_quirk = (kind == SOM_NEIGH_MEXICAN_HAT && compact && topology == SOM_TOPO_RECTANGULAR) ? 1 : 0;
    const double dd = 2.0 * std_coeff * std_coeff * sigma * sigma;   // neighborhoods.py:19
    P.inv_d = (float)(1.0 / dd);
    P.two_over_d = (float)(2.0 / dd);
    P.eta = (float)eta;
    const size_t nxe = (size_t)3 * gx * gx, nye = (size_t)gy * gy;
    P.tx = tables; P.ty = tables + nxe; P.mx = tables + nxe + nye; P.my = tables + 2 * nxe + nye;
}

inline int launch_neigh_apply(const float *S, const float *c, int gx, int gy, int d, int topology, int kind,
                              double sigma, double eta, double std_coeff, int compact,
                              float *num, float *den, float *tables, float *scratch, float *partials, size_t partial_floats,
                              int sm_count, cudaStream_t st, const double *sched = nullptr, const int *epoch = nullptr) {
    NeighParams P;
    neigh_params(P, gx, gy, d, topology, kind, sigma, eta, std_coeff, compact, tables);
    P.sched = sched; P.epoch = epoch;
    const double dd = 2.0 * std_coeff * std_coeff * sigma * sigma;   // neighborhoods.py:19
    const size_t nxe = (size_t)3 * gx * gx, nye = (size_t)gy * gy;
    float *tx = tables, *ty = tx + nxe, *mx = ty + nye, *my = mx + nxe;
    const int tot = (int)(nxe + nye);
    int rc = check_cuda(launch_pdl(neigh_tables_kernel, dim3((tot + 255) / 256), dim3(256), 0, st, gx, gy, kind, P.compact,
                                   P.shifted, sigma, dd, tx, ty, mx, my, sched, epoch, std_coeff),
                        "neigh_tables_kernel launch");
    if (rc) return rc;
    const int K = gx * gy;
    if (scratch != nullptr && neigh_is_separable(topology, kind, gx, gy)) {
        // plane q = 1 of TX is the unshifted (rectangular) factor table
        const float *TX = tx + (size_t)gx * gx, *TY = ty;
        float *T = scratch, *Tc = scratch + (size_t)K * d;
        // pass 1 (over bj): T[bi, j, :] = sum_bj TY[bj][j] S[bi, bj, :]      in: (gx, gy, D)
        if ((rc = launch_axis_contract(S, TY, gy, gx, gy, gy, d, 1.f, nullptr, nullptr, T, st))) return rc;
        if ((rc = launch_axis_contract(c, TY, gy, gx, gy, gy, 1, 1.f, nullptr, nullptr, Tc, st))) return rc;
        // pass 2 (over bi): num[i, (j, :)] = eta * sum_bi TX[bi][i] T[bi, (j, :)]   in: (1, gx, gy*D)
        if ((rc = launch_axis_contract(T, TX, gx, 1, gx, gx, (int64_t)gy * d, P.eta, sched, epoch, num, st))) return rc;
        return launch_axis_contract(Tc, TX, gx, 1, gx, gx, gy, P.eta, sched, epoch, den, st);
    }
    SliceGeom g = neigh_slice_geom(K, d, sm_count);
    if (g.slices > 1 && (partials == nullptr || partial_floats < (size_t)g.slices * ((size_t)K * d + K))) {
        g.slices = 1; g.b_per_slice = (int)round_up(K, NB_K);    // no room for partial blocks: one slice (still exact)
    }
    const bool big = g.tm == 128;
    dim3 grid((unsigned)ceil_div(K, g.tm), (unsigned)ceil_div(d, g.tn), (unsigned)g.slices);
    if (big)
        rc = check_cuda(launch_pdl(neigh_apply_kernel<8, 8>, grid, dim3(NB_THREADS), 0, st, P, S, c, num, den, partials, g.b_per_slice),
                        "neigh_apply_kernel<8,8> launch");
    else
        rc = check_cuda(launch_pdl(neigh_apply_kernel<4, 4>, grid, dim3(NB_THREADS), 0, st, P, S, c, num, den, partials, g.b_per_slice),
                        "neigh_apply_kernel<4,4> launch");
    if (rc || g.slices == 1) return rc;
    const size_t block = (size_t)K * d + K;
    int blocks = (int)ceil_div((int64_t)block, 256);
    if (blocks > sm_count * 8) blocks = sm_count * 8;
    return check_cuda(launch_pdl(slice_reduce_kernel, dim3(blocks), dim3(256), 0, st, (const float *)partials, g.slices, K, d, num, den),
                      "slice_reduce_kernel launch");
}

}  // namespace somb200
