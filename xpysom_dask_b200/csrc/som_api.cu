// C ABI of libsom_b200.so — see include/som_b200.h for the contract and the
// reference lines each entry point replaces.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "bmu_simt.cuh"
#include "bmu_tc.cuh"
#include "bmu_tc2.cuh"
#include "bmu_tc3.cuh"
#include "accumulate.cuh"
#include "neigh.cuh"
#include "peer.cuh"
#include "bmu_filter.cuh"
#include "misc.cuh"
#include "epoch_tail.cuh"

namespace somb200 {

static thread_local char g_err[512] = "";

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_cuda(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return 0;
    set_error("%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    return (int)e;
}

struct DevInfo { int ok = 0, sm = 0, cc = 0; size_t smem_optin = 0; };

static int device_info(DevInfo &out) {
    static thread_local DevInfo cache[64];
    int dev = 0;
    SOM_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) dev = 0;
    if (!cache[dev].ok) {
        cudaDeviceProp p;
        SOM_CUDA(cudaGetDeviceProperties(&p, dev));
        cache[dev].sm = p.multiProcessorCount;
        cache[dev].cc = p.major * 10 + p.minor;
        cache[dev].smem_optin = p.sharedMemPerBlockOptin;
        cache[dev].ok = 1;
    }
    out = cache[dev];
    SOM_REQUIRE(out.cc >= 100, SOM_E_NODEVICE, "libsom_b200 needs an sm_100a device, found cc %d", out.cc);
    return 0;
}

static int launch_tc(int use, const float *X, int64_t n, int d, int64_t ldx, const float *xscale, int k,
                     const WsLayout &L, uint8_t *ws, int32_t *bmu, float *best, const AccTarget &T, int sm_count,
                     cudaStream_t st) {
    if (use == SOM_ALGO_TC_3XF16)
        return tc3::launch_bmu_tc3(X, n, d, ldx, xscale, k, L, ws, bmu, best, T, sm_count, st);
    return tc2::launch_bmu_tc2(X, n, d, ldx, k, L, ws, bmu, best, T, sm_count, st);
}

// the accumulator the ABI hands over: acc_replicas(k, d) copies of [S: k * acc_ld(d) words | counts: k words]
// (an accumulator in peer memory, som_b200_peer_accumulator, is used as ONE copy: every replica would be read by every rank)
static AccTarget acc_target(uint64_t *acc_dev, const float *qscale_dev, int k, int d) {
    AccTarget T;
    T.S = reinterpret_cast<unsigned long long *>(acc_dev);
    T.cnt = T.S + (size_t)k * acc_ld(d);
    T.qscale = qscale_dev;
    T.rep_words = acc_words_one(k, d);
    T.reps = peer_lookup(acc_dev) ? 1 : acc_replicas(k, d);
    return T;
}

static bool known_dist(int k) { return k >= SOM_DIST_EUCLIDEAN && k <= SOM_DIST_NORM_P; }

// AUTO: a tensor-core kernel for the two contraction distances whenever TMA can address X -- the
// fp16-split one when the row scales are available (short rows, D <= 16, in its packed mode), the TF32 one otherwise.
static int pick_algo_rule(int algo, int dist_kind, int d, bool tma_ok, bool has_xscale) {
    const bool contraction = dist_kind == SOM_DIST_EUCLIDEAN || dist_kind == SOM_DIST_COSINE;
    if (algo == SOM_ALGO_AUTO) {
        if (!(contraction && tma_ok && d >= 8)) return SOM_ALGO_SIMT_FP32;
        return has_xscale ? SOM_ALGO_TC_3XF16 : SOM_ALGO_TC_3XTF32;
    }
    return algo;
}
static int pick_algo(int algo, int dist_kind, const float *X, int64_t n, int d, int64_t ldx, const float *xscale) {
    return pick_algo_rule(algo, dist_kind, d, tc::shape_ok(X, n, d, ldx), xscale != nullptr);
}

}  // namespace somb200

using namespace somb200;

extern "C" {

int som_b200_abi_version(void) { return SOM_B200_ABI_VERSION; }

int som_b200_pick_algo(int algo, int dist_kind, int d, int rows_tma_addressable, int has_row_scales) {
    return pick_algo_rule(algo, dist_kind, d, rows_tma_addressable != 0, has_row_scales != 0);
}

const char *som_b200_last_error(void) { return g_err; }

int som_b200_device_info(int *sm_count, int *cc, size_t *smem_per_block_optin) {
    DevInfo di;
    int rc = device_info(di);
    if (rc) return rc;
    if (sm_count) *sm_count = di.sm;
    if (cc) *cc = di.cc;
    if (smem_per_block_optin) *smem_per_block_optin = di.smem_optin;
    return 0;
}

size_t som_b200_workspace_bytes(int k, int d) {
    if (k <= 0 || d <= 0) return 0;
    return ws_layout(k, d).total;
}

size_t som_b200_shard_workspace_bytes(int64_t n, int k, int d) {
    if (k <= 0 || d <= 0 || n < 0) return 0;
    return ws_layout(k, d).total + (size_t)round_up(n * 4, 1024);
}

size_t som_b200_neigh_table_floats(int gx, int gy) {
    if (gx <= 0 || gy <= 0) return 0;
    return neigh_table_floats(gx, gy);
}

size_t som_b200_neigh_scratch_floats(int gx, int gy, int d) {
    if (gx <= 0 || gy <= 0 || d <= 0) return 0;
    return neigh_table_floats(gx, gy) + neigh_separable_floats(gx, gy, d) + neigh_partial_floats(gx, gy, d);
}

// the caller's neighbourhood scratch: [factor tables | intermediates of the separable path | per-slice partial blocks];
// a buffer that ends early simply turns the later parts off (direct kernel / one slice)
struct NeighScratch { float *sep = nullptr, *partials = nullptr; size_t partial_floats = 0; };
static NeighScratch carve_scratch(float *base, size_t floats, int gx, int gy, int d) {
    NeighScratch n;
    size_t off = neigh_table_floats(gx, gy);
    const size_t sepf = neigh_separable_floats(gx, gy, d);
    if (floats >= off + sepf) { n.sep = base + off; off += sepf; } else return n;
    if (floats > off) { n.partials = base + off; n.partial_floats = floats - off; }
    return n;
}

static SplitOut split_out(const WsLayout &L, uint8_t *ws) {
    SplitOut O;
    O.whi = reinterpret_cast<float *>(ws + L.whi_off); O.wlo = reinterpret_cast<float *>(ws + L.wlo_off);
    O.w16hi = reinterpret_cast<__half *>(ws + L.w16hi_off); O.w16lo = reinterpret_cast<__half *>(ws + L.w16lo_off);
    O.wsinv = reinterpret_cast<float *>(ws + L.wsinv_off);
    O.wfold = reinterpret_cast<float *>(ws + L.wfold_off);
    O.d_pad = L.d_pad; O.d_pad64 = L.d_pad64;
    static const int no_fold = tc::env_int("SOM_B200_NO_FOLD");
    O.allow_fold = no_fold ? 0 : 1;
    return O;
}

int som_b200_prepare_codebook(const float *w_dev, int k, int d, int dist_kind, float p,
                              void *ws_dev, size_t ws_bytes, void *stream) {
    (void)p;
    SOM_REQUIRE(w_dev && ws_dev && k > 0 && d > 0, SOM_E_BADARG, "prepare_codebook: bad argument");
    SOM_REQUIRE(known_dist(dist_kind), SOM_E_BADARG, "prepare_codebook: unknown distance kind %d", dist_kind);
    const WsLayout L = ws_layout(k, d);
    SOM_REQUIRE(ws_bytes >= L.total, SOM_E_WORKSPACE, "prepare_codebook: workspace %zu < %zu bytes", ws_bytes, L.total);
    uint8_t *ws = static_cast<uint8_t *>(ws_dev);
    const bool split = dist_kind == SOM_DIST_EUCLIDEAN || dist_kind == SOM_DIST_COSINE;
    SOM_CUDA(cudaMemsetAsync(ws + L.done_off, 0, L.amax_off - L.done_off, (cudaStream_t)stream));   // grid barrier of the fused tail, statistics
    const int threads = 256, warps_per_block = threads / 32;
    const int blocks = (int)ceil_div(L.k_pad, warps_per_block);
    float *aux = reinterpret_cast<float *>(ws + L.aux_off), *amax = reinterpret_cast<float *>(ws + L.amax_off);
    unsigned int *gstat = reinterpret_cast<unsigned int *>(ws + L.gstat_off);
    int rc = check_cuda(launch_pdl(codebook_stats_kernel, dim3(blocks), dim3(threads), 0, (cudaStream_t)stream, w_dev, k, d,
                                   dist_kind, L.k_pad, aux, reinterpret_cast<float *>(ws + L.bias_off), amax, gstat),
                        "codebook_stats_kernel launch");
    if (rc || !split) return rc;
    return check_cuda(launch_pdl(codebook_split_kernel, dim3(blocks), dim3(threads), 0, (cudaStream_t)stream, w_dev, k, d,
                                 dist_kind, L.k_pad, split_out(L, ws), aux, reinterpret_cast<const float *>(ws + L.bias_off),
                                 amax, gstat),
                      "codebook_split_kernel launch");
}

int som_b200_prepare_samples(const float *x_dev, int64_t n, int d, int64_t ldx, float *xscale_dev, float *colmax_dev,
                             void *stream) {
    SOM_REQUIRE(n >= 0 && d > 0 && ldx >= d, SOM_E_BADARG, "prepare_samples: bad argument");
    if (n == 0) return 0;
    SOM_REQUIRE(x_dev && (xscale_dev || colmax_dev), SOM_E_BADARG, "prepare_samples: NULL pointer");
    const bool vec = (d % 4 == 0) && (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(x_dev) & 15) == 0);
    cudaStream_t st = (cudaStream_t)stream;
    if (xscale_dev) {
        int lpr = 1;
        while (lpr < 32 && lpr < d / 4) lpr <<= 1;
        int64_t blocks = ceil_div(n, vec ? 8 * (32 / lpr) : 8);
        if (blocks > 148 * 16) blocks = 148 * 16;
        if (blocks < 1) blocks = 1;
        if (vec) row_scale_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(x_dev, n, d, ldx, xscale_dev, lpr);
        else     row_scale_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(x_dev, n, d, ldx, xscale_dev, 32);
        SOM_CUDA(cudaGetLastError());
    }
    if (colmax_dev) {
        const int units = vec ? d / 4 : d;
        int L = 1;
        while (L < units && L < 256) L <<= 1;
        int64_t blocks = ceil_div(n, 256 / L * 16);           // ~16 rows per thread
        if (blocks > 148 * 8) blocks = 148 * 8;
        if (blocks < 1) blocks = 1;
        unsigned int *bits = reinterpret_cast<unsigned int *>(colmax_dev);
        if (vec) column_absmax_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(x_dev, n, d, ldx, bits, L);
        else     column_absmax_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(x_dev, n, d, ldx, bits, L);
        SOM_CUDA(cudaGetLastError());
    }
    return 0;
}

size_t som_b200_accum_words(int k, int d) {
    if (k <= 0 || d <= 0) return 0;
    return acc_words_one(k, d) * (size_t)acc_replicas(k, d);
}

int som_b200_accum_replicas(int k, int d) { return (k <= 0 || d <= 0) ? 0 : acc_replicas(k, d); }

int som_b200_accum_scales(const float *colmax_dev, int d, double n_total, float *qscale_dev, float *qinv_dev, void *stream) {
    SOM_REQUIRE(colmax_dev && qscale_dev && qinv_dev && d > 0 && n_total >= 0, SOM_E_BADARG, "accum_scales: bad argument");
    const int d_pad = (int)round_up(d, 4);
    accum_scales_kernel<<<(d_pad + 127) / 128, 128, 0, (cudaStream_t)stream>>>(colmax_dev, d, d_pad, n_total < 1 ? 1.0 : n_total,
                                                                             qscale_dev, qinv_dev);
    return check_cuda(cudaGetLastError(), "accum_scales_kernel launch");
}

int som_b200_accum_finalize(uint64_t *acc_dev, const float *qinv_dev, int k, int d, float *s_dev, float *c_dev, void *stream) {
    SOM_REQUIRE(acc_dev && qinv_dev && s_dev && c_dev && k > 0 && d > 0, SOM_E_BADARG, "accum_finalize: bad argument");
    DevInfo di;
    int rc = device_info(di);
    if (rc) return rc;
    const AccTarget T = acc_target(acc_dev, nullptr, k, d);
    // an accumulator handed out by som_b200_peer_accumulator: the sums run over all ranks (peer.cuh)
    PeerComm *pc = peer_lookup(acc_dev);
    SOM_REQUIRE(pc == nullptr || pc->words >= acc_words_one(k, d), SOM_E_SHAPE, "accum_finalize: the peer accumulator is too small");
    const PeerView V = pc ? peer_next_exchange(pc) : PeerView();
    int grid = grid_for((int64_t)k * acc_ld(d) + k, di.sm);
    if (pc && grid > di.sm) grid = di.sm;          // every block polls the flags: all of them co-resident
    return check_cuda(launch_pdl(accum_finalize_kernel, dim3(grid), dim3(256), 0,
                                 (cudaStream_t)stream, T.S, T.reps, T.rep_words, qinv_dev, k, d, acc_ld(d), s_dev, c_dev, 1, V),
                      "accum_finalize_kernel launch");
}

int som_b200_accum_fold_replicas(uint64_t *acc_dev, uint64_t *dst_dev, int k, int d, void *stream) {
    SOM_REQUIRE(acc_dev && k > 0 && d > 0, SOM_E_BADARG, "accum_fold_replicas: bad argument");
    const AccTarget T = acc_target(acc_dev, nullptr, k, d);
    unsigned long long *dst = dst_dev ? reinterpret_cast<unsigned long long *>(dst_dev) : T.S;
    if (T.reps <= 1 && dst == T.S) return 0;
    DevInfo di;
    int rc = device_info(di);
    if (rc) return rc;
    const int64_t all = (int64_t)k * acc_ld(d) + k;
    return check_cuda(launch_pdl(accum_fold_replicas_kernel, dim3(grid_for(all, di.sm)), dim3(256), 0, (cudaStream_t)stream, T.S,
                                 T.reps, T.rep_words, all, dst), "accum_fold_replicas_kernel launch");
}

int som_b200_accum_fold(uint64_t *acc_dev, const float *qinv_dev, int k, int d, double *sd_dev, void *stream) {
    SOM_REQUIRE(acc_dev && qinv_dev && sd_dev && k > 0 && d > 0, SOM_E_BADARG, "accum_fold: bad argument");
    DevInfo di;
    int rc = device_info(di);
    if (rc) return rc;
    const AccTarget T = acc_target(acc_dev, nullptr, k, d);
    accum_fold_kernel<<<grid_for((int64_t)k * acc_ld(d), di.sm), 256, 0, (cudaStream_t)stream>>>(
        T.S, T.reps, T.rep_words, qinv_dev, k, d, acc_ld(d), sd_dev, sd_dev + (size_t)k * d);
    return check_cuda(cudaGetLastError(), "accum_fold_kernel launch");
}

int som_b200_accum_finalize_f64(double *sd_dev, int k, int d, float *s_dev, float *c_dev, void *stream) {
    SOM_REQUIRE(sd_dev && s_dev && c_dev && k > 0 && d > 0, SOM_E_BADARG, "accum_finalize_f64: bad argument");
    DevInfo di;
    int rc = device_info(di);
    if (rc) return rc;
    accum_finalize_f64_kernel<<<grid_for((int64_t)k * d, di.sm), 256, 0, (cudaStream_t)stream>>>(
        sd_dev, sd_dev + (size_t)k * d, k, d, s_dev, c_dev);
    return check_cuda(cudaGetLastError(), "accum_finalize_f64_kernel launch");
}

int som_b200_bmu(const float *x_dev, int64_t n, int d, int64_t ldx, const float *xscale_dev, const float *w_dev, int k,
                 int dist_kind, float p, int algo, int32_t *bmu_dev, float *best_dev, void *ws_dev, size_t ws_bytes,
                 void *stream) {
    SOM_REQUIRE(w_dev && ws_dev && bmu_dev && k > 0 && d > 0 && n >= 0 && ldx >= d, SOM_E_BADARG, "bmu: bad argument");
    SOM_REQUIRE(known_dist(dist_kind), SOM_E_BADARG, "bmu: unknown distance kind %d", dist_kind);
    if (n == 0) return 0;
    SOM_REQUIRE(x_dev, SOM_E_BADARG, "bmu: x is NULL");
    const WsLayout L = ws_layout(k, d);
    SOM_REQUIRE(ws_bytes >= L.total, SOM_E_WORKSPACE, "bmu: workspace %zu < %zu bytes", ws_bytes, L.total);
    DevInfo di;
    int rc = device_info(di);
    if (rc) return rc;
    uint8_t *ws = static_cast<uint8_t *>(ws_dev);
    const int use = pick_algo(algo, dist_kind, x_dev, n, d, ldx, xscale_dev);
    if (use == SOM_ALGO_TC_3XTF32 || use == SOM_ALGO_TC_3XF16) {
        SOM_REQUIRE(dist_kind == SOM_DIST_EUCLIDEAN || dist_kind == SOM_DIST_COSINE, SOM_E_SHAPE,
                    "the tensor-core kernels compute contraction distances only (euclidean, cosine)");
        return launch_tc(use, x_dev, n, d, ldx, xscale_dev, k, L, ws, bmu_dev, best_dev, AccTarget(), di.sm,
                         (cudaStream_t)stream);
    }
    SOM_REQUIRE(use == SOM_ALGO_SIMT_FP32, SOM_E_BADARG, "bmu: unknown algo %d", algo);
    return launch_bmu_simt(x_dev, n, d, ldx, w_dev, k, dist_kind, p, reinterpret_cast<const float *>(ws + L.aux_off),
                           bmu_dev, best_dev, di.sm, (cudaStream_t)stream);
}

int som_b200_distances(const float *x_dev, int64_t n, int d, int64_t ldx, const float *w_dev, int k, int dist_kind,
                       float p, int mode, float *out_dev, void *ws_dev, size_t ws_bytes, void *stream) {
    SOM_REQUIRE(w_dev && ws_dev && out_dev && k > 0 && d > 0 && n >= 0 && ldx >= d, SOM_E_BADARG, "distances: bad argument");
    SOM_REQUIRE(known_dist(dist_kind) && mode >= 0 && mode <= 2, SOM_E_BADARG, "distances: unknown kind / mode");
    if (n == 0) return 0;
    SOM_REQUIRE(x_dev, SOM_E_BADARG, "distances: x is NULL");
    const WsLayout L = ws_layout(k, d);
    SOM_REQUIRE(ws_bytes >= L.total, SOM_E_WORKSPACE, "distances: workspace %zu < %zu bytes", ws_bytes, L.total);
    // |w|^2 per neuron: aux holds it for every kind except cosine (1/|w|); bias holds it only for euclidean.
    // Recompute into the amax slot, which no BMU kernel reads.
    uint8_t *ws = static_cast<uint8_t *>(ws_dev);
    float *wsq = reinterpret_cast<float *>(ws + L.amax_off);
    row_sq_kernel<<<(unsigned)ceil_div(k, 8), 256, 0, (cudaStream_t)stream>>>(w_dev, k, d, wsq);
    int rc = check_cuda(cudaGetLastError(), "row_sq_kernel launch");
    if (rc) return rc;
    return launch_dist_matrix(x_dev, n, d, ldx, w_dev, k, dist_kind, p, mode, wsq, out_dev, (cudaStream_t)stream);
}

int som_b200_top2(const float *x_dev, int64_t n, int d, int64_t ldx, const float *w_dev, int k, int32_t *top2_dev,
                  void *ws_dev, size_t ws_bytes, void *stream) {
    SOM_REQUIRE(w_dev && ws_dev && top2_dev && k > 0 && d > 0 && n >= 0 && ldx >= d, SOM_E_BADARG, "top2: bad argument");
    if (n == 0) return 0;
    SOM_REQUIRE(x_dev, SOM_E_BADARG, "top2: x is NULL");
    const WsLayout L = ws_layout(k, d);
    SOM_REQUIRE(ws_bytes >= L.total, SOM_E_WORKSPACE, "top2: workspace %zu < %zu bytes", ws_bytes, L.total);
    DevInfo di;
    int rc = device_info(di);
    if (rc) return rc;
    uint8_t *ws = static_cast<uint8_t *>(ws_dev);
    float *wsq = reinterpret_cast<float *>(ws + L.amax_off);        // as som_b200_distances: no BMU kernel reads this slot
    row_sq_kernel<<<(unsigned)ceil_div(k, 8), 256, 0, (cudaStream_t)stream>>>(w_dev, k, d, wsq);
    if ((rc = check_cuda(cudaGetLastError(), "row_sq_kernel launch"))) return rc;
    return launch_top2_simt(x_dev, n, d, ldx, w_dev, k, wsq, top2_dev, di.sm, (cudaStream_t)stream);
}

/* ---- one-pass filter + exact refine (bmu_filter.cuh) ---- */
int som_b200_filter_eligible(const float *x_dev, int64_t n, int d, int64_t ldx, int k, int dist_kind) {
    return flt::filter_eligible(x_dev, n, d, ldx, k, dist_kind) ? 1 : 0;
}

size_t som_b200_filter_workspace_bytes(int64_t n, int k, int d) {
    if (n <= 0 || k <= 0 || d <= 0) return 0;
    return flt::filter_layout(n, k, d).total;
}

size_t som_b200_filter_overflow_offset(int64_t n, int k, int d) {
    if (n <= 0 || k <= 0 || d <= 0) return 0;
    return flt::filter_layout(n, k, d).ovf_off;
}

int som_b200_filter_prepare_samples(const float *x_dev, int64_t n, int d, int64_t ldx, void *fws_dev, size_t fws_bytes, void *stream) {
    SOM_REQUIRE(x_dev && fws_dev && n > 0 && d > 0 && ldx >= d, SOM_E_BADARG, "filter_prepare_samples: bad argument");
    SOM_REQUIRE(flt::filter_eligible(x_dev, n, d, ldx, 1 << 20, SOM_DIST_EUCLIDEAN), SOM_E_SHAPE,
                "filter_prepare_samples: shape not eligible (needs 256 <= d <= 1024, d %% 4 == 0, 16-byte aligned rows, n >= 4096)");
    SOM_REQUIRE(fws_bytes >= flt::filter_layout(n, 1, d).cand_off, SOM_E_WORKSPACE, "filter_prepare_samples: workspace too small");
    DevInfo di;
    int rc = device_info(di);
    if (rc) return rc;
    return flt::filter_prepare_samples(x_dev, n, d, ldx, static_cast<uint8_t *>(fws_dev), di.sm, (cudaStream_t)stream);
}

int som_b200_bmu_filter(const float *x_dev, int64_t n, int d, int64_t ldx, const float *w_dev, int k, int32_t *bmu_dev,
                        const float *qscale_dev, uint64_t *acc_dev, void *fws_dev, size_t fws_bytes, void *stream) {
    SOM_REQUIRE(x_dev && w_dev && bmu_dev && fws_dev, SOM_E_BADARG, "bmu_filter: NULL pointer");
    SOM_REQUIRE(flt::filter_eligible(x_dev, n, d, ldx, k, SOM_DIST_EUCLIDEAN), SOM_E_SHAPE, "bmu_filter: shape not eligible");
    SOM_REQUIRE(fws_bytes >= flt::filter_layout(n, k, d).total, SOM_E_WORKSPACE, "bmu_filter: workspace %zu < %zu bytes", fws_bytes,
                flt::filter_layout(n, k, d).total);
    DevInfo di;
    int rc = device_info(di);
    if (rc) return rc;
    SOM_REQUIRE((acc_dev == nullptr) == (qscale_dev == nullptr), SOM_E_BADARG, "bmu_filter: accumulator and scales go together");
    SOM_REQUIRE(acc_dev == nullptr || ((reinterpret_cast<uintptr_t>(acc_dev) & 15) == 0 && (reinterpret_cast<uintptr_t>(qscale_dev) & 15) == 0),
                SOM_E_SHAPE, "bmu_filter: the accumulator and the scales must be 16-byte aligned");
    const AccTarget T = acc_dev ? acc_target(acc_dev, qscale_dev, k, d) : AccTarget();
    return flt::launch_bmu_filter(x_dev, n, d, ldx, w_dev, k, static_cast<uint8_t *>(fws_dev), bmu_dev, T, di.sm, (cudaStream_t)stream);
}

int som_b200_accumulate(const float *x_dev, int64_t n, int d, int64_t ldx, const int32_t *bmu_dev, int k,
                        const float *qscale_dev, uint64_t *acc_dev, void *stream) {
    SOM_REQUIRE(qscale_dev && acc_dev && k > 0 && d > 0 && n >= 0 && ldx >= d, SOM_E_BADARG, "accumulate: bad argument");
    if (n == 0) return 0;
    SOM_REQUIRE(x_dev && bmu_dev, SOM_E_BADARG, "accumulate: NULL input");
    SOM_REQUIRE((reinterpret_cast<uintptr_t>(acc_dev) & 15) == 0 && (reinterpret_cast<uintptr_t>(qscale_dev) & 15) == 0,
                SOM_E_SHAPE, "accumulate: the accumulator and the scales must be 16-byte aligned");
    DevInfo di;
    int rc = device_info(di);
    if (rc) return rc;
    return launch_accumulate(x_dev, n, d, ldx, bmu_dev, k, acc_target(acc_dev, qscale_dev, k, d), di.sm, (cudaStream_t)stream);
}

int som_b200_epoch_accumulate(const float *x_dev, int64_t n, int d, int64_t ldx, const float *xscale_dev,
                              const float *w_dev, int k, int dist_kind, float p, int algo, const float *qscale_dev,
                              uint64_t *acc_dev, int32_t *bmu_dev, void *ws_dev, size_t ws_bytes, void *stream) {
    SOM_REQUIRE(n >= 0 && k > 0 && d > 0 && ldx >= d, SOM_E_BADARG, "epoch_accumulate: bad argument");
    if (n == 0) return 0;
    SOM_REQUIRE(x_dev && w_dev && qscale_dev && acc_dev && ws_dev, SOM_E_BADARG, "epoch_accumulate: NULL pointer");
    SOM_REQUIRE((reinterpret_cast<uintptr_t>(acc_dev) & 15) == 0 && (reinterpret_cast<uintptr_t>(qscale_dev) & 15) == 0,
                SOM_E_SHAPE, "epoch_accumulate: the accumulator and the scales must be 16-byte aligned");
    SOM_REQUIRE(known_dist(dist_kind), SOM_E_BADARG, "epoch_accumulate: unknown distance kind %d", dist_kind);
    const WsLayout L = ws_layout(k, d);
    const int use = pick_algo(algo, dist_kind, x_dev, n, d, ldx, xscale_dev);
    if ((use == SOM_ALGO_TC_3XTF32 || use == SOM_ALGO_TC_3XF16) &&
        (dist_kind == SOM_DIST_EUCLIDEAN || dist_kind == SOM_DIST_COSINE)) {
        // one kernel: contraction + argmin + per-BMU sums (X read from HBM once)
        SOM_REQUIRE(ws_bytes >= L.total, SOM_E_WORKSPACE, "epoch_accumulate: workspace %zu < %zu bytes", ws_bytes, L.total);
        DevInfo di;
        int rc = device_info(di);
        if (rc) return rc;
        return launch_tc(use, x_dev, n, d, ldx, xscale_dev, k, L, static_cast<uint8_t *>(ws_dev), bmu_dev, nullptr,
                         acc_target(acc_dev, qscale_dev, k, d), di.sm, (cudaStream_t)stream);
    }
    int32_t *bmu = bmu_dev;
    if (!bmu) {
        const size_t need = L.total + (size_t)round_up(n * 4, 1024);
        SOM_REQUIRE(ws_dev && ws_bytes >= need, SOM_E_WORKSPACE,
                    "epoch_accumulate: workspace %zu < %zu bytes (no bmu buffer given)", ws_bytes, need);
        bmu = reinterpret_cast<int32_t *>(static_cast<uint8_t *>(ws_dev) + L.total);
    }
    int rc = som_b200_bmu(x_dev, n, d, ldx, xscale_dev, w_dev, k, dist_kind, p, algo, bmu, nullptr, ws_dev, ws_bytes, stream);
    if (rc) return rc;
    return som_b200_accumulate(x_dev, n, d, ldx, bmu, k, qscale_dev, acc_dev, stream);
}

static int neigh_check(int topology, int neigh_kind, int gx, int gy, int compact_support) {
    SOM_REQUIRE(topology == SOM_TOPO_RECTANGULAR || topology == SOM_TOPO_HEXAGONAL, SOM_E_BADARG,
                "unknown topology %d", topology);
    SOM_REQUIRE(neigh_kind >= SOM_NEIGH_GAUSSIAN && neigh_kind <= SOM_NEIGH_TRIANGLE, SOM_E_BADARG,
                "unknown neighbourhood %d", neigh_kind);
    // combinations the reference itself rejects
    SOM_REQUIRE(!(neigh_kind == SOM_NEIGH_TRIANGLE && topology == SOM_TOPO_HEXAGONAL), SOM_E_SHAPE,
                "triangle is not available on a hexagonal map (xpysom.py:272-279)");
    SOM_REQUIRE(!(neigh_kind == SOM_NEIGH_MEXICAN_HAT && compact_support && topology == SOM_TOPO_RECTANGULAR && gx != gy),
                SOM_E_SHAPE, "mexican_hat with compact_support broadcasts (n,gx)*(n,gy) in the reference "
                "(neighborhoods.py:69-71): needs gx == gy");
    return 0;
}

int som_b200_neigh_apply(const float *s_dev, const float *c_dev, int gx, int gy, int d, int topology,
                         int neigh_kind, double sigma, double eta, double std_coeff, int compact_support,
                         float *num_dev, float *den_dev, float *tables_dev, size_t tables_floats, void *stream) {
    SOM_REQUIRE(s_dev && c_dev && num_dev && den_dev && tables_dev && gx > 0 && gy > 0 && d > 0, SOM_E_BADARG,
                "neigh_apply: bad argument");
    int rc = neigh_check(topology, neigh_kind, gx, gy, compact_support);
    if (rc) return rc;
    SOM_REQUIRE(sigma != 0.0 && std_coeff != 0.0, SOM_E_BADARG, "neigh_apply: sigma and std_coeff must be non-zero");
    SOM_REQUIRE(tables_floats >= neigh_table_floats(gx, gy), SOM_E_WORKSPACE,
                "neigh_apply: scratch of %zu floats < %zu", tables_floats, neigh_table_floats(gx, gy));
    DevInfo di;
    if ((rc = device_info(di))) return rc;
    const NeighScratch ns = carve_scratch(tables_dev, tables_floats, gx, gy, d);
    return launch_neigh_apply(s_dev, c_dev, gx, gy, d, topology, neigh_kind, sigma, eta, std_coeff, compact_support,
                              num_dev, den_dev, tables_dev, ns.sep, ns.partials, ns.partial_floats, di.sm, (cudaStream_t)stream);
}

// same as som_b200_neigh_apply, with sigma and eta read ON THE DEVICE from sched_dev[2e], sched_dev[2e+1],
// e = *epoch_dev, so that one captured CUDA graph can be replayed for every epoch
int som_b200_neigh_apply_sched(const float *s_dev, const float *c_dev, int gx, int gy, int d, int topology,
                               int neigh_kind, const double *sched_dev, const int *epoch_dev, double std_coeff,
                               int compact_support, float *num_dev, float *den_dev, float *tables_dev,
                               size_t tables_floats, void *stream) {
    SOM_REQUIRE(s_dev && c_dev && num_dev && den_dev && tables_dev && sched_dev && epoch_dev && gx > 0 && gy > 0 && d > 0,
                SOM_E_BADARG, "neigh_apply_sched: bad argument");
    int rc = neigh_check(topology, neigh_kind, gx, gy, compact_support);
    if (rc) return rc;
    SOM_REQUIRE(std_coeff != 0.0, SOM_E_BADARG, "neigh_apply_sched: std_coeff must be non-zero");
    SOM_REQUIRE(tables_floats >= neigh_table_floats(gx, gy), SOM_E_WORKSPACE,
                "neigh_apply_sched: scratch of %zu floats < %zu", tables_floats, neigh_table_floats(gx, gy));
    DevInfo di;
    if ((rc = device_info(di))) return rc;
    const NeighScratch ns = carve_scratch(tables_dev, tables_floats, gx, gy, d);
    return launch_neigh_apply(s_dev, c_dev, gx, gy, d, topology, neigh_kind, 1.0, 1.0, std_coeff, compact_support,
                              num_dev, den_dev, tables_dev, ns.sep, ns.partials, ns.partial_floats, di.sm, (cudaStream_t)stream,
                              sched_dev, epoch_dev);
}

// Fused epoch tail (epoch_tail.cuh) when the map is small enough for one co-resident grid; otherwise the same
// work as separate launches.  Either way, on return (stream order): W holds the merged codebook, the workspace
// holds its statistics and operand copies (as after som_b200_prepare_codebook), S and c hold the fp32 per-BMU sums
// of the epoch just finished and the exact accumulator (when one was given) is zero again.
int som_b200_epoch_tail(uint64_t *acc_dev, const float *qinv_dev, float *s_dev, float *c_dev, float *w_dev, int gx, int gy,
                        int d, int topology, int neigh_kind, double sigma, double eta, double std_coeff, int compact_support,
                        int dist_kind, float p, float *num_dev, float *den_dev, float *tables_dev, size_t tables_floats,
                        void *ws_dev, size_t ws_bytes, void *stream) {
    SOM_REQUIRE(s_dev && c_dev && w_dev && num_dev && den_dev && tables_dev && ws_dev, SOM_E_BADARG, "epoch_tail: NULL pointer");
    SOM_REQUIRE(acc_dev == nullptr || qinv_dev != nullptr, SOM_E_BADARG, "epoch_tail: an accumulator needs its inverse scales");
    SOM_REQUIRE(gx > 0 && gy > 0 && d > 0 && sigma != 0.0 && std_coeff != 0.0, SOM_E_BADARG, "epoch_tail: bad argument");
    SOM_REQUIRE(known_dist(dist_kind), SOM_E_BADARG, "epoch_tail: unknown distance kind %d", dist_kind);
    const int K = gx * gy;
    const WsLayout L = ws_layout(K, d);
    SOM_REQUIRE(ws_bytes >= L.total, SOM_E_WORKSPACE, "epoch_tail: workspace %zu < %zu bytes", ws_bytes, L.total);
    SOM_REQUIRE(tables_floats >= neigh_table_floats(gx, gy), SOM_E_WORKSPACE, "epoch_tail: tables buffer too small");
    int rc = neigh_check(topology, neigh_kind, gx, gy, compact_support);
    if (rc) return rc;
    DevInfo di;
    if ((rc = device_info(di))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t *ws = static_cast<uint8_t *>(ws_dev);
    const NeighScratch ns = carve_scratch(tables_dev, tables_floats, gx, gy, d);
    const bool separable = ns.sep != nullptr && neigh_is_separable(topology, neigh_kind, gx, gy);
    // Measured on B200 (same box, bench.py --steps 10): fused 0.426 vs 0.438 ms per epoch at config 2, 0.996 vs 1.003 at
    // config 3; with the 128x128 apply tiles (more than 64 features) the persistent grid LOSES 2.5 % at config 5, so
    // those maps keep the separate launches (experiments builds: SOM_B200_TAIL_SEPARATE=1 forces them everywhere).
    static const int no_fuse = tc::env_int("SOM_B200_TAIL_SEPARATE");
    const bool wide = d > 64 && K >= 512;
    const bool fused = !no_fuse && !separable && !wide && (int64_t)K * d <= (int64_t)1 << 20;
    auto separate_launches = [&]() -> int {
        int r;
        if (acc_dev && (r = som_b200_accum_finalize(acc_dev, qinv_dev, K, d, s_dev, c_dev, stream))) return r;
        if ((r = launch_neigh_apply(s_dev, c_dev, gx, gy, d, topology, neigh_kind, sigma, eta, std_coeff, compact_support,
                                    num_dev, den_dev, tables_dev, ns.sep, ns.partials, ns.partial_floats, di.sm, st))) return r;
        if ((r = som_b200_merge(w_dev, num_dev, den_dev, K, d, stream))) return r;
        return som_b200_prepare_codebook(w_dev, K, d, dist_kind, p, ws_dev, ws_bytes, stream);
    };
    if (!fused) return separate_launches();
    TailArgs A;
    neigh_params(A.P, gx, gy, d, topology, neigh_kind, sigma, eta, std_coeff, compact_support, tables_dev);
    A.sigma = sigma; A.dd = 2.0 * std_coeff * std_coeff * sigma * sigma;
    A.S = s_dev; A.c = c_dev; A.num = num_dev; A.den = den_dev; A.W = w_dev;
    const AccTarget T = acc_target(acc_dev, nullptr, K, d);
    A.Si = acc_dev ? T.S : nullptr; A.qinv = qinv_dev; A.lds = acc_ld(d); A.reps = T.reps; A.rep_words = T.rep_words;
    A.k = K; A.d = d; A.dist_kind = dist_kind; A.k_pad = L.k_pad;
    A.aux = reinterpret_cast<float *>(ws + L.aux_off); A.bias = reinterpret_cast<float *>(ws + L.bias_off);
    A.amax = reinterpret_cast<float *>(ws + L.amax_off); A.gstat = reinterpret_cast<unsigned int *>(ws + L.gstat_off);
    A.split = split_out(L, ws);
    A.do_split = (dist_kind == SOM_DIST_EUCLIDEAN || dist_kind == SOM_DIST_COSINE) ? 1 : 0;
    A.bar = reinterpret_cast<unsigned int *>(ws + L.done_off + 256);
    A.tiles_m = (int)ceil_div(K, 64); A.tiles_n = (int)ceil_div(d, 64);
    static int blocks_per_sm = 0;
    if (!blocks_per_sm) {
        int nb = 0;
        SOM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, epoch_tail_kernel<4, 4>, NB_THREADS, 0));
        SOM_REQUIRE(nb >= 1, SOM_E_NODEVICE, "epoch_tail: kernel does not fit an SM");
        blocks_per_sm = nb > 2 ? 2 : nb;
    }
    const int grid = di.sm * blocks_per_sm;
    const int gxy = A.tiles_m * A.tiles_n;
    int slices = grid / gxy;                                     // one wave of tiles over the persistent grid
    const int max_slices = (int)ceil_div(K, 4 * NB_K);           // at least 64 BMUs per slice
    if (slices > max_slices) slices = max_slices;
    if (slices < 1) slices = 1;
    // every slice writes its own partial block; without room for them the reduction is not sliced
    const size_t block = (size_t)K * d + K;
    if (slices > 1 && ns.partial_floats / block < (size_t)slices) slices = (int)(ns.partial_floats / block);
    if (slices < 1) slices = 1;
    A.partials = ns.partials;
    A.b_per_slice = (int)round_up(ceil_div(K, slices), NB_K);
    A.slices = (int)ceil_div(K, A.b_per_slice);
    PeerComm *pc = peer_lookup(acc_dev);
    SOM_REQUIRE(pc == nullptr || pc->words >= acc_words_one(K, d), SOM_E_SHAPE, "epoch_tail: the peer accumulator is too small");
    const unsigned seq_before = pc ? pc->seq : 0u;
    if (pc) A.peer = peer_next_exchange(pc);
    void *args[] = {&A};
    // cooperative launch: all CTAs co-resident or the launch fails -- the grid barriers cannot deadlock
    const cudaError_t e = cudaLaunchCooperativeKernel((const void *)epoch_tail_kernel<4, 4>, dim3(grid), dim3(NB_THREADS), args, 0, st);
    if (e == cudaErrorCooperativeLaunchTooLarge || e == cudaErrorNotSupported) {
        (void)cudaGetLastError();                   // the GPU is shared / partitioned: same work, separate launches
        if (pc) pc->seq = seq_before;               // (the exchange was not launched)
        return separate_launches();
    }
    return check_cuda(e, "epoch_tail_kernel launch");
}

__global__ void epoch_advance_kernel(int *epoch) { *epoch += 1; }

int som_b200_epoch_advance(int *epoch_dev, void *stream) {
    SOM_REQUIRE(epoch_dev, SOM_E_BADARG, "epoch_advance: NULL pointer");
    epoch_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(epoch_dev);
    return check_cuda(cudaGetLastError(), "epoch_advance_kernel launch");
}

int som_b200_merge(float *w_dev, const float *num_dev, const float *den_dev, int k, int d, void *stream) {
    SOM_REQUIRE(w_dev && num_dev && den_dev && k > 0 && d > 0, SOM_E_BADARG, "merge: bad argument");
    const int64_t tot = (int64_t)k * d;
    int blocks = (int)ceil_div(tot, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    SOM_CUDA(launch_pdl(merge_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, w_dev, num_dev, den_dev, k, d));
    return check_cuda(cudaGetLastError(), "merge_kernel launch");
}

int som_b200_peer_create(size_t acc_words, int world, int rank, void **comm_out, void *handle_out_64_bytes) {
    PeerComm *c = nullptr;
    const int rc = peer_create(acc_words, world, rank, &c, handle_out_64_bytes);
    if (rc == 0) *comm_out = c;
    return rc;
}

int som_b200_peer_connect(void *comm, const void *all_handles) {
    return peer_connect(static_cast<PeerComm *>(comm), all_handles);
}

uint64_t *som_b200_peer_accumulator(void *comm) {
    PeerComm *c = static_cast<PeerComm *>(comm);
    if (!c || !c->connected) return nullptr;
    return reinterpret_cast<uint64_t *>(c->acc(c->next_par()));
}

int som_b200_peer_destroy(void *comm) { return peer_destroy(static_cast<PeerComm *>(comm)); }

int som_b200_quantize(const float *x_dev, int64_t n, int d, int64_t ldx, const float *w_dev, int k,
                      const int32_t *bmu_dev, float *q_dev, float *err_dev, void *stream) {
    SOM_REQUIRE(w_dev && k > 0 && d > 0 && n >= 0 && ldx >= d, SOM_E_BADARG, "quantize: bad argument");
    if (n == 0) return 0;
    SOM_REQUIRE(x_dev && bmu_dev, SOM_E_BADARG, "quantize: NULL input");
    int64_t blocks = ceil_div(n, 8);
    if (blocks > 148 * 32) blocks = 148 * 32;
    quantize_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x_dev, n, d, ldx, w_dev, bmu_dev, q_dev, err_dev);
    return check_cuda(cudaGetLastError(), "quantize_kernel launch");
}

int som_b200_distance_map(const float *w_dev, int gx, int gy, int d, int topology, float *um_dev, void *stream) {
    SOM_REQUIRE(w_dev && um_dev && gx > 0 && gy > 0 && d > 0, SOM_E_BADARG, "distance_map: bad argument");
    const int K = gx * gy;
    distance_map_kernel<<<(int)ceil_div(K, 8), 256, 0, (cudaStream_t)stream>>>(w_dev, gx, gy, d, topology, um_dev);
    return check_cuda(cudaGetLastError(), "distance_map_kernel launch");
}

int som_b200_debug_timeline(long long *host_out, int n) {
    SOM_REQUIRE(host_out && n > 0 && n <= 8 * 256, SOM_E_BADARG, "debug_timeline: bad argument");
#ifdef SOM_B200_EXPERIMENTS
    if (tc::env_int("SOM_B200_DBG") == 8) {          // scatter-warp timeline (6 stamps per batch)
        SOM_CUDA(cudaMemcpyFromSymbol(host_out, g_scat_dbg, (size_t)(n < 8 * 64 ? n : 8 * 64) * sizeof(long long)));
        return 0;
    }
#endif
    SOM_CUDA(cudaMemcpyFromSymbol(host_out, tc::g_dbg, (size_t)n * sizeof(long long)));
    return 0;
}

int som_b200_train_host(const float *x_host, int64_t n, int64_t ldx, float *w_host, const som_b200_train_config *cfg,
                        const double *sigma_per_epoch, const double *eta_per_epoch, int n_epochs) {
    SOM_REQUIRE(x_host && w_host && cfg && sigma_per_epoch && eta_per_epoch && n > 0 && n_epochs >= 0, SOM_E_BADARG,
                "train_host: bad argument");
    const int d = cfg->d, K = cfg->gx * cfg->gy;
    SOM_REQUIRE(d > 0 && K > 0 && ldx >= d, SOM_E_BADARG, "train_host: bad shape");
    struct Bufs {
        float *x = nullptr, *w = nullptr, *sc = nullptr, *nd = nullptr, *tab = nullptr, *xs = nullptr, *q = nullptr;
        uint64_t *acc = nullptr;
        uint8_t *ws = nullptr; cudaStream_t st = nullptr;
        ~Bufs() { cudaFree(x); cudaFree(w); cudaFree(sc); cudaFree(nd); cudaFree(tab); cudaFree(xs); cudaFree(q); cudaFree(acc);
                  cudaFree(ws); if (st) cudaStreamDestroy(st); }
    } b;
    const int64_t dld = round_up(d, 4);                 // device row stride: 16-byte aligned rows for TMA / float4
    const size_t ws_bytes = som_b200_shard_workspace_bytes(n, K, d);
    const size_t words = som_b200_accum_words(K, d);
    const size_t dq = (size_t)round_up(d, 4);
    SOM_CUDA(cudaStreamCreateWithFlags(&b.st, cudaStreamNonBlocking));
    SOM_CUDA(cudaMalloc(&b.x, (size_t)n * dld * 4));
    SOM_CUDA(cudaMalloc(&b.w, (size_t)K * d * 4));
    SOM_CUDA(cudaMalloc(&b.sc, ((size_t)K * d + K) * 4));
    SOM_CUDA(cudaMalloc(&b.nd, ((size_t)K * d + K) * 4));
    SOM_CUDA(cudaMalloc(&b.tab, som_b200_neigh_scratch_floats(cfg->gx, cfg->gy, d) * 4));
    SOM_CUDA(cudaMalloc(&b.ws, ws_bytes));
    SOM_CUDA(cudaMalloc(&b.xs, (size_t)n * 4));
    SOM_CUDA(cudaMalloc(&b.q, 3 * dq * 4));             // column maxima | 2^q | 2^-q
    SOM_CUDA(cudaMalloc(&b.acc, words * 8));
    if (dld != d) SOM_CUDA(cudaMemsetAsync(b.x, 0, (size_t)n * dld * 4, b.st));
    SOM_CUDA(cudaMemcpy2DAsync(b.x, dld * 4, x_host, ldx * 4, (size_t)d * 4, (size_t)n, cudaMemcpyHostToDevice, b.st));
    SOM_CUDA(cudaMemcpyAsync(b.w, w_host, (size_t)K * d * 4, cudaMemcpyHostToDevice, b.st));
    SOM_CUDA(cudaMemsetAsync(b.q, 0, 3 * dq * 4, b.st));
    SOM_CUDA(cudaMemsetAsync(b.acc, 0, words * 8, b.st));
    float *S = b.sc, *c = b.sc + (size_t)K * d, *num = b.nd, *den = b.nd + (size_t)K * d;
    float *colmax = b.q, *qscale = b.q + dq, *qinv = b.q + 2 * dq;
    int rc;
    if ((rc = som_b200_prepare_samples(b.x, n, d, dld, b.xs, colmax, b.st))) return rc;
    if ((rc = som_b200_accum_scales(colmax, d, (double)n, qscale, qinv, b.st))) return rc;
    // epoch = BMU search + exact per-BMU sums -> everything else (som_b200_epoch_tail leaves the workspace prepared for
    // the next search and the accumulator cleared)
    if (n_epochs > 0 && (rc = som_b200_prepare_codebook(b.w, K, d, cfg->dist_kind, cfg->p, b.ws, ws_bytes, b.st))) return rc;
    for (int e = 0; e < n_epochs; ++e) {
        if ((rc = som_b200_epoch_accumulate(b.x, n, d, dld, b.xs, b.w, K, cfg->dist_kind, cfg->p, cfg->algo, qscale, b.acc,
                                            nullptr, b.ws, ws_bytes, b.st))) return rc;
        if ((rc = som_b200_epoch_tail(b.acc, qinv, S, c, b.w, cfg->gx, cfg->gy, d, cfg->topology, cfg->neigh_kind,
                                      sigma_per_epoch[e], eta_per_epoch[e], cfg->std_coeff, cfg->compact_support, cfg->dist_kind,
                                      cfg->p, num, den, b.tab, som_b200_neigh_scratch_floats(cfg->gx, cfg->gy, d), b.ws, ws_bytes,
                                      b.st)))
            return rc;
    }
    SOM_CUDA(cudaMemcpyAsync(w_host, b.w, (size_t)K * d * 4, cudaMemcpyDeviceToHost, b.st));
    SOM_CUDA(cudaStreamSynchronize(b.st));
    return 0;
}

}  // extern "C"
