/*
 * som_b200.h — C ABI of the B200-native batch-SOM epoch (libsom_b200.so).
 *
 * Drop-in boundary for ONE hot path of XPySom-Dask: the training epoch
 *   XPySom.train -> XPySom._update -> DistanceFunction + argmin + neighborhood
 *   + dot -> XPySom._merge_updates       (reference xpysom_dask/xpysom.py:420-577)
 * and the inference calls that reuse its BMU search (winner / quantization /
 * quantization_error / distance_map, xpysom.py:370-408, 620-707, 788-817).
 *
 * Conventions
 *   - plain C, no torch / C++ types; every function returns 0 on success, a
 *     positive cudaError_t value on a CUDA failure, or a negative SOM_E_* code
 *     for a rejected argument.  No exceptions cross the boundary.
 *     som_b200_last_error() returns a thread-local message for the last failure.
 *   - "dev" pointers are device pointers owned by the caller (torch tensors in
 *     the Python host class); "host" pointers are ordinary host memory.
 *   - every device entry point takes the CUDA stream to enqueue on
 *     (`void*` == cudaStream_t; NULL is the legacy default stream) and does not
 *     synchronise.
 *   - codebook W is (K, D) fp32 row-major, K = gx*gy, neuron (i,j) at flat
 *     index k = i*gy + j (C-order reshape of the reference's (x, y, D) tensor,
 *     distances.py:185, xpysom.py:240).  Samples X are (n, D) fp32 with a row
 *     stride of ldx floats.
 *   - accumulators: S (K, D) = per-BMU sample sums, c (K) = per-BMU counts.
 *     The reference's numerator/denominator (xpysom.py:436-441) are
 *     num = eta * H^T S, den = eta * H^T c with H[b,k] = h(bmu=b, neuron=k);
 *     the identity is checked in tests/test_oracle_golden.py.
 *   - S and c are ACCUMULATED EXACTLY, as 64-bit fixed-point integers (one power-of-two scale per feature
 *     column, from the column's largest magnitude and the total sample count): integer additions are
 *     associative, so the sums do not depend on the order of the GPU's atomics, on tiling, or on how the
 *     samples are sharded over GPUs (an integer all-reduce of the accumulator gives every rank the bits one
 *     GPU would have computed).  The accumulator is `som_b200_accum_words(k, d)` uint64 words:
 *     `som_b200_accum_replicas(k, d)` copies (4 for accumulators of up to 4 MB, else 1; each copy padded to an even
 *     number of words) of [S: k rows of (d rounded up to even) words | counts: k words], 16-byte aligned, zero before
 *     the first accumulate of an epoch.  The kernels spread their thread blocks over the copies -- samples that share
 *     a BMU would otherwise serialise on the same L2 lines -- and the finalize sums them (integers: same bits).  som_b200_accum_finalize / som_b200_epoch_tail round it ONCE to the fp32
 *     S, c that the neighbourhood apply reads, and clear it.
 */
#ifndef SOM_B200_H
#define SOM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SOM_B200_ABI_VERSION 2

/* activation distances: DistanceFunction table, distances.py:162-170 */
enum som_dist {
    SOM_DIST_EUCLIDEAN = 0, /* 'euclidean'  -2 x.w + |w|^2   distances.py:11-23 */
    SOM_DIST_COSINE    = 1, /* 'cosine'     1 - x.w/(|x||w|) distances.py:45-59 */
    SOM_DIST_MANHATTAN = 2, /* 'manhattan'  sum|x-w|         distances.py:109-158 */
    SOM_DIST_CHEBYSHEV = 3, /* max|x-w| — extension named by north_star, not in the reference */
    SOM_DIST_NORM_P    = 4  /* 'norm_p'     sum|x-w|^p       distances.py:61-107 */
};

/* neighbourhood functions: xpysom.py:255-283, neighborhoods.py:14-130 */
enum som_neigh {
    SOM_NEIGH_GAUSSIAN    = 0,
    SOM_NEIGH_MEXICAN_HAT = 1,
    SOM_NEIGH_BUBBLE      = 2,
    SOM_NEIGH_TRIANGLE    = 3  /* rectangular only, xpysom.py:268-279 */
};

enum som_topology { SOM_TOPO_RECTANGULAR = 0, SOM_TOPO_HEXAGONAL = 1 }; /* xpysom.py:196-206 */

/* which BMU kernel to run */
enum som_algo {
    SOM_ALGO_AUTO      = 0, /* tensor-core kernel when the shape allows, else SIMT */
    SOM_ALGO_SIMT_FP32 = 1, /* shared-memory tiled SIMT, plain fp32 FMA chains */
    SOM_ALGO_TC_3XTF32 = 2, /* tcgen05 kind::tf32 MMA, 3-term hi/lo split (fp32-accurate), TMA staged */
    SOM_ALGO_TC_3XF16  = 3  /* tcgen05 kind::f16 MMA at twice the TF32 rate: fp16 hi/lo split of exactly
                               power-of-two-scaled operands, same 22-bit accuracy; needs xscale_dev */
};

/* negative return codes */
#define SOM_E_BADARG   (-1)  /* NULL pointer, non-positive size, unknown enum */
#define SOM_E_SHAPE    (-2)  /* shape not supported by the requested algo */
#define SOM_E_WORKSPACE (-3) /* workspace too small */
#define SOM_E_NODEVICE (-4)  /* no CUDA device / wrong architecture */

int         som_b200_abi_version(void);
const char *som_b200_last_error(void);

/* Device properties the host class needs (SM count, cc major*10+minor). */
int som_b200_device_info(int *sm_count, int *cc, size_t *smem_per_block_optin);

/* Bytes of scratch som_b200_prepare_codebook / _bmu / _epoch_accumulate need
 * for a codebook of K neurons x D features (prepared operand copies, |w|^2). */
size_t som_b200_workspace_bytes(int k, int d);

/* Which BMU kernel `algo` resolves to for this distance / feature count (SOM_ALGO_AUTO: a tensor-core kernel for
 * the two contraction distances whenever TMA can address the rows and d >= 8: the fp16-split one when row scales
 * are available, the TF32 one otherwise; SIMT for everything else). */
int som_b200_pick_algo(int algo, int dist_kind, int d, int rows_tma_addressable, int has_row_scales);

/* One-time statistics of an uploaded sample matrix (one pass each; valid as long as X is unchanged -- the host
 * class computes them once per upload, not per epoch).  Either output may be NULL.
 *   xscale_dev (n floats): per-row power-of-two scales for SOM_ALGO_TC_3XF16, 2^a_r with
 *                          max_c |x[r,c]| * 2^a_r in [2^14, 2^15);
 *   colmax_dev (d floats): per-column largest magnitude, MAX-ACCUMULATED into the buffer (zero it before the first
 *                          part of an upload; sharded runs all-reduce it with MAX): the input of
 *                          som_b200_accum_scales. */
int som_b200_prepare_samples(const float *x_dev, int64_t n, int d, int64_t ldx, float *xscale_dev, float *colmax_dev,
                             void *stream);

/* Scales of the exact accumulation: qscale[c] = 2^q_c, qinv[c] = 2^-q_c (each d rounded up to a multiple of 4
 * floats, 16-byte aligned), q_c = 62 - (exponent of colmax[c] + 1) - ceil(log2 n_total): n_total samples (over
 * all shards) cannot overflow 63 bits.  Every shard of one job must use the same scales. */
int    som_b200_accum_scales(const float *colmax_dev, int d, double n_total, float *qscale_dev, float *qinv_dev, void *stream);
size_t som_b200_accum_words(int k, int d);
int    som_b200_accum_replicas(int k, int d);
/* Sharded runs: dst_dev == NULL adds replicas 1.. into replica 0 (and clears them) -- then all-reduce only the first
 * som_b200_accum_words(k, d) / som_b200_accum_replicas(k, d) words; dst_dev = a peer accumulator
 * (som_b200_peer_accumulator) adds ALL replicas into it and clears them: on data with hot BMUs, accumulate into a local
 * replicated accumulator, fold it into the peer accumulator, then som_b200_epoch_tail on the peer accumulator. */
int    som_b200_accum_fold_replicas(uint64_t *acc_dev, uint64_t *dst_dev, int k, int d, void *stream);

/* Exact accumulator -> fp32 S (k, d) and c (k), one rounding per element; the accumulator is cleared. */
int som_b200_accum_finalize(uint64_t *acc_dev, const float *qinv_dev, int k, int d, float *s_dev, float *c_dev, void *stream);
/* Epochs whose samples arrive in several parts with different column scales (chunked uploads, blocks streamed from
 * host memory): after each part, fold the accumulator into running fp64 sums sd_dev (k*d + k doubles, zero before
 * the first part; the accumulator is cleared); after the last part round them once to fp32 (sd_dev is cleared). */
int som_b200_accum_fold(uint64_t *acc_dev, const float *qinv_dev, int k, int d, double *sd_dev, void *stream);
int som_b200_accum_finalize_f64(double *sd_dev, int k, int d, float *s_dev, float *c_dev, void *stream);

/* Q: per-epoch codebook preparation.  Replaces the |w|^2 cache of
 * xpysom.py:529-539 and, for the tensor-core kernels, writes the operand
 * copies of the scaled codebook W' (-2 W or -W/|W|) into the workspace: the
 * TF32 hi/lo split (with the bias folded into three spare feature columns
 * when the last 32-feature block has them) and the fp16 hi/lo split times an
 * exact power-of-two scale (one for the codebook, or one per neuron when the
 * neurons' magnitudes spread over more than 2^12).  Must be called after
 * every change of W and before _bmu / _epoch_accumulate with the same ws
 * (som_b200_epoch_tail does it for the codebook it has just merged). */
int som_b200_prepare_codebook(const float *w_dev, int k, int d, int dist_kind, float p,
                              void *ws_dev, size_t ws_bytes, void *stream);

/* W: BMU search.  Replaces XPySom._winner (xpysom.py:410-417): activation
 * distance + first-minimum argmin over the K neurons, fused, the (n,K) matrix
 * is never materialised.  bmu_dev (n) receives flat indices.  best_dev may be
 * NULL; otherwise it receives the kernel's winning score (distance up to the
 * row-constant terms the argmin does not need). */
/* xscale_dev: output of som_b200_prepare_samples, or NULL (then SOM_ALGO_AUTO never picks the fp16
 * kernel and SOM_ALGO_TC_3XF16 is rejected). */
int som_b200_bmu(const float *x_dev, int64_t n, int d, int64_t ldx, const float *xscale_dev,
                 const float *w_dev, int k, int dist_kind, float p, int algo,
                 int32_t *bmu_dev, float *best_dev,
                 void *ws_dev, size_t ws_bytes, void *stream);

/* Full distance matrix out_dev (n, K), for the callers that want it (activate xpysom.py:323-354,
 * distance_from_weights :647-671, topographic_error :709-746).  mode 0: the activation distance as the
 * reference defines it (euclidean = partial -2x.w+|w|^2, cosine = 1 - nan_to_num(sim), L1 / Linf / Lp sums);
 * mode 1: the Euclidean distance sqrt(|x-w|^2) with nan_to_num (distances.py:33-43) whatever dist_kind;
 * mode 2: the full squared Euclidean distance (-2x.w+|w|^2)+|x|^2 of 'euclidean_no_opt' (distances.py:25-31). */
int som_b200_distances(const float *x_dev, int64_t n, int d, int64_t ldx,
                       const float *w_dev, int k, int dist_kind, float p, int mode,
                       float *out_dev, void *ws_dev, size_t ws_bytes, void *stream);

/* Best and second-best matching unit of every row on the Euclidean distance (sqrt form, nan_to_num), fused: what
 * topographic_error needs (xpysom.py:709-746 argsorts the whole (n, K) matrix and keeps two columns).
 * top2_dev receives 2 n flat indices: [2r] = best, [2r+1] = second best. */
int som_b200_top2(const float *x_dev, int64_t n, int d, int64_t ldx, const float *w_dev, int k,
                  int32_t *top2_dev, void *ws_dev, size_t ws_bytes, void *stream);

/* U (first half): S[bmu[r], :] += X[r, :], c[bmu[r]] += 1 for the n rows, exactly (see "accumulators" above).
 * Replaces the sample side of g^T X and sum(g) in XPySom._update (xpysom.py:434-441). */
int som_b200_accumulate(const float *x_dev, int64_t n, int d, int64_t ldx,
                        const int32_t *bmu_dev, int k,
                        const float *qscale_dev, uint64_t *acc_dev, void *stream);

/* W+U fused: som_b200_bmu followed by som_b200_accumulate on one shard of rows
 * — the unit of distributed work (`_update` on one Dask block, xpysom.py:551).
 * bmu_dev may be NULL if the caller does not want the indices; then
 * ws must also hold n int32 (see som_b200_shard_workspace_bytes). */
int som_b200_epoch_accumulate(const float *x_dev, int64_t n, int d, int64_t ldx, const float *xscale_dev,
                              const float *w_dev, int k, int dist_kind, float p, int algo,
                              const float *qscale_dev, uint64_t *acc_dev, int32_t *bmu_dev,
                              void *ws_dev, size_t ws_bytes, void *stream);
size_t som_b200_shard_workspace_bytes(int64_t n, int k, int d);

/* U (second half) + N: num = eta * H^T S, den = eta * H^T c with the
 * neighbourhood evaluated on the fly for every (bmu, neuron) pair.  Replaces
 * neighborhoods.py:14-130 and xpysom.py:434-441.  num (K,D) and den (K) are
 * overwritten.  tables_dev is scratch of som_b200_neigh_table_floats(gx,gy)
 * floats (tables_floats says how many were given).  Returns SOM_E_SHAPE for the combinations the reference rejects
 * (triangle on a hexagonal map; mexican_hat + compact_support on a rectangular
 * map with gx != gy, whose broadcast raises in neighborhoods.py:69-71). */
int som_b200_neigh_apply(const float *s_dev, const float *c_dev, int gx, int gy, int d,
                         int topology, int neigh_kind, double sigma, double eta,
                         double std_coeff, int compact_support,
                         float *num_dev, float *den_dev, float *tables_dev, size_t tables_floats, void *stream);
/* minimum scratch (factor tables only: the direct K^2 D kernel, reduction over BMUs not sliced) */
size_t som_b200_neigh_table_floats(int gx, int gy);
/* scratch that also holds the intermediates of the two-pass separable path (rectangular gaussian / bubble /
 * triangle on maps of >= 4096 neurons: 2 K (gx+gy) D flops instead of 2 K^2 D) and the per-slice partial blocks
 * of the direct kernel (small maps slice the reduction over BMUs to fill the GPU; the slices are summed in slice
 * order, never with atomics, so num / den are bit-reproducible) */
size_t som_b200_neigh_scratch_floats(int gx, int gy, int d);

/* Graph-replay variant of som_b200_neigh_apply: sigma and the learning rate are read on the device,
 * sigma = sched_dev[2e], eta = sched_dev[2e+1] with e = *epoch_dev, so ONE captured CUDA graph of the epoch
 * (prepare, BMU+accumulate, all-reduce, apply, merge, som_b200_epoch_advance) serves every epoch. */
int som_b200_neigh_apply_sched(const float *s_dev, const float *c_dev, int gx, int gy, int d,
                               int topology, int neigh_kind, const double *sched_dev, const int *epoch_dev,
                               double std_coeff, int compact_support,
                               float *num_dev, float *den_dev, float *tables_dev, size_t tables_floats, void *stream);
/* *epoch_dev += 1 (last node of the epoch graph) */
int som_b200_epoch_advance(int *epoch_dev, void *stream);

/* M: W <- den != 0 ? num/den : W   (XPySom._merge_updates, xpysom.py:446-455). */
int som_b200_merge(float *w_dev, const float *num_dev, const float *den_dev,
                   int k, int d, void *stream);

/* Everything between two BMU searches in one call: som_b200_accum_finalize (when acc_dev != NULL; with NULL, s_dev
 * and c_dev must already hold the epoch's fp32 sums) + som_b200_neigh_apply (K4) + som_b200_merge (M) +
 * som_b200_prepare_codebook of the NEW codebook (Q, xpysom.py:529-539, for the next epoch); the accumulator is
 * left cleared (xpysom.py:516-527).  On small maps (direct neighbourhood kernel with 64x64 tiles: at most 64 features or
 * fewer than 512 neurons, K*D <= 2^20) this is ONE cooperative kernel with grid-wide barriers between the
 * phases instead of ten stream operations (csrc/epoch_tail.cuh); otherwise it issues the separate launches.
 * Same results either way.  The workspace must have been prepared before (som_b200_prepare_codebook), as it
 * is for the BMU search that precedes this call.  The caller's epoch becomes:
 * som_b200_epoch_accumulate -> (all-reduce of [S | c]) -> som_b200_epoch_tail. */
int som_b200_epoch_tail(uint64_t *acc_dev, const float *qinv_dev, float *s_dev, float *c_dev, float *w_dev,
                        int gx, int gy, int d,
                        int topology, int neigh_kind, double sigma, double eta, double std_coeff,
                        int compact_support, int dist_kind, float p,
                        float *num_dev, float *den_dev, float *tables_dev, size_t tables_floats,
                        void *ws_dev, size_t ws_bytes, void *stream);

/* ---- long rows: one tensor-core pass + exact refinement (csrc/bmu_filter.cuh) -------------------
 * For Euclidean maps with 256 <= d <= 1024 features and >= 1024 neurons the three-pass contraction is bound by the
 * tensor pipe.  This path runs ONE fp16 pass on centred operands that yields, per (sample, neuron), an interval
 * guaranteed to contain the score; neurons whose lower bound exceeds the row's smallest upper bound cannot be the
 * BMU; the few survivors are re-scored exactly (fp64 sum of squared differences of the caller's fp32 arrays) and the
 * first minimum wins.  Replaces _winner / _activate (xpysom.py:336-354,410-417) for those shapes.
 *   som_b200_filter_prepare_samples   once per upload: centre, fp16 copy of the samples, per-row statistics
 *   som_b200_bmu_filter               BMUs of all rows (the codebook side is prepared inside, every call); bmu_dev holds the
 *                                     previous epoch's BMUs on entry (seeds of the bounds; -1 = none); with acc_dev /
 *                                     qscale_dev non-NULL the refine also adds every row it resolved to the exact accumulator
 * After som_b200_bmu_filter the workspace holds, at byte som_b200_filter_overflow_offset: int32 = rows whose candidate
 * lists overflowed (their bmu is -1: the caller re-does them with som_b200_bmu), and at +8 a uint64 = candidates
 * re-scored in that call (long lists make the refinement dearer than the two tensor passes it saves: the caller's
 * policy input).  An epoch is then: som_b200_bmu_filter with the accumulator, then som_b200_bmu + som_b200_accumulate on
 * the overflowed rows only. */
int    som_b200_filter_eligible(const float *x_dev, int64_t n, int d, int64_t ldx, int k, int dist_kind);
size_t som_b200_filter_workspace_bytes(int64_t n, int k, int d);
size_t som_b200_filter_overflow_offset(int64_t n, int k, int d);
int    som_b200_filter_prepare_samples(const float *x_dev, int64_t n, int d, int64_t ldx, void *fws_dev, size_t fws_bytes, void *stream);
int    som_b200_bmu_filter(const float *x_dev, int64_t n, int d, int64_t ldx, const float *w_dev, int k, int32_t *bmu_dev,
                           const float *qscale_dev, uint64_t *acc_dev, void *fws_dev, size_t fws_bytes, void *stream);

/* ---- sharded path: the one exchange step, fused into the epoch tail ---------------------------
 * The reference sums the per-block partial updates with Dask (`sum(...)` over the delayed `_update`
 * results, xpysom.py:545-558).  With one process per GPU that is ONE sum of the exact accumulators
 * [S | counts] across the ranks per epoch: an all-reduce of the integer buffer by the caller (NCCL),
 * or -- single node, small maps, where that all-reduce is pure latency -- accumulators that live in
 * NVLink peer memory (CUDA IPC) and are summed over the ranks by the finalize phase itself:
 *   every rank creates a communicator for `acc_words` (som_b200_accum_words) 64-bit words, the 64-byte
 *   handles are exchanged by the host (any transport), every rank connects;
 *   per epoch every rank passes som_b200_peer_accumulator(comm) as `acc_dev` to
 *   som_b200_epoch_accumulate and then to som_b200_epoch_tail (or som_b200_accum_finalize), which
 *   recognise it, wait for all ranks' accumulators of this exchange and read them over NVLink.  The
 *   pointer alternates between two buffers, so ask again every epoch.  Integer sums are exact:
 *   every rank gets the same bits, equal to one GPU holding all the rows.
 * All ranks must make the same sequence of calls.  Before som_b200_peer_destroy the caller must make
 * sure (a barrier over the ranks) that no peer is still inside an exchange. */
int som_b200_peer_create(size_t acc_words, int world, int rank, void **comm_out, void *handle_out_64_bytes);
int som_b200_peer_connect(void *comm, const void *all_handles_world_x_64_bytes);
uint64_t *som_b200_peer_accumulator(void *comm);
int som_b200_peer_destroy(void *comm);

/* quantization / quantization_error support (xpysom.py:620-707): for each row,
 * q_dev (n,D) <- W[bmu[r]] if q_dev != NULL, and err_dev (n) <- ||x_r - W[bmu[r]]||_2
 * if err_dev != NULL. */
int som_b200_quantize(const float *x_dev, int64_t n, int d, int64_t ldx,
                      const float *w_dev, int k, const int32_t *bmu_dev,
                      float *q_dev, float *err_dev, void *stream);

/* distance_map (xpysom.py:788-817): um (gx*gy) <- sum of Euclidean distances
 * to the 8 (rectangular) or 6 (hexagonal) grid neighbours, NOT normalised;
 * the host divides by the maximum. */
int som_b200_distance_map(const float *w_dev, int gx, int gy, int d, int topology,
                          float *um_dev, void *stream);

/* Experiments builds only (-DSOM_B200_EXPERIMENTS): with SOM_B200_DBG=9 the fp16 kernel stamps clock64() at its
 * pipeline hand-off points; this copies the first n stamps (8 per tile) to host memory (zeros otherwise). */
int som_b200_debug_timeline(long long *host_out, int n);

/* Whole-job entry with HOST buffers: what a maintainer of the reference would
 * bind from XPySom.train (xpysom.py:458-594) with numpy arrays — uploads the
 * samples once, runs epochs [iter_beg, iter_end) with the given per-epoch
 * sigma / learning-rate schedule (length iter_end - iter_beg, computed by the
 * reference's own decays.py on the host), downloads the codebook.  Single GPU
 * (the current device).  Synchronous. */
typedef struct som_b200_train_config {
    int    gx, gy, d;
    int    topology, neigh_kind, dist_kind, algo, compact_support;
    float  p;
    double std_coeff;
} som_b200_train_config;

int som_b200_train_host(const float *x_host, int64_t n, int64_t ldx,
                        float *w_host, const som_b200_train_config *cfg,
                        const double *sigma_per_epoch, const double *eta_per_epoch,
                        int n_epochs);

#ifdef __cplusplus
}
#endif
#endif /* SOM_B200_H */
