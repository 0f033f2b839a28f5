"""Device engine: torch tensors for memory and streams, libsom_b200 for every kernel.

One engine instance drives one GPU.  All methods enqueue on torch's current
CUDA stream and take / return torch tensors living on that GPU; nothing here
computes with torch ops on the hot path.  Without a CUDA device or without the
shared object the constructor raises — there is no CPU fallback.
"""
import ctypes

import torch

from . import _lib


class CudaEngine:
    name = "cuda"

    def __init__(self, device=None):
        if not torch.cuda.is_available():
            raise _lib.SomB200Error(
                "xpysom_dask_b200 needs a CUDA device (B200, sm_100a); torch.cuda.is_available() is False. "
                "There is no CPU fallback.")
        self.lib = _lib.load()
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device)
        with torch.cuda.device(self.device):
            sm, cc = ctypes.c_int(), ctypes.c_int()
            smem = ctypes.c_size_t()
            _lib.check(self.lib.som_b200_device_info(ctypes.byref(sm), ctypes.byref(cc), ctypes.byref(smem)),
                       "som_b200_device_info")
        self.sm_count, self.cc = sm.value, cc.value
        self.launches = 0          # kernels enqueued through this engine (bench.py reports it)

    # -- helpers ---------------------------------------------------------------
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _p(t):
        if t is None or isinstance(t, ctypes.c_void_p):        # (a raw device address: peer accumulators)
            return t
        return ctypes.c_void_p(t.data_ptr())

    def empty(self, *shape, dtype=torch.float32):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    def zeros(self, *shape, dtype=torch.float32):
        return torch.zeros(*shape, dtype=dtype, device=self.device)

    def to_device(self, host_tensor):
        """Host (pinned if possible) -> device copy on the current stream."""
        return host_tensor.to(self.device, non_blocking=True)

    def workspace(self, n, k, d):
        nbytes = self.lib.som_b200_shard_workspace_bytes(int(n), int(k), int(d))
        return torch.empty(nbytes, dtype=torch.uint8, device=self.device)

    def neigh_tables(self, gx, gy, d=None):
        """Scratch of the neighbourhood apply: factor tables, plus (when d is given) the intermediates of
        the two-pass separable path."""
        if d is None:
            return self.empty(self.lib.som_b200_neigh_table_floats(int(gx), int(gy)))
        return self.empty(self.lib.som_b200_neigh_scratch_floats(int(gx), int(gy), int(d)))

    # -- kernels -----------------------------------------------------------------
    def prepare_codebook(self, w, dist_kind, p, ws):
        k, d = w.shape
        with torch.cuda.device(self.device):
            _lib.check(self.lib.som_b200_prepare_codebook(self._p(w), k, d, dist_kind, float(p), self._p(ws),
                                                          ws.numel(), self._stream()), "som_b200_prepare_codebook")
        self.launches += 1

    def prepare_samples(self, x, want_scale=True, out=None, colmax=None):
        """One-time statistics of an uploaded sample matrix: the per-row power-of-two scales of the fp16-split kernel
        (want_scale) and the per-column largest magnitudes the exact accumulation derives its scales from
        (max-accumulated into `colmax` when one is passed: the parts of one upload share it).  -> (xscale | None, colmax)"""
        n, d = x.shape
        xs = None
        if want_scale:
            xs = out if out is not None else self.empty(max(n, 1))
        if colmax is None:
            colmax = self.zeros((d + 3) // 4 * 4)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.som_b200_prepare_samples(self._p(x), n, d, x.stride(0), self._p(xs), self._p(colmax),
                                                         self._stream()), "som_b200_prepare_samples")
        self.launches += 2 if want_scale else 1
        return xs, colmax

    def accum_scales(self, colmax, d, n_total):
        """(qscale, qinv): 2^q_c and 2^-q_c per feature column for the exact accumulation of n_total samples."""
        dq = (d + 3) // 4 * 4
        q = self.empty(2 * dq)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.som_b200_accum_scales(self._p(colmax), d, float(n_total), self._p(q[:dq]), self._p(q[dq:]),
                                                      self._stream()), "som_b200_accum_scales")
        self.launches += 1
        return q[:dq], q[dq:]

    def accumulator(self, k, d):
        """Zeroed exact accumulator [S | counts] of som_b200_accum_words(k, d) 64-bit words."""
        return self.zeros(self.lib.som_b200_accum_words(int(k), int(d)), dtype=torch.int64)

    def accum_one_copy(self, acc, k, d):
        """The part of a local accumulator that has to travel in an all-reduce: its replicas folded into the first."""
        reps = self.lib.som_b200_accum_replicas(int(k), int(d))
        if reps <= 1:
            return acc
        with torch.cuda.device(self.device):
            _lib.check(self.lib.som_b200_accum_fold_replicas(self._p(acc), None, int(k), int(d), self._stream()),
                       "som_b200_accum_fold_replicas")
        self.launches += 1
        return acc[:acc.numel() // reps]

    def accum_fold_into(self, acc, dst, k, d):
        """All replicas of the local accumulator added into dst (a peer accumulator) and cleared."""
        with torch.cuda.device(self.device):
            _lib.check(self.lib.som_b200_accum_fold_replicas(self._p(acc), self._p(dst), int(k), int(d), self._stream()),
                       "som_b200_accum_fold_replicas")
        self.launches += 1

    def accum_finalize(self, acc, qinv, k, d, s, c):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.som_b200_accum_finalize(self._p(acc), self._p(qinv), k, d, self._p(s), self._p(c),
                                                        self._stream()), "som_b200_accum_finalize")
        self.launches += 1

    def accum_fold(self, acc, qinv, k, d, sd):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.som_b200_accum_fold(self._p(acc), self._p(qinv), k, d, self._p(sd), self._stream()),
                       "som_b200_accum_fold")
        self.launches += 1

    def accum_finalize_f64(self, sd, k, d, s, c):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.som_b200_accum_finalize_f64(self._p(sd), k, d, self._p(s), self._p(c), self._stream()),
                       "som_b200_accum_finalize_f64")
        self.launches += 1

    def bmu(self, x, w, dist_kind, p, algo, ws, bmu_out=None, best_out=None, xscale=None):
        n, d = x.shape
        k = w.shape[0]
        if bmu_out is None:
            bmu_out = self.empty(n, dtype=torch.int32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.som_b200_bmu(self._p(x), n, d, x.stride(0), self._p(xscale), self._p(w), k, dist_kind,
                                             float(p), algo,
                                             self._p(bmu_out), self._p(best_out), self._p(ws), ws.numel(),
                                             self._stream()), "som_b200_bmu")
        self.launches += 1
        return bmu_out

    def distances(self, x, w, dist_kind, p, mode, ws):
        """(n, K) distance matrix: mode 0 = activation distance, 1 = Euclidean distance (sqrt), 2 = squared Euclidean."""
        n, d = x.shape
        k = w.shape[0]
        out = self.empty(n, k)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.som_b200_distances(self._p(x), n, d, x.stride(0), self._p(w), k, dist_kind, float(p),
                                                   int(mode), self._p(out), self._p(ws), ws.numel(), self._stream()),
                       "som_b200_distances")
        self.launches += 2
        return out

    def top2(self, x, w, ws):
        """(n, 2) int32: best and second-best unit of every row on the Euclidean distance (fused, no (n, K) matrix)."""
        n, d = x.shape
        out = self.empty(n, 2, dtype=torch.int32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.som_b200_top2(self._p(x), n, d, x.stride(0), self._p(w), w.shape[0], self._p(out),
                                              self._p(ws), ws.numel(), self._stream()), "som_b200_top2")
        self.launches += 2
        return out

    def accumulate(self, x, bmu, k, qscale, acc):
        n, d = x.shape
        with torch.cuda.device(self.device):
            _lib.check(self.lib.som_b200_accumulate(self._p(x), n, d, x.stride(0), self._p(bmu), k, self._p(qscale),
                                                    self._p(acc), self._stream()), "som_b200_accumulate")
        self.launches += 1

    def epoch_accumulate(self, x, w, dist_kind, p, algo, qscale, acc, ws, bmu_out=None, xscale=None):
        n, d = x.shape
        k = w.shape[0]
        with torch.cuda.device(self.device):
            _lib.check(self.lib.som_b200_epoch_accumulate(self._p(x), n, d, x.stride(0), self._p(xscale), self._p(w), k,
                                                          dist_kind, float(p), algo, self._p(qscale), self._p(acc),
                                                          self._p(bmu_out), self._p(ws), ws.numel(), self._stream()),
                       "som_b200_epoch_accumulate")
        self.launches += 1

    # -- long rows: one tensor-core pass + exact refinement (csrc/bmu_filter.cuh) ---------------------------------
    def filter_eligible(self, x, k, dist_kind):
        n, d = x.shape
        return bool(self.lib.som_b200_filter_eligible(self._p(x), n, d, x.stride(0), int(k), int(dist_kind)))

    def filter_workspace(self, x, k):
        """Workspace of the filter path for the samples x (prepared: centre, fp16 copy, row statistics)."""
        n, d = x.shape
        fws = torch.empty(self.lib.som_b200_filter_workspace_bytes(n, int(k), d), dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.som_b200_filter_prepare_samples(self._p(x), n, d, x.stride(0), self._p(fws), fws.numel(),
                                                                self._stream()), "som_b200_filter_prepare_samples")
        self.launches += 3
        return fws

    def filter_stats(self, fws, n, k, d):
        """(rows whose candidate lists overflowed -- their bmu is -1 --, candidates re-scored) of the last bmu_filter call.
        Synchronises."""
        off = self.lib.som_b200_filter_overflow_offset(int(n), int(k), int(d))
        v = fws[off:off + 16].view(torch.int64).cpu()
        return int(v[0].item()) & 0xffffffff, int(v[1].item())

    def bmu_filter(self, x, w, fws, bmu_out=None, qscale=None, acc=None):
        """BMUs through the one-pass filter + refinement; bmu_out holds the previous epoch's BMUs on entry (-1: none).
        With qscale / acc the refine also adds every resolved row to the exact accumulator (overflowed rows, -1, not)."""
        n, d = x.shape
        if bmu_out is None:
            bmu_out = torch.full((n,), -1, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.som_b200_bmu_filter(self._p(x), n, d, x.stride(0), self._p(w), w.shape[0], self._p(bmu_out),
                                                    self._p(qscale), self._p(acc), self._p(fws), fws.numel(), self._stream()),
                       "som_b200_bmu_filter")
        self.launches += 3
        return bmu_out

    def neigh_apply(self, s, c, gx, gy, d, topology, neigh_kind, sigma, eta, std_coeff, compact, num, den, tables):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.som_b200_neigh_apply(self._p(s), self._p(c), gx, gy, d, topology, neigh_kind,
                                                     float(sigma), float(eta), float(std_coeff), int(bool(compact)),
                                                     self._p(num), self._p(den), self._p(tables), tables.numel(),
                                                     self._stream()),
                       "som_b200_neigh_apply")
        self.launches += 2

    def epoch_tail(self, acc, qinv, s, c, w, gx, gy, d, topology, neigh_kind, sigma, eta, std_coeff, compact, dist_kind, p,
                   num, den, tables, ws):
        """accum_finalize (acc is not None) + neigh_apply + merge + prepare_codebook(new W): one cooperative kernel on
        small maps.  The accumulator is left cleared; s, c hold the epoch's fp32 sums."""
        with torch.cuda.device(self.device):
            _lib.check(self.lib.som_b200_epoch_tail(self._p(acc), self._p(qinv), self._p(s), self._p(c), self._p(w), gx, gy, d,
                                                    topology, neigh_kind,
                                                    float(sigma), float(eta), float(std_coeff), int(bool(compact)),
                                                    dist_kind, float(p), self._p(num), self._p(den), self._p(tables),
                                                    tables.numel(), self._p(ws), ws.numel(), self._stream()),
                       "som_b200_epoch_tail")
        self.launches += 1

    def neigh_apply_sched(self, s, c, gx, gy, d, topology, neigh_kind, sched, epoch, std_coeff, compact, num, den, tables):
        """neigh_apply with sigma / eta read on the device from sched[2e], sched[2e+1], e = epoch[0]."""
        with torch.cuda.device(self.device):
            _lib.check(self.lib.som_b200_neigh_apply_sched(self._p(s), self._p(c), gx, gy, d, topology, neigh_kind,
                                                           self._p(sched), self._p(epoch), float(std_coeff),
                                                           int(bool(compact)), self._p(num), self._p(den),
                                                           self._p(tables), tables.numel(), self._stream()),
                       "som_b200_neigh_apply_sched")
        self.launches += 2

    def epoch_advance(self, epoch):
        with torch.cuda.device(self.device):
            _lib.check(self.lib.som_b200_epoch_advance(self._p(epoch), self._stream()), "som_b200_epoch_advance")
        self.launches += 1

    def merge(self, w, num, den):
        k, d = w.shape
        with torch.cuda.device(self.device):
            _lib.check(self.lib.som_b200_merge(self._p(w), self._p(num), self._p(den), k, d, self._stream()),
                       "som_b200_merge")
        self.launches += 1

    def quantize(self, x, w, bmu, want_q=False, want_err=True):
        n, d = x.shape
        q = self.empty(n, d) if want_q else None
        err = self.empty(n) if want_err else None
        with torch.cuda.device(self.device):
            _lib.check(self.lib.som_b200_quantize(self._p(x), n, d, x.stride(0), self._p(w), w.shape[0], self._p(bmu),
                                                  self._p(q), self._p(err), self._stream()), "som_b200_quantize")
        self.launches += 1
        return q, err

    def distance_map(self, w, gx, gy, topology):
        um = self.empty(gx * gy)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.som_b200_distance_map(self._p(w), gx, gy, w.shape[1], topology, self._p(um),
                                                      self._stream()), "som_b200_distance_map")
        self.launches += 1
        return um
