"""`XPySom` — the reference's model API, with the batch-SOM epoch on B200 kernels.

Host-side mirror of ``xpysom_dask/xpysom.py:72-892``: same constructor
arguments, defaults, validation errors and method names, so it drops in as the
compute backend for the training path (``train`` -> ``_update`` ->
``_merge_updates``) and for the inference calls that reuse its BMU search.
What changed underneath:

* numpy/CuPy dispatch (``xp``) and Dask scheduling (``use_dask``,
  ``dask_chunks``) are accepted and ignored; every array operation of the epoch
  runs in ``libsom_b200.so`` (hand-written sm_100a CUDA) through a C ABI;
* ``n_parallel`` chunking (xpysom.py:560-569) becomes kernel tiling;
* Dask row-blocks (xpysom.py:545-558) become shards: one process per GPU, each
  calling ``train`` on its own rows with ``process_group`` set, and ONE
  all-reduce of the (K*D + K) per-BMU sums per epoch; the codebook stays
  replicated.

``self._weights`` stays a host numpy array of shape (x, y, input_len) that
callers may assign to between calls, exactly as with the reference
(tests.py:31-33); it is re-uploaded at every public call.
"""
from collections import Counter, defaultdict
from warnings import warn

import os
import weakref

import numpy as np
import torch

from . import _lib
from .decays import DECAY_FUNCTIONS

_NEIGHBORHOODS = {
    "rectangular": ("gaussian", "mexican_hat", "bubble", "triangle"),
    "hexagonal": ("gaussian", "mexican_hat", "bubble"),          # xpysom.py:271-279
}
_DISTANCES = ("euclidean", "euclidean_no_opt", "manhattan", "manhattan_no_opt", "cosine", "norm_p",
              "norm_p_no_opt", "chebyshev")                       # distances.py:162-170 + chebyshev
_UNPICKLED_SHARDED = 'sharded (unpickled)'
_DEFAULT_N_PARALLEL = 148 * 2048     # SMs x max threads/SM on B200: the rule of utils.py:4-11


def _as_f32_matrix(data):
    """Any array-like / torch tensor -> 2-D float32 torch tensor (no copy when possible)."""
    if not isinstance(data, (torch.Tensor, np.ndarray)) and hasattr(data, '__dlpack__'):
        # device arrays of other libraries (CuPy -- the reference's own GPU input, xpysom.py:487-510 -- JAX, ...)
        # come in through DLPack without a copy
        data = torch.from_dlpack(data)
    if isinstance(data, torch.Tensor):
        t = data
        if t.dtype != torch.float32:
            t = t.to(torch.float32)
    else:
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(data, dtype=np.float32)))
    if t.dim() == 1:
        t = t.unsqueeze(0)
    if t.dim() != 2:
        raise ValueError("data must be a 2-D (samples, features) matrix")
    return t


class XPySom:
    def __init__(self, x, y, input_len, sigma=0, sigmaN=1, learning_rate=0.5, learning_rateN=0.01,
                 decay_function='exponential', neighborhood_function='gaussian', std_coeff=0.5,
                 topology='rectangular', activation_distance='euclidean', activation_distance_kwargs={},
                 random_seed=None, n_parallel=0, compact_support=False, xp=None, use_dask=False,
                 dask_chunks='auto', *, device=None, algo='auto', process_group=None, use_cuda_graph=False,
                 max_resident_bytes=None):
        """Same arguments as the reference constructor (xpysom.py:73-82).

        Keyword-only additions: ``device`` (CUDA device of this process),
        ``algo`` ('auto' | 'tc' | 'simt': which BMU kernel), ``process_group``
        (a torch.distributed group, or True for the default group: this process
        holds one shard of the samples) and ``use_cuda_graph`` (replay one
        captured CUDA graph per epoch instead of launching the epoch's ~10 kernels one by one; capture
        costs a few milliseconds, so it only pays off for runs of several hundred epochs on small maps) and
        ``max_resident_bytes`` (host arrays larger than this -- default: 80 % of the free device memory -- are
        streamed through two block buffers every epoch instead of being uploaded once).
        """
        if sigma >= x or sigma >= y:
            warn('Warning: sigma is too high for the dimension of the map.')
        self._random_generator = np.random.RandomState(random_seed)
        self.xp = xp                       # accepted for signature compatibility, unused
        self.use_dask = False              # Dask scheduling is replaced by GPU shards
        self.dask_chunks = dask_chunks
        self._learning_rate = learning_rate
        self._learning_rateN = learning_rateN
        self._sigma = min(x, y) / 2 if sigma == 0 else sigma
        self._std_coeff = std_coeff
        self._sigmaN = sigmaN
        self._input_len = input_len

        # bit-compatible initial codebook: same generator, same call order (xpysom.py:167,189-190)
        self._weights = self._random_generator.rand(x, y, input_len) * 2 - 1
        self._weights /= np.linalg.norm(self._weights, axis=-1, keepdims=True)

        self._neigx = np.arange(x)
        self._neigy = np.arange(y)
        if topology not in _NEIGHBORHOODS:
            raise ValueError('%s not supported only hexagonal and rectangular available' % topology)
        self.topology = topology
        self._xx, self._yy = np.meshgrid(self._neigx, self._neigy)
        self._xx = self._xx.astype(float)
        self._yy = self._yy.astype(float)
        if topology == 'hexagonal':
            self._xx[::-2] -= 0.5          # xpysom.py:206
            if neighborhood_function in ['triangle']:
                warn('triangle neighborhood function does not take in account hexagonal topology')

        if decay_function not in DECAY_FUNCTIONS:
            raise ValueError('%s not supported. Functions available: %s'
                             % (decay_function, ', '.join(DECAY_FUNCTIONS.keys())))
        self._decay_function_name = decay_function
        self.compact_support = compact_support
        if neighborhood_function not in _NEIGHBORHOODS[topology]:
            raise ValueError('%s not supported. Functions available: %s'
                             % (neighborhood_function, ', '.join(_NEIGHBORHOODS[topology])))
        self.neighborhood_func_name = neighborhood_function
        if activation_distance not in _DISTANCES:
            raise ValueError('%s not supported. Distances available: %s'
                             % (activation_distance, ', '.join(_DISTANCES)))
        self._activation_distance_name = activation_distance
        self._activation_distance_kwargs = dict(activation_distance_kwargs)
        if n_parallel == 0:
            n_parallel = _DEFAULT_N_PARALLEL
        self._n_parallel = n_parallel      # kept for API parity; tiling is internal to the kernels

        if algo not in _lib.ALGO:
            raise ValueError("algo must be one of %s" % ', '.join(_lib.ALGO))
        self._algo = algo
        self._device = device
        self._process_group = process_group
        self._engine = None                # the CudaEngine of this process, created on first use
        self._use_cuda_graph = bool(use_cuda_graph)
        self._max_resident_bytes = max_resident_bytes
        self._profile = False              # bench.py: record CUDA events around the BMU / accumulate kernels
        self._profile_events = []
        self.stats = {}

    # ------------------------------------------------------------------ plumbing
    @property
    def _decay_function(self):
        return DECAY_FUNCTIONS[self._decay_function_name]

    def _get_engine(self):
        if self._engine is None:
            from .engine import CudaEngine
            self._engine = CudaEngine(self._device)
        return self._engine

    def _dist_kind(self, euclidean_only=False):
        if euclidean_only:
            return _lib.DIST['euclidean'], 2.0
        name = self._activation_distance_name
        p = float(self._activation_distance_kwargs.get('p', 2))
        return _lib.DIST[name], p

    def _wants_xscale(self, dist_kind):
        return self._algo in ('auto', 'tc16') and dist_kind in (_lib.DIST['euclidean'], _lib.DIST['cosine'])

    def _group(self):
        pg = self._process_group
        if pg is None or pg is False:
            return None
        import torch.distributed as dist
        if not dist.is_initialized():
            raise RuntimeError("this model trains one shard of a sharded job (process_group was set%s) but "
                               "torch.distributed is not initialised" % (", then pickled" if pg == _UNPICKLED_SHARDED else ""))
        # a pickled sharded model comes back attached to the default group (group handles do not pickle)
        return dist.group.WORLD if (pg is True or pg == _UNPICKLED_SHARDED) else pg

    def _peer_accumulator(self, eng, group, words):
        """The cached peer-memory accumulators of a sharded model (peer.py), or None when NCCL should all-reduce.
        Collective: every rank takes the same decision (same sizes, and the set-up itself agrees by all-reduce)."""
        from . import peer
        import torch.distributed as dist
        if (words * 8 > peer.PEER_MAX_BYTES or dist.get_world_size(group) < 2 or getattr(eng, 'name', '') != 'cuda'
                or os.environ.get('SOM_B200_PEER', '1') == '0'):
            return None
        key = (id(group), words, str(eng.device))
        cached = getattr(self, '_peer_cache', None)
        if cached is None or cached[0] != key:
            if cached is not None:
                cached[1].close()
            cached = (key, peer.PeerAccumulator(eng, group, words))
            self._peer_cache = cached
        return cached[1] if cached[1].active else None

    def _shape(self):
        gx, gy, d = self._weights.shape
        return gx, gy, d

    def _weights_to_device(self, eng):
        gx, gy, d = self._shape()
        w = torch.from_numpy(np.ascontiguousarray(self._weights, dtype=np.float32).reshape(gx * gy, d))
        return eng.to_device(w)

    def _data_to_device(self, eng, data):
        """Samples -> fp32 device matrix whose rows start on 16-byte boundaries (row stride a
        multiple of 4 floats), which is what TMA and the float4 paths need.  A CUDA tensor that
        already satisfies this is used in place (no copy)."""
        t = _as_f32_matrix(data)
        n, d = t.shape
        ok = (t.device == eng.device and t.stride(1) == 1 and t.stride(0) >= d and t.stride(0) % 4 == 0
              and t.data_ptr() % 16 == 0)
        if ok:
            return t
        ld = (d + 3) // 4 * 4
        buf = torch.empty((n, ld), dtype=torch.float32, device=eng.device)
        view = buf[:, :d]
        view.copy_(t, non_blocking=True)
        return view

    def _upload_in_chunks(self, eng, host):
        """Host matrix -> device matrix (16-byte aligned rows) in up to 16 chunks on a side stream.
        Returns the device view and [(row_begin, row_end, event recorded when the chunk has landed)]."""
        n, d = host.shape
        ld = (d + 3) // 4 * 4
        buf = torch.empty((n, ld), dtype=torch.float32, device=eng.device)
        x = buf[:, :d]
        nchunks = int(min(16, max(2, host.numel() * 4 // (16 << 20))))
        rows = -(-n // nchunks)
        rows = (rows + 255) // 256 * 256           # whole 256-row tiles per chunk
        copy_stream = torch.cuda.Stream(device=eng.device)
        copy_stream.wait_stream(torch.cuda.current_stream(eng.device))
        buf.record_stream(copy_stream)
        chunks = []
        with torch.cuda.stream(copy_stream):
            for r0 in range(0, n, rows):
                r1 = min(n, r0 + rows)
                x[r0:r1].copy_(host[r0:r1], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                chunks.append((r0, r1, ev))
        return x, chunks

    def _check_input_len(self, data):
        """xpysom.py:361-367"""
        data_len = len(data[0])
        if self._input_len != data_len:
            raise ValueError('Received %d features, expected %d.' % (data_len, self._input_len))

    def _check_iteration_number(self, num_iteration):
        if num_iteration < 1:
            raise ValueError('num_iteration must be > 1')

    # ------------------------------------------------------------------ training
    def _sample_stats(self, eng, x, want_scale, group, cache_key=None):
        """Per-upload statistics of a device-resident sample matrix: row scales of the fp16-split kernel and the column
        scales of the exact accumulation (one pass each, engine.prepare_samples).  Sharded runs agree on the column
        scales through one MAX all-reduce of the column maxima and one SUM of the row counts.  A caller-owned device
        tensor that has not been written since the last call keeps its statistics (torch's version counter catches
        in-place changes; the cache holds a weak reference, so a freed tensor whose address is reused cannot hit it)."""
        cached = getattr(self, '_stats_cache', None)
        gkey = id(group) if group is not None else None
        if (cache_key is not None and cached is not None and cached[0]() is cache_key and cached[1] == cache_key._version
                and cached[2] == (gkey, want_scale)):
            return cached[3]
        n, d = x.shape
        if n == 0 and group is None:
            stats = (None, eng.zeros((d + 3) // 4 * 4), eng.zeros((d + 3) // 4 * 4))
        else:
            xscale, colmax = eng.prepare_samples(x, want_scale) if n > 0 else (None, eng.zeros((d + 3) // 4 * 4))
            n_total = n
            if group is not None:
                import torch.distributed as dist
                cnt = torch.tensor([float(n)], dtype=torch.float64, device=colmax.device)
                dist.all_reduce(colmax, op=dist.ReduceOp.MAX, group=group)
                dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
                n_total = int(cnt.item())
            qscale, qinv = eng.accum_scales(colmax, d, max(n_total, 1))
            stats = (xscale, qscale, qinv)
        self._stats_cache = ((weakref.ref(cache_key), cache_key._version, (gkey, want_scale), stats)
                             if cache_key is not None else None)
        return stats

    # ---- long rows: one tensor-core pass + exact refinement (csrc/bmu_filter.cuh) -------------------------------
    _FILTER_PROBE_ROWS = 8192
    # re-scored candidates per row above which the three-pass kernel is cheaper / fraction of rows whose lists may overflow
    _FILTER_MAX_CANDIDATES = float(os.environ.get('SOM_B200_FILTER_MAXC', '24'))
    _FILTER_MAX_OVERFLOW = float(os.environ.get('SOM_B200_FILTER_MAXO', '0.005'))

    def _filter_state(self, eng, x, K, dist_kind, cache_key=None):
        """Workspaces of the filter path for the resident samples x (None when the shape is not eligible or another
        kernel was asked for): the prepared samples, and a prepared slice of them on which every epoch that is not
        sure the filter pays first PROBES the current codebook (candidates per row, overflowed lists)."""
        if (self._algo != 'auto' or getattr(eng, 'name', '') != 'cuda' or x.shape[0] == 0
                or os.environ.get('SOM_B200_FILTER', '1') == '0' or not eng.filter_eligible(x, K, dist_kind)):
            return None
        cached = getattr(self, '_filter_cache', None)
        if (cache_key is not None and cached is not None and cached[0]() is cache_key and cached[1] == cache_key._version
                and cached[2] == K):
            return cached[3]
        n = x.shape[0]
        npr = min(n, self._FILTER_PROBE_ROWS)
        try:               # ~3.8 KB of workspace per row (fp16 copy of the samples, candidate lists)
            st = {'fws': eng.filter_workspace(x, K), 'n_probe': npr, 'on': None,
                  'probe': eng.filter_workspace(x[:npr], K) if npr < n else None,
                  'probe_bmu': torch.full((npr,), -1, dtype=torch.int32, device=eng.device)}
        except torch.cuda.OutOfMemoryError:
            torch.cuda.empty_cache()
            return None    # no room for it: the three-pass kernel needs no extra memory
        self._filter_cache = (weakref.ref(cache_key), cache_key._version, K, st) if cache_key is not None else None
        return st

    def _filter_good(self, ovf, evals, rows):
        return ovf <= self._FILTER_MAX_OVERFLOW * rows and evals <= self._FILTER_MAX_CANDIDATES * max(rows - ovf, 1)

    def _filter_epoch(self, eng, st, x, w, bmu, ws, dist_kind, p, xscale, qscale, acc):
        """BMUs of all rows of x through the filter path, into ``bmu``; False when this epoch should take the three-pass
        kernel instead (the probe, or the previous epoch, found the candidate lists too long for the current map)."""
        n, d = x.shape
        K = w.shape[0]
        if not st['on'] and st['probe'] is not None:
            if st.get('skip', 0) > 0:            # the last probe was far from paying: do not even probe this epoch
                st['skip'] -= 1
                return False
            npr = st['n_probe']
            eng.bmu_filter(x[:npr], w, st['probe'], st['probe_bmu'])
            ovf, ev = eng.filter_stats(st['probe'], npr, K, d)
            st['on'] = self._filter_good(ovf, ev, npr)
            self.stats['filter_probe'] = (ovf / npr, ev / max(npr - ovf, 1))
            if not st['on']:
                if ev > 2.0 * self._FILTER_MAX_CANDIDATES * max(npr - ovf, 1) or ovf > 4 * self._FILTER_MAX_OVERFLOW * npr:
                    st['skip'] = 2
                return False
        eng.bmu_filter(x, w, st['fws'], bmu, qscale=qscale, acc=acc)      # (the refine accumulates the rows it resolves)
        ovf, ev = eng.filter_stats(st['fws'], n, K, d)
        if ovf:            # rows whose lists overflowed: the three-pass kernel and the accumulate on just those rows
            rows = torch.nonzero(bmu < 0).squeeze(1)
            xg = x.index_select(0, rows)
            xs_g = eng.prepare_samples(xg, True)[0]
            bmu_g = eng.bmu(xg, w, dist_kind, p, _lib.ALGO['tc16'], ws, xscale=xs_g)
            eng.accumulate(xg, bmu_g, K, qscale, acc)
            bmu.index_copy_(0, rows, bmu_g)
        st['on'] = self._filter_good(ovf, ev, n)
        self.stats['filter_epochs'] = self.stats.get('filter_epochs', 0) + 1
        self.stats['filter_last'] = (ovf / n, ev / max(n - ovf, 1))
        return True

    def _device_budget(self, eng, nbytes=None):
        """Bytes of samples this process may keep resident (the rest of the job streams through two block buffers).
        With ``nbytes`` given the answer only has to be right about ``nbytes <= budget``: a matrix below 1/8 of what
        torch has not reserved of the device's memory is accepted without asking the driver (cudaMemGetInfo takes a
        driver-wide lock: 0.1 - 1 ms per train() call when something else is polling the GPU)."""
        if self._max_resident_bytes is not None:
            return int(self._max_resident_bytes)
        if getattr(eng, 'name', '') != 'cuda':
            return 1 << 62
        if nbytes is not None:
            total = torch.cuda.get_device_properties(eng.device).total_memory
            if nbytes * 8 <= total - torch.cuda.memory_reserved(eng.device):
                return int(nbytes)
        free, _ = torch.cuda.mem_get_info(eng.device)
        return int(free * 0.8)

    def train(self, data, num_epochs, iter_beg=0, iter_end=None, verbose=False):
        """Batch-SOM training, epochs [iter_beg, iter_end) of a num_epochs schedule
        (xpysom.py:458-594).  With ``process_group`` set, ``data`` is this
        process's shard of the samples.

        Where the samples live decides how an epoch reads them:
          * device tensors, and host arrays that fit, are RESIDENT: uploaded once, one fused BMU + accumulate kernel per
            epoch.  Host arrays of >= 32 MB go up in chunks on a copy stream and the first epoch consumes each chunk as
            it lands;
          * host arrays larger than the device budget are STREAMED every epoch through two block buffers (pinned
            registration, copy of block i+1 under the compute of block i) -- what the reference's chunk loop
            (xpysom.py:560-569) and Dask blocks (:546) do for data of any size."""
        if iter_end is None:
            iter_end = num_epochs
        eng = self._get_engine()
        cuda = getattr(eng, 'name', '') == 'cuda'
        gx, gy, d = self._shape()
        K = gx * gy
        dist_kind, p = self._dist_kind()
        algo = _lib.ALGO[self._algo]
        topo = _lib.TOPO[self.topology]
        neigh = _lib.NEIGH[self.neighborhood_func_name]
        group = self._group()
        want_scale = self._wants_xscale(dist_kind)

        n_ep = iter_end - iter_beg
        prof = self._profile_events if getattr(self, '_profile', False) else None
        graphed = self._use_cuda_graph and prof is None and not verbose and n_ep >= 3 and cuda
        host = _as_f32_matrix(data)
        if host.shape[1] != d:
            raise ValueError('Received %d features, expected %d.' % (host.shape[1], d))
        n = host.shape[0]
        on_host = host.device.type == 'cpu' and cuda
        mode = 'resident'
        if on_host and n_ep >= 1 and n > 0:
            if host.numel() * 4 > self._device_budget(eng, host.numel() * 4):
                mode = 'stream'
            elif not graphed and host.numel() * 4 >= (32 << 20):
                mode = 'chunked'
        uploaded = self._upload_in_chunks(eng, host) if mode == 'chunked' else None     # the copies start first

        w = self._weights_to_device(eng)
        acc = eng.accumulator(K, d)          # exact [S | counts], 64-bit fixed point (cleared by every epoch tail)
        sc = eng.empty(K * d + K)            # fp32 [S | c] the neighbourhood apply reads
        nd = eng.empty(K * d + K)            # [num | den]
        S, c = sc[:K * d], sc[K * d:]
        num, den = nd[:K * d], nd[K * d:]
        ws = eng.workspace(0, K, d)
        tables = eng.neigh_tables(gx, gy, d)

        def schedule(t):
            eta_t = self._decay_function(self._learning_rate, self._learning_rateN, t, num_epochs)
            sig_t = self._decay_function(self._sigma, self._sigmaN, t, num_epochs)   # same rule (xpysom.py:541-543)
            return sig_t, eta_t

        def all_reduce(t):
            # the one exchange step of the sharded path.  Resident epochs reduce the INTEGER accumulator: the sum is
            # exact, so every rank -- and a single GPU holding all the rows -- ends up with the same bits.
            if group is not None:
                import torch.distributed as dist
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)

        def tail(t, from_acc, qinv):
            sig, eta = schedule(t)
            eng.epoch_tail(from_acc, qinv, S, c, w, gx, gy, d, topo, neigh, sig, eta, self._std_coeff,
                           self.compact_support, dist_kind, p, num, den, tables, ws)
            if verbose:
                print('\r [ %d / %d ]' % (t + 1, num_epochs), end='')

        if mode == 'stream':
            self._train_streamed(eng, host, w, acc, S, c, ws, dist_kind, p, algo, want_scale, group, all_reduce, tail,
                                 iter_beg, iter_end, K, d)
        else:
            if mode == 'chunked':
                x, chunks = uploaded
            else:
                x, chunks = self._data_to_device(eng, host), None
            bmu = eng.empty(n, dtype=torch.int32)
            if n_ep > 0:
                eng.prepare_codebook(w, dist_kind, p, ws)
                bmu.fill_(-1)               # 'no BMU yet' (the filter path seeds its bounds with the previous epoch's)
            first = iter_beg
            if chunks is not None:
                # first epoch: every chunk is searched and accumulated as it lands, with its own column scales, and
                # folded into running fp64 sums (chunk order: deterministic); the statistics of the whole matrix fall
                # out of the same pass for the epochs after it
                sd = eng.zeros(K * d + K, dtype=torch.float64)
                xscale = eng.empty(n) if want_scale else None
                colmax = eng.zeros((d + 3) // 4 * 4)
                cur = torch.cuda.current_stream(eng.device)
                for r0, r1, landed in chunks:
                    cur.wait_event(landed)
                    xs_c, cm_c = eng.prepare_samples(x[r0:r1], want_scale, out=xscale[r0:r1] if want_scale else None)
                    qs_c, qi_c = eng.accum_scales(cm_c, d, r1 - r0)
                    eng.epoch_accumulate(x[r0:r1], w, dist_kind, p, algo, qs_c, acc, ws, bmu_out=bmu[r0:r1], xscale=xs_c)
                    eng.accum_fold(acc, qi_c, K, d, sd)
                    torch.maximum(colmax, cm_c, out=colmax)
                all_reduce(sd)
                eng.accum_finalize_f64(sd, K, d, S, c)
                tail(first, None, None)
                first += 1
                n_total = n
                if group is not None:
                    import torch.distributed as dist
                    cnt = torch.tensor([float(n)], dtype=torch.float64, device=colmax.device)
                    dist.all_reduce(colmax, op=dist.ReduceOp.MAX, group=group)
                    dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
                    n_total = int(cnt.item())
                qscale, qinv = eng.accum_scales(colmax, d, max(n_total, 1))
            else:
                xscale, qscale, qinv = self._sample_stats(eng, x, want_scale, group, cache_key=x if x is data else None)

            if graphed:
                # One CUDA graph of the whole epoch, replayed n_ep - 1 times: sigma / eta live in a device-side
                # schedule indexed by a device-side epoch counter, so the captured launches never change.
                sched = eng.to_device(torch.tensor([[float(v) for v in schedule(t)] for t in range(iter_beg, iter_end)],
                                                   dtype=torch.float64))
                epoch_idx = eng.zeros(1, dtype=torch.int32)

                def epoch_body():
                    eng.prepare_codebook(w, dist_kind, p, ws)
                    if n > 0:
                        eng.epoch_accumulate(x, w, dist_kind, p, algo, qscale, acc, ws, bmu_out=bmu, xscale=xscale)
                    all_reduce(eng.accum_one_copy(acc, K, d))
                    eng.accum_finalize(acc, qinv, K, d, S, c)
                    eng.neigh_apply_sched(S, c, gx, gy, d, topo, neigh, sched, epoch_idx, self._std_coeff,
                                          self.compact_support, num, den, tables)
                    eng.merge(w, num, den)
                    eng.epoch_advance(epoch_idx)

                launches0 = eng.launches
                epoch_body()                               # first epoch eagerly: warms up every kernel / NCCL
                per_epoch = eng.launches - launches0
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph):
                    epoch_body()                           # captured, not executed
                for _ in range(n_ep - 1):
                    graph.replay()
                eng.launches = launches0 + per_epoch * n_ep    # kernels actually executed (capture launches none)
            else:
                # epoch = BMU search + exact per-BMU sums (one fused kernel) -> sum of the shards' integer accumulators
                # -> everything else (eng.epoch_tail: finalize, apply, merge, preparation of the new codebook).
                # The sum over the shards: on one node and small maps the accumulators live in NVLink peer memory and
                # the tail's finalize phase reads all of them itself (peer.py); otherwise NCCL all-reduces the integers.
                pacc = (self._peer_accumulator(eng, group, acc.numel() // max(1, eng.lib.som_b200_accum_replicas(K, d)))
                        if (group is not None and first < iter_end and cuda) else None)
                flt = self._filter_state(eng, x, K, dist_kind, cache_key=x if x is data else None) if first < iter_end else None
                # Peer accumulators exist in one copy (every replica would be read by every rank), so samples that share
                # a BMU serialise on its L2 lines again.  When the previous train() call saw hot BMUs (one neuron with more
                # than 8x the mean count) the epoch accumulates into the local replicated accumulator instead and one
                # small kernel folds it into the peer accumulator before the tail (+4 us per epoch, -35 % on blob data).
                via_local = (pacc is not None and getattr(self, '_hot_bmus', False) and cuda
                             and eng.lib.som_b200_accum_replicas(K, d) > 1)
                for t in range(first, iter_end):
                    a = pacc.current() if pacc is not None else acc
                    a_peer = a
                    if via_local:
                        a = acc
                    if prof is not None:
                        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
                        ev[0].record()
                    if n > 0 and flt is not None and self._filter_epoch(eng, flt, x, w, bmu, ws, dist_kind, p, xscale, qscale, a):
                        pass                # BMUs from the one-pass filter + refinement, accumulated by the refine kernel
                    elif n > 0:             # a rank may hold an EMPTY shard: it still joins the exchange and the tail
                        eng.epoch_accumulate(x, w, dist_kind, p, algo, qscale, a, ws, bmu_out=bmu, xscale=xscale)
                    if via_local:
                        eng.accum_fold_into(acc, a_peer, K, d)
                        a = a_peer
                    if prof is not None:
                        ev[1].record()
                        prof.append(ev)
                    if pacc is None and group is not None:
                        all_reduce(eng.accum_one_copy(acc, K, d))
                    tail(t, a, qinv)
                if pacc is not None:
                    pacc.fence()            # nobody is still reading this rank's accumulators when train() returns
                self._bmu_last = bmu if (first < iter_end and n > 0) else None     # BMUs of the last epoch (device)
                if pacc is not None and first < iter_end:      # c: the last epoch's counts over ALL shards (same on every rank)
                    cmax, csum = float(c.max().item()), float(c.sum().item())
                    self._hot_bmus = cmax > 8.0 * max(csum, 1.0) / K

        self._weights = w.cpu().numpy().reshape(gx, gy, d)      # synchronises; fp32 like xpysom.py:580-583
        if verbose:
            print('\n quantization error:', self.quantization_error(data))
        return self

    def _train_streamed(self, eng, host, w, acc, S, c, ws, dist_kind, p, algo, want_scale, group, all_reduce, tail,
                        iter_beg, iter_end, K, d):
        """Out-of-core epochs: the samples stay in (page-locked) host memory and every epoch streams them through two
        device block buffers; block i+1 is copied while block i is searched and accumulated.  Each block has its own
        column scales (computed when it is first seen, then kept: they are a few hundred bytes) and is folded into
        running fp64 sums in block order, so the result does not depend on timing."""
        n = host.shape[0]
        ld = (d + 3) // 4 * 4
        budget = max(self._device_budget(eng), 4 * 256 * ld * 4)
        rows = max(256, (budget // (2 * ld * 4)) // 256 * 256)
        rows = min(rows, (n + 255) // 256 * 256)
        nblocks = -(-n // rows)
        host = host.contiguous()
        registered = False
        if not host.is_pinned():
            rc = torch.cuda.cudart().cudaHostRegister(host.data_ptr(), host.numel() * 4, 0)
            registered = int(rc) == 0
        bufs = [torch.empty((rows, ld), dtype=torch.float32, device=eng.device) for _ in range(min(2, nblocks))]
        if ld != d:
            for b in bufs:
                b.zero_()
        copy_stream = torch.cuda.Stream(device=eng.device)
        cur = torch.cuda.current_stream(eng.device)
        sd = eng.zeros(K * d + K, dtype=torch.float64)
        xscale_all = eng.empty(n) if want_scale else None
        bmu = eng.empty(rows, dtype=torch.int32)
        block_scales = [None] * nblocks
        freed = [None, None]                       # event: the compute that last read buffer j has finished
        self.stats['streamed_blocks'] = nblocks
        self.stats['streamed_block_rows'] = rows
        try:
            eng.prepare_codebook(w, dist_kind, p, ws)
            for t in range(iter_beg, iter_end):
                landed = [None] * nblocks

                def issue(i):
                    j = i % len(bufs)
                    r0, r1 = i * rows, min(n, (i + 1) * rows)
                    with torch.cuda.stream(copy_stream):
                        if freed[j] is not None:
                            copy_stream.wait_event(freed[j])
                        bufs[j][:r1 - r0, :d].copy_(host[r0:r1], non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(copy_stream)
                    landed[i] = ev

                copy_stream.wait_stream(cur)
                issue(0)
                for i in range(nblocks):
                    if i + 1 < nblocks:
                        issue(i + 1)
                    j = i % len(bufs)
                    r0, r1 = i * rows, min(n, (i + 1) * rows)
                    xb = bufs[j][:r1 - r0, :d]
                    cur.wait_event(landed[i])
                    if block_scales[i] is None:
                        xs_b, cm_b = eng.prepare_samples(xb, want_scale, out=xscale_all[r0:r1] if want_scale else None)
                        block_scales[i] = eng.accum_scales(cm_b, d, r1 - r0)
                    xs_b = xscale_all[r0:r1] if want_scale else None
                    qs_b, qi_b = block_scales[i]
                    eng.epoch_accumulate(xb, w, dist_kind, p, algo, qs_b, acc, ws, bmu_out=bmu[:r1 - r0], xscale=xs_b)
                    eng.accum_fold(acc, qi_b, K, d, sd)
                    ev = torch.cuda.Event()
                    ev.record(cur)
                    freed[j] = ev
                all_reduce(sd)
                eng.accum_finalize_f64(sd, K, d, S, c)
                tail(t, None, None)
            torch.cuda.current_stream(eng.device).synchronize()
        finally:
            if registered:
                torch.cuda.cudart().cudaHostUnregister(host.data_ptr())

    def train_batch(self, data, num_iteration, verbose=False):
        """Compatibility with MiniSom, alias for train (xpysom.py:597-599)."""
        return self.train(data, num_iteration, verbose=verbose)

    def train_random(self, data, num_iteration, verbose=False):
        """Compatibility with MiniSom (xpysom.py:602-605): batch SOM has no sample order."""
        print("WARNING: due to batch SOM algorithm, random order is not supported. Falling back to train_batch.")
        return self.train(data, num_iteration, verbose=verbose)

    # ------------------------------------------------------------------ inference
    def _bmu_flat(self, data, euclidean_only=False):
        """Device BMU search -> (flat int32 indices on device, x, w, engine)."""
        eng = self._get_engine()
        gx, gy, d = self._shape()
        dist_kind, p = self._dist_kind(euclidean_only)
        w = self._weights_to_device(eng)
        x = self._data_to_device(eng, data)
        ws = eng.workspace(0, gx * gy, d)
        eng.prepare_codebook(w, dist_kind, p, ws)
        xscale = eng.prepare_samples(x, True)[0] if (self._wants_xscale(dist_kind) and x.shape[0] > 0) else None
        bmu = eng.bmu(x, w, dist_kind, p, _lib.ALGO[self._algo], ws, xscale=xscale)
        return bmu, x, w, eng

    def _distance_matrix(self, data, mode):
        eng = self._get_engine()
        gx, gy, d = self._shape()
        dist_kind, p = self._dist_kind()
        w = self._weights_to_device(eng)
        x = self._data_to_device(eng, data)
        ws = eng.workspace(0, gx * gy, d)
        return eng.distances(x, w, dist_kind, p, mode, ws)

    def activate(self, x):
        """Activation map of x: the (n, K) matrix of activation distances (xpysom.py:323-354)."""
        # 'euclidean_no_opt' is the full squared distance (distances.py:25-31); 'euclidean' drops the row term |x|^2
        mode = 2 if self._activation_distance_name == 'euclidean_no_opt' else 0
        return self._distance_matrix(x, mode).cpu().numpy()

    def distance_from_weights(self, data, weights_gpu=None):
        """d[i, j] = Euclidean distance between data[i] and the j-th weight (xpysom.py:647-671; the second
        argument is ignored there as well)."""
        return self._distance_matrix(data, 1).cpu().numpy()

    def topographic_error(self, data):
        """Share of samples whose best and second-best matching units are not adjacent (xpysom.py:709-746);
        Euclidean distances whatever the activation distance, like the reference."""
        self._check_input_len(data)
        if np.prod(self._weights.shape) == 1:
            warn('The topographic error is not defined for a 1-by-1 map.')
            return np.nan
        eng = self._get_engine()
        gx, gy, d = self._shape()
        w = self._weights_to_device(eng)
        x = self._data_to_device(eng, data)
        # best and second-best unit per row from ONE fused kernel (the reference argsorts the (n, K) matrix)
        b2 = eng.top2(x, w, eng.workspace(0, gx * gy, d)).cpu().numpy().astype(np.int64)
        bx, by = np.unravel_index(b2, (gx, gy))
        if self.topology == 'rectangular':
            return ((np.abs(np.diff(bx)) > 1) | (np.abs(np.diff(by)) > 1)).mean().item()
        # the reference indexes its (gy, gx) meshgrids with (i, j) here (xpysom.py:742-743); kept as is
        ex, ey = self._xx[bx, by], self._yy[bx, by]
        dxdy = np.hstack([np.diff(ex), np.diff(ey)])
        return (np.linalg.norm(dxdy, axis=1) > 1.5).mean().item()

    def winner(self, x):
        """Coordinates of the winning neuron(s) (xpysom.py:370-408): a tuple for a
        1-D sample, a list of tuples for a 2-D batch."""
        single = (np.ndim(x) == 1) if not isinstance(x, torch.Tensor) else (x.dim() == 1)
        bmu, _, _, _ = self._bmu_flat(x)
        flat = bmu.cpu().numpy().astype(np.int64)
        gy = self._weights.shape[1]
        wi, wj = flat // gy, flat % gy
        if single:
            return (wi[0].item(), wj[0].item())
        return list(zip(wi, wj))

    def predict(self, data):
        """Flat index of the winner of every sample (xpysom.py:608-617)."""
        bmu, _, _, _ = self._bmu_flat(data)
        return bmu.cpu().numpy().astype(np.int64)

    def quantization(self, data):
        """Code-book vector of every sample's Euclidean BMU (xpysom.py:620-645)."""
        self._check_input_len(data)
        bmu, _, _, _ = self._bmu_flat(data, euclidean_only=True)
        flat = bmu.cpu().numpy().astype(np.int64)
        gx, gy, d = self._shape()
        return np.array(self._weights).reshape(gx * gy, d)[flat]

    def quantization_error(self, data):
        """Mean Euclidean distance between each sample and its BMU (xpysom.py:673-707)."""
        self._check_input_len(data)
        bmu, x, w, eng = self._bmu_flat(data, euclidean_only=True)
        _, err = eng.quantize(x, w, bmu, want_q=False, want_err=True)
        return (err.sum(dtype=torch.float64) / err.numel()).item()

    def distance_map(self):
        """Normalised sum of distances to the grid neighbours (xpysom.py:788-817)."""
        eng = self._get_engine()
        gx, gy, _ = self._shape()
        w = self._weights_to_device(eng)
        um = eng.distance_map(w, gx, gy, _lib.TOPO[self.topology]).cpu().numpy().astype(np.float64)
        return (um / um.max()).reshape(gx, gy)

    def activation_response(self, data):
        """How many times each neuron wins (xpysom.py:819-829)."""
        self._check_input_len(data)
        gx, gy, _ = self._shape()
        flat = self.predict(data)
        return np.bincount(flat, minlength=gx * gy).astype(float).reshape(gx, gy)

    def win_map(self, data):
        """(i,j) -> list of the samples mapped there (xpysom.py:831-840)."""
        self._check_input_len(data)
        winmap = defaultdict(list)
        for sample, win in zip(data, self.winner(data)):
            winmap[win].append(sample)
        return winmap

    def labels_map(self, data, labels):
        """(i,j) -> Counter of the labels mapped there (xpysom.py:842-865)."""
        self._check_input_len(data)
        if not len(data) == len(labels):
            raise ValueError('data and labels must have the same length.')
        winmap = defaultdict(list)
        for win, lab in zip(self.winner(data), labels):
            winmap[win].append(lab)
        return {pos: Counter(v) for pos, v in winmap.items()}

    # ------------------------------------------------------------------ host-side conveniences
    def get_weights(self):
        """xpysom.py:286-288"""
        return self._weights

    def get_euclidean_coordinates(self):
        """xpysom.py:291-305"""
        return self._xx.T, self._yy.T

    def convert_map_to_euclidean(self, xy):
        """xpysom.py:308-320"""
        return self._xx.T[xy], self._yy.T[xy]

    def random_weights_init(self, data):
        """Codebook <- random samples, same generator draws as xpysom.py:749-759."""
        self._check_input_len(data)
        gx, gy, _ = self._shape()
        for i in range(gx):
            for j in range(gy):
                self._weights[i, j] = data[self._random_generator.randint(len(data))]

    def pca_weights_init(self, data):
        """Codebook spans the first two principal components (xpysom.py:762-785)."""
        if self._input_len == 1:
            raise ValueError('The data needs at least 2 features for pca initialization')
        self._check_input_len(data)
        if len(self._neigx) == 1 or len(self._neigy) == 1:
            warn('PCA initialization inappropriate:One of the dimensions of the map is 1.')
        pc_length, pc = np.linalg.eig(np.cov(np.transpose(data)))
        order = np.argsort(-pc_length)
        for i, c1 in enumerate(np.linspace(-1, 1, len(self._neigx))):
            for j, c2 in enumerate(np.linspace(-1, 1, len(self._neigy))):
                self._weights[i, j] = c1 * pc[order[0]] + c2 * pc[order[1]]

    # ------------------------------------------------------------------ pickling (xpysom.py:868-892)
    def __getstate__(self):
        state = self.__dict__.copy()
        state['_engine'] = None            # device handles are rebuilt on demand
        state['_process_group'] = None if self._process_group in (None, False) else _UNPICKLED_SHARDED
        state.pop('_stats_cache', None)
        state.pop('_peer_cache', None)
        state.pop('_filter_cache', None)
        state.pop('_bmu_last', None)
        state.pop('_hot_bmus', None)
        state['_profile_events'] = []
        state['xp'] = None
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
