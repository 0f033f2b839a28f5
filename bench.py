#!/usr/bin/env python
"""Benchmark of the batch-SOM training epoch (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--workload c2|c3|c4|c5] [--impl reference]

A step is one training epoch over the (resident) synthetic samples of the named workload.  Under torchrun
(N > 1) every rank holds its own shard of the same size (weak scaling) and the per-BMU sums are all-reduced
once per epoch.  Rank 0 prints ONE JSON line:

  * top level: the headline workload (c2 unless --workload says otherwise): `value` device-resident, `e2e` through
    XPySom.train with pinned host samples copied in every step, `roofline` of the fused BMU kernel, `cpu_baseline`
    and a `parity` report (N = 1), `replica_check` (N > 1);
  * `workloads`: the same device-resident measurement (value, ms_per_step, kernel_ms, roofline) for the other named
    configurations (c3, c4, c5: the per-GPU shards of BASELINE.json configs[2..4]) and for the headline shape on the
    64-blob mixture (hot BMUs), at the same N, inside the same driver-timed command.

`--impl reference` times the reference's own CPU implementation of the path on the host cores: the UNMODIFIED
reference package from baseline/_ref when it is importable (kind "reference"), else the oracle port; both the plain
`use_dask=False` loop with all BLAS threads and a Dask-shaped run (one row block per worker process calling the
reference's pickled `_update`, partials summed, merge on the parent: xpysom.py:545-558 -- Dask itself is not
installed in this image).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
HOST_CORES = len(os.sched_getaffinity(0))
if "reference" in sys.argv or int(os.environ.get("WORLD_SIZE", "1")) == 1:
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU legs must see the box's cores.  Set BEFORE numpy
    # (OpenBLAS) is first imported; threadpoolctl re-asserts it at run time.
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(HOST_CORES)

import numpy as np  # noqa: E402

# SURVEY 8d synthetic workloads (BASELINE.json configs[1..4]).  `n` is the number of rows PER GPU (weak
# scaling): c2 is the metric's own single-GPU configuration; c3/c4/c5 are the 1/8 shards of the named
# multi-GPU configurations (16M, 8M, 4M rows over 8 GPUs), so that N=8 reproduces the named sizes.
WORKLOADS = {
    "c2": dict(name="synthetic 1Mx64 f32, 32x32 map, gaussian/euclidean", n=1_000_000, d=64, gx=32, gy=32, kw={}),
    "c3": dict(name="synthetic 16Mx16 f32 over 8 GPUs (2M rows per GPU), 40x40 map, linear decay", n=2_000_000, d=16,
               gx=40, gy=40, kw=dict(decay_function="linear")),
    "c4": dict(name="synthetic 8Mx784 f32 over 8 GPUs (1M rows per GPU), 100x100 map", n=1_000_000, d=784,
               gx=100, gy=100, kw={}),
    "c5": dict(name="synthetic 4Mx128 f32 over 8 GPUs (500k rows per GPU), 50x50 hexagonal, mexican_hat, cosine",
               n=500_000, d=128, gx=50, gy=50,
               kw=dict(topology="hexagonal", neighborhood_function="mexican_hat", activation_distance="cosine")),
}
TOTAL_EPOCHS = 100   # length of the decay schedule the timed epochs are taken from (raised to warmup + steps if needed)
MIN_TIMED_S = 0.25   # the K-step block is repeated until the device-timed region is at least this long


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            m = json.load(f)
        return dict(hbm_gbs=m["hbm_gbs"], bf16=m["bf16_tflops"], bf16_sustained=m.get("bf16_tflops_sustained"),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))


def synth(n, d, seed):
    """U[0,1) iid float32 -- SURVEY 8d distribution (i): throughput, worst-case near-ties."""
    rng = np.random.RandomState(seed)
    out = np.empty((n, d), dtype=np.float32)
    step = max(1, (1 << 24) // d)
    for s in range(0, n, step):
        out[s:s + step] = rng.random_sample((min(step, n - s), d)).astype(np.float32)
    return out


def synth_device(torch, n, d, seed, dev, blobs=False):
    """The same two distributions generated on the GPU (the sub-records never need the samples on the host)."""
    g = torch.Generator(device=dev).manual_seed(seed)
    if not blobs:
        return torch.rand(n, d, generator=g, device=dev)
    # SURVEY 8d distribution (ii): 64-blob Gaussian mixture, centres U[0,1)^D, sigma 0.1
    centres = torch.rand(64, d, generator=g, device=dev)
    x = torch.empty(n, d, device=dev)
    step = max(1, (1 << 26) // d)
    for s in range(0, n, step):
        m = min(step, n - s)
        lab = torch.randint(0, 64, (m,), generator=g, device=dev)
        x[s:s + m] = centres[lab] + 0.1 * torch.randn(m, d, generator=g, device=dev)
    return x


# ------------------------------------------------------------------------ CPU arms
def _blas_threads(n):
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=n)
    except Exception:
        pass


def _reference_module():
    """The unmodified reference package, if it travelled with the repo (baseline/_ref, installed by
    __graft_entry__.build()) or is named by $SOM_REFERENCE; None otherwise (the oracle port is timed instead)."""
    for cand in (os.environ.get("SOM_REFERENCE"), os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "xpysom_dask")):
            if cand not in sys.path:
                sys.path.insert(0, cand)
            try:
                import contextlib
                import io
                with contextlib.redirect_stdout(io.StringIO()):      # the package prints at import when CuPy is missing
                    import xpysom_dask
                return xpysom_dask
            except Exception:
                continue
    return None


class CpuSom:
    """One CPU implementation of the path behind a tiny common interface: the reference itself or the oracle port."""

    def __init__(self, wl, cores):
        self.wl, self.cores = wl, cores
        self.ref = _reference_module()
        self.kind = "reference" if self.ref is not None else "port"
        kw = dict(wl["kw"])
        if self.ref is not None:
            self.som = self.ref.XPySom(wl["gx"], wl["gy"], wl["d"], random_seed=0, xp=np, n_parallel=cores * 500, **kw)
        else:
            from oracle import som_oracle as so
            self.so = so
            self.spec = so.SomSpec(gx=wl["gx"], gy=wl["gy"], dim=wl["d"], random_seed=0, n_parallel=cores * 500, **kw)
            self.w = np.asarray(so.init_weights(self.spec), dtype=np.float32)

    def epoch_plain(self, x, t, T):
        """use_dask=False path: the chunk loop of xpysom.py:560-569 with all BLAS threads."""
        if self.ref is not None:
            self.som.train(x, T, iter_beg=t, iter_end=t + 1)
        else:
            self.w = self.so.epoch(self.spec, x, self.w, t, T)

    # -- the Dask-shaped run -------------------------------------------------------------------------
    def schedule(self, t, T):
        if self.ref is not None:
            s = self.som
            return (s._decay_function(s._learning_rate, s._learning_rateN, t, T),
                    s._decay_function(s._sigma, s._sigmaN, t, T))
        sp = self.spec
        return (self.so.decay_value(sp.decay_function, sp.learning_rate, sp.learning_rateN, t, T),
                self.so.decay_value(sp.decay_function, sp.sigma, sp.sigmaN, t, T))


_POOL_STATE = {}


def _pool_init(blocks_x):
    _blas_threads(1)                      # one BLAS thread per worker, as LocalCluster(threads_per_worker=1)
    _POOL_STATE["x"] = blocks_x           # fork: the parent's sample blocks are shared, not copied or re-sent


def _pool_update(args):
    """One task of the graph of xpysom.py:545-558: `_update` of the pickled model on one row block."""
    impl, blk, weights, eta, sig = args
    x = _POOL_STATE["x"][blk]
    if hasattr(impl, "_update"):          # the reference object, unpickled in this worker (xpysom.py:868-892)
        if impl._activation_distance.can_cache:
            impl._sq_weights_gpu = np.power(weights.reshape(-1, weights.shape[2]), 2).sum(axis=1, keepdims=True)
        else:
            impl._sq_weights_gpu = None
        return impl._update(x, weights, eta, sig)
    so, spec = impl
    return so.update_block(spec, x, weights, eta, sig)


def dask_shaped_rate(cpu, x, epochs, warm, T):
    """`multiprocessing.Pool(cores)`, one row block per task through the reference's own pickled `_update`,
    partials reduced with Python's sum, `_merge_updates` on the parent -- the graph of xpysom.py:545-558."""
    import multiprocessing as mp
    cores = cpu.cores
    rows = -(-len(x) // cores)
    blocks = [x[s:s + rows] for s in range(0, len(x), rows)]
    ctx = mp.get_context("fork")
    if cpu.ref is not None:
        impl = cpu.som
        w = np.asarray(impl._weights, dtype=np.float32)
        merge = impl._merge_updates
    else:
        impl = (cpu.so, cpu.spec)
        w = cpu.w
        merge = lambda ww, num, den: cpu.so.merge(ww, num, den)   # noqa: E731
    with ctx.Pool(cores, initializer=_pool_init, initargs=(blocks,)) as pool:
        def one(t, w):
            eta, sig = cpu.schedule(t, T)
            parts = pool.map(_pool_update, [(impl, b, w, eta, sig) for b in range(len(blocks))])
            num = sum(p[0] for p in parts)
            den = sum(p[1] for p in parts)
            return np.asarray(merge(w, num, den), dtype=np.float32)
        for t in range(warm):
            w = one(t, w)
        t0 = time.perf_counter()
        for t in range(warm, warm + epochs):
            w = one(t, w)
        dt = time.perf_counter() - t0
    return len(x) * epochs / dt, dt


def cpu_rates(wl, rows, epochs, warm, T, dask_shaped=True):
    """Both CPU arms on the first `rows` rows of the workload.  Returns a dict for the JSON line."""
    _blas_threads(HOST_CORES)
    x = synth(rows, wl["d"], 0)
    cpu = CpuSom(wl, HOST_CORES)
    for t in range(warm):
        cpu.epoch_plain(x, t, T)
    t0 = time.perf_counter()
    for t in range(warm, warm + epochs):
        cpu.epoch_plain(x, t, T)
    dt = time.perf_counter() - t0
    arms = {"plain_all_blas_threads": {"value": rows * epochs / dt, "seconds": dt, "blas_threads": HOST_CORES}}
    if dask_shaped:
        try:
            rate, ddt = dask_shaped_rate(CpuSom(wl, HOST_CORES), x, epochs, warm, T)
            arms["dask_shaped_process_pool"] = {"value": rate, "seconds": ddt, "workers": HOST_CORES,
                                                "blas_threads_per_worker": 1}
        except Exception as e:                       # never lose the whole line to the emulation
            arms["dask_shaped_process_pool"] = {"error": repr(e)[:200]}
    best = max(a["value"] for a in arms.values() if "value" in a)
    return dict(value=best, unit="samples*epochs/s", cores=HOST_CORES, kind=cpu.kind, arms=arms,
                sample="%d epochs (after %d warm-up) over the first %d rows; %s, numpy %s; value = the faster of the "
                       "plain use_dask=False loop (all BLAS threads) and the Dask-shaped run (one row block per worker "
                       "process, reference _update pickled to the workers, partials summed, merge on the parent; Dask "
                       "itself is not installed)" % (epochs, warm, rows,
                                                     "UNMODIFIED reference package from baseline/_ref" if cpu.kind == "reference"
                                                     else "oracle port of the reference (oracle/som_oracle.py)", np.__version__))


def run_reference(args, wl, rank):
    if rank != 0:
        return
    rows = min(wl["n"], max(20_000, int(2.0e10 / (wl["gx"] * wl["gy"] * wl["d"]))))   # ~2 s per epoch and arm on 16 cores
    T = TOTAL_EPOCHS
    t0 = time.perf_counter()
    cb = cpu_rates(wl, rows, args.steps, args.warmup, T)
    wall = time.perf_counter() - t0
    rate = cb["value"]
    line = {
        "impl": "reference", "metric": "SOM training samples*epochs/sec", "value": rate, "unit": "samples*epochs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * rows / rate,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "rows_per_step": rows, "map": "%dx%d" % (wl["gx"], wl["gy"]),
                   "features": wl["d"], "note": "bounded sample of the workload: the rate is linear in the rows"},
        "cpu_baseline": cb,
        "e2e": {"value": rate, "unit": "samples*epochs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": wall,
    }
    print(json.dumps(line), flush=True)


def measure_tf32_peak(torch, dev):
    """Dense TF32 tensor throughput of this GPU (TFLOP/s): the roofline denominator of the TF32 kernel (SURVEY 8d asks
    for it to be measured on the box; MEASURED_PEAKS.json has bf16 only).  cuBLAS, 8192^3, best of 5 after warm-up."""
    try:
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        m = 8192
        a = torch.randn(m, m, device=dev)
        b = torch.randn(m, m, device=dev)
        for _ in range(3):
            torch.matmul(a, b)
        best = 0.0
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            torch.matmul(a, b)
            e1.record()
            torch.cuda.synchronize()
            best = max(best, 2.0 * m ** 3 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
        torch.backends.cuda.matmul.allow_tf32 = old
        del a, b
        return best
    except Exception:
        return None


# ------------------------------------------------------------------------ GPU arm
class GpuBench:
    def __init__(self, args, rank, world, local_rank):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.args, self.rank, self.world = args, rank, world
        torch.cuda.set_device(local_rank)
        self.dev = torch.device("cuda", local_rank)
        self.local_rank = local_rank
        if world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.pk = peaks()
        self._tf32 = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def maxr(self, vals):
        t = self.torch.tensor(vals, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t.tolist()

    def make_som(self, wl):
        from xpysom_dask_b200 import XPySom
        return XPySom(wl["gx"], wl["gy"], wl["d"], random_seed=0, algo=self.args.algo, device=self.dev,
                      use_cuda_graph=self.args.cuda_graph, process_group=True if self.world > 1 else None, **wl["kw"])

    def timed_blocks(self, som, x_dev, steps, warmup, min_s):
        """`steps` epochs in ONE train() call = one block, CUDA events around it, barrier + synchronize on both sides;
        the block is repeated until the summed device time reaches min_s.  Returns (median block ms, all block ms,
        kernels launched per block)."""
        torch = self.torch
        eng = som._get_engine()
        som.train(x_dev, TOTAL_EPOCHS, iter_beg=0, iter_end=warmup)
        blocks, launches = [], 0
        while True:
            self.barrier()
            l0 = eng.launches
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            som.train(x_dev, TOTAL_EPOCHS, iter_beg=warmup, iter_end=warmup + steps)
            e1.record()
            self.barrier()
            launches = eng.launches - l0
            (ms,) = self.maxr([e0.elapsed_time(e1)])
            blocks.append(ms)
            if sum(blocks) >= 1e3 * min_s or len(blocks) >= 40:
                break
        return float(np.median(blocks)), blocks, launches

    def kernel_pass(self, som, x_dev, steps, warmup):
        """The dominant kernel bracketed by CUDA events on its own stream: the same epochs again, one by one."""
        torch = self.torch
        som._profile = True
        som._profile_events = []
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        p0.record()
        som.train(x_dev, TOTAL_EPOCHS, iter_beg=warmup, iter_end=warmup + steps)
        p1.record()
        self.barrier()
        eager_ms = p0.elapsed_time(p1)
        bmu_ms = float(np.mean([ev[0].elapsed_time(ev[1]) for ev in som._profile_events]))
        som._profile = False
        bmu_ms, eager_ms = self.maxr([bmu_ms, eager_ms])
        return bmu_ms, eager_ms

    def roofline(self, wl, n, bmu_ms, step_ms, eager_step_ms, wl_key):
        torch = self.torch
        from xpysom_dask_b200 import _lib
        d, K = wl["d"], wl["gx"] * wl["gy"]
        dist_name = wl["kw"].get("activation_distance", "euclidean")
        use = _lib.load().som_b200_pick_algo(_lib.ALGO[self.args.algo], _lib.DIST[dist_name], d, 1, 1)
        pk = self.pk
        flops = 2.0 * n * K * d                       # algorithmic flops of one BMU launch
        if use == _lib.ALGO["simt"]:
            kernel, peak, peak_note = "bmu_simt_kernel + accumulate_kernel", None, "SIMT fp32"
        elif use == _lib.ALGO["tc16"]:
            kernel = "bmu_tc3_kernel (tcgen05 kind::f16, 3-term fp16 split, cta_group::2, argmin + per-BMU accumulate fused)"
            peak, peak_note = pk["bf16"], "dense bf16/fp16 tensor rate bf16_tflops from %s" % pk["source"]
        else:
            kernel = "bmu_tc2_kernel (tcgen05 kind::tf32, 3-term TF32 split, cta_group::2, argmin + per-BMU accumulate fused)"
            if self._tf32 is None:
                self._tf32 = measure_tf32_peak(torch, self.dev) or 0.0
            if self._tf32:
                peak, peak_note = self._tf32, ("dense TF32 rate measured here with a cuBLAS 8192^3 TF32 GEMM (MEASURED_PEAKS.json "
                                               "holds bf16 only: %.0f TFLOP/s)" % pk["bf16"])
            else:
                peak, peak_note = pk["bf16"] / 2.0, "dense TF32 rate taken as half of bf16_tflops from %s" % pk["source"]
        ach = flops / (bmu_ms * 1e-3) / 1e12
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                tj = json.load(f)
            traffic = tj.get(wl_key)
            traffic_src = (tj.get("source") or {}).get(wl_key)
        return {
            "kernel": kernel, "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
            "frac": (ach / peak) if peak else None, "traffic": traffic,
            "traffic_source": ("ncu --set full capture of the same kernel and shape, not measured in this run: %s" % traffic_src)
                              if traffic_src else None,
            "note": "achieved = 2*n*K*D algorithmic flops / CUDA-event time of the fused BMU kernel, averaged over the same "
                    "epochs launched one by one right after the timed region; peak = %s; the kernel executes 3x the algorithmic "
                    "flops (hi/lo split for fp32 accuracy), so its attainable ceiling is frac 0.333" % peak_note,
            "frac_of_3pass_ceiling": (ach / (peak / 3.0)) if peak else None,
            "kernel_ms": bmu_ms, "step_ms": step_ms, "step_ms_epochs_launched_one_by_one": eager_step_ms,
            "hbm": {"achieved": 4.0 * n * d / (step_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "note": "sample-read bytes 4*D per sample-epoch / step time"},
        }

    def sub_record(self, key, wl, blobs, steps, warmup, min_s):
        """Device-resident measurement of one more workload at this N (samples generated on the GPU)."""
        torch = self.torch
        n = self.args.rows or wl["n"]
        x_dev = synth_device(torch, n, wl["d"], 100 + self.rank, self.dev, blobs=blobs)
        som = self.make_som(wl)
        ms, blocks, launches = self.timed_blocks(som, x_dev, steps, warmup, min_s)
        bmu_ms, eager_ms = self.kernel_pass(som, x_dev, steps, warmup)
        rec = None
        if self.rank == 0:
            step_ms = ms / steps
            rec = {"workload": wl["name"] + (" -- 64-blob Gaussian mixture (hot BMUs)" if blobs else "")
                               + (" -- epochs %d..%d of the %d-epoch schedule" % (warmup, warmup + steps - 1, TOTAL_EPOCHS)
                                  if warmup > 3 else ""),
                   "value": n * self.world * steps / (ms * 1e-3), "unit": "samples*epochs/s", "n_gpus": self.world,
                   "rows_per_gpu": n, "steps": steps, "warmup": warmup, "ms_per_step": step_ms,
                   "block_ms": blocks, "gpu_launches": launches,
                   "roofline": self.roofline(wl, n, bmu_ms, step_ms, eager_ms / steps, key)}
            self.note_filter(som, rec)
        del som, x_dev
        torch.cuda.empty_cache()
        return rec

    @staticmethod
    def note_filter(som, rec):
        """Long rows (D >= 256) may take the one-pass filter + refinement path (csrc/bmu_filter.cuh) when a probe of the
        current codebook says it pays: say how many of the epochs run so far did, and what the kernel time covers then."""
        st = getattr(som, "stats", {})
        if "filter_probe" not in st and "filter_last" not in st:
            return
        rec["filter_path"] = {"epochs_on_it_so_far": st.get("filter_epochs", 0),
                              "last_full_run_overflow_frac_and_candidates_per_row": st.get("filter_last"),
                              "last_probe_overflow_frac_and_candidates_per_row": st.get("filter_probe"),
                              "policy": "used when <= 24 candidates per row survive the one-pass bounds and <= 0.5 % of the "
                                        "rows overflow their lists (probe on 8192 rows, or the previous epoch)"}
        if st.get("filter_epochs", 0):
            r = rec["roofline"]
            r["kernel"] = ("bmu_filter_kernel (tcgen05 kind::f16, ONE pass on centred fp16 operands, interval epilogue) + "
                           "bmu_refine_kernel + accumulate_kernel on the epochs that took the filter path; " + r["kernel"] +
                           " on the others")
            r["note"] += ("; kernel_ms here spans seed + filter + refine + the host's read of the statistics + exact "
                          "accumulate on filter epochs, whose ceiling is frac 1.0 (one tensor pass)")

    def replica_check(self, som, wl, x_dev):
        """N > 1 (row G of SURVEY 8a, reference semantics xpysom.py:546-558: any chunking gives the same sum):
        (a) the codebook every rank holds after the timed training is bit-identical; (b) one sharded epoch over
        rows gathered from all ranks equals the same epoch run by ONE GPU on the gathered rows."""
        torch, dist = self.torch, self.dist
        from xpysom_dask_b200 import XPySom
        w = torch.from_numpy(np.ascontiguousarray(som._weights, dtype=np.float32)).to(self.dev)
        lo, hi = w.clone(), w.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        identical = bool(torch.equal(lo, hi))
        per = 100_000 // self.world
        mine = x_dev[:per].contiguous()
        allrows = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(allrows, mine)
        gathered = torch.cat(allrows)
        w0 = lo.cpu().numpy().reshape(som._weights.shape)
        sharded = self.make_som(wl)
        sharded._weights = w0.copy()
        sharded.train(mine, TOTAL_EPOCHS, iter_beg=5, iter_end=6)
        single = XPySom(wl["gx"], wl["gy"], wl["d"], random_seed=0, algo=self.args.algo, device=self.dev, **wl["kw"])
        single._weights = w0.copy()
        single.train(gathered, TOTAL_EPOCHS, iter_beg=5, iter_end=6)
        a, b = np.asarray(sharded._weights, np.float64), np.asarray(single._weights, np.float64)
        rel = float(np.abs(a - b).max() / np.abs(b).max())
        rel_t = torch.tensor([rel], dtype=torch.float64, device=self.dev)
        dist.all_reduce(rel_t, op=dist.ReduceOp.MAX)
        return {"ranks_bit_identical_after_timed_training": identical,
                "sharded_vs_single_gpu_epoch_rel_err": rel_t.item(), "bitwise_equal_to_single_gpu": rel_t.item() == 0.0,
                "rows": per * self.world,
                "note": "one epoch from the same W_t: %d ranks x %d rows with the per-epoch all-reduce vs ONE GPU on the "
                        "gathered rows" % (self.world, per)}

    def parity_report(self, wl, x_host, rows):
        """N = 1: the cpu_baseline leg's oracle epoch doubles as the checker of the SAME epoch on the GPU (teacher-forced
        from the initial codebook): BMU agreement under the stated near-tie rule and codebook error."""
        from oracle import som_oracle as so
        from xpysom_dask_b200 import XPySom
        spec = so.SomSpec(gx=wl["gx"], gy=wl["gy"], dim=wl["d"], random_seed=0, n_parallel=HOST_CORES * 500, **wl["kw"])
        x = x_host[:rows].numpy()
        w0 = np.asarray(so.init_weights(spec), dtype=np.float32)
        t0 = time.perf_counter()
        w_ref, bmu_ref = so.epoch(spec, x, w0, 0, TOTAL_EPOCHS, return_bmu=True)
        dt = time.perf_counter() - t0
        som = XPySom(wl["gx"], wl["gy"], wl["d"], random_seed=0, algo=self.args.algo, device=self.dev, **wl["kw"])
        bmu = som.predict(x)
        som.train(x, TOTAL_EPOCHS, iter_beg=0, iter_end=1)
        _, _, gap, scale = so.top2_gap(spec, x, w0)
        clear = gap > 1e-6 * scale
        mism = bmu != bmu_ref
        S, c = so.sums_by_bmu(bmu, x, spec.K)
        sig = float(so.decay_value(spec.decay_function, spec.sigma, spec.sigmaN, 0, TOTAL_EPOCHS))
        H = so.neighborhood_table(spec, sig).astype(np.float64)
        num, den = H.T @ S, H.T @ c
        w_same = np.where(den[:, None] != 0, num / den[:, None], w0.reshape(spec.K, -1)).reshape(w0.shape)
        rel = lambda a, b: float(np.abs(np.asarray(a, np.float64) - b).max() / np.abs(b).max())   # noqa: E731
        return {"rows": rows, "epsilon": 1e-6, "near_tie_rate": float((~clear).mean()), "bmu_mismatch_rate": float(mism.mean()),
                "bmu_mismatches_outside_near_tie_band": int((mism & clear).sum()),
                "codebook_rel_err_vs_oracle_epoch": rel(som._weights, np.asarray(w_ref, np.float64)),
                "codebook_rel_err_vs_oracle_update_from_gpu_bmus": rel(som._weights, w_same),
                "oracle": "oracle/som_oracle.py epoch (teacher-forced from the initial codebook, t = 0)"}, rows / dt, dt

    def run(self, wl_key):
        torch = self.torch
        args, rank, world = self.args, self.rank, self.world
        wl = WORKLOADS[wl_key]
        n, d, gx, gy = args.rows or wl["n"], wl["d"], wl["gx"], wl["gy"]
        K = gx * gy
        x_host = torch.from_numpy(synth(n, d, seed=rank)).pin_memory()
        x_dev = x_host.to(self.dev)
        som = self.make_som(wl)
        sampler = ClockSampler(self.local_rank)
        if rank == 0:
            sampler.start()       # BEFORE the first barrier: spawning nvidia-smi takes ~1 ms
        ms, blocks, launches = self.timed_blocks(som, x_dev, args.steps, args.warmup, MIN_TIMED_S)
        bmu_ms, eager_ms = self.kernel_pass(som, x_dev, args.steps, args.warmup)

        # ---- end to end: host (pinned) samples in, codebook out, every step -------------------
        som.train(x_host, TOTAL_EPOCHS, iter_beg=0, iter_end=1)
        self.barrier()
        t0 = time.perf_counter()
        for s in range(args.steps):
            som.train(x_host, TOTAL_EPOCHS, iter_beg=args.warmup + s, iter_end=args.warmup + s + 1)
        self.barrier()
        (e2e_ms,) = self.maxr([1e3 * (time.perf_counter() - t0)])
        clocks = sampler.stop() if rank == 0 else None

        replica = self.replica_check(som, wl, x_dev) if world > 1 else None
        self._top_som_stats = type("S", (), {"stats": dict(som.stats)})()
        del som
        subs = {}
        if not args.no_extra:
            others = [k for k in ("c3", "c4", "c5") if k != wl_key]
            for k in others:
                w2 = WORKLOADS[k]
                heavy = w2["gx"] * w2["gy"] * w2["d"] > 4_000_000
                subs[k] = self.sub_record(k, w2, False, 10 if heavy else min(args.steps, 20), 3, 0.0 if heavy else 0.1)
            subs[wl_key + "_blobs"] = self.sub_record(wl_key, wl, True, min(args.steps, 20), 3, 0.1)
            if wl_key != "c4":     # the north-star shape later in the schedule, where the map has differentiated
                subs["c4_late"] = self.sub_record("c4", WORKLOADS["c4"], False, 10, 40, 0.0)

        if rank != 0:
            return
        total = n * world
        step_ms = ms / args.steps
        roof = self.roofline(wl, n, bmu_ms, step_ms, eager_ms / args.steps, wl_key)
        f16 = "kind::f16" in roof["kernel"]
        line = {
            "metric": "SOM training samples*epochs/sec", "value": total * args.steps / (ms * 1e-3), "unit": "samples*epochs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": step_ms,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": ("f32 (tensor-core contraction on a 3-term %s split)" % ("fp16" if f16 else "tf32"))
                     if roof["peak"] else "f32",
            "data": "synthetic",
            "config": {"workload": wl["name"], "rows_per_gpu": n, "map": "%dx%d" % (gx, gy), "features": d,
                       "algo": args.algo, "parallelism": "dp%d" % world,
                       "l2": "inputs larger than L2 (%.0f MB of samples per GPU per epoch)" % (4e-6 * n * d),
                       "timing": "the %d-step block (one train() call) was timed %d times back to back (CUDA events, barrier + "
                                 "synchronize on both sides, max over ranks): ms_per_step is the median block / %d; device-timed "
                                 "region %.3f s" % (args.steps, len(blocks), args.steps, sum(blocks) * 1e-3)},
            "block_ms": blocks,
            "clocks": clocks,
            "e2e": {"value": total * args.steps / (e2e_ms * 1e-3), "unit": "samples*epochs/s",
                    "h2d_bytes_per_step": 4 * n * d + 4 * K * d, "d2h_bytes_per_step": 4 * K * d,
                    "note": "XPySom.train(pinned host samples, one epoch per call): H2D of the samples and codebook, "
                            "D2H of the codebook inside the timed region"},
            "gpu_launches": launches,
            "roofline": roof,
            "workloads": subs,
        }
        self.note_filter(self._top_som_stats, line) if getattr(self, "_top_som_stats", None) is not None else None
        if replica is not None:
            line["replica_check"] = replica
        if world == 1 and not args.no_cpu:
            rows = min(n, max(20_000, int(6.0e10 / (K * d))))
            try:
                par, rate1, dt1 = self.parity_report(wl, x_host, rows)
                line["parity"] = par
            except Exception as e:
                line["parity"] = {"error": repr(e)[:300]}
            cb = cpu_rates(wl, rows, 1, 0, TOTAL_EPOCHS)
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--algo", default="auto", choices=["auto", "tc16", "tc", "simt"])
    ap.add_argument("--rows", type=int, default=0, help="override rows per GPU (debug)")
    ap.add_argument("--cuda-graph", action="store_true", help="replay one captured CUDA graph per epoch")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline / parity leg")
    ap.add_argument("--no-extra", action="store_true", help="skip the `workloads` sub-records")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    global TOTAL_EPOCHS
    TOTAL_EPOCHS = max(TOTAL_EPOCHS, max(args.warmup, 3) + args.steps + 1)     # the timed epochs stay inside the schedule
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl, rank)
        return
    if args.warmup < 3:
        args.warmup = 3                    # timing rule: W >= 3
    gb = GpuBench(args, rank, world, local_rank)
    gb.run(args.workload)
    if world > 1:
        gb.dist.destroy_process_group()


if __name__ == "__main__":
    main()
