#!/usr/bin/env python
"""Benchmark of the batch-SOM training epoch (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--workload c2|c3|c4|c5] [--impl reference]

A step is one training epoch over the (resident) synthetic samples of the named
workload.  Under torchrun (N > 1) every rank holds its own shard of the same
size (weak scaling) and the per-BMU sums are all-reduced once per epoch.
Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# SURVEY §8d synthetic workloads (BASELINE.json configs[1..4]).  `n` is the number of rows PER GPU (weak
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
                                          "--format=csv,noheader,nounits", "-lms", "20"],
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
    """U[0,1) iid float32 — SURVEY §8d distribution (i): throughput, worst-case near-ties."""
    rng = np.random.RandomState(seed)
    out = np.empty((n, d), dtype=np.float32)
    step = max(1, (1 << 24) // d)
    for s in range(0, n, step):
        out[s:s + step] = rng.random_sample((min(step, n - s), d)).astype(np.float32)
    return out


# ------------------------------------------------------------------------ CPU arm
def oracle_rate(wl, rows, epochs, warm=0):
    """The reference's numpy path restated (oracle/som_oracle.py), all host BLAS threads."""
    from oracle import som_oracle as so
    kw = dict(wl["kw"])
    spec = so.SomSpec(gx=wl["gx"], gy=wl["gy"], dim=wl["d"], random_seed=0,
                      n_parallel=len(os.sched_getaffinity(0)) * 500, **kw)
    x = synth(rows, wl["d"], 0)
    w = np.asarray(so.init_weights(spec), dtype=np.float32)
    for t in range(warm):
        w = so.epoch(spec, x, w, t, TOTAL_EPOCHS)
    t0 = time.perf_counter()
    for t in range(warm, warm + epochs):
        w = so.epoch(spec, x, w, t, TOTAL_EPOCHS)
    dt = time.perf_counter() - t0
    return rows * epochs / dt, dt


def run_reference(args, wl, rank):
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    rows = min(wl["n"], max(20_000, int(2.5e10 / (wl["gx"] * wl["gy"] * wl["d"]))))   # ~3 s per epoch on 8 cores
    rate, dt = oracle_rate(wl, rows, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": "SOM training samples*epochs/sec", "value": rate, "unit": "samples*epochs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "rows_per_step": rows, "map": "%dx%d" % (wl["gx"], wl["gy"]),
                   "features": wl["d"]},
        "cpu_baseline": {"value": rate, "unit": "samples*epochs/s", "cores": cores, "kind": "port",
                         "sample": "%d epochs over the first %d rows (numpy %s, reference algorithm restated in "
                                   "oracle/som_oracle.py; Dask is not installed, use_dask=False path)"
                                   % (args.steps, rows, np.__version__)},
        "e2e": {"value": rate, "unit": "samples*epochs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def measure_tf32_peak(dev):
    """Dense TF32 tensor throughput of this GPU (TFLOP/s): the roofline denominator of the TF32 kernel (SURVEY 8d asks
    for it to be measured on the box; MEASURED_PEAKS.json has bf16 only).  cuBLAS, 8192^3, best of 5 after warm-up."""
    import torch
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
def run_gpu(args, wl, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from xpysom_dask_b200 import XPySom

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, d, gx, gy = wl["n"], wl["d"], wl["gx"], wl["gy"]
    K = gx * gy
    if args.rows:
        n = args.rows

    x_host = torch.from_numpy(synth(n, d, seed=rank)).pin_memory()
    x_dev = x_host.to(dev)
    som = XPySom(gx, gy, d, random_seed=0, algo=args.algo, device=dev, use_cuda_graph=args.cuda_graph,
                 process_group=True if world > 1 else None, **wl["kw"])
    eng = som._get_engine()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput: K epochs in ONE train() call ------------------------
    som.train(x_dev, TOTAL_EPOCHS, iter_beg=0, iter_end=args.warmup)             # warm-up epochs
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()       # BEFORE the barrier: spawning nvidia-smi takes ~1 ms, which the other ranks would
    barrier()                 # otherwise spend waiting for rank 0 inside their first all-reduce of the timed region
    launches0 = eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    som.train(x_dev, TOTAL_EPOCHS, iter_beg=args.warmup, iter_end=args.warmup + args.steps)
    e1.record()
    barrier()
    launches = eng.launches - launches0
    ms = e0.elapsed_time(e1)

    # ---- the dominant kernel, bracketed by CUDA events on its own stream: the same `steps` epochs again,
    # launched one by one (events cannot sit inside a replayed graph)
    som._profile = True
    som._profile_events = []
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    som.train(x_dev, TOTAL_EPOCHS, iter_beg=args.warmup, iter_end=args.warmup + args.steps)
    p1.record()
    barrier()
    eager_ms = p0.elapsed_time(p1)
    bmu_ms = float(np.mean([ev[0].elapsed_time(ev[1]) for ev in som._profile_events]))
    som._profile = False

    # ---- end to end: host (pinned) samples in, codebook out, every step -------------------
    som.train(x_host, TOTAL_EPOCHS, iter_beg=0, iter_end=1)
    barrier()
    t0 = time.perf_counter()
    for s in range(args.steps):
        som.train(x_host, TOTAL_EPOCHS, iter_beg=args.warmup + s, iter_end=args.warmup + s + 1)
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0)
    # the sampler covered every measured section (timed epochs, the per-kernel pass, the end-to-end loop): the
    # device-resident region alone lasts a few milliseconds, less than one nvidia-smi sampling period
    clocks = sampler.stop() if rank == 0 else None

    t = torch.tensor([ms, e2e_ms, bmu_ms, eager_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, bmu_ms, eager_ms = t.tolist()

    if rank == 0:
        pk = peaks()
        total = n * world
        value = total * args.steps / (ms * 1e-3)
        flops = 2.0 * n * K * d                       # algorithmic flops of one BMU launch
        dist_name = wl["kw"].get("activation_distance", "euclidean")
        contraction = dist_name in ("euclidean", "cosine") and args.algo != "simt"
        # which tensor-core kernel AUTO resolves to (som_api.cu: pick_algo)
        f16 = contraction and (args.algo == "tc16" or (args.algo == "auto" and d > 32))
        if not contraction:
            kernel, peak, peak_note = "bmu_simt_kernel + accumulate_kernel", None, "SIMT fp32"
        elif f16:
            kernel = "bmu_tc3_kernel (tcgen05 kind::f16, 3-term fp16 split, cta_group::2, argmin + per-BMU accumulate fused)"
            peak, peak_note = pk["bf16"], "dense bf16/fp16 tensor rate bf16_tflops from %s" % pk["source"]
        else:
            kernel = "bmu_tc2_kernel (tcgen05 kind::tf32, 3-term TF32 split, cta_group::2, argmin + per-BMU accumulate fused)"
            tf32 = measure_tf32_peak(dev)
            if tf32:
                peak, peak_note = tf32, ("dense TF32 rate measured here with a cuBLAS 8192^3 TF32 GEMM (MEASURED_PEAKS.json "
                                         "holds bf16 only: %.0f TFLOP/s)" % pk["bf16"])
            else:
                peak, peak_note = pk["bf16"] / 2.0, "dense TF32 rate taken as half of bf16_tflops from %s" % pk["source"]
        ach = flops / (bmu_ms * 1e-3) / 1e12
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            with open(tp) as f:
                traffic = json.load(f).get(args.workload)
        roofline = {
            "kernel": kernel,
            "bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
            "frac": (ach / peak) if peak else None,
            "traffic": traffic,
            "note": "achieved = 2*n*K*D algorithmic flops / CUDA-event time of the fused BMU kernel, averaged over the "
                    "same epochs launched the same way right after the timed region; peak = %s; the kernel executes 3x the algorithmic flops (hi/lo split for fp32 accuracy), "
                    "so its attainable ceiling is frac 0.333" % peak_note,
            "frac_of_3pass_ceiling": (ach / (peak / 3.0)) if peak else None,
            "kernel_ms": bmu_ms, "step_ms": ms / args.steps, "step_ms_unfused_launches": eager_ms / args.steps,
            "hbm": {"achieved": 4.0 * n * d / ((ms / args.steps) * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "note": "sample-read bytes 4*D per sample-epoch / step time"},
        }
        line = {
            "metric": "SOM training samples*epochs/sec", "value": value, "unit": "samples*epochs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32 (tensor-core contraction on a 3-term %s split)" % ("fp16" if f16 else "tf32") if contraction else "f32",
            "data": "synthetic",
            "config": {"workload": wl["name"], "rows_per_gpu": n, "map": "%dx%d" % (gx, gy), "features": d,
                       "algo": args.algo, "parallelism": "dp%d" % world,
                       "l2": "inputs larger than L2 (%.0f MB of samples per GPU per epoch)" % (4e-6 * n * d)},
            "clocks": clocks,
            "e2e": {"value": total * args.steps / (e2e_ms * 1e-3), "unit": "samples*epochs/s",
                    "h2d_bytes_per_step": 4 * n * d + 4 * K * d, "d2h_bytes_per_step": 4 * K * d,
                    "note": "XPySom.train(pinned host samples, one epoch per call): H2D of the samples and codebook, "
                            "D2H of the codebook inside the timed region"},
            "gpu_launches": launches,
            "roofline": roofline,
        }
        if world == 1 and not args.no_cpu:
            cores = len(os.sched_getaffinity(0))
            rows = min(n, max(20_000, int(1.0e11 / (K * d))))
            rate, dt = oracle_rate(wl, rows, 1)
            line["cpu_baseline"] = {"value": rate, "unit": "samples*epochs/s", "cores": cores, "kind": "port",
                                    "sample": "1 epoch over the first %d rows, %.1f s (numpy %s, all BLAS threads)"
                                              % (rows, dt, np.__version__)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--algo", default="auto", choices=["auto", "tc16", "tc", "simt"])
    ap.add_argument("--rows", type=int, default=0, help="override rows per GPU (debug)")
    ap.add_argument("--cuda-graph", action="store_true", help="replay one captured CUDA graph per epoch")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    global TOTAL_EPOCHS
    TOTAL_EPOCHS = max(TOTAL_EPOCHS, max(args.warmup, 3) + args.steps)     # the timed epochs stay inside the schedule
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl, rank)
        return
    if args.warmup < 3:
        args.warmup = 3                    # timing rule: W >= 3
    run_gpu(args, wl, rank, world, local_rank)


if __name__ == "__main__":
    main()
