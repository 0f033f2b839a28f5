"""GPU bring-up diagnostics: kernel agreement and timings (not a test, not the bench).

    python tools/first_light.py simt|tc|epoch
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import som_oracle as so                       # noqa: E402
from xpysom_dask_b200 import _lib                         # noqa: E402
from xpysom_dask_b200.engine import CudaEngine            # noqa: E402


def ev_time(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def data(n, d, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.rand(n, d, generator=g, device="cuda", dtype=torch.float32)


def stage_bmu(eng, algo, shapes, dist="euclidean"):
    for n, d, k in shapes:
        x = data(n, d, 1)
        w = data(k, d, 2)
        ws = eng.workspace(0, k, d)
        eng.prepare_codebook(w, _lib.DIST[dist], 2.0, ws)
        best = eng.empty(n)
        xsc = eng.prepare_samples(x) if algo == "tc16" else None
        bmu = eng.bmu(x, w, _lib.DIST[dist], 2.0, _lib.ALGO[algo], ws, best_out=best, xscale=xsc)
        torch.cuda.synchronize()
        # fp64 truth on a subset of rows
        m = min(n, 4096)
        xs, wd = x[:m].double(), w.double()
        if dist == "euclidean":
            dd = (wd * wd).sum(1)[None, :] - 2 * xs @ wd.T
        else:
            dd = -(xs @ wd.T) / wd.norm(dim=1)[None, :]
        truth = dd.argmin(1)
        srt = dd.sort(dim=1).values
        gap = (srt[:, 1] - srt[:, 0]) / ((xs * xs).sum(1) + srt[:, 0].abs())
        mism = (bmu[:m].long() != truth)
        worst = gap[mism].max().item() if mism.any() else 0.0
        score_err = (best[:m].double() - dd.min(1).values).abs().max().item() if dist == "euclidean" else float("nan")
        ms = ev_time(lambda: eng.bmu(x, w, _lib.DIST[dist], 2.0, _lib.ALGO[algo], ws, bmu_out=bmu, xscale=xsc))
        tf = 2.0 * n * k * d / (ms * 1e-3) / 1e12
        print("[%s %s] n=%d d=%d K=%d: mismatch vs fp64 %d/%d (worst rel gap %.2e), |score err| %.2e, %.3f ms, %.1f TFLOP/s algorithmic"
              % (algo, dist, n, d, k, int(mism.sum()), m, worst, score_err, ms, tf), flush=True)


def stage_epoch(eng):
    from xpysom_dask_b200 import XPySom
    for (n, d, gx, gy, kw) in [(1_000_000, 64, 32, 32, {}),
                               (4_000_000, 16, 40, 40, dict(decay_function="linear")),
                               (200_000, 784, 100, 100, {}),
                               (1_000_000, 128, 50, 50, dict(topology="hexagonal", neighborhood_function="mexican_hat",
                                                             activation_distance="cosine"))]:
        x = data(n, d, 3)
        for algo in ("tc16", "tc", "simt"):
            som = XPySom(gx, gy, d, random_seed=0, algo=algo, **kw)
            som.train(x, 100, iter_beg=0, iter_end=2)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            som._profile, som._profile_events = True, []
            som.train(x, 100, iter_beg=2, iter_end=5)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) / 3
            b = np.mean([e[0].elapsed_time(e[1]) for e in som._profile_events])
            print("[epoch %s] n=%d d=%d K=%d %s: %.2f ms/epoch (bmu+accumulate %.2f) -> %.3e samples*epochs/s"
                  % (algo, n, d, gx * gy, kw, dt * 1e3, b, n / dt), flush=True)


if __name__ == "__main__":
    stage = sys.argv[1]
    eng = CudaEngine("cuda:0")
    print("device", torch.cuda.get_device_name(0), "SMs", eng.sm_count, "cc", eng.cc, flush=True)
    small = [(128, 32, 256), (1000, 64, 1024), (4096, 16, 1600), (3000, 100, 300), (2048, 784, 1000)]
    big = [(1_000_000, 64, 1024), (2_000_000, 16, 1600), (100_000, 784, 10_000), (500_000, 128, 2500)]
    if stage == "simt":
        stage_bmu(eng, "simt", small + big[:1])
    elif stage == "tc16":
        stage_bmu(eng, "tc16", small)
        stage_bmu(eng, "tc16", small[:3], dist="cosine")
        stage_bmu(eng, "tc16", big)
    elif stage == "tc":
        stage_bmu(eng, "tc", small)
        stage_bmu(eng, "tc", small[:3], dist="cosine")
        stage_bmu(eng, "tc", big)
    elif stage == "epoch":
        stage_epoch(eng)
