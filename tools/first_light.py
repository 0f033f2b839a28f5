"""Bring-up check of the BMU kernels on small shapes: every tensor-core variant (fp16 packed / folded / resident /
streaming, TF32) against the SIMT kernel, mismatch rate and kernel time.  Run under `timeout`: a protocol bug traps
after a few seconds, it must not eat the GPU budget."""
import sys
import time

import torch

sys.path.insert(0, ".")
from xpysom_dask_b200 import _lib                      # noqa: E402
from xpysom_dask_b200.engine import CudaEngine         # noqa: E402

eng = CudaEngine("cuda:0")
shapes = [(20000, 64, 1024, "euclidean"), (20000, 16, 1600, "euclidean"), (20000, 12, 300, "cosine"),
          (20000, 100, 900, "euclidean"), (6000, 784, 2000, "euclidean"), (20000, 40, 1024, "cosine"),
          (1_000_000, 64, 1024, "euclidean"), (2_000_000, 16, 1600, "euclidean")]
if len(sys.argv) > 1:
    shapes = [tuple(int(v) if i < 3 else v for i, v in enumerate(a.split(","))) for a in sys.argv[1:]]
for n, d, K, dist in shapes:
    g = torch.Generator(device="cuda").manual_seed(n + d)
    x = torch.rand(n, d, generator=g, device="cuda")
    w = torch.rand(K, d, generator=g, device="cuda")
    ws = eng.workspace(0, K, d)
    eng.prepare_codebook(w, _lib.DIST[dist], 2.0, ws)
    xs, colmax = eng.prepare_samples(x)
    ref = eng.bmu(x, w, _lib.DIST[dist], 2.0, _lib.ALGO["simt"], ws)
    torch.cuda.synchronize()
    for algo in ("tc16", "tc"):
        t0 = time.time()
        got = eng.bmu(x, w, _lib.DIST[dist], 2.0, _lib.ALGO[algo], ws, xscale=xs)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            eng.bmu(x, w, _lib.DIST[dist], 2.0, _lib.ALGO[algo], ws, xscale=xs)
        e1.record()
        torch.cuda.synchronize()
        mism = (got != ref).float().mean().item()
        print("n=%d d=%d K=%d %s %-4s mismatch vs simt %.2e   %.3f ms" % (n, d, K, dist, algo, mism, e0.elapsed_time(e1) / 3),
              flush=True)
    # fused accumulate, exact
    qs, qi = eng.accum_scales(colmax, d, n)
    acc = eng.accumulator(K, d)
    bmu = eng.empty(n, dtype=torch.int32)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    eng.epoch_accumulate(x, w, _lib.DIST[dist], 2.0, _lib.ALGO["auto"], qs, acc, ws, bmu_out=bmu, xscale=xs)
    acc.zero_()
    e0.record()
    eng.epoch_accumulate(x, w, _lib.DIST[dist], 2.0, _lib.ALGO["auto"], qs, acc, ws, bmu_out=bmu, xscale=xs)
    e1.record()
    torch.cuda.synchronize()
    S_, c_ = eng.empty(K, d), eng.empty(K)
    eng.accum_finalize(acc, qi, K, d, S_, c_)             # (sums the accumulator's replicas)
    cnt = int(c_.sum(dtype=torch.float64).item())
    print("   fused auto: %.3f ms, counts sum %d (n=%d), bmu mismatch vs simt %.2e" %
          (e0.elapsed_time(e1), cnt, n, (bmu != ref).float().mean().item()), flush=True)
