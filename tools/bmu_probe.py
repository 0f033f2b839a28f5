"""Run the BMU kernel (optionally fused with the accumulate) a few times on one shape; for ncu.

    python tools/bmu_probe.py N D K [fused] [reps]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from xpysom_dask_b200 import _lib                          # noqa: E402
from xpysom_dask_b200.engine import CudaEngine             # noqa: E402

if os.environ.get("SOM_TOOL_LIB"):          # tuning: time another build of the library (tools/variants/*.so)
    _lib.LIB_PATH = os.environ["SOM_TOOL_LIB"]
n, d, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
fused = len(sys.argv) > 4 and sys.argv[4] == "fused"
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 5
algo = sys.argv[6] if len(sys.argv) > 6 else "tc"
eng = CudaEngine("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.rand(n, d, generator=g, device="cuda")
w = torch.rand(k, d, generator=g, device="cuda")
ws = eng.workspace(0, k, d)
eng.prepare_codebook(w, 0, 2.0, ws)
bmu = eng.empty(n, dtype=torch.int32)
xs, colmax = eng.prepare_samples(x, want_scale=algo in ("tc16", "auto"))
qs, qi = eng.accum_scales(colmax, d, n)
acc = eng.accumulator(k, d)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(reps + 2):
    if i == 2:
        e0.record()
    if fused:
        eng.epoch_accumulate(x, w, 0, 2.0, _lib.ALGO[algo], qs, acc, ws, bmu_out=bmu, xscale=xs)
    else:
        eng.bmu(x, w, 0, 2.0, _lib.ALGO[algo], ws, bmu_out=bmu, xscale=xs)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(algo, "n=%d d=%d K=%d fused=%s: %.3f ms, %.1f TFLOP/s algorithmic, %.3e elements/s"
      % (n, d, k, fused, ms, 2.0 * n * k * d / ms / 1e9, n * k / ms * 1e3))
if os.environ.get("SOM_B200_DBG") == "9":
    import ctypes
    import numpy as np
    buf = np.zeros(8 * 64, dtype=np.int64)
    _lib.check(eng.lib.som_b200_debug_timeline(buf.ctypes.data_as(ctypes.c_void_p), buf.size), "timeline")
    t = buf.reshape(64, 8)
    t0 = t[4:, 0].min()
    print("tile |  MMA: wait_acc  acc_free  ops_ready  committed |  EPI: wait  tile_ready  drained   (cycles, leader CTA 0)")
    for i in range(8, 40):
        print("%4d | %9d %9d %9d %9d | %9d %9d %9d" % ((i,) + tuple(int(v - t0) for v in t[i, :7])))
if os.environ.get("SOM_B200_DBG") == "8":
    import ctypes
    import numpy as np
    buf = np.zeros(8 * 64, dtype=np.int64)
    _lib.check(eng.lib.som_b200_debug_timeline(buf.ctypes.data_as(ctypes.c_void_p), buf.size), "timeline")
    t = buf.reshape(64, 8)
    t0 = t[8, 0]
    print("batch | start  loads_issued  buffers_free  stored  fenced  sent   (cycles, scatter warp 0 of CTA 0)")
    for i in range(8, 28):
        print("%4d | %8d %8d %8d %8d %8d %8d" % ((i,) + tuple(int(v - t0) for v in t[i, :6])))
