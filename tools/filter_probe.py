"""Filter + refine BMU path against the three-pass kernel and fp64 truth; timing.   python tools/filter_probe.py N D K [data]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from xpysom_dask_b200 import _lib                          # noqa: E402
from xpysom_dask_b200.engine import CudaEngine             # noqa: E402

n, d, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
kind = sys.argv[4] if len(sys.argv) > 4 else "uniform"
eng = CudaEngine("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.rand(n, d, generator=g, device="cuda")
if kind == "smooth":      # a young, smooth map: every neuron close to the data mean
    gx = int(round(k ** 0.5))
    u = torch.linspace(-1, 1, gx, device="cuda")
    a, b = torch.randn(d, generator=g, device="cuda") * 0.02, torch.randn(d, generator=g, device="cuda") * 0.02
    w = (0.5 + u[:, None, None] * a + u[None, :, None] * b).reshape(-1, d)[:k].contiguous()
    w = w + 1e-4 * torch.randn(w.shape, generator=g, device="cuda")
else:
    w = torch.rand(k, d, generator=g, device="cuda")
k = w.shape[0]
assert eng.filter_eligible(x, k, 0), "shape not eligible"
fws = eng.filter_workspace(x, k)
bmu_f = torch.full((n,), -1, dtype=torch.int32, device='cuda')
eng.bmu_filter(x, w, fws, bmu_f)
ov0, ev0 = eng.filter_stats(fws, n, k, d)
print('unseeded sweep: overflow rows %d, %.2f candidates re-scored per row' % (ov0, ev0 / max(n - ov0, 1)))
bmu_f.clamp_(min=0)
eng.bmu_filter(x, w, fws, bmu_f)          # second call: seeded with the first call's BMUs
ovf, evals = eng.filter_stats(fws, n, k, d)
print("filter: overflow rows %d of %d, %.2f candidates re-scored per row" % (ovf, n, evals / max(n - ovf, 1)))
# fp64 truth on a sample of rows
idx = torch.randperm(n, generator=torch.Generator().manual_seed(1))[:2048].cuda()
xs = x[idx].double()
d2 = (xs * xs).sum(1, keepdim=True) - 2 * xs @ w.double().T + (w.double() ** 2).sum(1)[None, :]
truth = d2.argmin(1).int()
mis = (bmu_f[idx] != truth) & (bmu_f[idx] >= 0)
print("vs fp64 truth on %d rows: %d mismatches" % (len(idx), int(mis.sum())))
if mis.any():
    r = torch.nonzero(mis)[:5, 0]
    for i in r.tolist():
        a, b = int(bmu_f[idx][i]), int(truth[i])
        print("   row %d: filter %d (%.9e) truth %d (%.9e)" % (int(idx[i]), a, float(d2[i, a]), b, float(d2[i, b])))
# against the three-pass kernel
ws = eng.workspace(0, k, d)
eng.prepare_codebook(w, 0, 2.0, ws)
xsc, colmax = eng.prepare_samples(x, True)
bmu_t = eng.bmu(x, w, 0, 2.0, _lib.ALGO["tc16"], ws, xscale=xsc)
print("vs three-pass kernel: mismatch rate %.3e" % float(((bmu_t != bmu_f) & (bmu_f >= 0)).float().mean()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, fn in (("filter + refine", lambda: eng.bmu_filter(x, w, fws, bmu_f)),
                 ("three-pass", lambda: eng.bmu(x, w, 0, 2.0, _lib.ALGO["tc16"], ws, bmu_out=bmu_t, xscale=xsc))):
    fn(); torch.cuda.synchronize()
    e0.record()
    for _ in range(3):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print("%-16s %.3f ms  (%.1f TFLOP/s algorithmic)" % (name, ms, 2.0 * n * k * d / ms / 1e9))
# candidate statistics
L = _lib.load()
meta_off = None
