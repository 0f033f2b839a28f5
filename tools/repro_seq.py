import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from xpysom_dask_b200 import _lib
from xpysom_dask_b200.engine import CudaEngine
eng = CudaEngine("cuda:0")
g = torch.Generator(device="cuda").manual_seed(0)
mode = sys.argv[1]
shapes = [(1000000, 64, 1024), (2000000, 16, 1600), (100000, 784, 10000), (500000, 128, 2500), (2000000, 16, 1600)]
for (n, d, k) in shapes:
    x = torch.rand(n, d, generator=g, device="cuda"); w = torch.rand(k, d, generator=g, device="cuda")
    ws = eng.workspace(0, k, d); eng.prepare_codebook(w, 0, 2.0, ws)
    xs = eng.prepare_samples(x)
    best = eng.empty(n) if mode == "best" else None
    bmu = eng.empty(n, dtype=torch.int32)
    for i in range(6):
        eng.bmu(x, w, 0, 2.0, _lib.ALGO["tc16"], ws, bmu_out=bmu, best_out=best if i == 0 else None, xscale=xs)
        if mode == "sync":
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    print("ok", n, d, k, int(bmu.sum().item()), flush=True)
    if mode == "truth":
        m = 4096
        xs64, wd = x[:m].double(), w.double()
        dd = (wd * wd).sum(1)[None, :] - 2 * xs64 @ wd.T
        print("truth mism", int((dd.argmin(1) != bmu[:m].long()).sum()), flush=True)
