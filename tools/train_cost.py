"""Wall time of XPySom.train on resident samples for several epoch counts (fixed cost per call vs marginal epoch)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from xpysom_dask_b200 import XPySom

dev = torch.device("cuda:0")
n, d, gx, gy = 1000000, 64, 32, 32
x = torch.rand(n, d, device=dev)
som = XPySom(gx, gy, d, random_seed=0, device=dev)
som.train(x, 1000, iter_beg=0, iter_end=3)
torch.cuda.synchronize()
for ne in (0, 1, 2, 5, 10, 20, 40):
    ts = []
    for rep in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        som.train(x, 1000, iter_beg=3, iter_end=3 + ne)
        torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    print("train %2d epochs: %.3f ms (min of 5), %.3f ms/epoch" % (ne, min(ts) * 1e3, min(ts) * 1e3 / max(ne, 1)))

# where the fixed cost goes (host side of one call)
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(5):
        som.train(x, 1000, iter_beg=3, iter_end=5)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cpu_time_total", row_limit=22, max_name_column_width=50))
