"""Where the time of one host-fed train() call goes (config 2: 1M x 64 pinned host samples, one epoch per call).

    python tools/e2e_probe.py [rows] [d] [gx] [gy]
"""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from xpysom_dask_b200 import XPySom  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
d = int(sys.argv[2]) if len(sys.argv) > 2 else 64
gx = int(sys.argv[3]) if len(sys.argv) > 3 else 32
gy = int(sys.argv[4]) if len(sys.argv) > 4 else 32
dev = torch.device("cuda", 0)
x_host = torch.from_numpy(np.random.RandomState(0).random_sample((n, d)).astype(np.float32)).pin_memory()


def wall(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(1e3 * (time.perf_counter() - t0))
    return float(np.median(ts)), float(np.min(ts))


buf = torch.empty((n, d), dtype=torch.float32, device=dev)
print("H2D one copy            : median %.3f ms (min %.3f)" % wall(lambda: buf.copy_(x_host, non_blocking=True)))


def chunked():
    rows = -(-n // 16)
    for r0 in range(0, n, rows):
        buf[r0:r0 + rows].copy_(x_host[r0:r0 + rows], non_blocking=True)


print("H2D 16 chunks           : median %.3f ms (min %.3f)" % wall(chunked))
print("mem_get_info            : median %.3f ms (min %.3f)" % wall(lambda: torch.cuda.mem_get_info(dev)))
som = XPySom(gx, gy, d, random_seed=0, device=dev)
x_dev = x_host.to(dev)
print("train(device, 1 epoch)  : median %.3f ms (min %.3f)" % wall(lambda: som.train(x_dev, 30, iter_beg=3, iter_end=4)))
print("train(host, 1 epoch)    : median %.3f ms (min %.3f)" % wall(lambda: som.train(x_host, 30, iter_beg=3, iter_end=4)))
print("train(host, 2 epochs)   : median %.3f ms (min %.3f)" % wall(lambda: som.train(x_host, 30, iter_beg=3, iter_end=5)))

from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        som.train(x_host, 30, iter_beg=3, iter_end=4)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
print(prof.key_averages().table(sort_by="cpu_time_total", row_limit=20, max_name_column_width=60))
