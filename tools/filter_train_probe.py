"""Per-epoch statistics of the filter path inside XPySom.train on the config-4 shape.   python tools/filter_train_probe.py [rows] [data]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from xpysom_dask_b200 import XPySom   # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
kind = sys.argv[2] if len(sys.argv) > 2 else "uniform"
T = int(sys.argv[3]) if len(sys.argv) > 3 else 30
if kind == "blobs":
    import som_testutil as U
    x = torch.from_numpy(U.blobs(n, 784, seed=0)).cuda()
else:
    x = torch.rand((n, 784), generator=torch.Generator(device="cuda").manual_seed(0), device="cuda")
som = XPySom(100, 100, 784, random_seed=0, device="cuda:0")
XPySom._FILTER_MAX_CANDIDATES = float(os.environ.get("MAXC", "12"))
for t in range(0, min(T, int(os.environ.get("EPOCHS", "12")))):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    som.train(x, T, iter_beg=t, iter_end=t + 1)
    torch.cuda.synchronize(); dt = 1e3 * (time.perf_counter() - t0)
    print("epoch %2d: %.2f ms  probe(ovf frac, cand/row) %s  full %s  filter epochs so far %d"
          % (t, dt, som.stats.get("filter_probe"), som.stats.get("filter_last"), som.stats.get("filter_epochs", 0)), flush=True)
    som.stats.pop("filter_probe", None); som.stats.pop("filter_last", None)
