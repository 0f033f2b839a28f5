import os, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from xpysom_dask_b200 import _lib
if os.environ.get("SOM_TOOL_LIB"): _lib.LIB_PATH = os.environ["SOM_TOOL_LIB"]
from xpysom_dask_b200 import XPySom
x = torch.rand((1_000_000, 64), device="cuda")
som = XPySom(32, 32, 64, random_seed=0, device="cuda:0")
som.train(x, 1000, iter_beg=0, iter_end=5)
ts = []
for rep in range(6):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); som.train(x, 1000, iter_beg=5, iter_end=55); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / 50)
    time.sleep(0.5)
print(os.environ.get("SOM_TOOL_LIB"), " ".join("%.4f" % t for t in ts))
