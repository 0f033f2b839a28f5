import sys, time, subprocess
import numpy as np, torch
sys.path.insert(0, ".")
from xpysom_dask_b200 import XPySom
dev = torch.device("cuda", 0)
n, d = 1_000_000, 64
x_host = torch.from_numpy(np.random.RandomState(0).random_sample((n, d)).astype(np.float32)).pin_memory()
som = XPySom(32, 32, d, random_seed=0, device=dev)
def loop(tag):
    som.train(x_host, 30, iter_beg=0, iter_end=1)
    torch.cuda.synchronize()
    ts = []
    for s in range(20):
        t0 = time.perf_counter()
        som.train(x_host, 30, iter_beg=3 + s, iter_end=4 + s)
        ts.append(1e3 * (time.perf_counter() - t0))
    print(tag, "mean %.3f median %.3f min %.3f max %.3f" % (np.mean(ts), np.median(ts), np.min(ts), np.max(ts)), flush=True)
loop("quiet      ")
for ms in ("20", "50", "200"):
    p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", ms], stdout=subprocess.DEVNULL)
    time.sleep(0.3)
    loop("smi -lms %-3s" % ms)
    p.terminate(); p.wait()
loop("quiet again")
