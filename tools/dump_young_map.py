"""Train the config-4 map for a few epochs of the 100-epoch schedule on uniform data and save the codebook (+ a sample of
rows) for the CPU analysis of candidate-set sizes (tools/filter_candidates.py --map ...)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from xpysom_dask_b200 import XPySom
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400_000
x = torch.rand((n, 784), generator=torch.Generator(device="cuda").manual_seed(0), device="cuda")
som = XPySom(100, 100, 784, random_seed=0, device="cuda:0")
os.makedirs("gpurun_out", exist_ok=True)
done = 0
for t in (8,):
    som.train(x, 100, iter_beg=done, iter_end=t)
    done = t
    np.save("gpurun_out/young_w_t%d.npy" % t, np.asarray(som._weights, dtype=np.float16 if False else np.float32))
np.save("gpurun_out/young_x.npy", x[:1024].cpu().numpy())
print("saved", som.stats)
