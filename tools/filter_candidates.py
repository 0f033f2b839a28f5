"""how many neurons survive a 1-pass fp16 filter at config 4 (K = 10^4, D = 784)?

A 1-pass contraction x_hi . w_hi (fp16 inputs, fp32 accumulate) costs a third of the 3-term split.  It can only
replace it if the neurons whose approximate score is within the RIGOROUS error bound of the approximate minimum are
few enough to be re-scored exactly.  Bound per (row, neuron), from the actual rounding residuals (Cauchy-Schwarz):
    |s_k - s^_k| <= E_k = |x_lo| |w_hi,k| + |x_hi| |w_lo,k| + |x_lo| |w_lo,k|        (score s_k = |w_k|^2 - 2 x . w_k)
candidates = { k : s^_k - E_k <= min_j (s^_j + E_j) }.
Measured for uniform and 64-blob data, on the random initial codebook and on a young (smooth) map after 3 epochs, raw and
with the data's column means subtracted from samples and codebook (the Euclidean BMU is translation invariant).

    python tools/filter_candidates.py            (CPU only; ~2 minutes)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import som_oracle as so  # noqa: E402
import som_testutil as U             # noqa: E402

GX = GY = 100
D = 784
N_TRAIN, N_EVAL = 6000, 1500


def split16(a):
    """fp16 hi/lo split of rows scaled by a power of two (as the kernel does): returns hi, lo in float64, unscaled."""
    amax = np.abs(a).max(axis=1, keepdims=True)
    amax[amax == 0] = 1.0
    sc = 2.0 ** (14 - np.floor(np.log2(amax)))
    hi = (a * sc).astype(np.float16).astype(np.float64)
    lo = (a * sc - hi).astype(np.float16).astype(np.float64)
    return hi / sc, lo / sc


def candidates(x, w):
    xh, xl = split16(x)
    wp = -2.0 * w
    wh, wl = split16(wp)
    bias = (w * w).sum(1)
    s_hat = xh @ wh.T + bias[None, :]
    nxh, nxl = np.linalg.norm(xh, axis=1), np.linalg.norm(xl, axis=1)
    nwh, nwl = np.linalg.norm(wh, axis=1), np.linalg.norm(wl, axis=1)
    E = nxl[:, None] * nwh[None, :] + nxh[:, None] * nwl[None, :] + nxl[:, None] * nwl[None, :]
    upper = (s_hat + E).min(axis=1, keepdims=True)
    cnt = ((s_hat - E) <= upper).sum(axis=1)
    # the statistical (non-rigorous) picture: where does the exact BMU rank in the approximate ordering?
    s = x @ wp.T + bias[None, :]
    true = s.argmin(1)
    rank = (s_hat < s_hat[np.arange(len(x)), true][:, None]).sum(1)
    return cnt, rank


def report(tag, x, w):
    cnt, rank = candidates(x.astype(np.float64), w.astype(np.float64))
    print("%-58s candidates: median %5d  p90 %5d  max %5d  <=32: %5.1f%%   | exact BMU's rank in the 1-pass order: max %d"
          % (tag, np.median(cnt), np.percentile(cnt, 90), cnt.max(), 100.0 * (cnt <= 32).mean(), rank.max()), flush=True)


for data_name in ("uniform", "blobs"):
    rng = np.random.RandomState(0)
    xs = (rng.random_sample((N_TRAIN, D)).astype(np.float32) if data_name == "uniform" else U.blobs(N_TRAIN, D, seed=0))
    spec = so.SomSpec(gx=GX, gy=GY, dim=D, random_seed=0, n_parallel=4000)
    w = np.asarray(so.init_weights(spec), dtype=np.float32)
    xe = xs[:N_EVAL]
    mu = xs.mean(0)
    for stage in ("initial codebook", "after 3 of 10 epochs"):
        if stage != "initial codebook":
            for t in range(3):
                w = np.asarray(so.epoch(spec, xs, w, t, 10), dtype=np.float32)
        wf = w.reshape(-1, D)
        report("%s, %s, raw" % (data_name, stage), xe, wf)
        report("%s, %s, column means subtracted" % (data_name, stage), xe - mu, wf - mu)
