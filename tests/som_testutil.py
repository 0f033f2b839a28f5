"""Shared helpers for the GPU parity tests (oracle = checker only)."""
import json
import os

import numpy as np

from oracle import som_oracle as so

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# Stated near-tie epsilon of the BMU parity criterion (BASELINE.json north_star):
# a row counts as a near tie when (d2 - d1) <= EPS * scale on the reference's own
# fp32 distances, scale = |x|^2 + |d1| (euclidean partial distance), 1 (cosine),
# |d1| (L1 / Linf / Lp sums).  Outside that band BMUs must be bit-exact.
EPS = 1e-6


def blobs(n, d, seed, centres=64, spread=0.1):
    rng = np.random.RandomState(seed)
    c = rng.rand(centres, d)
    return (c[rng.randint(centres, size=n)] + spread * rng.randn(n, d)).astype(np.float32)


def uniform(n, d, seed):
    return np.random.RandomState(seed).rand(n, d).astype(np.float32)


def load(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def load_json(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def spec_from_case(c):
    kw = dict(c["kwargs"])
    adk = kw.pop("activation_distance_kwargs", {})
    return so.SomSpec(gx=c["gx"], gy=c["gy"], dim=c["D"], n_parallel=c["n_parallel"],
                      random_seed=c["seed"], p=adk.get("p", 2), **kw)


def bmu_parity(spec, x, w, bmu_gpu, eps=EPS):
    """Compare GPU BMUs with the oracle's.  Returns a dict with the mismatch count
    outside the near-tie band (must be 0), the near-tie rate and the raw mismatch rate."""
    b_ref, d1, gap, scale = so.top2_gap(spec, x, w)
    clear = gap > eps * scale
    mism = np.asarray(bmu_gpu) != b_ref
    worst = float((gap[mism] / scale[mism]).max()) if mism.any() else 0.0
    return dict(bad=int((mism & clear).sum()), near_tie_rate=float((~clear).mean()),
                mismatch_rate=float(mism.mean()), worst_rel_gap=worst, n=len(b_ref))


def codebook_rel_err(w_gpu, w_ref):
    """max-abs difference over max-abs reference value (the survey's definition)."""
    w_ref = np.asarray(w_ref, dtype=np.float64)
    return float(np.abs(np.asarray(w_gpu, dtype=np.float64) - w_ref).max() / np.abs(w_ref).max())
