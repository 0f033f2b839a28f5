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


def eps_for(d):
    """The stated epsilon as a function of the feature count: 1e-6 up to 64 features, growing like sqrt(D / 64)
    above (D = 784: 3.5e-6).  The band has to cover the rounding noise of the REFERENCE's own fp32 sgemm, whose
    dot products of D terms carry an error ~ u sqrt(D): at D = 784 numpy disagrees with its own fp64 evaluation on
    rows whose fp32 gap is up to ~1.3e-6 of the scale (measured, tests/test_gpu_config_parity.py)."""
    return EPS * max(1.0, (d / 64.0) ** 0.5)


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


def bmu_parity(spec, x, w, bmu_gpu, eps=None):
    """Compare GPU BMUs with the oracle's.  Returns a dict with the mismatch count
    outside the near-tie band (must be 0), the near-tie rate and the raw mismatch rate."""
    if eps is None:
        eps = eps_for(x.shape[1])
    b_ref, d1, gap, scale = so.top2_gap(spec, x, w)
    clear = gap > eps * scale
    mism = np.asarray(bmu_gpu) != b_ref
    worst = float((gap[mism] / scale[mism]).max()) if mism.any() else 0.0
    return dict(bad=int((mism & clear).sum()), near_tie_rate=float((~clear).mean()),
                mismatch_rate=float(mism.mean()), worst_rel_gap=worst, n=len(b_ref), eps=eps)


def fp64_regret(spec, x, w, bmu_gpu, bmu_ref):
    """On the rows where the GPU and the reference picked different units: the fp64 score of each pick minus the
    fp64 minimum over all units, relative to the epsilon's scale.  Says which side the rounding noise sits on."""
    sel = np.nonzero(np.asarray(bmu_gpu) != np.asarray(bmu_ref))[0]
    if len(sel) == 0:
        return dict(rows=0, gpu_worst=0.0, ref_worst=0.0, gpu_better=0)
    sel = sel[:4096]
    xs = x[sel].astype(np.float64)
    wf = w.reshape(-1, w.shape[-1]).astype(np.float64)
    if spec.activation_distance == "cosine":
        nw = np.linalg.norm(wf, axis=1)
        nx = np.linalg.norm(xs, axis=1)
        with np.errstate(divide="ignore", invalid="ignore"):
            d = 1.0 - np.nan_to_num((xs @ wf.T) / (nx[:, None] * nw[None, :]))
        scale = np.ones(len(sel))
    else:
        d = -2.0 * xs @ wf.T + (wf * wf).sum(1)[None, :]
        scale = (xs * xs).sum(1) + np.abs(d.min(axis=1))
    r = np.arange(len(sel))
    best = d.min(axis=1)
    g = (d[r, np.asarray(bmu_gpu)[sel]] - best) / scale
    f = (d[r, np.asarray(bmu_ref)[sel]] - best) / scale
    return dict(rows=int(len(sel)), gpu_worst=float(g.max()), ref_worst=float(f.max()), gpu_better=int((g < f).sum()))



def codebook_rel_err(w_gpu, w_ref):
    """max-abs difference over max-abs reference value (the survey's definition)."""
    w_ref = np.asarray(w_ref, dtype=np.float64)
    return float(np.abs(np.asarray(w_gpu, dtype=np.float64) - w_ref).max() / np.abs(w_ref).max())


def separable_factors(spec, sigma):
    """Per-axis factors of a product-form neighbourhood on a rectangular map, read off the oracle's own function:
    h((bi,bj),(i,j)) = A[bi,i] * B[bj,j] / h00   (gaussian, bubble, triangle; neighborhoods.py:33,112,130)."""
    gx, gy = spec.gx, spec.gy
    zx, zy = np.zeros(gx, dtype=np.int64), np.zeros(gy, dtype=np.int64)
    A = np.asarray(so.neighborhood(spec, np.arange(gx), zx, sigma), dtype=np.float64)[:, :, 0]
    B = np.asarray(so.neighborhood(spec, zy, np.arange(gy), sigma), dtype=np.float64)[:, 0, :]
    h00 = float(A[0, 0])
    return A, B, h00
