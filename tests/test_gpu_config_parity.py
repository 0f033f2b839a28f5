"""Parity of the CUDA path against the oracle AT THE MAP / FEATURE SIZES of BASELINE.json's configs 2-5.

SURVEY.md 8d fixes the subsets: the first 100k rows of C2 / C3 / C5 and the first 20k rows of C4, 64-blob
Gaussian mixture data (the parity distribution), so that the numpy oracle finishes in seconds.  Every case is
teacher-forced (the same W_t goes into both sides, one epoch, reference xpysom.py:458-594 driven through
iter_beg / iter_end, :515) at t in {0, T/2, T-1} of a T-epoch schedule; W_t for t > 0 is the codebook the GPU
itself reached after t free-running epochs, i.e. a trained, smooth map -- the regime where neighbouring
neurons are nearly identical and near-ties are frequent.

Asserted, with `algo=auto` (the kernel the product picks):
  * BMUs bit-exact wherever the oracle's own top-2 gap exceeds eps(D) * scale (som_testutil.eps_for: 1e-6 up to
    64 features, 1e-6 sqrt(D / 64) above -- the reference's own fp32 dot products carry that much noise); the
    near-tie rate and the raw mismatch rate are printed and written to gpurun_out/parity_configs.json;
  * codebook after the epoch within 1e-5 of the oracle's update evaluated (fp64) from the GPU's own BMUs --
    this isolates accumulate + neighbourhood + merge and holds whatever the near-tie rate is;
  * codebook within 1e-4 of the oracle's own epoch whenever no BMU differs.  When BMUs inside the near-tie
    band did flip, the 1e-4 criterion is not well posed (one sample moving between two adjacent neurons
    changes their means by ~1/count): the error is REPORTED next to the flip count instead of skipped, and it
    is explained by the flips alone: every flipped row lies inside the band, and the oracle's update evaluated from
    the GPU's assignment (= the reference's assignment with exactly those rows re-assigned) reproduces the GPU
    codebook to 1e-5.
"""
import json
import os

import numpy as np
import pytest

from oracle import som_oracle as so
import som_testutil as U

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
T_SCHEDULE = 10

CONFIGS = {
    # name: rows of the parity subset, D, gx, gy, constructor kwargs (BASELINE.json configs[1..4])
    "c2": (100_000, 64, 32, 32, {}),
    "c3": (100_000, 16, 40, 40, dict(decay_function="linear")),
    "c4": (20_000, 784, 100, 100, {}),
    "c5": (100_000, 128, 50, 50, dict(topology="hexagonal", neighborhood_function="mexican_hat",
                                      activation_distance="cosine")),
}
_REPORT = {}


def _oracle_update_from_bmus(spec, data, w_in, bmu, t):
    """W_{t+1} the reference's formulas give for a GIVEN assignment of samples to BMUs (fp64)."""
    sig = float(so.decay_value(spec.decay_function, spec.sigma, spec.sigmaN, t, T_SCHEDULE))
    S, c = so.sums_by_bmu(np.asarray(bmu), data, spec.K)
    if spec.topology == "rectangular" and spec.neighborhood_function != "mexican_hat":
        # product-form neighbourhoods factor per axis (neighborhoods.py:33,112,130):
        # h((bi,bj),(i,j)) = A[bi,i] * B[bj,j] / h00 with A, B, h00 read off the reference formula itself --
        # avoids the K x K table at K = 10^4 (checked against the full table in test_oracle_golden.py)
        A, B, h00 = U.separable_factors(spec, sig)
        S3 = S.reshape(spec.gx, spec.gy, -1)
        num = np.einsum("ai,abd->ibd", A, S3, optimize=True)
        num = (np.einsum("bj,ibd->ijd", B, num, optimize=True) / h00).reshape(spec.K, -1)
        den = (A.T @ c.reshape(spec.gx, spec.gy) @ B / h00).reshape(spec.K)
    else:
        H = so.neighborhood_table(spec, sig).astype(np.float64)
        num, den = H.T @ S, H.T @ c
    w_flat = w_in.reshape(spec.K, -1).astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        out = np.where(den[:, None] != 0, num / den[:, None], w_flat)
    return out.reshape(w_in.shape), den


def _rel(a, b, mask=None):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    diff = np.abs(a - b)
    if mask is not None:
        diff = diff[mask]
    return float(diff.max() / np.abs(b).max())


@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_config_scale_teacher_forced_parity(name):
    from xpysom_dask_b200 import XPySom
    rows, d, gx, gy, kw = CONFIGS[name]
    data = U.blobs(rows, d, seed=0)
    spec = so.SomSpec(gx=gx, gy=gy, dim=d, random_seed=0, n_parallel=4000, **kw)
    som = XPySom(gx, gy, d, random_seed=0, algo="auto", **kw)
    np.testing.assert_array_equal(som._weights, so.init_weights(spec))       # same initial codebook on both sides
    x_dev = torch.from_numpy(data).cuda()
    lines = []
    done = 0
    for t in (0, T_SCHEDULE // 2, T_SCHEDULE - 1):
        if t > done:                                   # free-running GPU epochs up to t: a trained, smooth W_t
            som.train(x_dev, T_SCHEDULE, iter_beg=done, iter_end=t)
            done = t
        w_in = np.asarray(som._weights, dtype=np.float32).copy()
        # GPU: BMUs for W_t, then one epoch from W_t
        bmu_gpu = som.predict(x_dev)
        som.train(x_dev, T_SCHEDULE, iter_beg=t, iter_end=t + 1)
        # the BMUs the training epoch itself used (same W_t).  On long rows they come from the one-pass filter + exact fp64
        # refinement, which may settle a near-tie differently from predict()'s three-pass kernel
        if getattr(som, "_bmu_last", None) is not None:
            bmu_gpu = som._bmu_last.cpu().numpy()
        w_gpu = som._weights.copy()
        som._weights = w_in.copy()                     # later t continue from the SAME trajectory
        # oracle: the reference's epoch from the same W_t, its BMUs and its top-2 gaps
        w_ref, bmu_ref = so.epoch(spec, data, w_in, t, T_SCHEDULE, return_bmu=True)
        r = U.bmu_parity(spec, data, w_in, bmu_gpu)
        flips = int((bmu_gpu != bmu_ref).sum())
        reg = U.fp64_regret(spec, data, w_in, bmu_gpu, bmu_ref)
        w_same, den = _oracle_update_from_bmus(spec, data, w_in, bmu_gpu, t)
        # the mexican hat's denominators cross zero: compare where |den| is not tiny (as the golden-epoch tests do)
        ok = None
        if spec.neighborhood_function == "mexican_hat":
            ok = (np.abs(den) > 1e-3 * np.abs(den).max()).reshape(gx, gy)
        err_same = _rel(w_gpu, w_same, ok)
        err_ref = _rel(w_gpu, w_ref, ok)
        line = dict(config=name, t=t, rows=rows, K=gx * gy, D=d, near_tie_rate=r["near_tie_rate"],
                    mismatch_rate=r["mismatch_rate"], mismatch_outside_band=r["bad"], flipped_rows=flips,
                    worst_mismatching_rel_gap=r["worst_rel_gap"], eps=r["eps"], fp64_regret_on_flipped_rows=reg,
                    codebook_rel_err_vs_reference_epoch=err_ref,
                    codebook_rel_err_vs_same_bmu_oracle=err_same)
        lines.append(line)
        print("\n[parity %s t=%d] near-tie rate %.3e  raw mismatch %.3e (%d rows, worst rel gap %.2e)  outside band %d  |  "
              "codebook rel err: %.2e vs reference epoch, %.2e vs same-BMU oracle"
              % (name, t, r["near_tie_rate"], r["mismatch_rate"], flips, r["worst_rel_gap"], r["bad"], err_ref, err_same))
        if flips:
            print("    flipped rows in fp64: GPU pick within %.2e of the true minimum, reference pick within %.2e; GPU closer on "
                  "%d of %d" % (reg["gpu_worst"], reg["ref_worst"], reg["gpu_better"], reg["rows"]))
        assert r["bad"] == 0, line
        # (mexican hat: num / den with den crossing zero amplifies the fp32 rounding of the apply -- the reference's own
        # epoch and an fp64 re-derivation of it differ by 1e-5 there)
        assert err_same <= (1e-4 if spec.neighborhood_function == "mexican_hat" else 1e-5), line
        if flips == 0:
            assert err_ref <= 1e-4, line
        else:
            # every flipped row sits inside the near-tie band (bad == 0 above) and the same-BMU check has just shown
            # that accumulate / neighbourhood / merge are exact to 1e-5, so err_ref is the effect of re-assigning
            # `flips` near-tie samples and nothing else; it is reported, not bounded (at C4's 2 samples per neuron one
            # re-assigned sample moves a neuron by half the distance between two samples)
            assert reg["gpu_worst"] <= r["eps"], line        # the GPU's pick is a true near-minimum in fp64
    _REPORT[name] = lines
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "parity_configs.json"), "w") as f:
            json.dump(_REPORT, f, indent=1)


@pytest.mark.parametrize("algo", ["simt", "tc", "tc16"])
def test_cosine_zero_rows_and_zero_neurons(algo):
    """distances.py:57: a zero sample or a zero neuron gives similarity nan -> 0, i.e. distance exactly 1.  A zero
    sample therefore ties on every neuron (argmin = neuron 0); a zero neuron wins exactly the samples whose
    similarity to every other neuron is negative."""
    from xpysom_dask_b200 import XPySom
    rng = np.random.RandomState(4)
    n, d, gx, gy = 3000, 40, 9, 8
    x = rng.randn(n, d).astype(np.float32)                  # both signs: negative similarities exist
    x[::97] = 0.0
    spec = so.SomSpec(gx=gx, gy=gy, dim=d, activation_distance="cosine", random_seed=2)
    w = rng.randn(gx, gy, d).astype(np.float32)
    w[2, 3] = 0.0
    w[0, 0] = 0.0
    som = XPySom(gx, gy, d, activation_distance="cosine", random_seed=2, algo=algo)
    som._weights = w.copy()
    bmu = som.predict(x)
    r = U.bmu_parity(spec, x, w, bmu)
    assert r["bad"] == 0, r
    assert (bmu[::97] == 0).all()                            # zero samples: every distance is 1, first index wins
    ref = so.bmu_flat(spec, x, w)
    zero_wins = np.isin(ref, [0, 2 * gy + 3])
    assert zero_wins.sum() > 0 and (np.isin(bmu, [0, 2 * gy + 3]) == zero_wins).mean() > 0.999
    # and a training epoch with those rows / neurons stays finite and matches the oracle's epoch
    som.train(x, 4, iter_beg=0, iter_end=1)
    w_ref = so.epoch(spec, x, w, 0, 4)
    assert np.isfinite(som._weights).all()
    assert U.codebook_rel_err(som._weights, w_ref) < 1e-3   # a handful of near-tie rows may flip on this random map


def test_chunked_host_upload_matches_device_resident():
    """Host arrays of >= 32 MB are uploaded in chunks on a copy stream and the FIRST epoch consumes each chunk as it
    lands (xpysom_dask_b200/xpysom.py: _upload_in_chunks).  Same BMUs and the same codebook as the device-resident
    path, for the first epoch and for the epochs after it."""
    from xpysom_dask_b200 import XPySom
    n, d, gx, gy = 200_003, 64, 16, 16                         # 51 MB, ragged against the 256-row chunk granularity
    data = U.blobs(n, d, seed=3)
    assert data.nbytes >= 32 << 20
    for n_ep in (1, 3):
        a = XPySom(gx, gy, d, random_seed=1)
        b = XPySom(gx, gy, d, random_seed=1)
        a.train(torch.from_numpy(data).cuda(), 5, iter_beg=0, iter_end=n_ep)
        b.train(data, 5, iter_beg=0, iter_end=n_ep)                       # numpy host array -> chunked upload
        tol = 1e-5 if n_ep == 1 else 2e-3         # free-running epochs amplify last-bit differences
        assert U.codebook_rel_err(b._weights, a._weights) < tol, n_ep
    pinned = torch.from_numpy(data).pin_memory()
    c = XPySom(gx, gy, d, random_seed=1)
    c.train(pinned, 5, iter_beg=0, iter_end=1)
    a1 = XPySom(gx, gy, d, random_seed=1)
    a1.train(torch.from_numpy(data).cuda(), 5, iter_beg=0, iter_end=1)
    assert U.codebook_rel_err(c._weights, a1._weights) < 1e-5
