"""Pin the CPU oracle (oracle/som_oracle.py) to the reference.

The fixtures in tests/golden were produced by the real reference
(oracle/make_golden.py).  On the same numpy build the oracle must reproduce
them exactly; across numpy/BLAS builds a few ulp of slack is allowed on
GEMM-backed quantities.  The scalar known answers come from the reference's
own test_distances.py:92-154.
"""
import json
import os

import numpy as np
import pytest

from oracle import som_oracle as so
import som_testutil as U

G = os.path.join(os.path.dirname(__file__), "golden")


def _load(name):
    return np.load(os.path.join(G, name), allow_pickle=False)


# ------------------------------------------------------------------ distances
DIST_FUNCS = {
    "euclidean": so.dist_euclidean_part,
    "euclidean_sq": so.dist_euclidean_sq,
    "euclidean_sqrt": so.dist_euclidean,
    "cosine": so.dist_cosine,
    "manhattan": so.dist_manhattan,
    "norm_p2": lambda x, w: so.dist_norm_p(x, w, 2),
    "norm_p3": lambda x, w: so.dist_norm_p(x, w, 3),
    "norm_p4": lambda x, w: so.dist_norm_p(x, w, 4),
}


@pytest.mark.parametrize("fn", sorted(DIST_FUNCS))
def test_distances_match_reference_outputs(fn):
    g = _load("distances.npz")
    for ci in range(int(g["n_cases"])):
        for dt in ("float64", "float32"):
            tag = "c%d_%s" % (ci, dt)
            with np.errstate(all="ignore"):
                got = DIST_FUNCS[fn](g[tag + "_x"], g[tag + "_w"])
            want = g[tag + "_" + fn]
            assert got.dtype == want.dtype and got.shape == want.shape
            tol = 1e-12 if dt == "float64" else 2e-6
            np.testing.assert_allclose(got, want, rtol=tol, atol=tol)


SCALAR = {
    "euclidean": lambda x, w: -2 * np.dot(x, w) + np.dot(w, w),
    "euclidean_sq": lambda x, w: np.sum((x - w) ** 2),
    "euclidean_sqrt": lambda x, w: np.linalg.norm(x - w),
    "cosine": lambda x, w: 1 - np.nan_to_num(np.dot(x, w) / (np.linalg.norm(x) * np.linalg.norm(w))),
    "manhattan": lambda x, w: np.linalg.norm(x - w, ord=1),
    "norm_p2": lambda x, w: np.sum((x - w) ** 2),
    "norm_p3": lambda x, w: np.sum(np.abs(x - w) ** 3),
    "norm_p4": lambda x, w: np.sum(np.abs(x - w) ** 4),
    "chebyshev": lambda x, w: np.max(np.abs(x - w)),
}


def _binary_inputs():
    """All binary vectors of length 1..3 against each other plus seeded fuzz
    (the input family of the reference's test_distances.py:37-88)."""
    cases = []
    for l in range(1, 4):
        vecs = [[(v >> b) & 1 for b in range(l)] for v in range(2 ** l)]
        cases.append((np.array(vecs, float), np.array(vecs, float)))
        cases.append((np.array(vecs[:1], float), np.array(vecs, float)))
        cases.append((np.array(vecs, float), np.array(vecs[::2], float)))
    rng = np.random.RandomState(0)
    for n in (2, 7):
        for m in (3, 11):
            for l in (5, 13):
                cases.append((rng.rand(n, l), rng.rand(m, l)))
    return cases


@pytest.mark.parametrize("fn", sorted(SCALAR))
def test_distances_scalar_known_answers(fn):
    f = dict(DIST_FUNCS, chebyshev=so.dist_chebyshev)[fn]
    for x, w in _binary_inputs():
        with np.errstate(all="ignore"):
            got = f(x, w)
            want = np.array([[SCALAR[fn](a, b) for b in w] for a in x])
        np.testing.assert_almost_equal(got, want)      # 7 decimals, as the reference's test


# ------------------------------------------------------------- neighbourhoods
def test_neighborhoods_match_reference_tables():
    g = _load("neighborhoods.npz")
    cases = json.load(open(os.path.join(G, "neighborhoods.json")))
    assert len(cases) > 100
    for c in cases:
        spec_kw = dict(gx=c["gx"], gy=c["gy"], dim=3, sigma=1.0, neighborhood_function=c["fn"],
                       topology=c["topology"], compact_support=c["compact"], std_coeff=0.5)
        spec = so.SomSpec(**spec_kw)
        bi, bj = np.unravel_index(np.arange(spec.K), (spec.gx, spec.gy))
        sigma = np.float64(c["sigma"]) if c["sigma_is_np"] else c["sigma"]
        if c["error"]:
            with pytest.raises(ValueError):
                so.neighborhood(spec, bi, bj, sigma)
            continue
        h = so.neighborhood(spec, bi, bj, sigma)
        want = g[c["key"]]
        assert str(h.dtype) == c["dtype"], c["key"]
        np.testing.assert_array_equal(h, want, err_msg=c["key"])


def test_hex_coordinate_rule():
    g = _load("api.npz")
    xx, yy = so.hex_coords(6, 5, "hexagonal")
    np.testing.assert_array_equal(xx.T, g["hex_xx"])
    np.testing.assert_array_equal(yy.T, g["hex_yy"])
    # closed form: neuron (i, j) at (i - 0.5*[(gy-1-j) even], j)
    for i in range(6):
        for j in range(5):
            assert xx[j, i] == i - 0.5 * ((5 - 1 - j) % 2 == 0)


# --------------------------------------------------------------------- decays
def test_decays_match_reference():
    rows = json.load(open(os.path.join(G, "decays.json")))
    for r in rows:
        v = so.decay_value(r["kind"], r["v0"], r["vN"], r["t"], r["T"])
        assert float(v) == pytest.approx(r["value"], rel=1e-15, abs=0)
        assert isinstance(v, np.floating) == r["is_np"]


# --------------------------------------------------------------------- epochs
def _spec_from(c):
    kw = dict(c["kwargs"])
    adk = kw.pop("activation_distance_kwargs", {})
    return so.SomSpec(gx=c["gx"], gy=c["gy"], dim=c["D"], n_parallel=c["n_parallel"],
                      random_seed=c["seed"], p=adk.get("p", 2), **kw)


EPOCHS = json.load(open(os.path.join(G, "epochs.json")))


@pytest.mark.parametrize("case", EPOCHS, ids=[c["name"] for c in EPOCHS])
def test_epoch_teacher_forced_matches_reference(case):
    g = _load("epochs.npz")
    spec = _spec_from(case)
    name = case["name"]
    np.testing.assert_array_equal(so.init_weights(spec), g[name + "_w_init"])
    data = g[name + "_data"]
    for t in case["steps"]:
        w_in = g["%s_t%d_w_in" % (name, t)]
        with np.errstate(all="ignore"):
            w_out, bmu = so.epoch(spec, data, w_in, t, case["T"], return_bmu=True)
        want = g["%s_t%d_w_out" % (name, t)]
        assert w_out.dtype == want.dtype
        np.testing.assert_array_equal(bmu, g["%s_t%d_bmu" % (name, t)])
        np.testing.assert_allclose(w_out, want, rtol=2e-6, atol=1e-7)


@pytest.mark.parametrize("case", EPOCHS[:4], ids=[c["name"] for c in EPOCHS[:4]])
def test_free_running_train_matches_reference(case):
    g = _load("epochs.npz")
    spec = _spec_from(case)
    name = case["name"]
    with np.errstate(all="ignore"):
        w = so.train(spec, g[name + "_data"], g[name + "_w_init"], case["T"])
    np.testing.assert_allclose(w, g[name + "_w_final"], rtol=1e-4, atol=1e-5)


def test_iris_config1():
    g = _load("iris.npz")
    spec = so.SomSpec(gx=7, gy=7, dim=4, sigma=3, learning_rate=0.5, random_seed=10, n_parallel=4000)
    np.testing.assert_array_equal(so.init_weights(spec), g["w_init"])
    for t in (0, 50, 99):
        w_out, bmu = so.epoch(spec, g["data"], g["t%d_w_in" % t], t, 100, return_bmu=True)
        np.testing.assert_array_equal(bmu, g["t%d_bmu" % t])
        np.testing.assert_allclose(w_out, g["t%d_w_out" % t], rtol=2e-6, atol=1e-7)
    w = g["w_final"]
    assert so.quantization_error(spec, g["data"], w) == pytest.approx(float(g["qe_final"]), rel=1e-6)
    np.testing.assert_allclose(so.distance_map(spec, w), g["distance_map_final"], rtol=1e-12)
    np.testing.assert_array_equal(so.quantization(spec, g["data"], w), g["quantization_final"])
    np.testing.assert_array_equal(np.array(so.winner(spec, g["data"], w)), g["winner_final"])


def test_dask_graph_shape_blocks():
    """xpysom.py:545-558: per-block _update + Python sum + merge."""
    g = _load("blocks.npz")
    spec = so.SomSpec(gx=8, gy=8, dim=16, random_seed=2, n_parallel=100000)
    for nb in (1, 2, 4, 8):
        w = so.block_partials(spec, g["data"], g["w_in"], 3, 10, nb)
        np.testing.assert_allclose(w, g["w_out_nb%d" % nb], rtol=2e-6, atol=1e-7)


def test_identity_num_is_Ht_S():
    """num = H^T S, den = H^T c   — the algebra the CUDA path uses."""
    g = _load("epochs.npz")
    for case in EPOCHS:
        spec = _spec_from(case)
        name = case["name"]
        t = case["steps"][-1]
        data, w_in = g[name + "_data"], g["%s_t%d_w_in" % (name, t)]
        bmu = g["%s_t%d_bmu" % (name, t)]
        sig = so.decay_value(spec.decay_function, spec.sigma, spec.sigmaN, t, case["T"])
        S, c = so.sums_by_bmu(bmu, data, spec.K)
        H = so.neighborhood_table(spec, float(sig)).astype(np.float64)
        num, den = H.T @ S, H.T @ c
        with np.errstate(all="ignore"):
            w = np.where(den[:, None] != 0, num / den[:, None], w_in.reshape(spec.K, -1))
        want = g["%s_t%d_w_out" % (name, t)].reshape(spec.K, -1)
        scale = np.abs(want).max()
        # mexican hat denominators cross zero: compare where |den| is not tiny
        ok = np.abs(den) > 1e-3 * np.abs(den).max()
        assert np.abs(w - want)[ok].max() / scale < 1e-4, name


# ------------------------------------------------------------------------ api
def test_api_known_answers():
    g = _load("api.npz")
    spec = so.SomSpec(gx=5, gy=5, dim=1, std_coeff=1)
    w = g["fake_w"]
    assert so.dist_euclidean_part(np.array([[5.0]]), w.reshape(-1, 1)).argmin() == 13   # tests.py:66
    assert so.quantization_error(spec, [[5], [2]], w) == 0.0                           # tests.py:78
    assert so.quantization_error(spec, [[4], [1]], w) == 1.0                           # tests.py:79
    q = so.quantization(spec, np.array([[4], [2]]), w)
    assert q[0] == 5.0 and q[1] == 2.0                                                 # tests.py:93-96
    np.testing.assert_array_equal(np.array(so.winner(spec, [[5.0], [2.0]], w)), g["winner_5_2"])
    d = g["dfw_data"]
    np.testing.assert_array_equal(so.dist_euclidean(d, w.reshape(-1, 1)), g["dfw"])
    assert so.topographic_error(spec, [[5]], g["topo_w"]) == 0.0                       # tests.py:89
    assert so.topographic_error(spec, [[15]], g["topo_w"]) == 1.0                      # tests.py:90
    s1 = so.SomSpec(gx=5, gy=5, dim=2, sigma=1.0, learning_rate=0.5, random_seed=1, n_parallel=4000)
    np.testing.assert_array_equal(so.init_weights(s1), g["seed1_w_init"])
    wt = so.train(s1, g["seed1_data"], g["seed1_w_init"], 10)
    np.testing.assert_allclose(wt, g["seed1_w_trained"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(so.distance_map(s1, g["seed1_w_trained"]), g["seed1_distance_map"], rtol=1e-12)
    assert so.topographic_error(s1, g["seed1_data"], g["seed1_w_trained"]) == float(g["seed1_topo"])
    s2 = so.SomSpec(gx=6, gy=5, dim=3, topology="hexagonal", random_seed=3)
    np.testing.assert_allclose(so.distance_map(s2, g["hex_w"]), g["hex_distance_map"], rtol=1e-12)
    assert so.quantization_error(s2, g["hex_data"], g["hex_w"]) == pytest.approx(float(g["hex_qe"]), rel=1e-6)
    s2q = so.SomSpec(gx=6, gy=6, dim=3, topology="hexagonal", random_seed=4)
    assert so.topographic_error(s2q, g["hex_data"], g["hexsq_w"]) == float(g["hexsq_topo"])


def test_unknown_names_raise():
    with pytest.raises(ValueError):
        so.SomSpec(gx=5, gy=5, dim=1, neighborhood_function="boooom")
    with pytest.raises(ValueError):
        so.activation_distance("ridethewave", np.zeros((1, 1)), np.zeros((1, 1)))
    with pytest.raises(ValueError):
        so.SomSpec(gx=5, gy=5, dim=1, topology="hexagonal", neighborhood_function="triangle")


def test_separable_factors_reproduce_the_full_table():
    """som_testutil.separable_factors (used by the config-scale GPU parity tests to avoid a K x K table at K = 10^4):
    h((bi,bj),(i,j)) = A[bi,i] B[bj,j] / h00 for the product-form neighbourhoods of neighborhoods.py:33,112,130."""
    for fn, compact in [("gaussian", False), ("gaussian", True), ("bubble", False), ("triangle", False), ("triangle", True)]:
        for sig in (3.3, np.float64(1.7), 0.9):
            spec = so.SomSpec(gx=9, gy=6, dim=3, neighborhood_function=fn, compact_support=compact)
            H = so.neighborhood_table(spec, sig).astype(np.float64).reshape(9, 6, 9, 6)
            A, B, h00 = U.separable_factors(spec, sig)
            H2 = np.einsum("ai,bj->abij", A, B) / h00
            assert np.abs(H - H2).max() <= 1e-7 * np.abs(H).max(), (fn, compact, sig)
