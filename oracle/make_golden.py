"""Generate tests/golden/*.npz by running the REAL reference (/root/reference).

TEST INFRASTRUCTURE ONLY.  Run in the build container (the reference tree does
not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python oracle/make_golden.py

The reference is imported read-only, with xp=numpy (CuPy and Dask are absent, so
its import prints two warnings).  Everything written is plain arrays plus a JSON
case list, so the fixtures are usable without the reference.
"""
import contextlib
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden")
REF = os.environ.get("SOM_REFERENCE", "/root/reference")

sys.dont_write_bytecode = True
sys.path.insert(0, REF)
with contextlib.redirect_stdout(io.StringIO()):
    from xpysom_dask import XPySom                      # noqa: E402
    from xpysom_dask import distances as rd             # noqa: E402
    from xpysom_dask import neighborhoods as rn         # noqa: E402
    from xpysom_dask import decays as rdec              # noqa: E402


def blobs(n, d, seed, centres=16, spread=0.1):
    """Gaussian mixture used for parity data."""
    rng = np.random.RandomState(seed)
    c = rng.rand(centres, d)
    x = c[rng.randint(centres, size=n)] + spread * rng.randn(n, d)
    return x.astype(np.float32)


# ---------------------------------------------------------------- distances
def gen_distances():
    out = {}
    rng = np.random.RandomState(0)
    cases = []
    for n, m, l in [(2, 3, 5), (7, 11, 13), (33, 49, 4), (64, 100, 17), (5, 1, 1)]:
        x = rng.rand(n, l)
        w = rng.rand(m, l)
        cases.append((x, w))
    # a case with zero rows / zero neurons for the cosine nan_to_num rule
    x = rng.rand(6, 8); x[2] = 0
    w = rng.rand(9, 8); w[4] = 0
    cases.append((x, w))
    for ci, (x, w) in enumerate(cases):
        for dt in (np.float64, np.float32):
            xa, wa = x.astype(dt), w.astype(dt)
            tag = "c%d_%s" % (ci, np.dtype(dt).name)
            out[tag + "_x"] = xa
            out[tag + "_w"] = wa
            with np.errstate(all="ignore"):
                out[tag + "_euclidean"] = rd.euclidean_squared_distance_part(xa, wa, xp=np)
                out[tag + "_euclidean_sq"] = rd.euclidean_squared_distance(xa, wa, xp=np)
                out[tag + "_euclidean_sqrt"] = rd.euclidean_distance(xa, wa, xp=np)
                out[tag + "_cosine"] = rd.cosine_distance(xa, wa, xp=np)
                out[tag + "_manhattan"] = rd.manhattan_distance(xa, wa, xp=np)
                for p in (2, 3, 4):
                    out[tag + "_norm_p%d" % p] = rd.norm_p_power_distance(xa, wa, p=p, xp=np)
    out["n_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(OUT, "distances.npz"), **out)


# ------------------------------------------------------------ neighbourhoods
def gen_neighborhoods():
    out = {}
    manifest = []
    maps = [(5, 5), (9, 6), (6, 9), (7, 7)]
    sigmas = [1.0, 2.5, np.float64(1.7)]
    for gx, gy in maps:
        for topo in ("rectangular", "hexagonal"):
            for fn in ("gaussian", "mexican_hat", "bubble", "triangle"):
                if topo == "hexagonal" and fn == "triangle":
                    continue
                for compact in (False, True):
                    if fn == "bubble" and compact:
                        continue
                    for si, sigma in enumerate(sigmas):
                        with contextlib.redirect_stdout(io.StringIO()):
                            som = XPySom(gx, gy, 3, sigma=1.0, neighborhood_function=fn,
                                         topology=topo, compact_support=compact,
                                         std_coeff=0.5, random_seed=1, xp=np, n_parallel=100)
                        bi, bj = np.unravel_index(np.arange(gx * gy), (gx, gy))
                        try:
                            h = som.neighborhood((bi, bj), sigma, xp=np)
                            err = ""
                        except ValueError as e:      # mexican_hat_rect + compact on gx != gy
                            h = np.zeros(0)
                            err = "ValueError"
                        key = "h_%dx%d_%s_%s_c%d_s%d" % (gx, gy, topo[:3], fn, compact, si)
                        out[key] = h
                        manifest.append(dict(key=key, gx=gx, gy=gy, topology=topo, fn=fn,
                                             compact=compact, sigma=float(sigma),
                                             sigma_is_np=isinstance(sigma, np.floating),
                                             dtype=str(h.dtype), error=err))
    np.savez_compressed(os.path.join(OUT, "neighborhoods.npz"), **out)
    with open(os.path.join(OUT, "neighborhoods.json"), "w") as f:
        json.dump(manifest, f, indent=0)


# -------------------------------------------------------------------- decays
def gen_decays():
    rows = []
    for kind, fn in (("exponential", rdec.exponential_decay), ("asymptotic", rdec.asymptotic_decay),
                     ("linear", rdec.linear_decay)):
        for v0, vN in ((0.5, 0.01), (16.0, 1.0), (3.0, 0.0), (20.0, 1)):
            for T in (1, 2, 10, 37):
                for t in sorted({0, 1, T // 2, T - 1}):
                    if t < 0:
                        continue
                    v = fn(v0, vN, t, T)
                    rows.append(dict(kind=kind, v0=v0, vN=vN, t=t, T=T, value=float(v),
                                     is_np=isinstance(v, np.floating)))
    with open(os.path.join(OUT, "decays.json"), "w") as f:
        json.dump(rows, f, indent=0)


# -------------------------------------------------------------------- epochs
EPOCH_CASES = [
    # name, gx, gy, D, N, kwargs, T, epochs to teacher-force
    ("rect_gauss_euc", 8, 8, 16, 1500, dict(), 10, (0, 4, 9)),
    ("rect_gauss_euc_lin", 10, 7, 16, 1500, dict(decay_function="linear"), 8, (0, 3, 7)),
    ("rect_gauss_euc_asym", 6, 9, 5, 800, dict(decay_function="asymptotic"), 6, (0, 5)),
    ("rect_gauss_euc_compact", 8, 8, 12, 1000, dict(compact_support=True), 6, (0, 5)),
    ("rect_mex_euc", 8, 8, 16, 1500, dict(neighborhood_function="mexican_hat"), 10, (0, 9)),
    ("rect_mex_euc_compact", 7, 7, 9, 900, dict(neighborhood_function="mexican_hat", compact_support=True,
                                                 decay_function="linear"), 6, (0, 5)),
    ("rect_bubble_euc", 8, 8, 16, 1500, dict(neighborhood_function="bubble"), 10, (0, 9)),
    ("rect_tri_euc", 9, 6, 16, 1500, dict(neighborhood_function="triangle", decay_function="linear"), 10, (0, 9)),
    ("rect_tri_euc_compact", 9, 6, 8, 700, dict(neighborhood_function="triangle", compact_support=True), 5, (0, 4)),
    ("hex_gauss_euc", 8, 8, 16, 1500, dict(topology="hexagonal"), 10, (0, 9)),
    ("hex_gauss_euc_odd", 7, 9, 16, 1500, dict(topology="hexagonal", decay_function="linear"), 10, (0, 9)),
    ("hex_gauss_compact", 6, 8, 10, 900, dict(topology="hexagonal", compact_support=True), 6, (0, 5)),
    ("hex_mex_cos", 10, 10, 32, 2000, dict(topology="hexagonal", neighborhood_function="mexican_hat",
                                           activation_distance="cosine"), 10, (0, 9)),
    ("hex_mex_compact", 6, 7, 8, 800, dict(topology="hexagonal", neighborhood_function="mexican_hat",
                                           compact_support=True), 6, (0, 5)),
    ("hex_bubble_man", 8, 6, 12, 1200, dict(topology="hexagonal", neighborhood_function="bubble",
                                            activation_distance="manhattan"), 6, (0, 5)),
    ("rect_gauss_cos", 8, 8, 24, 1500, dict(activation_distance="cosine"), 10, (0, 9)),
    ("rect_gauss_man", 8, 8, 7, 1500, dict(activation_distance="manhattan"), 10, (0, 9)),
    ("rect_gauss_p3", 6, 6, 6, 600, dict(activation_distance="norm_p", activation_distance_kwargs={"p": 3}), 5, (0, 4)),
    ("rect_gauss_p4", 6, 6, 6, 600, dict(activation_distance="norm_p", activation_distance_kwargs={"p": 4}), 5, (0, 4)),
    ("rect_gauss_euc_D33", 12, 12, 33, 3000, dict(), 10, (0, 9)),
    ("rect_gauss_euc_D130", 16, 16, 130, 2000, dict(sigma=4.0), 10, (0, 9)),
    ("rect_gauss_euc_std1", 5, 5, 3, 400, dict(std_coeff=1.0, sigma=2.0, sigmaN=0.5, learning_rate=1.0), 4, (0, 3)),
]


def gen_epochs():
    out = {}
    manifest = []
    for ci, (name, gx, gy, D, N, kw, T, steps) in enumerate(EPOCH_CASES):
        data = blobs(N, D, seed=100 + ci)
        with contextlib.redirect_stdout(io.StringIO()):
            som = XPySom(gx, gy, D, random_seed=7 + ci, xp=np, n_parallel=512, **kw)
        w0 = som._weights.copy()
        out[name + "_data"] = data
        out[name + "_w_init"] = w0
        # free run to get realistic W_t at the teacher-forced epochs
        w_t = {}
        w = np.asarray(w0, dtype=np.float32)
        for t in range(T):
            if t in steps:
                w_t[t] = w.copy()
            som._weights = w
            with contextlib.redirect_stdout(io.StringIO()):
                som.train(data, T, iter_beg=t, iter_end=t + 1)
            w = som._weights
            if t in steps:
                out["%s_t%d_w_in" % (name, t)] = w_t[t]
                out["%s_t%d_w_out" % (name, t)] = w.copy()
                som._weights = w_t[t]
                win = som.winner(data)
                flat = np.array([i * gy + j for i, j in win], dtype=np.int32)
                out["%s_t%d_bmu" % (name, t)] = flat
                som._weights = w
        out[name + "_w_final"] = w.copy()
        manifest.append(dict(name=name, gx=gx, gy=gy, D=D, N=N, T=T, steps=list(steps),
                             seed=7 + ci, n_parallel=512,
                             kwargs={k: v for k, v in kw.items()}))
    np.savez_compressed(os.path.join(OUT, "epochs.npz"), **out)
    with open(os.path.join(OUT, "epochs.json"), "w") as f:
        json.dump(manifest, f, indent=0)


# ---------------------------------------------------------------------- iris
def gen_iris():
    """Config 1 of BASELINE.json: Iris 150x4, 7x7 rectangular, gaussian, euclidean."""
    raw = np.genfromtxt(os.path.join(REF, "examples", "iris.csv"), delimiter=",", usecols=(0, 1, 2, 3))
    data = raw.astype(np.float32)
    with contextlib.redirect_stdout(io.StringIO()):
        som = XPySom(7, 7, 4, sigma=3, learning_rate=0.5, neighborhood_function="gaussian",
                     random_seed=10, xp=np, n_parallel=4000)
    out = dict(data=data, w_init=som._weights.copy())
    T = 100
    w = np.asarray(som._weights, dtype=np.float32)
    for t in range(T):
        som._weights = w
        if t in (0, 50, 99):
            out["t%d_w_in" % t] = w.copy()
            win = som.winner(data)
            out["t%d_bmu" % t] = np.array([i * 7 + j for i, j in win], dtype=np.int32)
        som.train(data, T, iter_beg=t, iter_end=t + 1)
        w = som._weights
        if t in (0, 50, 99):
            out["t%d_w_out" % t] = w.copy()
    out["w_final"] = w.copy()
    out["qe_final"] = np.array(som.quantization_error(data))
    out["distance_map_final"] = som.distance_map()
    out["quantization_final"] = som.quantization(data)
    win = som.winner(data)
    out["winner_final"] = np.array(win, dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "iris.npz"), **out)


# ----------------------------------------------------------------------- api
def gen_api():
    """Known answers of the reference's own unittest fixture (tests.py:22-33 etc.)."""
    out = {}
    with contextlib.redirect_stdout(io.StringIO()):
        som = XPySom(5, 5, 1, std_coeff=1, xp=np)
    w = np.zeros((5, 5, 1)); w[2, 3] = 5.0; w[1, 1] = 2.0
    som._weights = w
    out["fake_w"] = w
    out["activate_5"] = som.activate(5.0)
    out["winner_5_2"] = np.array(som.winner([[5.0], [2.0]]))
    out["qe_5_2"] = np.array(som.quantization_error([[5], [2]]))
    out["qe_4_1"] = np.array(som.quantization_error([[4], [1]]))
    out["quant_4_2"] = som.quantization(np.array([[4], [2]]))
    d = np.arange(-5, 5).reshape(-1, 1)
    out["dfw_data"] = d
    out["dfw"] = som.distance_from_weights(d, None)
    w2 = w.copy(); w2[2, 4] = 6.0; w2[4, 4] = 15.0; w2[0, 0] = 14.0
    som._weights = w2
    out["topo_w"] = w2
    out["topo_5"] = np.array(som.topographic_error([[5]]))
    out["topo_15"] = np.array(som.topographic_error([[15]]))
    # seeded init + short training (tests.py:98-121)
    with contextlib.redirect_stdout(io.StringIO()):
        s1 = XPySom(5, 5, 2, sigma=1.0, learning_rate=0.5, random_seed=1, xp=np)
    out["seed1_w_init"] = s1._weights.copy()
    np.random.seed(1234)
    data = np.random.rand(100, 2)
    out["seed1_data"] = data
    with contextlib.redirect_stdout(io.StringIO()):
        s1.train(data, 10)
    out["seed1_w_trained"] = s1._weights.copy()
    out["seed1_qe"] = np.array(s1.quantization_error(data))
    out["seed1_distance_map"] = s1.distance_map()
    out["seed1_topo"] = np.array(s1.topographic_error(data))
    # hexagonal distance_map / topographic error
    with contextlib.redirect_stdout(io.StringIO()):
        s2 = XPySom(6, 5, 3, topology="hexagonal", random_seed=3, xp=np)
    out["hex_w"] = s2._weights.copy()
    out["hex_distance_map"] = s2.distance_map()
    dd = np.random.RandomState(5).rand(50, 3)
    out["hex_data"] = dd
    # topographic_error on a hexagonal map indexes the (gy, gx) meshgrids with
    # (i, j) (xpysom.py:742-743): IndexError unless the map is square.
    with contextlib.redirect_stdout(io.StringIO()):
        s2q = XPySom(6, 6, 3, topology="hexagonal", random_seed=4, xp=np)
    out["hexsq_w"] = s2q._weights.copy()
    out["hexsq_topo"] = np.array(s2q.topographic_error(dd))
    out["hex_qe"] = np.array(s2.quantization_error(dd))
    xx, yy = s2.get_euclidean_coordinates()
    out["hex_xx"] = xx
    out["hex_yy"] = yy
    # pca / distance_map known answers (tests.py:129-139)
    with contextlib.redirect_stdout(io.StringIO()):
        s3 = XPySom(2, 2, 2, xp=np)
    s3.pca_weights_init(np.array([[1., 0.], [0., 1.], [1., 0.], [0., 1.]]))
    out["pca_w"] = s3._weights.copy()
    np.savez_compressed(os.path.join(OUT, "api.npz"), **out)


# --------------------------------------------------------------- dask branch
def gen_blocks():
    """The Dask branch (xpysom.py:545-558) cannot run (no dask); its graph is
    `_update` per row block + Python sum + `_merge_updates`, which we can drive
    directly on the reference object."""
    out = {}
    data = blobs(3000, 16, seed=55)
    with contextlib.redirect_stdout(io.StringIO()):
        som = XPySom(8, 8, 16, random_seed=2, xp=np, n_parallel=100000)
    w = np.asarray(som._weights, dtype=np.float32)
    from xpysom_dask.decays import exponential_decay
    T, t = 10, 3
    eta = exponential_decay(0.5, 0.01, t, T)
    sig = exponential_decay(som._sigma, 1, t, T)
    som._sq_weights_gpu = np.power(w.reshape(-1, 16), 2).sum(axis=1, keepdims=True)
    for nb in (1, 2, 4, 8):
        rows = -(-len(data) // nb)
        parts = [som._update(data[s:s + rows], w, eta, sig) for s in range(0, len(data), rows)]
        num = sum(p[0] for p in parts)
        den = sum(p[1] for p in parts)
        out["w_out_nb%d" % nb] = som._merge_updates(w, num, den)
    out["data"] = data
    out["w_in"] = w
    np.savez_compressed(os.path.join(OUT, "blocks.npz"), **out)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    gen_distances()
    gen_neighborhoods()
    gen_decays()
    gen_epochs()
    gen_iris()
    gen_api()
    gen_blocks()
    tot = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT))
    print("golden fixtures written to", os.path.normpath(OUT), "total bytes", tot)
