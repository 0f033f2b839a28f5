"""Parity of the CUDA path (through the C ABI) against the oracle and the golden fixtures.

BMU indices: bit-exact wherever the reference's own top-2 gap exceeds EPS*scale
(tests/som_testutil.py); the near-tie rate is printed.  Codebooks: teacher-forced, one
epoch from the same W_t, max-abs error / max-abs value <= 1e-4 (fp32).
"""
import ctypes

import numpy as np
import pytest

from oracle import som_oracle as so
import som_testutil as U

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

CODEBOOK_TOL = 1e-4


@pytest.fixture(scope="module")
def eng():
    from xpysom_dask_b200.engine import CudaEngine
    return CudaEngine("cuda:0")


def _gpu_bmu(eng, x, w, dist, p, algo):
    from xpysom_dask_b200 import _lib
    n, d = x.shape
    ld = (d + 3) // 4 * 4                      # TMA needs a 16-byte row stride (the host class pads the same way)
    buf = torch.full((n, ld), 7.0, dtype=torch.float32, device="cuda")     # garbage in the padding on purpose
    buf[:, :d] = torch.from_numpy(x).cuda()
    xd = buf[:, :d]
    wd = torch.from_numpy(np.ascontiguousarray(w.reshape(-1, w.shape[-1]), dtype=np.float32)).cuda()
    ws = eng.workspace(0, wd.shape[0], wd.shape[1])
    eng.prepare_codebook(wd, _lib.DIST[dist], p, ws)
    xs = eng.prepare_samples(xd)[0] if algo in ("tc16", "auto") else None
    bmu = eng.bmu(xd, wd, _lib.DIST[dist], p, _lib.ALGO[algo], ws, xscale=xs)
    torch.cuda.synchronize()
    return bmu.cpu().numpy()


BMU_SHAPES = [
    # n, d, gx, gy
    (150, 4, 7, 7),          # Iris shape (config 1)
    (3000, 16, 8, 8),
    (5000, 64, 32, 32),      # config-2 map
    (2500, 100, 20, 15),     # K, D not multiples of the tile sizes
    (1000, 784, 25, 20),     # config-4 feature count
    (4097, 128, 50, 50),     # config-5 map, ragged n
    (1, 16, 5, 5),           # single sample
]


@pytest.mark.parametrize("n,d,gx,gy", BMU_SHAPES)
@pytest.mark.parametrize("dist", ["euclidean", "cosine"])
@pytest.mark.parametrize("algo", ["simt", "tc", "tc16"])
def test_bmu_contraction_distances(eng, n, d, gx, gy, dist, algo):
    x = U.blobs(n, d, seed=n + d)
    spec = so.SomSpec(gx=gx, gy=gy, dim=d, activation_distance=dist, random_seed=d)
    w = so.init_weights(spec).astype(np.float32) * 0.5 + 0.5 * U.uniform(gx * gy, d, 3).reshape(gx, gy, d)
    bmu = _gpu_bmu(eng, x, w, dist, 2.0, algo)
    r = U.bmu_parity(spec, x, w, bmu)
    print("\n[bmu %s/%s n=%d d=%d K=%d] near-tie rate %.2e, raw mismatch %.2e, worst mismatching rel gap %.2e"
          % (dist, algo, n, d, gx * gy, r["near_tie_rate"], r["mismatch_rate"], r["worst_rel_gap"]))
    assert r["bad"] == 0, r


# feature counts around the edges of the tensor-core K blocks: the TF32 kernel folds the bias into columns
# d..d+2 of the last 32-feature block when it has 3 spare columns (29: yes, crossing a 16-byte chunk; 30: no;
# 33 / 61: fold in a second block), both kernels skip the 8 / 16-column MMA steps that hold only padding
@pytest.mark.parametrize("d", [5, 8, 13, 17, 29, 30, 31, 33, 47, 61, 62, 65, 129])
@pytest.mark.parametrize("dist", ["euclidean", "cosine"])
@pytest.mark.parametrize("algo", ["tc", "tc16"])
def test_bmu_tensor_core_block_edges(eng, d, dist, algo):
    n, gx, gy = 1777, 19, 15                     # K = 285: one full and one ragged 256-neuron tile
    x = U.blobs(n, d, seed=3 * d)
    spec = so.SomSpec(gx=gx, gy=gy, dim=d, activation_distance=dist, random_seed=d)
    w = so.init_weights(spec).astype(np.float32) * 0.5 + 0.5 * U.uniform(gx * gy, d, 7).reshape(gx, gy, d)
    bmu = _gpu_bmu(eng, x, w, dist, 2.0, algo)
    r = U.bmu_parity(spec, x, w, bmu)
    assert r["bad"] == 0, r


@pytest.mark.parametrize("n,d,gx,gy", [(2000, 7, 8, 8), (1500, 33, 12, 10), (3000, 64, 16, 16)])
@pytest.mark.parametrize("dist,p", [("manhattan", 1.0), ("chebyshev", 0.0), ("norm_p", 3.0), ("norm_p", 2.0),
                                    ("norm_p", 4.0), ("norm_p", 2.5), ("euclidean", 2.0), ("cosine", 2.0)])
def test_bmu_simt_all_distances(eng, n, d, gx, gy, dist, p):
    x = U.blobs(n, d, seed=11 * n + d)
    spec = so.SomSpec(gx=gx, gy=gy, dim=d, activation_distance=dist, p=p, random_seed=1)
    w = U.uniform(gx * gy, d, 5).reshape(gx, gy, d)
    if dist == "norm_p" and p % 2 == 0:
        # the reference's even-p path is a binomial expansion with cancellation
        # (distances.py:77-96); compare against the direct definition instead
        spec_direct = so.SomSpec(gx=gx, gy=gy, dim=d, activation_distance="norm_p_no_opt", p=p, random_seed=1)
        spec_direct.activation_distance = "norm_p_no_opt"
        ref_spec = spec_direct
    else:
        ref_spec = spec
    bmu = _gpu_bmu(eng, x, w, dist, p, "simt")
    r = U.bmu_parity(ref_spec, x.astype(np.float64) if dist == "norm_p" and p == 2.5 else x, w if p != 2.5 else w.astype(np.float64), bmu)
    print("\n[bmu %s p=%g n=%d d=%d] near-tie %.2e mismatch %.2e" % (dist, p, n, d, r["near_tie_rate"], r["mismatch_rate"]))
    assert r["bad"] == 0, r


@pytest.mark.parametrize("algo", ["tc", "tc16"])
def test_bmu_wide_dynamic_range(eng, algo):
    """Rows and neurons whose magnitudes span 12 orders of magnitude, plus zero rows: the per-row /
    per-neuron power-of-two scaling must keep the fp16 split as accurate as the TF32 one."""
    rng = np.random.RandomState(3)
    n, d, gx, gy = 4000, 96, 16, 16
    x = U.blobs(n, d, seed=5)
    x *= (10.0 ** rng.uniform(-6, 6, size=(n, 1))).astype(np.float32)      # per-row scale
    x[::500] = 0.0                                                           # zero samples
    x[:, ::7] *= 1e-4                                                        # tiny features
    spec = so.SomSpec(gx=gx, gy=gy, dim=d, random_seed=1)
    w = U.uniform(gx * gy, d, 2).reshape(gx, gy, d) * (10.0 ** rng.uniform(-6, 6, size=(gx, gy, 1))).astype(np.float32)
    w[3, 3] = 0.0
    bmu = _gpu_bmu(eng, x, w, "euclidean", 2.0, algo)
    r = U.bmu_parity(spec, x, w, bmu)
    print("\n[bmu %s wide range] near-tie %.2e mismatch %.2e worst %.2e" % (algo, r["near_tie_rate"], r["mismatch_rate"], r["worst_rel_gap"]))
    assert r["bad"] == 0, r


def test_unaligned_rows_fall_back_to_simt(eng):
    """AUTO must pick the SIMT kernel when TMA cannot address X (row stride not a multiple of 16 B)."""
    from xpysom_dask_b200 import _lib
    x = U.blobs(1000, 30, seed=1)
    spec = so.SomSpec(gx=9, gy=9, dim=30, random_seed=2)
    w = so.init_weights(spec).astype(np.float32)
    bmu = _gpu_bmu(eng, x, w, "euclidean", 2.0, "auto")
    assert U.bmu_parity(spec, x, w, bmu)["bad"] == 0
    xd = torch.from_numpy(x).cuda()            # contiguous (1000, 30): rows not 16-byte aligned
    wd = torch.from_numpy(w.reshape(81, 30)).cuda()
    ws = eng.workspace(0, 81, 30)
    eng.prepare_codebook(wd, 0, 2.0, ws)
    with pytest.raises(_lib.SomB200Error):
        eng.bmu(xd, wd, 0, 2.0, _lib.ALGO["tc"], ws)


def _exact_sums(eng, xd, bd, K):
    """S (K, D) and c (K) through the C ABI: column maxima -> scales -> exact accumulate -> finalize."""
    n, d = xd.shape
    _, colmax = eng.prepare_samples(xd, want_scale=False)
    qscale, qinv = eng.accum_scales(colmax, d, n)
    acc = eng.accumulator(K, d)
    eng.accumulate(xd, bd, K, qscale, acc)
    S, c = eng.empty(K, d), eng.empty(K)
    raw = acc.clone()
    eng.accum_finalize(acc, qinv, K, d, S, c)
    torch.cuda.synchronize()
    assert int(acc.abs().max()) == 0                       # finalize leaves the accumulator cleared
    # small accumulators are kept in several copies that the finalize sums (include/som_b200.h): compare the sum
    reps = eng.lib.som_b200_accum_replicas(K, d)
    return S.cpu().numpy(), c.cpu().numpy(), raw.view(reps, -1).sum(0).cpu().numpy()


def test_accumulate_matches_oracle_sums(eng):
    """Exact fixed-point accumulation (csrc/accumulate.cuh): equal to the fp64 segmented sums to fp32 rounding, exact
    counts, bit-identical from launch to launch, vector and scalar paths, rows longer than one 128-column piece."""
    for n, d, K in [(5000, 16, 64), (3001, 30, 50), (20000, 64, 1024), (777, 5, 13000), (1500, 300, 40), (900, 131, 7)]:
        x = U.blobs(n, d, seed=n)
        x[:, ::3] *= 1e-3                                   # columns of different magnitudes: per-column scales
        bmu = np.random.RandomState(n).randint(K, size=n).astype(np.int32)
        S_ref, c_ref = so.sums_by_bmu(bmu, x, K)
        ld = (d + 3) // 4 * 4
        buf = torch.zeros((n, ld), device="cuda")
        buf[:, :d] = torch.from_numpy(x).cuda()
        for xd in (buf[:, :d], torch.from_numpy(x).cuda()):          # 16-byte aligned rows / packed rows
            bd = torch.from_numpy(bmu).cuda()
            S, c, raw = _exact_sums(eng, xd, bd, K)
            np.testing.assert_array_equal(c, c_ref)
            assert np.abs(S - S_ref).max() <= 1e-7 * np.abs(S_ref).max() + 1e-30, (n, d, K)
            col = np.abs(S_ref).max(axis=0)
            assert (np.abs(S - S_ref).max(axis=0) <= 2e-7 * col + 1e-30).all(), (n, d, K)   # per column, not only overall
            S2, c2, raw2 = _exact_sums(eng, xd, bd, K)
            np.testing.assert_array_equal(raw, raw2)               # integers: the same bits every time
            np.testing.assert_array_equal(S, S2)
        # the order of the rows does not matter either
        perm = np.random.RandomState(1).permutation(n)
        S3, _, raw3 = _exact_sums(eng, torch.from_numpy(x[perm]).cuda(), torch.from_numpy(bmu[perm]).cuda(), K)
        np.testing.assert_array_equal(raw, raw3)


def _apply_neigh(eng, case, sigma, S, c, eta=0.37):
    from xpysom_dask_b200 import _lib
    gx, gy = case["gx"], case["gy"]
    K, d = gx * gy, S.shape[1]
    Sd, cd = torch.from_numpy(S).cuda(), torch.from_numpy(c).cuda()
    num, den = eng.empty(K, d), eng.empty(K)
    eng.neigh_apply(Sd, cd, gx, gy, d, _lib.TOPO[case["topology"]], _lib.NEIGH[case["fn"]], sigma, eta, 0.5,
                    case["compact"], num, den, eng.neigh_tables(gx, gy))
    torch.cuda.synchronize()
    return num.cpu().numpy(), den.cpu().numpy()


def test_neigh_apply_matches_reference_tables(eng):
    """num = eta H^T S, den = eta H^T c with H taken from the reference's own neighbourhood
    functions (golden tables for every BMU of 5x5, 9x6, 6x9, 7x7 maps, both topologies,
    compact support on/off, Python-float and numpy-float sigma)."""
    from xpysom_dask_b200 import _lib
    g = U.load("neighborhoods.npz")
    cases = U.load_json("neighborhoods.json")
    rng = np.random.RandomState(0)
    worst = 0.0
    for case in cases:
        K = case["gx"] * case["gy"]
        S = rng.randn(K, 9).astype(np.float32)
        c = rng.randint(1, 50, size=K).astype(np.float32)
        if case["error"]:
            with pytest.raises(_lib.SomB200Error):
                _apply_neigh(eng, case, case["sigma"], S, c)
            continue
        H = g[case["key"]].reshape(K, K).astype(np.float64)
        num, den = _apply_neigh(eng, case, case["sigma"], S, c)
        num_ref, den_ref = 0.37 * H.T @ S.astype(np.float64), 0.37 * H.T @ c.astype(np.float64)
        e1 = np.abs(num - num_ref).max() / max(np.abs(num_ref).max(), 1e-30)
        e2 = np.abs(den - den_ref).max() / max(np.abs(den_ref).max(), 1e-30)
        worst = max(worst, e1, e2)
        assert e1 < 5e-6 and e2 < 5e-6, (case["key"], e1, e2)
    print("\n[neigh_apply] %d reference tables, worst rel err %.2e" % (len(cases), worst))


@pytest.mark.parametrize("fn,compact", [("gaussian", False), ("gaussian", True), ("bubble", False), ("triangle", False),
                                        ("triangle", True), ("mexican_hat", False)])
@pytest.mark.parametrize("topology", ["rectangular", "hexagonal"])
def test_neigh_apply_large_map_separable_path(eng, fn, compact, topology):
    """Maps of >= 1024 neurons: rectangular product-form neighbourhoods of >= 4096 neurons take the two-pass separable path,
    everything else the direct kernel; both must agree with the oracle's H^T S."""
    from xpysom_dask_b200 import _lib
    if topology == "hexagonal" and fn == "triangle":
        pytest.skip("rejected by the reference")
    gx, gy, d = 70, 60, 12
    K = gx * gy
    rng = np.random.RandomState(1)
    S = rng.randn(K, d).astype(np.float32)
    c = rng.randint(0, 30, size=K).astype(np.float32)
    S[c == 0] = 0
    spec = so.SomSpec(gx=gx, gy=gy, dim=d, sigma=1.0, neighborhood_function=fn, topology=topology, compact_support=compact)
    for sigma in (7.3, 1.0):
        H = so.neighborhood_table(spec, sigma).astype(np.float64)
        Sd, cd = torch.from_numpy(S).cuda(), torch.from_numpy(c).cuda()
        num, den = eng.empty(K, d), eng.empty(K)
        eng.neigh_apply(Sd, cd, gx, gy, d, _lib.TOPO[topology], _lib.NEIGH[fn], sigma, 0.5, 0.5, compact, num, den,
                        eng.neigh_tables(gx, gy, d))
        torch.cuda.synchronize()
        num_ref, den_ref = 0.5 * H.T @ S.astype(np.float64), 0.5 * H.T @ c.astype(np.float64)
        e1 = np.abs(num.cpu().numpy() - num_ref).max() / np.abs(num_ref).max()
        e2 = np.abs(den.cpu().numpy() - den_ref).max() / np.abs(den_ref).max()
        assert e1 < 1e-5 and e2 < 1e-5, (fn, compact, topology, sigma, e1, e2)


@pytest.mark.parametrize("fn,compact,topology", [("mexican_hat", False, "hexagonal"), ("mexican_hat", True, "hexagonal"),
                                                 ("gaussian", True, "hexagonal"), ("bubble", False, "rectangular")])
def test_neigh_apply_wide_tile_variant(eng, fn, compact, topology):
    """More than 64 features on maps of >= 512 neurons run the 128 x 128-tile instantiation of the direct kernel
    (8 x 8 outputs per thread); ragged K and D, empty BMUs, sliced reduction (atomics) included."""
    from xpysom_dask_b200 import _lib
    gx, gy, d = 27, 23, 100                      # K = 621: 4 full + 1 ragged 128-neuron tile, d = 100 of 128
    K = gx * gy
    rng = np.random.RandomState(5)
    S = rng.randn(K, d).astype(np.float32)
    c = rng.randint(0, 9, size=K).astype(np.float32)
    S[c == 0] = 0
    spec = so.SomSpec(gx=gx, gy=gy, dim=d, sigma=1.0, neighborhood_function=fn, topology=topology, compact_support=compact)
    for sigma in (5.1, 0.9):
        H = so.neighborhood_table(spec, sigma).astype(np.float64)
        Sd, cd = torch.from_numpy(S).cuda(), torch.from_numpy(c).cuda()
        num, den = eng.empty(K, d), eng.empty(K)
        eng.neigh_apply(Sd, cd, gx, gy, d, _lib.TOPO[topology], _lib.NEIGH[fn], sigma, 0.5, 0.5, compact, num, den,
                        eng.neigh_tables(gx, gy, d))
        torch.cuda.synchronize()
        num_ref, den_ref = 0.5 * H.T @ S.astype(np.float64), 0.5 * H.T @ c.astype(np.float64)
        e1 = np.abs(num.cpu().numpy() - num_ref).max() / np.abs(num_ref).max()
        e2 = np.abs(den.cpu().numpy() - den_ref).max() / np.abs(den_ref).max()
        assert e1 < 1e-5 and e2 < 1e-5, (fn, compact, topology, sigma, e1, e2)


@pytest.mark.parametrize("gx,gy,d,fn,topology,dist", [
    (32, 32, 64, "gaussian", "rectangular", "euclidean"),       # config-2 map: fused kernel, 64x64 tiles, sliced
    (27, 23, 100, "mexican_hat", "hexagonal", "cosine"),        # wide features: the entry issues the separate launches
    (27, 23, 50, "mexican_hat", "hexagonal", "cosine"),         # fused, hexagonal mexican hat, ragged K and D
    (7, 7, 4, "bubble", "rectangular", "manhattan"),            # tiny map, SIMT distance (no operand copies)
    (40, 40, 16, "triangle", "rectangular", "euclidean"),       # config-3 map
    (70, 60, 12, "gaussian", "rectangular", "euclidean"),       # separable path: the entry falls back to separate launches
])
def test_epoch_tail_matches_separate_launches(eng, gx, gy, d, fn, topology, dist):
    """som_b200_epoch_tail (one cooperative kernel on small maps) == accum_finalize + neigh_apply + merge +
    prepare_codebook: same codebook (bit for bit: no atomics anywhere), a workspace that finds the same BMUs, a
    cleared accumulator."""
    from xpysom_dask_b200 import _lib
    K = gx * gy
    rng = np.random.RandomState(gx + d)
    W0 = rng.rand(K, d).astype(np.float32)
    X = torch.from_numpy(rng.rand(3000, (d + 3) // 4 * 4).astype(np.float32)).cuda()[:, :d]
    b0 = torch.from_numpy(rng.randint(0, max(1, K // 2), size=3000).astype(np.int32)).cuda()   # half the units stay empty
    dk, tk, nk = _lib.DIST[dist], _lib.TOPO[topology], _lib.NEIGH[fn]
    _, colmax = eng.prepare_samples(X, want_scale=False)
    qscale, qinv = eng.accum_scales(colmax, d, 3000)
    res = []
    for fused in (False, True):
        w = torch.from_numpy(W0).cuda()
        acc = eng.accumulator(K, d)
        eng.accumulate(X, b0, K, qscale, acc)
        Sd, cd, num, den = eng.empty(K, d), eng.empty(K), eng.empty(K, d), eng.empty(K)
        ws, tables = eng.workspace(0, K, d), eng.neigh_tables(gx, gy, d)
        eng.prepare_codebook(w, dk, 2.0, ws)          # as in training: the tail follows a BMU search on a prepared workspace
        if fused:
            eng.epoch_tail(acc, qinv, Sd, cd, w, gx, gy, d, tk, nk, 2.3, 0.4, 0.5, False, dk, 2.0, num, den, tables, ws)
        else:
            eng.accum_finalize(acc, qinv, K, d, Sd, cd)
            eng.neigh_apply(Sd, cd, gx, gy, d, tk, nk, 2.3, 0.4, 0.5, False, num, den, tables)
            eng.merge(w, num, den)
            eng.prepare_codebook(w, dk, 2.0, ws)
        xs = eng.prepare_samples(X)[0] if dist in ("euclidean", "cosine") else None
        bmu = eng.bmu(X, w, dk, 2.0, _lib.ALGO["auto"], ws, xscale=xs)
        torch.cuda.synchronize()
        res.append((w.cpu().numpy(), bmu.cpu().numpy(), int(acc.abs().max())))
    (wa, ba, _), (wb, bb, amax) = res
    assert amax == 0                         # the accumulator is handed back cleared
    # the two routes slice the reduction over BMUs differently (fixed order each): last-bit differences only
    assert U.codebook_rel_err(wb, wa) < 2e-6
    assert (ba != bb).mean() < 2e-3
    # and each route is bit-reproducible: the same call again lands on the same codebook
    w = torch.from_numpy(W0).cuda()
    acc = eng.accumulator(K, d)
    eng.accumulate(X, b0, K, qscale, acc)
    eng.prepare_codebook(w, dk, 2.0, ws)
    eng.epoch_tail(acc, qinv, Sd, cd, w, gx, gy, d, tk, nk, 2.3, 0.4, 0.5, False, dk, 2.0, num, den, tables, ws)
    torch.cuda.synchronize()
    np.testing.assert_array_equal(w.cpu().numpy(), wb)


def test_neigh_apply_skips_empty_bmus(eng):
    case = dict(gx=6, gy=5, topology="rectangular", fn="gaussian", compact=False)
    S = np.zeros((30, 4), np.float32); c = np.zeros(30, np.float32)
    S[7] = [1, 2, 3, 4]; c[7] = 2
    num, den = _apply_neigh(eng, case, 1.3, S, c, eta=1.0)
    spec = so.SomSpec(gx=6, gy=5, dim=4)
    H = so.neighborhood_table(spec, 1.3)
    np.testing.assert_allclose(num, H.T @ S, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(den, H.T @ c, rtol=1e-5, atol=1e-7)


EPOCHS = U.load_json("epochs.json")


@pytest.mark.parametrize("algo", ["simt", "tc", "tc16", "auto"])
@pytest.mark.parametrize("case", EPOCHS, ids=[c["name"] for c in EPOCHS])
def test_epoch_teacher_forced_vs_reference(case, algo):
    """W_t from the reference -> one epoch on the GPU -> compare with the reference's W_{t+1}."""
    from xpysom_dask_b200 import XPySom
    g = U.load("epochs.npz")
    name = case["name"]
    spec = U.spec_from_case(case)
    dist = case["kwargs"].get("activation_distance", "euclidean")
    if algo in ("tc", "tc16") and dist not in ("euclidean", "cosine"):
        pytest.skip("tensor-core kernels: contraction distances only")
    data = g[name + "_data"]
    som = XPySom(case["gx"], case["gy"], case["D"], random_seed=case["seed"], algo=algo, **case["kwargs"])
    np.testing.assert_array_equal(som._weights, g[name + "_w_init"])      # bit-compatible init (xpysom.py:189-190)
    for t in case["steps"]:
        w_in = g["%s_t%d_w_in" % (name, t)]
        som._weights = w_in.copy()
        bmu = som.predict(data)
        r = U.bmu_parity(spec, data, w_in, bmu)
        assert r["bad"] == 0, (name, t, r)
        som.train(data, case["T"], iter_beg=t, iter_end=t + 1)
        assert som._weights.dtype == np.float32 and som._weights.shape == w_in.shape
        want = g["%s_t%d_w_out" % (name, t)]
        if "mex" in name:
            # the mexican hat's denominators cross zero: compare where |den| is not tiny
            sig = so.decay_value(spec.decay_function, spec.sigma, spec.sigmaN, t, case["T"])
            H = so.neighborhood_table(spec, float(sig)).astype(np.float64)
            _, c = so.sums_by_bmu(g["%s_t%d_bmu" % (name, t)], data, spec.K)
            den = H.T @ c
            ok = (np.abs(den) > 1e-3 * np.abs(den).max()).reshape(case["gx"], case["gy"])
            err = np.abs(som._weights - want)[ok].max() / np.abs(want).max()
        else:
            err = U.codebook_rel_err(som._weights, want)
        assert err <= CODEBOOK_TOL, (name, t, err, r)


def test_iris_config1_free_run():
    """Config 1 of BASELINE.json end to end: 100 epochs free-running on Iris, compared through the
    aggregate the reference's own tests use (quantization error), plus teacher-forced epochs."""
    from xpysom_dask_b200 import XPySom
    g = U.load("iris.npz")
    data = g["data"]
    som = XPySom(7, 7, 4, sigma=3, learning_rate=0.5, neighborhood_function="gaussian", random_seed=10)
    np.testing.assert_array_equal(som._weights, g["w_init"])
    for t in (0, 50, 99):
        som._weights = g["t%d_w_in" % t].copy()
        som.train(data, 100, iter_beg=t, iter_end=t + 1)
        assert U.codebook_rel_err(som._weights, g["t%d_w_out" % t]) <= CODEBOOK_TOL
    som = XPySom(7, 7, 4, sigma=3, learning_rate=0.5, neighborhood_function="gaussian", random_seed=10)
    som.train(data, 100)
    qe = som.quantization_error(data)
    print("\n[iris] QE gpu %.6f  reference %.6f  codebook rel err after 100 free epochs %.2e"
          % (qe, float(g["qe_final"]), U.codebook_rel_err(som._weights, g["w_final"])))
    assert qe == pytest.approx(float(g["qe_final"]), rel=2e-2)
    som._weights = g["w_final"].copy()
    assert som.quantization_error(data) == pytest.approx(float(g["qe_final"]), rel=1e-5)
    np.testing.assert_allclose(som.distance_map(), g["distance_map_final"], rtol=1e-5, atol=1e-6)
    q = som.quantization(data)
    assert (np.abs(q - g["quantization_final"]).max(axis=1) == 0).mean() > 0.98
    win = np.array(som.winner(data))
    assert (win == g["winner_final"]).all(axis=1).mean() > 0.98


def test_api_known_answers_from_reference_tests():
    """tests.py:22-33,49-96 of the reference, on the GPU backend."""
    from xpysom_dask_b200 import XPySom
    g = U.load("api.npz")
    som = XPySom(5, 5, 1, std_coeff=1)
    for i in range(5):
        for j in range(5):
            np.testing.assert_almost_equal(1.0, np.linalg.norm(som._weights[i, j]))
    som._weights = g["fake_w"].copy()
    assert som.winner(np.array([5.0])) == (2, 3)
    assert som.winner(np.array([[5.0], [2.0]])) == [(2, 3), (1, 1)]
    wm = som.win_map([[5.0], [2.0]])
    assert wm[(2, 3)][0] == [5.0] and wm[(1, 1)][0] == [2.0]
    lm = som.labels_map([[5.0], [2.0]], ['a', 'b'])
    assert lm[(2, 3)]['a'] == 1 and lm[(1, 1)]['b'] == 1
    with pytest.raises(ValueError):
        som.labels_map([[5.0]], ['a', 'b'])
    resp = som.activation_response([[5.0], [2.0]])
    assert resp[2, 3] == 1 and resp[1, 1] == 1
    assert som.quantization_error([[5], [2]]) == 0.0
    assert som.quantization_error([[4], [1]]) == 1.0
    q = som.quantization(np.array([[4], [2]]))
    assert q[0] == 5.0 and q[1] == 2.0
    som2 = XPySom(2, 2, 2, random_seed=1)
    som2._weights = np.array([[[1., 0.], [0., 1.]], [[1., 0.], [0., 1.]]])
    np.testing.assert_array_equal(som2.distance_map(), np.array([[1., 1.], [1., 1.]]))
    s1 = XPySom(5, 5, 2, sigma=1.0, learning_rate=0.5, random_seed=1)
    np.testing.assert_array_equal(s1._weights, g["seed1_w_init"])
    q1 = s1.quantization_error(g["seed1_data"])
    s1.train(g["seed1_data"], 10)
    assert q1 > s1.quantization_error(g["seed1_data"])                      # tests.py:111-121
    s1b = XPySom(5, 5, 2, sigma=1.0, learning_rate=0.5, random_seed=1)
    s1b.train_random(g["seed1_data"], 10)
    np.testing.assert_array_almost_equal(s1._weights, s1b._weights, decimal=4)   # tests.py:98-109
    s1._weights = g["seed1_w_trained"].copy()
    np.testing.assert_allclose(s1.distance_map(), g["seed1_distance_map"], rtol=1e-5)
    hexs = XPySom(6, 5, 3, topology="hexagonal", random_seed=3)
    np.testing.assert_array_equal(hexs._weights, g["hex_w"])
    np.testing.assert_allclose(hexs.distance_map(), g["hex_distance_map"], rtol=1e-5)
    assert hexs.quantization_error(g["hex_data"]) == pytest.approx(float(g["hex_qe"]), rel=1e-5)


def test_cuda_graph_replay_matches_eager_launches():
    """use_cuda_graph=True replays one captured graph per epoch with sigma / eta read from a device-side
    schedule; it must land where the kernel-by-kernel path lands.  The comparison is made fully
    deterministic: small-integer samples make the per-BMU sums exact in fp32 whatever the order of the
    accumulation (they are exact anyway), so the two runs cannot drift apart chaotically the way free-running runs
    do."""
    from xpysom_dask_b200 import XPySom
    rng = np.random.RandomState(21)
    centres = rng.randint(0, 8, size=(24, 32))
    data = (centres[rng.randint(24, size=8000)] + rng.randint(0, 2, size=(8000, 32))).astype(np.float32)
    for kw in (dict(), dict(topology="hexagonal", neighborhood_function="mexican_hat", decay_function="linear")):
        a = XPySom(8, 8, 32, sigma=2.0, random_seed=4, use_cuda_graph=False, **kw)
        b = XPySom(8, 8, 32, sigma=2.0, random_seed=4, use_cuda_graph=True, **kw)
        a.train(data, 6)
        b.train(data, 6)
        assert U.codebook_rel_err(b._weights, a._weights) < 1e-6
        # schedule bookkeeping: resuming mid-schedule through the graph path
        a2 = XPySom(8, 8, 32, sigma=2.0, random_seed=4, use_cuda_graph=True, **kw)
        a2.train(data, 6, iter_beg=0, iter_end=3)
        a2.train(data, 6, iter_beg=3, iter_end=6)
        assert U.codebook_rel_err(a2._weights, a._weights) < 1e-6


def test_row_scales_cached_per_device_tensor_and_refreshed_on_in_place_change():
    """The per-row power-of-two scales are computed once per uploaded / caller-owned device tensor, not per
    train() call; an in-place change of the tensor (torch's version counter) must recompute them."""
    from xpysom_dask_b200 import XPySom
    x = torch.rand(6000, 48, device="cuda")
    som = XPySom(6, 6, 48, sigma=1.5, random_seed=1)
    eng = som._get_engine()
    calls = []
    orig = eng.prepare_samples
    eng.prepare_samples = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
    som.train(x, 4, iter_beg=0, iter_end=2)
    som.train(x, 4, iter_beg=2, iter_end=3)
    assert len(calls) == 1
    x.mul_(1024.0)                                   # same storage, new contents: stale scales would overflow fp16
    w_before = som._weights.copy()
    som.train(x, 4, iter_beg=3, iter_end=4)
    assert len(calls) == 2
    ref = XPySom(6, 6, 48, sigma=1.5, random_seed=1)
    ref._weights = w_before.copy()
    ref.train(x.clone(), 4, iter_beg=3, iter_end=4)
    assert U.codebook_rel_err(som._weights, ref._weights) < 1e-5
    som.train(x.cpu().numpy(), 4, iter_beg=3, iter_end=4)      # host input: a fresh upload every call
    assert len(calls) == 3


def test_train_on_empty_shard_leaves_codebook_untouched():
    """No rows: no BMU search, zero sums, den == 0 everywhere -> _merge_updates keeps W (xpysom.py:451-455)."""
    from xpysom_dask_b200 import XPySom
    som = XPySom(6, 5, 8, random_seed=3)
    w0 = som._weights.astype(np.float32).copy()
    som.train(np.zeros((0, 8), dtype=np.float32), 3)
    np.testing.assert_array_equal(som._weights, w0)


def test_activate_distance_from_weights_topographic_error():
    """tests.py:66-90 of the reference and its own outputs on seeded maps (golden api.npz)."""
    from xpysom_dask_b200 import XPySom
    g = U.load("api.npz")
    som = XPySom(5, 5, 1, std_coeff=1)
    som._weights = g["fake_w"].copy()
    act = som.activate(np.array([5.0]))
    assert act.argmin() == 13                                                   # tests.py:66-67
    np.testing.assert_allclose(act.ravel(), np.asarray(g["activate_5"]).ravel(), rtol=1e-6)
    d = g["dfw_data"]
    dist = som.distance_from_weights(d, None)
    w = som._weights.reshape(-1, 1)
    for i in range(len(d)):
        for j in range(len(w)):
            assert dist[i][j] == np.linalg.norm(d[i] - w[j])                    # exact, tests.py:69-75
    som._weights = g["topo_w"].copy()
    assert som.topographic_error([[5]]) == 0.0                                  # tests.py:89
    assert som.topographic_error([[15]]) == 1.0                                 # tests.py:90
    s1 = XPySom(5, 5, 2, sigma=1.0, learning_rate=0.5, random_seed=1)
    s1._weights = g["seed1_w_trained"].copy()
    assert s1.topographic_error(g["seed1_data"]) == pytest.approx(float(g["seed1_topo"]), abs=0.011)
    hq = XPySom(6, 6, 3, topology="hexagonal", random_seed=4)
    np.testing.assert_array_equal(hq._weights, g["hexsq_w"])
    assert hq.topographic_error(g["hex_data"]) == pytest.approx(float(g["hexsq_topo"]), abs=0.021)
    with pytest.raises(IndexError):                      # the reference's (i, j) indexing of (gy, gx) grids
        XPySom(6, 5, 3, topology="hexagonal", random_seed=3).topographic_error(g["hex_data"])
    # every activation distance against the oracle's matrix
    x = U.blobs(300, 20, seed=2)
    for dist_name, p in (("euclidean", 2), ("cosine", 2), ("manhattan", 1), ("chebyshev", 0), ("norm_p", 3)):
        sm = XPySom(7, 6, 20, activation_distance=dist_name, activation_distance_kwargs={"p": p}, random_seed=3)
        ref_name = "norm_p_no_opt" if dist_name == "norm_p" else dist_name
        want = so.activation_distance(ref_name, x, np.asarray(sm._weights, np.float32).reshape(42, 20), None, p)
        got = sm.activate(x)
        assert np.abs(got - want).max() / np.abs(want).max() < 2e-6, dist_name


def test_train_host_c_abi_matches_class():
    """The whole-job C entry with HOST buffers (what the reference would bind) == the Python class.
    Teacher-forced, one epoch per call from the same W_t: free-running epochs amplify any BMU flip
    chaotically, for the reference itself as well."""
    from xpysom_dask_b200 import XPySom, _lib
    from xpysom_dask_b200.decays import exponential_decay
    lib = _lib.load()
    n, d, gx, gy, T = 6000, 24, 10, 9, 4
    data = U.blobs(n, d, seed=9)
    for algo in ("simt", "auto", "tc16"):
        som = XPySom(gx, gy, d, random_seed=5, algo=algo)
        for t in range(T):
            w_t = np.ascontiguousarray(som._weights, dtype=np.float32)
            som.train(data, T, iter_beg=t, iter_end=t + 1)
            cfg = _lib.TrainConfig(gx, gy, d, 0, 0, 0, _lib.ALGO[algo], 0, 2.0, 0.5)
            sig = (ctypes.c_double * 1)(float(exponential_decay(min(gx, gy) / 2, 1, t, T)))
            eta = (ctypes.c_double * 1)(float(exponential_decay(0.5, 0.01, t, T)))
            w = w_t.copy()
            rc = lib.som_b200_train_host(data.ctypes.data_as(ctypes.c_void_p), n, d, w.ctypes.data_as(ctypes.c_void_p),
                                         ctypes.byref(cfg), sig, eta, 1)
            _lib.check(rc, "som_b200_train_host")
            assert U.codebook_rel_err(w, som._weights) < 1e-5, (algo, t)
    # several epochs in one call run and stay finite
    w = w_t.copy()
    sig = (ctypes.c_double * 3)(2.0, 1.5, 1.0)
    eta = (ctypes.c_double * 3)(0.5, 0.3, 0.1)
    _lib.check(lib.som_b200_train_host(data.ctypes.data_as(ctypes.c_void_p), n, d, w.ctypes.data_as(ctypes.c_void_p),
                                       ctypes.byref(cfg), sig, eta, 3), "som_b200_train_host")
    assert np.isfinite(w).all()


def test_errors_cross_the_abi_as_codes(eng):
    from xpysom_dask_b200 import _lib
    lib = eng.lib
    assert lib.som_b200_merge(None, None, None, 4, 4, None) == -1
    assert b"merge" in lib.som_b200_last_error()
    w = eng.zeros(16, 8)
    small = torch.empty(16, dtype=torch.uint8, device="cuda")
    assert lib.som_b200_prepare_codebook(eng._p(w), 16, 8, 0, 2.0, eng._p(small), 16, None) == -3
    assert lib.som_b200_prepare_codebook(eng._p(w), 16, 8, 99, 2.0, eng._p(small), 16, None) == -1


# ---- long rows: one tensor-core pass + refinement (csrc/bmu_filter.cuh) ---------------------------------------------
def _filter_bmus(eng, xd, wd, seed_bmu=None):
    n, d = xd.shape
    K = wd.shape[0]
    assert eng.filter_eligible(xd, K, 0)
    fws = eng.filter_workspace(xd, K)
    bmu = torch.full((n,), -1, dtype=torch.int32, device="cuda") if seed_bmu is None else seed_bmu.clone()
    eng.bmu_filter(xd, wd, fws, bmu)
    ovf, evals = eng.filter_stats(fws, n, K, d)
    return bmu, ovf, evals


@pytest.mark.parametrize("n,d,gx,gy", [(4500, 256, 32, 32), (5000, 784, 40, 50), (4099, 300, 33, 32), (6000, 1024, 32, 33)])
def test_bmu_filter_matches_oracle(eng, n, d, gx, gy):
    """The filter + refine path picks, on every row, a neuron inside the stated band of the reference's minimum; run
    unseeded (no previous BMUs), seeded with its own result, and seeded with arbitrary neurons (any neuron's score is a
    valid upper bound): the same BMUs every time."""
    x = U.blobs(n, d, seed=n + d)
    spec = so.SomSpec(gx=gx, gy=gy, dim=d, random_seed=d)
    w = so.init_weights(spec).astype(np.float32) * 0.5 + 0.5 * U.uniform(gx * gy, d, 3).reshape(gx, gy, d)
    xd, wd = torch.from_numpy(x).cuda(), torch.from_numpy(w.reshape(gx * gy, d)).cuda()
    bmu, ovf, evals = _filter_bmus(eng, xd, wd)
    assert ovf == 0 and int((bmu < 0).sum()) == 0
    r = U.bmu_parity(spec, x, w, bmu.cpu().numpy())
    print("\n[filter n=%d d=%d K=%d] near-tie rate %.2e, raw mismatch %.2e, %.2f candidates re-scored per row"
          % (n, d, gx * gy, r["near_tie_rate"], r["mismatch_rate"], evals / n))
    assert r["bad"] == 0, r
    again, _, evals2 = _filter_bmus(eng, xd, wd, seed_bmu=bmu)
    assert torch.equal(again, bmu) and evals2 <= evals
    junk = torch.randint(0, gx * gy, (n,), dtype=torch.int32, device="cuda")
    assert torch.equal(_filter_bmus(eng, xd, wd, seed_bmu=junk)[0], bmu)


def test_bmu_filter_overflow_rows_are_marked_and_training_fixes_them_up(eng):
    """A very smooth map (every neuron within 1e-4 of a plane through the data mean) has more live candidates than a
    list holds: those rows come back as -1 from the C entry, XPySom re-does them with the three-pass kernel (or skips the
    filter for the epoch), and the epoch still matches the oracle's update from its own BMUs."""
    from xpysom_dask_b200 import XPySom
    n, d, gx, gy = 6000, 512, 40, 40
    rng = np.random.RandomState(5)
    x = rng.random_sample((n, d)).astype(np.float32)
    u = np.linspace(-1, 1, gx, dtype=np.float32)
    a, b = rng.randn(d).astype(np.float32) * 0.02, rng.randn(d).astype(np.float32) * 0.02
    w = (0.5 + u[:, None, None] * a + u[None, :, None] * b + 1e-4 * rng.randn(gx, gy, d)).astype(np.float32)
    xd, wd = torch.from_numpy(x).cuda(), torch.from_numpy(w.reshape(gx * gy, d)).cuda()
    bmu, ovf, evals = _filter_bmus(eng, xd, wd)
    assert ovf == int((bmu < 0).sum())
    spec = so.SomSpec(gx=gx, gy=gy, dim=d, random_seed=0)
    keep = (bmu >= 0).cpu().numpy()
    r = U.bmu_parity(spec, x[keep], w, bmu.cpu().numpy()[keep])
    assert r["bad"] == 0, r
    for max_cand in (1e9, 12.0):                   # filter forced on (fix-up of the overflowed rows) / normal policy
        som = XPySom(gx, gy, d, random_seed=0)
        som._FILTER_MAX_CANDIDATES, som._FILTER_MAX_OVERFLOW = max_cand, (1.0 if max_cand > 1e6 else 0.005)
        som._weights = w.copy()
        som.train(xd, 10, iter_beg=4, iter_end=5)
        used = som._bmu_last.cpu().numpy()
        assert (used >= 0).all()
        assert U.bmu_parity(spec, x, w, used)["bad"] == 0
        S, c = so.sums_by_bmu(used, x, spec.K)
        sig = float(so.decay_value(spec.decay_function, spec.sigma, spec.sigmaN, 4, 10))
        H = so.neighborhood_table(spec, sig).astype(np.float64)
        num, den = H.T @ S, H.T @ c
        w_same = np.where(den[:, None] != 0, num / den[:, None], w.reshape(spec.K, -1)).reshape(w.shape)
        assert U.codebook_rel_err(som._weights, w_same) < 1e-5, max_cand
        assert som.stats.get("filter_epochs", 0) > 0        # (no probe on so few rows: the filter runs, then judges itself)
