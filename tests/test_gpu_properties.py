"""Size-independent properties of the CUDA path at the full map / feature sizes of BASELINE.json's configs
(where the numpy oracle would take minutes to hours): conservation of the per-BMU sums, agreement of the
three BMU kernels up to provable near-ties, determinism, exact invariance under power-of-two rescaling,
and the fixed point of the update."""
import numpy as np
import pytest

import som_testutil as U

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

CONFIGS = [
    # name, rows, D, gx, gy, dist      (rows: full for c2, a slice of the named row count for the others)
    ("c2", 1_000_000, 64, 32, 32, "euclidean"),
    ("c3", 2_000_000, 16, 40, 40, "euclidean"),
    ("c4", 60_000, 784, 100, 100, "euclidean"),
    ("c5", 500_000, 128, 50, 50, "cosine"),
]


@pytest.fixture(scope="module")
def eng():
    from xpysom_dask_b200.engine import CudaEngine
    return CudaEngine("cuda:0")


def _setup(eng, n, d, K, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.rand(n, d, generator=g, device="cuda")
    w = torch.rand(K, d, generator=g, device="cuda")
    return x, w


def _bmu(eng, x, w, dist, algo):
    from xpysom_dask_b200 import _lib
    ws = eng.workspace(0, w.shape[0], w.shape[1])
    eng.prepare_codebook(w, _lib.DIST[dist], 2.0, ws)
    xs = eng.prepare_samples(x)[0] if algo in ("tc16", "auto") else None
    return eng.bmu(x, w, _lib.DIST[dist], 2.0, _lib.ALGO[algo], ws, xscale=xs)


def _score64(x, w, idx, dist):
    """fp64 score of row r against neuron idx[r] (what each kernel minimises, in exact arithmetic)."""
    xr, wr = x.double(), w.double()[idx.long()]
    if dist == "euclidean":
        return (wr * wr).sum(1) - 2 * (xr * wr).sum(1)
    return -(xr * wr).sum(1) / wr.norm(dim=1)


@pytest.mark.parametrize("name,n,d,gx,gy,dist", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_kernels_agree_up_to_near_ties(eng, name, n, d, gx, gy, dist):
    x, w = _setup(eng, n, d, gx * gy)
    ref = _bmu(eng, x, w, dist, "simt")
    for algo in ("tc", "tc16"):
        got = _bmu(eng, x, w, dist, algo)
        diff = (got != ref).nonzero().flatten()
        rate = diff.numel() / n
        if diff.numel():
            sa = _score64(x[diff], w, got[diff], dist)
            sb = _score64(x[diff], w, ref[diff], dist)
            # euclidean: |x|^2 + |d| (the epsilon of the parity criterion); cosine: the kernels minimise
            # -x.w/|w|, which is |x| times the cosine similarity the criterion is stated on
            scale = (x[diff].double() ** 2).sum(1) + sb.abs() if dist == "euclidean" else x[diff].double().norm(dim=1)
            worst = ((sa - sb).abs() / scale).max().item()
        else:
            worst = 0.0
        print("\n[%s %s vs simt] disagreement %.2e of rows, worst fp64 relative score gap among them %.2e"
              % (name, algo, rate, worst))
        assert worst < 2e-6, (name, algo, worst)          # only provable near-ties may differ
        assert rate < 0.2


@pytest.mark.parametrize("name,n,d,gx,gy,dist", CONFIGS, ids=[c[0] for c in CONFIGS])
def test_fused_sums_conserve_the_samples(eng, name, n, d, gx, gy, dist):
    from xpysom_dask_b200 import _lib
    K = gx * gy
    x, w = _setup(eng, n, d, K, seed=1)
    ws = eng.workspace(0, K, d)
    eng.prepare_codebook(w, _lib.DIST[dist], 2.0, ws)
    xs, colmax = eng.prepare_samples(x)
    qscale, qinv = eng.accum_scales(colmax, d, n)
    acc, bmu = eng.accumulator(K, d), eng.empty(n, dtype=torch.int32)
    eng.epoch_accumulate(x, w, _lib.DIST[dist], 2.0, _lib.ALGO["auto"], qscale, acc, ws, bmu_out=bmu, xscale=xs)
    raw = acc.clone()
    S, c = eng.empty(K, d), eng.empty(K)
    eng.accum_finalize(acc, qinv, K, d, S, c)
    torch.cuda.synchronize()
    assert c.sum(dtype=torch.float64).item() == n                                   # every row counted once
    hist = torch.bincount(bmu.long(), minlength=K).float()
    assert torch.equal(hist, c)                                                      # counts match the BMUs
    # S is the segmented sum of the rows by BMU, to fp32 rounding of the final value (exact fixed-point sums)
    S_ref = torch.zeros(K, d, dtype=torch.float64, device="cuda").index_add_(0, bmu.long(), x.double())
    assert ((S.double() - S_ref).abs().max() / S_ref.abs().max()).item() < 1e-7
    col = x.sum(0, dtype=torch.float64)
    assert ((S.sum(0, dtype=torch.float64) - col).abs() / col.abs()).max().item() < 1e-6   # sum_b S[b] = sum_n x_n
    # the fused kernel twice: the same BMUs and the same accumulator, bit for bit
    eng.epoch_accumulate(x, w, _lib.DIST[dist], 2.0, _lib.ALGO["auto"], qscale, acc, ws, bmu_out=bmu, xscale=xs)
    torch.cuda.synchronize()
    assert torch.equal(acc, raw)


@pytest.mark.parametrize("algo", ["simt", "tc", "tc16"])
def test_bmu_is_deterministic_and_scale_invariant(eng, algo):
    n, d, K = 300_000, 64, 1024
    x, w = _setup(eng, n, d, K, seed=2)
    a = _bmu(eng, x, w, "euclidean", algo)
    b = _bmu(eng, x, w, "euclidean", algo)
    assert torch.equal(a, b)                                   # same launch twice: same BMUs
    for k in (-20, 7, 31):                                     # exact power-of-two rescaling of both operands
        s = 2.0 ** k
        assert torch.equal(_bmu(eng, x * s, w * s, "euclidean", algo), a), (algo, k)
    perm = torch.randperm(n, device="cuda")                    # BMUs do not depend on the row order / tiling
    assert torch.equal(_bmu(eng, x[perm].contiguous(), w, "euclidean", algo), a[perm])


def test_update_fixed_point_and_empty_neurons():
    """All samples identical -> every neuron with a non-zero neighbourhood weight lands exactly on that sample;
    bubble with sigma < 1 leaves the neurons nobody wins untouched (den == 0 branch of xpysom.py:451-455)."""
    from xpysom_dask_b200 import XPySom
    v = np.linspace(0.1, 0.9, 64, dtype=np.float32)
    data = np.tile(v, (5000, 1))
    som = XPySom(32, 32, 64, random_seed=0)
    som.train(data, 3, iter_beg=0, iter_end=1)
    np.testing.assert_allclose(som._weights, np.broadcast_to(v, som._weights.shape), rtol=1e-6)   # exact sums: no running-sum bias
    som = XPySom(16, 16, 64, sigma=0.9, sigmaN=0.9, neighborhood_function="bubble", random_seed=1)
    w0 = np.asarray(som._weights, dtype=np.float32).copy()
    som.train(data, 2, iter_beg=0, iter_end=1)
    moved = np.abs(som._weights - w0).max(axis=2) > 0
    assert moved.sum() == 1                                    # only the single BMU changes
    np.testing.assert_allclose(som._weights[moved][0], v, rtol=1e-6)


def test_full_size_config2_epochs_reduce_quantization_error():
    """Config 2 at full size, free-running: the aggregate the reference's own tests use (tests.py:111-121)."""
    from xpysom_dask_b200 import XPySom
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.rand(1_000_000, 64, generator=g, device="cuda")
    som = XPySom(32, 32, 64, random_seed=0)
    q0 = som.quantization_error(x)
    som.train(x, 10)
    q1 = som.quantization_error(x)
    assert np.isfinite(som._weights).all() and q1 < 0.8 * q0
    um = som.distance_map()
    assert um.shape == (32, 32) and um.max() == 1.0 and um.min() >= 0.0


def test_training_is_bit_reproducible_and_streaming_matches_resident():
    """Exact accumulation + slice-ordered neighbourhood sums: two free-running trainings give the same bits; samples
    streamed from host memory through two block buffers (out-of-core path) land where the resident path lands."""
    from xpysom_dask_b200 import XPySom
    data = U.blobs(60_000, 48, seed=11)
    x = torch.from_numpy(data).cuda()
    runs = []
    for _ in range(2):
        som = XPySom(24, 20, 48, random_seed=5)
        som.train(x, 6)
        runs.append(som._weights.copy())
    np.testing.assert_array_equal(runs[0], runs[1])
    # hot BMUs: most rows on a handful of units (atomics on a few lines) -- still the same bits twice
    hot = np.repeat(data[:40], 1500, axis=0)
    a = XPySom(24, 20, 48, random_seed=5); a.train(hot, 3)
    b = XPySom(24, 20, 48, random_seed=5); b.train(hot, 3)
    np.testing.assert_array_equal(a._weights, b._weights)
    # out-of-core: 60k x 48 floats = 11.5 MB of samples through blocks of ~1 MB, teacher-forced one epoch and three
    res = XPySom(24, 20, 48, random_seed=5)
    res.train(data, 6, iter_beg=0, iter_end=1)
    ooc = XPySom(24, 20, 48, random_seed=5, max_resident_bytes=2 << 20)
    ooc.train(data, 6, iter_beg=0, iter_end=1)
    assert ooc.stats["streamed_blocks"] >= 10
    assert U.codebook_rel_err(ooc._weights, res._weights) < 1e-6
    ooc.train(data, 6, iter_beg=1, iter_end=3)
    res.train(data, 6, iter_beg=1, iter_end=3)
    assert U.codebook_rel_err(ooc._weights, res._weights) < 2e-3       # free-running: last-bit differences amplify
    assert ooc.quantization_error(data) == pytest.approx(res.quantization_error(data), rel=1e-3)


@pytest.mark.gpu
def test_sharded_training_equals_one_gpu_bitwise_on_two_gpus():
    """Row G of SURVEY 8a on hardware (needs >= 2 GPUs in the box, skipped otherwise): tools/peer_check.py --quick trains
    five map types on 2 ranks with the exchange fused into the epoch tail (accumulators in NVLink peer memory) and with
    the NCCL all-reduce of the integer accumulator, and requires both to equal ONE GPU training on all the rows, bit for bit."""
    import os
    import subprocess
    import sys
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(root, "tools", "peer_check.py"), "--quick"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=500, cwd=root)
    assert out.returncode == 0 and "PEER CHECK PASSED" in out.stdout, out.stdout[-3000:] + out.stderr[-2000:]
