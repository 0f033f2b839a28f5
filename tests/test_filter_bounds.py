"""The interval arithmetic of the one-pass filter path (csrc/bmu_filter.cuh), restated in numpy and checked on the CPU.

The kernel treats the fp16 "hi x hi" contraction as an interval [s^ - E, s^ + E] around the centred score
s_k = |w_k - mu|^2 - 2 (x - mu).(w_k - mu), with E_k = A_r nwh_k + B_r nwl_k + 2^-22 bias_k.  These tests rebuild the
operands exactly as filter_samples_kernel / filter_codebook_kernel do (centring in fp32, power-of-two scaling, fp16
rounding) and verify, in fp64, that the interval always contains the score and that the survivors of the filter always
contain the true BMU -- for uniform data, blobs, a smooth young map and badly offset data.  (The CUDA kernels themselves
are checked against the oracle by tests/test_gpu_parity.py::test_bmu_filter_*.)"""
import numpy as np
import pytest

import som_testutil as U


def _pow2_scale(amax):
    amax = np.where(amax > 0, amax, 1.0)
    return 2.0 ** (14 - np.floor(np.log2(amax)))


def _split_rows(a32):
    """fp32 rows -> (hi, lo, scale): hi = fp16(a 2^s), lo = the exact fp32 residual, as the kernels compute them."""
    sc = _pow2_scale(np.abs(a32).max(axis=1, keepdims=True)).astype(np.float32)
    scaled = (a32 * sc).astype(np.float32)                      # exact: power-of-two scaling
    hi = scaled.astype(np.float16).astype(np.float32)
    lo = (scaled - hi).astype(np.float32)                       # exact in fp32 (Sterbenz-like: |lo| <= ulp_fp16 / 2)
    return hi.astype(np.float64), lo.astype(np.float64), sc.astype(np.float64)


def _intervals(x32, w32, mu32):
    xc = (x32 - mu32).astype(np.float32)                        # fl(x - mu)
    wc = (w32 - mu32).astype(np.float32)                        # fl(w - mu)
    xh, xl, xs = _split_rows(xc)
    wp = (-2.0 * wc).astype(np.float32)
    wh, wl, ws = _split_rows(wp)
    nxh, nxl = np.linalg.norm(xh, axis=1) / xs[:, 0], np.linalg.norm(xl, axis=1) / xs[:, 0]
    nwh, nwl = np.linalg.norm(wh, axis=1) / ws[:, 0], np.linalg.norm(wl, axis=1) / ws[:, 0]
    mun = np.linalg.norm(mu32.astype(np.float64))
    xn = np.linalg.norm(x32.astype(np.float64), axis=1)
    wn = np.linalg.norm(w32.astype(np.float64), axis=1)
    wcn = np.linalg.norm(wc.astype(np.float64), axis=1)
    A = 1.01 * (nxl + nxh * 2.0 ** -14 + (xn + mun) * 2.0 ** -23)
    B = 1.01 * (nxh + nxl)
    nwl_k = nwl + (wn + mun) * 2.0 ** -22 + wcn * 2.0 ** -22
    bias = (wc.astype(np.float64) ** 2).sum(1).astype(np.float32).astype(np.float64)
    s_hat = (xh / xs) @ (wh / ws).T + bias[None, :]             # what the tensor cores produce (exact products, fp64 sum here)
    E = A[:, None] * nwh[None, :] + B[:, None] * nwl_k[None, :] + 2.4e-7 * bias[None, :]
    # the score the interval has to contain: the centred score of the TRUE (unrounded) operands
    x64, w64, mu64 = x32.astype(np.float64), w32.astype(np.float64), mu32.astype(np.float64)
    s_true = ((w64 - mu64) ** 2).sum(1)[None, :] - 2.0 * (x64 - mu64) @ (w64 - mu64).T
    return s_hat, E, s_true


CASES = {
    "uniform": lambda rng, n, d: rng.random_sample((n, d)).astype(np.float32),
    "blobs": lambda rng, n, d: U.blobs(n, d, seed=3),
    "offset": lambda rng, n, d: (1000.0 + rng.random_sample((n, d))).astype(np.float32),       # |mu| >> spread
    "wide": lambda rng, n, d: (rng.standard_normal((n, d)) * 10.0 ** rng.uniform(-3, 3, size=(1, d))).astype(np.float32),
}


@pytest.mark.parametrize("data", sorted(CASES))
@pytest.mark.parametrize("codebook", ["random", "smooth"])
def test_interval_contains_the_score_and_survivors_contain_the_bmu(data, codebook):
    rng = np.random.RandomState(11)
    n, d, gx, gy = 160, 320, 24, 25
    x = CASES[data](rng, n, d)
    mu = x[: n // 2].mean(0).astype(np.float32)                 # a centre estimated from part of the rows, as the kernel's
    if codebook == "random":
        w = x[rng.randint(n, size=gx * gy)] + (0.05 * x.std(0) * rng.standard_normal((gx * gy, d))).astype(np.float32)
    else:                                                       # a young map: a smooth sheet through the data mean
        u, v = np.meshgrid(np.linspace(-1, 1, gx), np.linspace(-1, 1, gy), indexing="ij")
        a, b = x.std(0) * rng.standard_normal(d) * 0.1, x.std(0) * rng.standard_normal(d) * 0.1
        w = x.mean(0) + u.reshape(-1, 1) * a + v.reshape(-1, 1) * b
    w = w.astype(np.float32)
    s_hat, E, s_true = _intervals(x, w, mu)
    # the fp32 accumulation in TMEM and the fp32 evaluation of s^ are not modelled here (they have their own terms in E);
    # what is checked is the split / centring part of the bound, with those terms as head-room
    assert (np.abs(s_true - s_hat) <= E).all(), float((np.abs(s_true - s_hat) / E).max())
    U_row = (s_hat + E).min(axis=1, keepdims=True)
    survivors = (s_hat - E) <= U_row
    bmu = s_true.argmin(axis=1)
    assert survivors[np.arange(n), bmu].all()
    # and the bound is not vacuous: on centred data few neurons survive
    if data in ("uniform", "blobs") and codebook == "random":
        assert survivors.sum(1).mean() < 4.0
