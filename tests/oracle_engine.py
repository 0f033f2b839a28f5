"""CPU stand-in for CudaEngine, for tests of the HOST logic only (sharding, the per-epoch
all-reduce, API plumbing).  Every method is answered by the oracle; it is never shipped and
never selected by the product code (the tests assign it to the private `_engine` attribute of an XPySom instance)."""
import numpy as np
import torch

from oracle import som_oracle as so

_DIST = {0: "euclidean", 1: "cosine", 2: "manhattan", 3: "chebyshev", 4: "norm_p_no_opt"}
_NEIGH = {0: "gaussian", 1: "mexican_hat", 2: "bubble", 3: "triangle"}
_TOPO = {0: "rectangular", 1: "hexagonal"}


class OracleEngine:
    name = "oracle"

    def __init__(self):
        self.device = torch.device("cpu")
        self.launches = 0

    def empty(self, *shape, dtype=torch.float32):
        return torch.empty(*shape, dtype=dtype)

    def zeros(self, *shape, dtype=torch.float32):
        return torch.zeros(*shape, dtype=dtype)

    def to_device(self, t):
        return t

    def workspace(self, n, k, d):
        return torch.empty(1, dtype=torch.uint8)

    def neigh_tables(self, gx, gy, d=None):
        return torch.empty(1)

    def prepare_codebook(self, w, dist_kind, p, ws):
        pass

    def prepare_samples(self, x, want_scale=True, out=None, colmax=None):
        d = x.shape[1]
        cm = torch.zeros((d + 3) // 4 * 4) if colmax is None else colmax
        if x.shape[0]:
            cm[:d] = torch.maximum(cm[:d], x.abs().amax(dim=0))
        return None, cm

    def accum_scales(self, colmax, d, n_total):
        """The library's rule (csrc/accumulate.cuh): q_c = 62 - (exponent of the column maximum + 1) - ceil(log2 n_total)."""
        dq = (d + 3) // 4 * 4
        lg = int(np.ceil(np.log2(max(n_total, 1))))
        q = np.zeros(dq)
        cm = colmax.numpy()[:d].astype(np.float64)
        ok = cm > 0
        q[:d][ok] = 62 - (np.floor(np.log2(cm[ok])) + 1) - lg
        qs, qi = np.zeros(dq), np.zeros(dq)
        qs[:d], qi[:d] = 2.0 ** q[:d], 2.0 ** -q[:d]
        return torch.from_numpy(qs), torch.from_numpy(qi)           # float64 here: exact powers of two either way

    def accumulator(self, k, d):
        return torch.zeros(k * ((d + 1) // 2 * 2) + k, dtype=torch.int64)

    def accum_one_copy(self, acc, k, d):
        return acc                                                  # (one copy here: nothing to fold)

    def bmu(self, x, w, dist_kind, p, algo, ws, bmu_out=None, best_out=None, xscale=None):
        k, d = w.shape
        spec = so.SomSpec(gx=k, gy=1, dim=d, activation_distance="euclidean", p=p)
        spec.activation_distance = _DIST[dist_kind]
        flat = so.bmu_flat(spec, x.numpy(), w.numpy().reshape(k, 1, d))
        out = torch.from_numpy(flat.astype(np.int32))
        if bmu_out is not None:
            bmu_out.copy_(out)
            return bmu_out
        return out

    def accumulate(self, x, bmu, k, qscale, acc):
        """Fixed-point sums, as the CUDA path: rint(x * 2^q_c) added as 64-bit integers (order-independent)."""
        d = x.shape[1]
        lds = (d + 1) // 2 * 2
        xi = np.rint(x.numpy().astype(np.float64) * qscale.numpy()[:d]).astype(np.int64)
        S = acc.numpy()[:k * lds].reshape(k, lds)
        np.add.at(S[:, :d], bmu.numpy(), xi)
        acc.numpy()[k * lds:] += np.bincount(bmu.numpy(), minlength=k)

    def epoch_accumulate(self, x, w, dist_kind, p, algo, qscale, acc, ws, bmu_out=None, xscale=None):
        bmu = self.bmu(x, w, dist_kind, p, algo, ws, bmu_out=bmu_out)
        self.accumulate(x, bmu, w.shape[0], qscale, acc)

    def accum_finalize(self, acc, qinv, k, d, s, c):
        lds = (d + 1) // 2 * 2
        S = acc.numpy()[:k * lds].reshape(k, lds)[:, :d].astype(np.float64) * qinv.numpy()[:d]
        s.copy_(torch.from_numpy(S.astype(np.float32).ravel()))
        c.copy_(torch.from_numpy(acc.numpy()[k * lds:].astype(np.float32)))
        acc.zero_()

    def neigh_apply(self, s, c, gx, gy, d, topology, neigh_kind, sigma, eta, std_coeff, compact, num, den, tables):
        spec = so.SomSpec(gx=gx, gy=gy, dim=d, sigma=1.0, neighborhood_function=_NEIGH[neigh_kind],
                          topology=_TOPO[topology], std_coeff=std_coeff, compact_support=bool(compact))
        H = so.neighborhood_table(spec, float(sigma)).astype(np.float64) * float(eta)
        num.copy_(torch.from_numpy((H.T @ s.numpy().reshape(gx * gy, d).astype(np.float64)).astype(np.float32).ravel()))
        den.copy_(torch.from_numpy((H.T @ c.numpy().astype(np.float64)).astype(np.float32)))

    def epoch_tail(self, acc, qinv, s, c, w, gx, gy, d, topology, neigh_kind, sigma, eta, std_coeff, compact, dist_kind, p,
                   num, den, tables, ws):
        if acc is not None:
            self.accum_finalize(acc, qinv, gx * gy, d, s, c)
        self.neigh_apply(s, c, gx, gy, d, topology, neigh_kind, sigma, eta, std_coeff, compact, num, den, tables)
        self.merge(w, num, den)
        self.prepare_codebook(w, dist_kind, p, ws)

    def merge(self, w, num, den):
        k, d = w.shape
        out = so.merge(w.numpy(), num.numpy().reshape(k, d), den.numpy().reshape(k, 1))
        w.copy_(torch.from_numpy(out.astype(np.float32)))
