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

    def prepare_samples(self, x):
        return None

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

    def accumulate(self, x, bmu, k, s, c):
        S, cc = so.sums_by_bmu(bmu.numpy(), x.numpy(), k)
        s.view(k, -1).add_(torch.from_numpy(S.astype(np.float32)))
        c.add_(torch.from_numpy(cc.astype(np.float32)))

    def epoch_accumulate(self, x, w, dist_kind, p, algo, s, c, ws, bmu_out=None, xscale=None):
        bmu = self.bmu(x, w, dist_kind, p, algo, ws, bmu_out=bmu_out)
        self.accumulate(x, bmu, w.shape[0], s, c)

    def neigh_apply(self, s, c, gx, gy, d, topology, neigh_kind, sigma, eta, std_coeff, compact, num, den, tables):
        spec = so.SomSpec(gx=gx, gy=gy, dim=d, sigma=1.0, neighborhood_function=_NEIGH[neigh_kind],
                          topology=_TOPO[topology], std_coeff=std_coeff, compact_support=bool(compact))
        H = so.neighborhood_table(spec, float(sigma)).astype(np.float64) * float(eta)
        num.copy_(torch.from_numpy((H.T @ s.numpy().reshape(gx * gy, d).astype(np.float64)).astype(np.float32).ravel()))
        den.copy_(torch.from_numpy((H.T @ c.numpy().astype(np.float64)).astype(np.float32)))

    def epoch_tail(self, s, c, w, gx, gy, d, topology, neigh_kind, sigma, eta, std_coeff, compact, dist_kind, p,
                   num, den, tables, ws):
        self.neigh_apply(s, c, gx, gy, d, topology, neigh_kind, sigma, eta, std_coeff, compact, num, den, tables)
        self.merge(w, num, den)
        self.prepare_codebook(w, dist_kind, p, ws)
        s.zero_()
        c.zero_()

    def merge(self, w, num, den):
        k, d = w.shape
        out = so.merge(w.numpy(), num.numpy().reshape(k, d), den.numpy().reshape(k, 1))
        w.copy_(torch.from_numpy(out.astype(np.float32)))
