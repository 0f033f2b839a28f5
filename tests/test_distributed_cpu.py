"""The N>1 path on CPU: two `gloo` ranks, each holding one shard of the samples, one all-reduce
of the per-BMU sums [S | c] per epoch, replicated codebook (replaces the Dask block
graph of xpysom.py:545-558).  Device compute is answered by the oracle through the private `_engine`
attribute,
so this covers the host logic only: sharding, the collective, replicated apply/merge."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import som_testutil as U
from oracle import som_oracle as so


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, kw, shards, T, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle_engine import OracleEngine
    from xpysom_dask_b200 import XPySom
    som = XPySom(process_group=True, **kw)
    som._engine = OracleEngine()
    som.train(shards[rank], T)
    out[rank] = som._weights.copy()
    dist.destroy_process_group()


@pytest.mark.parametrize("kw", [
    dict(x=8, y=7, input_len=12, random_seed=3),
    dict(x=6, y=6, input_len=9, random_seed=4, topology="hexagonal", neighborhood_function="mexican_hat",
         activation_distance="cosine", decay_function="linear"),
])
def test_two_rank_shards_equal_single_process(kw):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle_engine import OracleEngine
    from xpysom_dask_b200 import XPySom
    data = U.blobs(1001, kw["input_len"], seed=8)              # ragged split: 501 + 500 rows
    T = 4
    single = XPySom(**kw)
    single._engine = OracleEngine()
    single.train(data, T)

    shards = [data[:501], data[501:]]
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, 2, port, kw, shards, T, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    np.testing.assert_array_equal(out[0], out[1])              # replicas stay bit-identical
    assert U.codebook_rel_err(out[0], single._weights) < 1e-5  # shard sums re-associate in fp32 only

    # and the host class on the oracle engine tracks the reference's own algorithm
    spec = so.SomSpec(gx=kw["x"], gy=kw["y"], dim=kw["input_len"], random_seed=kw["random_seed"], n_parallel=4000,
                      **{k: v for k, v in kw.items() if k not in ("x", "y", "input_len", "random_seed")})
    w_ref = so.epoch(spec, data, np.asarray(so.init_weights(spec), dtype=np.float32), 0, T)
    one = XPySom(**kw)
    one._engine = OracleEngine()
    one.train(data, T, iter_beg=0, iter_end=1)
    assert U.codebook_rel_err(one._weights, w_ref) < 1e-4


def test_empty_shard_on_one_rank():
    """Fewer rows than ranks (or an unlucky split): the rank without samples contributes zeros to the all-reduce
    and keeps its replica of the codebook identical (the reference's Dask graph accepts any chunking)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from oracle_engine import OracleEngine
    from xpysom_dask_b200 import XPySom
    kw = dict(x=5, y=4, input_len=6, random_seed=2)
    data = U.blobs(300, 6, seed=1)
    single = XPySom(**kw)
    single._engine = OracleEngine()
    single.train(data, 3)
    shards = [data, np.zeros((0, 6), dtype=np.float32)]
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    ctx = mp.get_context("spawn")
    procs = [ctx.Process(target=_worker, args=(r, 2, port, kw, shards, 3, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    np.testing.assert_array_equal(out[0], out[1])
    assert U.codebook_rel_err(out[0], single._weights) < 1e-6


def test_process_group_requires_initialised_backend():
    from oracle_engine import OracleEngine
    from xpysom_dask_b200 import XPySom
    som = XPySom(5, 5, 3, process_group=True)
    som._engine = OracleEngine()
    with pytest.raises(RuntimeError):
        som.train(np.zeros((4, 3), np.float32), 1)
