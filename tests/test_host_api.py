"""Host-side behaviour that needs no GPU: constructor contract, error behaviour, decays, the C-ABI
library loading with every declared symbol, and the loud failure when there is no CUDA device."""
import os
import pickle
import re
import warnings

import numpy as np
import pytest

import som_testutil as U

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    from xpysom_dask_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "som_b200.h")).read()
    declared = set(re.findall(r"\b(som_b200_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.som_b200_abi_version() == _lib.ABI_VERSION
    # pure host helpers may be called without a GPU
    assert lib.som_b200_workspace_bytes(1024, 64) >= 2 * 1024 * 64 * 4
    assert lib.som_b200_workspace_bytes(0, 64) == 0
    assert lib.som_b200_neigh_table_floats(10, 8) >= 2 * (3 * 100 + 64)
    assert lib.som_b200_shard_workspace_bytes(1000, 64, 16) >= lib.som_b200_workspace_bytes(64, 16) + 4000


def test_no_gpu_means_loud_failure_not_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from xpysom_dask_b200 import XPySom, _lib
    som = XPySom(5, 5, 2, random_seed=1)
    with pytest.raises(_lib.SomB200Error):
        som.train(np.zeros((3, 2), np.float32), 1)
    with pytest.raises(_lib.SomB200Error):
        som.winner(np.zeros((3, 2), np.float32))


def test_constructor_contract_matches_reference():
    from xpysom_dask_b200 import XPySom
    g = U.load("api.npz")
    s = XPySom(5, 5, 2, sigma=1.0, learning_rate=0.5, random_seed=1)
    np.testing.assert_array_equal(s._weights, g["seed1_w_init"])           # xpysom.py:167,189-190
    assert s._sigma == 1.0 and XPySom(8, 6, 2)._sigma == 3.0                # default min(x,y)/2, :178-181
    with pytest.raises(ValueError):
        XPySom(5, 5, 1, neighborhood_function='boooom')                     # tests.py:41-43
    with pytest.raises(ValueError):
        XPySom(5, 5, 1, activation_distance='ridethewave')                  # tests.py:45-47
    with pytest.raises(ValueError):
        XPySom(5, 5, 1, topology='triangular')
    with pytest.raises(ValueError):
        XPySom(5, 5, 1, decay_function='stepwise')
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        with pytest.raises(ValueError):
            XPySom(5, 5, 1, topology='hexagonal', neighborhood_function='triangle')   # :207-209, :272-279
        XPySom(5, 5, 1, sigma=5)                                            # :164-165
    assert any('sigma is too high' in str(x.message) for x in w)
    hexs = XPySom(6, 5, 3, topology="hexagonal", random_seed=3)
    xx, yy = hexs.get_euclidean_coordinates()
    np.testing.assert_array_equal(xx, g["hex_xx"])
    np.testing.assert_array_equal(yy, g["hex_yy"])
    # xp / use_dask / dask_chunks / n_parallel are accepted
    XPySom(4, 4, 2, xp=np, use_dask=True, dask_chunks=(10, 2), n_parallel=123)


def test_decays_match_reference_values():
    from xpysom_dask_b200.decays import DECAY_FUNCTIONS
    for r in U.load_json("decays.json"):
        v = DECAY_FUNCTIONS[r["kind"]](r["v0"], r["vN"], r["t"], r["T"])
        assert float(v) == r["value"]
        assert isinstance(v, np.floating) == r["is_np"]


def test_pca_and_random_init_and_pickle():
    from xpysom_dask_b200 import XPySom
    g = U.load("api.npz")
    s = XPySom(2, 2, 2)
    s.pca_weights_init(np.array([[1., 0.], [0., 1.], [1., 0.], [0., 1.]]))
    np.testing.assert_array_almost_equal(s._weights, g["pca_w"])           # tests.py:129-134
    s = XPySom(2, 2, 2, random_seed=1)
    s.random_weights_init(np.array([[1.0, .0]]))
    for w in s._weights:
        np.testing.assert_array_equal(w[0], np.array([1.0, .0]))           # tests.py:123-127
    s2 = pickle.loads(pickle.dumps(s))                                      # tests.py:145-150
    np.testing.assert_array_equal(s2._weights, s._weights)
    with pytest.raises(ValueError):
        XPySom(3, 3, 1).pca_weights_init(np.zeros((4, 1)))
    with pytest.raises(ValueError):
        s.quantization_error(np.zeros((4, 3)))                              # wrong feature count, :361-367


def test_input_adapters_numpy_list_torch_dlpack():
    """train / winner accept what the reference accepts (lists, numpy of any dtype, xpysom.py:485-510) plus
    torch tensors and any DLPack exporter (CuPy, the reference's GPU input); everything becomes a 2-D fp32
    matrix, 1-D samples become one row, anything else is rejected."""
    import torch
    from xpysom_dask_b200.xpysom import _as_f32_matrix

    class DlpackOnly:                      # stands in for a CuPy / JAX array: only the DLPack protocol
        def __init__(self, t):
            self._t = t

        def __dlpack__(self, *a, **k):
            return self._t.__dlpack__(*a, **k)

        def __dlpack_device__(self):
            return self._t.__dlpack_device__()

    base = np.arange(12, dtype=np.float64).reshape(4, 3)
    for data in (base, base.tolist(), base.astype(np.int32), torch.from_numpy(base),
                 DlpackOnly(torch.from_numpy(base.astype(np.float32)))):
        t = _as_f32_matrix(data)
        assert t.dtype == torch.float32 and tuple(t.shape) == (4, 3)
        np.testing.assert_array_equal(t.numpy(), base.astype(np.float32))
    shared = torch.arange(6, dtype=torch.float32).reshape(2, 3)
    assert _as_f32_matrix(DlpackOnly(shared)).data_ptr() == shared.data_ptr()        # no copy
    assert tuple(_as_f32_matrix([1.0, 2.0, 3.0]).shape) == (1, 3)
    with pytest.raises(ValueError):
        _as_f32_matrix(np.zeros((2, 2, 2)))
