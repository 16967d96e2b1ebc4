"""CPU: dgb_topo_window_stencil (host code of the library) against dg::create::window_stencil of the UNMODIFIED reference
(oracle/_ref/libdgref_ds.so): index setup is bit-exact -- row offsets, unsorted column indices with duplicates, values +-1."""
import ctypes as C
import os
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "libdgref_ds.so")


def ref_window(x0, x1, n, N, bc, window):
    L = C.CDLL(REF)
    ndim = len(N)
    rows = int(np.prod([n * v for v in N]))
    per = int(np.prod(window))
    pos, idx, val = np.empty(rows + 1, dtype=np.int32), np.empty(rows * per, dtype=np.int32), np.empty(rows * per)
    arr = lambda t, v: (t * ndim)(*v)
    nnz = L.ref_window_stencil(ndim, arr(C.c_double, x0), arr(C.c_double, x1), n, arr(C.c_int, N), arr(C.c_int, bc), arr(C.c_int, window),
                               pos.ctypes.data_as(C.c_void_p), idx.ctypes.data_as(C.c_void_p), val.ctypes.data_as(C.c_void_p))
    assert nnz == rows * per
    return pos, idx, val


@pytest.mark.parametrize("bc", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("window", [1, 3, 4, 5, 9])
def test_window_stencil_1d(bc, window):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/libdgref_ds.so not built (needs /root/reference)")
    from feltor_b200 import topology as T
    g = T.Grid([0.2], [1.7], 3, [7], [bc])
    got, want = T.window_stencil(g, window), ref_window([0.2], [1.7], 3, [7], [bc], [window])
    for a, b in zip(got, want):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("bcs", [(0, 0), (1, 4), (2, 3), (4, 1), (3, 0)])
@pytest.mark.parametrize("window", [(3, 3), (5, 3), (2, 4)])
def test_window_stencil_2d(bcs, window):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/libdgref_ds.so not built (needs /root/reference)")
    from feltor_b200 import topology as T
    g = T.Grid([0., -1.], [1., 2.], 3, [5, 4], list(bcs))
    got, want = T.window_stencil(g, list(window)), ref_window([0., -1.], [1., 2.], 3, [5, 4], list(bcs), list(window))
    for a, b in zip(got, want):
        assert np.array_equal(a, b)


def test_window_stencil_rejects_bad_window():
    from feltor_b200 import topology as T, DgbError
    g = T.Grid([0.], [1.], 3, [4], [0])
    with pytest.raises(DgbError):
        T.window_stencil(g, 0)
