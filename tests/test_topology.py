"""CPU: the host topology builders of libdgb200.so (feltor_b200/csrc/topology.cu) are bit-identical to the reference's
(committed fixtures from tests/golden/make_golden.py; live comparison when oracle/_ref/libdgref.so is present)."""
import numpy as np
import pytest
from feltor_b200 import topology as T
from util import same_bits


def test_fixtures(golden):
    for n in (2, 3):
        g = T.Grid([0.1], [2.3], n, [6], [0])
        assert same_bits(g.abscissas(0), golden[f"topo/absc/n{n}"])
        assert same_bits(g.weights1d(0), golden[f"topo/w1d/n{n}"])
        for bc in range(5):
            for d in range(3):
                m = T.derivative(0, g, bc, d)
                assert same_bits(m.data, golden[f"topo/dx/n{n}/bc{bc}/dir{d}/data"]), (n, bc, d)
                assert np.array_equal(m.cols_idx, golden[f"topo/dx/n{n}/bc{bc}/dir{d}/cols"])
                assert np.array_equal(m.data_idx, golden[f"topo/dx/n{n}/bc{bc}/dir{d}/didx"])
            assert same_bits(T.jump(0, g, bc).data, golden[f"topo/jump/n{n}/bc{bc}/data"])
    g = T.Grid([0, 0], [1, 2], 3, [4, 6], [0, 1])
    assert same_bits(g.weights(), golden["topo/w2d"])
    for kind, a, b in (("fast_projection", 1, 2), ("fast_interpolation", 1, 2), ("fast_projection", 3, 1)):
        m = getattr(T, kind)(1, g, a, b)
        assert same_bits(m.data, golden[f"topo/{kind}/{a}_{b}/data"]), kind
        assert np.array_equal(m.cols_idx, golden[f"topo/{kind}/{a}_{b}/cols"])
        assert np.array_equal(m.meta(), golden[f"topo/{kind}/{a}_{b}/meta"])
    for n in (3, 17):
        for w in range(4):
            assert same_bits(T.dlt(w, n), golden[f"topo/dlt/{w}/n{n}"])


def test_errors():
    import feltor_b200 as fb
    g = T.Grid([0, 0], [1, 1], 3, [7, 6], [0, 0])
    with pytest.raises(fb.DgbError):
        T.fast_projection(0, g, 1, 2)      # 7 cells not divisible by 2 (fast_interpolation.h:231)
    with pytest.raises(fb.DgbError):
        T.derivative(2, g, 0, 0)           # coord >= Nd (derivatives.h:50)
    with pytest.raises(fb.DgbError):
        T.dlt(0, 21)                       # n > 20 (dlt.h:72)


def test_live_reference(ref):
    if ref is None:
        pytest.skip("oracle/_ref/libdgref.so not built (reference tree absent)")

    def eq(r, m):
        return (r.data.shape == m.data.shape and same_bits(r.data, m.data) and np.array_equal(r.cols_idx, m.cols_idx)
                and np.array_equal(r.data_idx, m.data_idx)
                and (r.num_rows, r.num_cols, r.bpl, r.n, r.left_size, r.right_size)
                == (m.num_rows, m.num_cols, m.bpl, m.n, m.left_size, m.right_size))
    for n in range(1, 21):
        for w in range(4):
            assert same_bits(T.dlt(w, n), ref.dlt(w, n)), (w, n)
    for n in (1, 2, 3, 4, 5, 8):
        for N in (2, 5, 24):
            for bc in range(5):
                rg, g = ref.grid([0.1], [2.3], n, [N], [bc]), T.Grid([0.1], [2.3], n, [N], [bc])
                for d in range(3):
                    assert eq(ref.ell_create(rg, "derivative", 0, bc, d), T.derivative(0, g, bc, d)), (n, N, bc, d)
                assert eq(ref.ell_create(rg, "jump", 0, bc), T.jump(0, g, bc)), (n, N, bc)
    rg = ref.grid([0, 0, -1], [1, 2 * np.pi, 3], 3, [12, 8, 4], [0, 1, 4])
    g = T.Grid([0, 0, -1], [1, 2 * np.pi, 3], [3, 3, 1], [12, 8, 4], [0, 1, 4])
    assert same_bits(ref.weights(rg), g.weights())
    for u in range(3):
        assert same_bits(ref.abscissas(rg, u), g.abscissas(u))
        assert eq(ref.ell_create(rg, "derivative", u, 4, 1), T.derivative(u, g, 4, 1))
    for coord in (0, 1):
        for a, b in ((1, 2), (1, 4), (3, 1), (3, 2), (1, 1)):
            assert eq(ref.ell_create(rg, "fast_projection", coord, a=a, b=b), T.fast_projection(coord, g, a, b))
        for a, b in ((1, 2), (1, 4), (2, 1), (2, 3)):
            assert eq(ref.ell_create(rg, "fast_interpolation", coord, a=a, b=b), T.fast_interpolation(coord, g, a, b))
