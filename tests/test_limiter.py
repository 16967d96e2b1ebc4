"""CPU: dg::create::limiter_stencil / dg::CSRSlopeLimiter (inc/dg/topology/stencil.h:89-256, filter.h:288-336).  The host builder
of the library (dgb_topo_limiter_stencil) and the oracle's restatement of the functor are pinned to golden vectors of the
UNMODIFIED reference (tests/golden/limiter_golden.npz, made by tests/golden/make_golden_limiter.py) and, when oracle/_ref is built,
to the live reference on further cases."""
import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from oracle import orc
from util import same_bits

GOLD = np.load(os.path.join(ROOT, "tests", "golden", "limiter_golden.npz"))
REF = os.path.join(ROOT, "oracle", "_ref", "libdgref_ds.so")


def cases():
    import importlib.util
    spec = importlib.util.spec_from_file_location("_mgl", os.path.join(ROOT, "tests", "golden", "make_golden_limiter.py"))
    # only the table of cases is needed; the module loads libdgref_ds.so at import, so parse it instead of importing when absent
    src = open(spec.origin).read()
    ns = {}
    exec(src[src.index("CASES = "):src.index("if __name__")], ns)
    return ns["CASES"]


@pytest.mark.parametrize("case", cases(), ids=lambda c: c[0])
def test_limiter_stencil_builder_vs_golden(case):
    name, x0, x1, n, N, bc, direction, bound = case
    from feltor_b200 import topology as T
    g = T.Grid(x0, x1, n, N, bc)
    pos, idx, val = T.limiter_stencil(g, direction, bound)
    assert np.array_equal(pos, GOLD[name + "/pos"]) and np.array_equal(idx, GOLD[name + "/idx"])
    assert same_bits(val, GOLD[name + "/val"])


@pytest.mark.parametrize("case", cases(), ids=lambda c: c[0])
@pytest.mark.parametrize("mod", [0., 0.3])
def test_oracle_slope_limiter_vs_golden(case, mod):
    name = case[0]
    pos, idx, val, x = (GOLD[name + "/" + k] for k in ("pos", "idx", "val", "x"))
    y = np.full(x.size, np.nan)
    orc.csr_stencil(4, pos, idx, val, mod, x, y)
    want = GOLD[name + "/y_mod%g" % mod]
    assert same_bits(y, want)
    if mod == 0.:
        assert not np.array_equal(want, x), "the case limits nothing"


@pytest.mark.parametrize("seed", range(4))
def test_oracle_slope_limiter_vs_live_reference(seed):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/libdgref_ds.so not built (needs /root/reference)")
    import make_golden_limiter as M
    r = np.random.default_rng(seed)
    n, N = int(r.integers(2, 6)), [int(r.integers(3, 12)), int(r.integers(3, 12))]
    bc, direction = [int(r.integers(0, 5)), int(r.integers(0, 5))], int(r.integers(0, 2))
    pos, idx, val = M.ref_limiter([0., 0.], [1., 2.], n, N, bc, direction, bc[direction])
    from feltor_b200 import topology as T
    got = T.limiter_stencil(T.Grid([0., 0.], [1., 2.], n, N, bc), direction)
    assert np.array_equal(got[0], pos) and np.array_equal(got[1], idx) and same_bits(got[2], val)
    x = M.field(pos.size - 1, 100 + seed) * r.uniform(0.5, 2.)
    for mod in (0., 0.1):
        y = np.full(x.size, np.nan)
        orc.csr_stencil(4, pos, idx, val, mod, x, y)
        assert same_bits(y, M.ref_apply(pos, idx, val, mod, x))
