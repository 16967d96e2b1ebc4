"""CPU: dg::Elliptic3d in compute-in-2d mode and dg::Elliptic1d of the UNMODIFIED reference (oracle/_ref/libdgref.so) against
the C oracle's 2-d operator applied plane by plane -- pins the oracle's volume-form path (cylindrical grids: vol = 1/sqrt(1/R/R),
sigma = chi*vol, final division by vol) and shows that the 3-d class is the 2-d one on every plane.  Skipped without the
reference build."""
import numpy as np
import pytest
from oracle import orc
from util import same_bits, rng


@pytest.fixture(scope="module")
def R():
    from oracle import refwrap
    if not refwrap.available():
        pytest.skip("oracle/_ref/libdgref.so not built (needs /root/reference)")
    return refwrap


@pytest.mark.parametrize("cyl", [False, True])
@pytest.mark.parametrize("direction", [0, 1, 2])
@pytest.mark.parametrize("cwj", [False, True])
def test_elliptic3d_is_oracle_2d_per_plane(R, cyl, direction, cwj):
    from feltor_b200 import topology as T   # host-side topology only (no device call)
    x0, x1 = ([3., -1., 0.], [5., 1., 2 * np.pi]) if cyl else ([0., 0., 0.], [1., 2., 3.])
    N, bc = [9, 7, 4], [T.DIR, T.NEU if cyl else T.PER, T.PER]
    rg = R.grid(x0, x1, 3, N, bc)
    g2 = T.Grid(x0[:2], x1[:2], 3, N[:2], bc[:2])
    n2, nz = g2.size, N[2]
    n = n2 * nz
    r = rng(5 + direction)
    x, y0, chi = r.uniform(-1, 1, n), r.uniform(-1, 1, n), r.uniform(0.5, 2., n)
    mats = dict(leftx=T.derivative(0, g2, T.inverse_bc(bc[0]), T.inverse_dir(direction)),
                lefty=T.derivative(1, g2, T.inverse_bc(bc[1]), T.inverse_dir(direction)),
                rightx=T.derivative(0, g2, bc[0], direction), righty=T.derivative(1, g2, bc[1], direction),
                jumpx=T.jump(0, g2, bc[0]), jumpy=T.jump(1, g2, bc[1]))
    vol = None
    if cyl:
        Rr = np.ascontiguousarray(np.broadcast_to(g2.abscissas(0), (g2.shape(1), g2.shape(0))).reshape(-1))
        vol = 1. / np.sqrt((1. / Rr) / Rr)
    for alpha, beta in ((1., 0.), (-0.5, 0.3)):
        want, w, p = R.elliptic3d_symv(rg, cyl, direction, 0.7, cwj, chi, alpha, x, beta, y0)
        for k in range(nz):
            sl = slice(k * n2, (k + 1) * n2)
            sigma = chi[sl] * vol if cyl else chi[sl].copy()
            E = orc.Elliptic2d(mats, sigma=np.ascontiguousarray(sigma), vol=vol, jfactor=0.7, chi_weight_jump=cwj)
            y = y0[sl].copy()
            E.symv(alpha, np.ascontiguousarray(x[sl]), beta, y)
            assert same_bits(y, want[sl]), (cyl, direction, cwj, alpha, beta, k)
        assert same_bits(p, 1. / chi)
        w3 = T.Grid(x0, x1, [3, 3, 1], N, bc).weights()
        assert same_bits(w, w3 * np.tile(vol, nz) if cyl else w3)


@pytest.mark.parametrize("bcx", [0, 1, 2, 3, 4])
def test_elliptic1d_is_three_symv(R, bcx):
    """Elliptic1d::symv (elliptic.h:171-181) restated with the oracle's Ell symv and pointwiseDot"""
    from feltor_b200 import topology as T
    g = T.Grid([0.3], [2.1], 3, [21], [bcx])
    rg = R.grid([0.3], [2.1], 3, [21], [bcx])
    r = rng(3 + bcx)
    n = g.size
    x, y0, chi = r.uniform(-1, 1, n), r.uniform(-1, 1, n), r.uniform(0.5, 2., n)
    for direction in (0, 1, 2):
        left = T.derivative(0, g, T.inverse_bc(bcx), T.inverse_dir(direction))
        right, jump = T.derivative(0, g, bcx, direction), T.jump(0, g, bcx)
        for alpha, beta in ((1., 0.), (-0.5, 0.3)):
            want, w, p = R.elliptic1d_symv(rg, bcx, direction, 0.7, chi, alpha, x, beta, y0)
            t, y = np.zeros(n), y0.copy()
            orc.ell_symv(right, 1., x, 0., t)
            orc.pointwiseDot(1., t.copy(), chi, 0., t)
            orc.ell_symv(left, -alpha, t, beta, y)
            orc.ell_symv(jump, 0.7 * alpha, x, 1., y)
            assert same_bits(y, want), (bcx, direction, alpha, beta)
            assert same_bits(w, g.weights()) and same_bits(p, 1. / chi)


# ------------------------------------------------------------------ the same against committed fixtures (no reference build needed)
@pytest.fixture(scope="module")
def gold3():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "elliptic3d_golden.npz"))


@pytest.mark.parametrize("cyl", [0, 1])
@pytest.mark.parametrize("direction", [0, 2])
@pytest.mark.parametrize("cwj", [0, 1])
def test_elliptic3d_fixture(gold3, cyl, direction, cwj):
    from feltor_b200 import topology as T
    x0, x1 = ([3., -1., 0.], [5., 1., 2 * np.pi]) if cyl else ([0., 0., 0.], [1., 2., 3.])
    N, bc = [9, 7, 4], [T.DIR, T.NEU if cyl else T.PER, T.PER]
    g2 = T.Grid(x0[:2], x1[:2], 3, N[:2], bc[:2])
    n2, nz = g2.size, N[2]
    mats = dict(leftx=T.derivative(0, g2, T.inverse_bc(bc[0]), T.inverse_dir(direction)),
                lefty=T.derivative(1, g2, T.inverse_bc(bc[1]), T.inverse_dir(direction)),
                rightx=T.derivative(0, g2, bc[0], direction), righty=T.derivative(1, g2, bc[1], direction),
                jumpx=T.jump(0, g2, bc[0]), jumpy=T.jump(1, g2, bc[1]))
    vol = None
    if cyl:
        Rr = np.ascontiguousarray(np.broadcast_to(g2.abscissas(0), (g2.shape(1), g2.shape(0))).reshape(-1))
        vol = 1. / np.sqrt((1. / Rr) / Rr)
    x, y0, chi = gold3["x"], gold3["y0"], gold3["chi"]
    want = gold3[f"e3d/cyl{cyl}/dir{direction}/cwj{cwj}/y"]
    for k in range(nz):
        sl = slice(k * n2, (k + 1) * n2)
        sigma = chi[sl] * vol if cyl else chi[sl].copy()
        E = orc.Elliptic2d(mats, sigma=np.ascontiguousarray(sigma), vol=vol, jfactor=0.7, chi_weight_jump=bool(cwj))
        y = y0[sl].copy()
        E.symv(-0.5, np.ascontiguousarray(x[sl]), 0.3, y)
        assert same_bits(y, want[sl]), (cyl, direction, cwj, k)
    w3 = T.Grid(x0, x1, [3, 3, 1], N, bc).weights()
    assert same_bits(gold3[f"e3d/cyl{cyl}/weights"], w3 * np.tile(vol, nz) if cyl else w3)
    assert same_bits(gold3[f"e3d/cyl{cyl}/precond"], 1. / chi)


@pytest.mark.parametrize("bcx", [0, 1, 2, 3, 4])
def test_elliptic1d_fixture(gold3, bcx):
    from feltor_b200 import topology as T
    g = T.Grid([0.3], [2.1], 3, [21], [bcx])
    x, y0, chi = gold3["x1d"], gold3["y1d"], gold3["chi1d"]
    for direction in (0, 1, 2):
        left = T.derivative(0, g, T.inverse_bc(bcx), T.inverse_dir(direction))
        right, jump = T.derivative(0, g, bcx, direction), T.jump(0, g, bcx)
        t, y = np.zeros(g.size), y0.copy()
        orc.ell_symv(right, 1., x, 0., t)
        orc.pointwiseDot(1., t.copy(), chi, 0., t)
        orc.ell_symv(left, 0.5, t, 0.3, y)
        orc.ell_symv(jump, 0.7 * -0.5, x, 1., y)
        assert same_bits(y, gold3[f"e1d/bc{bcx}/dir{direction}/y"]), (bcx, direction)


# ------------------------------------------------------------------ full 3-d mode (z derivative, 3-d tensor product: elliptic.h:688-697)
class _OrcBackend:
    """dg-shaped primitives of the C oracle on numpy arrays (the GPU twin lives in tests/test_gpu_elliptic3d.py)"""

    def __init__(self):
        from backends import OracleBlas1
        self.b = OracleBlas1()

    def make(self, a): return np.array(a, dtype=np.float64, copy=True)
    def symv(self, m, a, x, b, y): orc.ell_symv(m, a, x, b, y)
    def tensor_multiply3d(self, lam, t, ins, mu, outs): orc.tensor_multiply3d(lam, t, list(ins), mu, list(outs))
    def tensor_multiply2d(self, lam, t, i0, i1, mu, o0, o1): orc.tensor_multiply2d(lam, t, i0, i1, mu, o0, o1)
    def pointwiseDot(self, *a): self.b.pointwiseDot(*a)
    def pointwiseDivide(self, *a): self.b.pointwiseDivide(*a)
    def axpbypgz(self, *a): self.b.axpbypgz(*a)
    def scal(self, x, a): self.b.scal(x, a)


E3_CASES = [(0, 0, 0, 0), (1, 2, 0, 0), (0, 1, 1, 0), (1, 0, 1, 0), (1, 2, 1, 1), (0, 2, 0, 1)]   # cyl, direction, cwj, compute_in_2d


def _e3_setup(cyl):
    x0, x1 = ([3., -1., 0.], [5., 1., 2 * np.pi]) if cyl else ([0., 0., 0.], [1., 2., 3.])
    return x0, x1, [9, 7, 5], [1, 4 if cyl else 0, 0]


@pytest.mark.parametrize("cyl,direction,cwj,in2d", E3_CASES)
def test_elliptic3d_full_mode_composition_vs_reference(R, cyl, direction, cwj, in2d):
    """the call-by-call restatement of Elliptic3d::symv (tests/util.elliptic3d_full_symv) on the C oracle equals the live
    reference class in its full 3-d mode bit for bit: pins the harness the GPU test runs on the C ABI"""
    from feltor_b200 import topology as T
    from util import elliptic3d_full_symv
    x0, x1, N, bc = _e3_setup(cyl)
    rg = R.grid(x0, x1, 3, N, bc)
    n = 9 * N[0] * N[1] * N[2]
    r = rng(11 + direction + 2 * cyl)
    x, y0, chi = r.uniform(-1, 1, n), r.uniform(-1, 1, n), r.uniform(0.5, 2., n)
    for alpha, beta in ((1., 0.), (-0.5, 0.3)):
        want = R.elliptic3d_symv_mode(rg, cyl, direction, 0.7, cwj, in2d, chi, alpha, x, beta, y0)
        y = y0.copy()
        elliptic3d_full_symv(_OrcBackend(), T, x0, x1, N, bc, direction, 0.7, bool(cwj), bool(cyl), chi, alpha, x, beta, y, bool(in2d))
        assert same_bits(y, want), (cyl, direction, cwj, in2d, alpha, beta)


@pytest.mark.parametrize("cyl,direction,cwj,in2d", E3_CASES)
def test_elliptic3d_full_mode_fixture(gold3, cyl, direction, cwj, in2d):
    """the same against committed outputs of the reference (tests/golden/make_golden_elliptic3d.py): no reference build needed"""
    from feltor_b200 import topology as T
    from util import elliptic3d_full_symv
    x0, x1, N, bc = _e3_setup(cyl)
    x, y0, chi = gold3["x_full"], gold3["y0_full"], gold3["chi_full"]
    y = y0.copy()
    elliptic3d_full_symv(_OrcBackend(), T, x0, x1, N, bc, direction, 0.7, bool(cwj), bool(cyl), chi, -0.5, x, 0.3, y, bool(in2d))
    assert same_bits(y, gold3[f"e3dfull/cyl{cyl}/dir{direction}/cwj{cwj}/in2d{in2d}/y"])
