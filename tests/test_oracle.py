"""CPU: pins the oracle (oracle/dgoracle.c) against (1) the reference's own golden vectors, (2) the committed
fixtures produced by the unmodified reference (tests/golden/make_golden.py) and (3) -- when
oracle/_ref/libdgref.so is present -- the live reference on random inputs."""
import numpy as np
import pytest
from oracle import orc
from util import bits, same_bits, blas1_sequence, BLAS1_GOLDEN, rng, wide
from backends import OracleBlas1, RefBlas1, np_make, np_get
import kat

OB = OracleBlas1()


def test_blas1_reference_goldens():
    """inc/dg/blas1_t.cpp:102-176; the reference accepts +-2 ulp, the oracle is exact"""
    out = blas1_sequence(OB, np_make, np_get)
    for name, g in BLAS1_GOLDEN:
        assert np.all(out[name] == g), name


def test_blas1_fixtures(golden):
    v = golden["blas1/in"]

    def run(name, f):
        w = [a.copy() for a in v]
        f(w)
        assert same_bits(np.stack(w), golden["blas1/" + name]), name
    run("axpby", lambda w: OB.axpby(0.7, w[0], -1.3, w[1]))
    run("axpbyz", lambda w: OB.axpby(0.7, w[0], -1.3, w[1], w[2]))
    run("axpbypgz", lambda w: OB.axpbypgz(0.7, w[0], -1.3, w[1], 0.4, w[2]))
    run("pdot", lambda w: OB.pointwiseDot(0.7, w[0], w[1], -1.3, w[2]))
    run("pdot_alias", lambda w: OB.pointwiseDot(0.7, w[0], w[1], -1.3, w[1]))
    run("pdot3", lambda w: OB.pointwiseDot(0.7, w[0], w[1], w[2], -1.3, w[3]))
    run("pdot2", lambda w: OB.pointwiseDot(0.7, w[0], w[1], -1.3, w[2], w[3], 0.4, w[4]))
    run("pdiv", lambda w: OB.pointwiseDivide(0.7, w[0], w[1], -1.3, w[2]))
    run("pdiv_alias", lambda w: OB.pointwiseDivide(0.7, w[2], w[1], -1.3, w[2]))
    run("tensor2d", lambda w: orc.tensor_multiply2d(w[0], (v[1], v[2], v[3], v[4]), w[1], w[2], 0.3, w[3], w[4]))


@pytest.mark.parametrize("case", list(kat.evaluation_cases()), ids=lambda c: c[0])
def test_dot_reference_goldens(case):
    """inc/dg/topology/evaluation_t.cpp:56-175 (tolerance < 2 as there)"""
    name, kind, ops, gold = case
    val, st = orc.dot2(*ops) if kind == "dot2" else orc.dot3(*ops)
    assert st == 0
    assert abs(int(bits([val])[0]) - gold) < 2, name


@pytest.mark.parametrize("name", ["small", "wide", "mid"])
def test_dot_fixtures(golden, name):
    x, w, y = (golden[f"dot/{name}/{k}"] for k in "xwy")
    a2, s2 = orc.exdot2(x, y)
    a3, s3 = orc.exdot3(x, w, y)
    assert s2 == 0 and s3 == 0
    assert np.array_equal(a2, golden[f"dot/{name}/acc2"]) and np.array_equal(a3, golden[f"dot/{name}/acc3"])
    assert same_bits([orc.round_acc(a2)], golden[f"dot/{name}/val2"])
    assert same_bits([orc.round_acc(a3)], golden[f"dot/{name}/val3"])


def test_dot_nonfinite_status():
    x = np.ones(10)
    x[3] = np.inf
    assert orc.dot2(x, x)[1] == 1
    x[3] = np.nan
    assert orc.dot3(x, x, x)[1] == 1
    big = np.full(4, 1e300)
    assert orc.dot2(big, big)[1] == 1  # product overflows although inputs are finite


@pytest.mark.parametrize("case", list(kat.derivative_cases(three_d=True)), ids=lambda c: c[0])
def test_derivative_reference_goldens(case):
    """inc/dg/topology/derivatives_t.cpp:54-133"""
    def symv(m, a, x, b, y): orc.ell_symv(m, a, x, b, y)
    got, gold, gh = kat.run_derivative_case(case, np_make, lambda x, y: orc.dot2(x, y)[0],
                                            lambda x, w, y: orc.dot3(x, w, y)[0], symv, orc.pointwiseDot_xy)
    assert abs(got - gold) < 2 or abs(got - gh) < 2, (case[0], got, gold)


def _ell_from_golden(golden):
    from feltor_b200 import topology as T
    g2 = T.Grid([0, 0.1], [np.pi, 2 * np.pi + 0.1], 3, [8, 6], [T.DIR, T.PER])
    g3 = T.Grid([0, 0.1, 1.], [np.pi, 2 * np.pi + 0.1, 2.], [3, 3, 1], [4, 3, 5], [T.DIR, T.PER, T.NEU_DIR])
    return T, g2, g3


def test_ell_symv_fixtures(golden):
    T, g2, g3 = _ell_from_golden(golden)
    x2, y2, x3 = golden["ell/x2"], golden["ell/y2"], golden["ell/x3"]
    for coord, bc in ((0, T.DIR), (1, T.PER), (0, T.NEU_DIR), (1, T.NEU)):
        for d in range(3):
            m = T.derivative(coord, g2, bc, d)
            for al, be in ((1., 0.), (-1., 1.), (0.5, -2.)):
                y = y2.copy()
                orc.ell_symv(m, al, x2, be, y)
                assert same_bits(y, golden[f"ell/d{coord}/bc{bc}/dir{d}/a{al}b{be}"]), (coord, bc, d, al, be)
        m = T.jump(coord, g2, bc)
        y = y2.copy()
        orc.ell_symv(m, 1., x2, 0., y)
        assert same_bits(y, golden[f"ell/j{coord}/bc{bc}"])
    for coord in range(3):
        m = T.derivative(coord, g3, g3.bc[coord], T.BACKWARD)
        y = np.full(g3.size, np.nan)  # beta == 0 must overwrite NaN (derivatives_t.cpp:134-156)
        orc.ell_symv(m, 1., x3, 0., y)
        assert same_bits(y, golden[f"ell/3d/d{coord}"])


def test_csr_fixtures(golden):
    pos, idx, val, x, y0 = (golden["csr/" + k] for k in ("pos", "idx", "val", "x", "y"))
    for al, be in ((1., 0.), (0.5, 1.), (-2., 0.25)):
        y = y0.copy()
        orc.csr_spmv(pos, idx, val, al, x, be, y)
        assert same_bits(y, golden[f"csr/a{al}b{be}"]), (al, be)


def _oracle_elliptic(T, g, bcx, bcy, d, jf, chi):
    mats = dict(leftx=T.derivative(0, g, T.inverse_bc(bcx), T.inverse_dir(d)),
                lefty=T.derivative(1, g, T.inverse_bc(bcy), T.inverse_dir(d)),
                rightx=T.derivative(0, g, bcx, d), righty=T.derivative(1, g, bcy, d),
                jumpx=T.jump(0, g, bcx), jumpy=T.jump(1, g, bcy))
    return orc.Elliptic2d(mats, sigma=chi.copy(), jfactor=jf)  # Cartesian: vol == 1 => sigma = chi*1


@pytest.mark.parametrize("tag,bcx,bcy,d,jf", [("dirper_fwd", 1, 0, 0, 1.0), ("neu_cen", 4, 0, 2, 0.1),
                                              ("dirneu_bwd", 2, 1, 1, 1.0)])
def test_elliptic_fixtures(golden, tag, bcx, bcy, d, jf):
    from feltor_b200 import topology as T
    g = T.Grid([0, 0], [np.pi, 2 * np.pi], 3, [10, 8], [bcx, bcy])
    chi = golden[f"elliptic/{tag}/chi"]
    E = _oracle_elliptic(T, g, bcx, bcy, d, jf, chi)
    x, y0 = golden[f"elliptic/{tag}/x"], golden[f"elliptic/{tag}/y0"]
    for al, be in ((1., 0.), (-0.5, 2.)):
        y = y0.copy()
        E.symv(al, x, be, y)
        assert same_bits(y, golden[f"elliptic/{tag}/a{al}b{be}"]), (tag, al, be)
    assert same_bits(g.weights(), golden[f"elliptic/{tag}/weights"])
    assert same_bits(1. / chi, golden[f"elliptic/{tag}/precond"])
    if tag == "dirper_fwd":
        b = golden[f"elliptic/{tag}/pcg_b"]
        xs = np.zeros(g.size)
        it = E.pcg_solve(xs, b, 1. / chi, g.weights(), 1e-8, 1.0, 1)
        assert it == int(golden[f"elliptic/{tag}/pcg_it"][0])
        assert same_bits(xs, golden[f"elliptic/{tag}/pcg_x"])


# ------------------------------------------------------------------------------------------------ live reference
def test_live_reference_random(ref):
    if ref is None:
        pytest.skip("oracle/_ref/libdgref.so not built (reference tree absent)")
    r = rng(7)
    RB = RefBlas1()
    for n in (1, 2, 63, 1000):
        v = [r.uniform(-3, 3, n) for _ in range(5)]
        for f in (lambda B, w: B.axpby(1.1, w[0], 0.3, w[1]), lambda B, w: B.axpbypgz(1.1, w[0], 0.3, w[1], -2., w[2]),
                  lambda B, w: B.pointwiseDot(1.1, w[0], w[1], 0.3, w[2], w[3], -2., w[4]),
                  lambda B, w: B.pointwiseDivide(1.1, w[0], w[1], 0.3, w[2])):
            a, b = [u.copy() for u in v], [u.copy() for u in v]
            f(OB, a)
            f(RB, b)
            assert same_bits(np.stack(a), np.stack(b))
        x, w, y = wide(r, n, -100, 100), wide(r, n), wide(r, n)
        a2, _ = ref.dot2(x, y)
        a3, _ = ref.dot3(x, w, y)
        assert np.array_equal(orc.normalize(a2), orc.exdot2(x, y)[0])
        assert np.array_equal(orc.normalize(a3), orc.exdot3(x, w, y)[0])
        assert same_bits([ref.round_acc(a3)], [orc.dot3(x, w, y)[0]])
