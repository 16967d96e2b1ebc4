"""Parity of the WALKER kernel (elliptic2d_walker_kernel, the kernel behind every headline number) against the oracle.

The automatic selection only picks the walker from ~400^2 cells on, so these tests pin the kernel with
dgb_elliptic2d_set_kernel(plan, DGB_ELLIPTIC_KERNEL_WALKER) and sweep every template / run-time variant on ragged grids the
oracle finishes in milliseconds: n in {2, 3} x five boundary-condition pairs x forward / backward / centered x (beta = 0,
beta != 0) x (volume form, none) x Helmholtz (chi, no chi) x the periodic-x seam (LDGSTS loader) x the fused dot of PCG,
all BITWISE against oracle/dgoracle.c; plus 512^2 / 1024^2 centered + Helmholtz cases against the live unmodified reference
(oracle/_ref/libdgref.so) where it has been built."""
import numpy as np
import pytest
from oracle import orc
from util import same_bits, rng

pytestmark = pytest.mark.gpu

GRIDS = ([37, 19], [61, 33], [33, 70])
BCS = ((0, 0), (1, 0), (4, 1), (2, 3), (3, 2))   # PER/DIR/NEU/DIR_NEU/NEU_DIR pairs; (0, *) has the periodic seam in x


@pytest.fixture(scope="module")
def G():
    import gpu_backend
    gpu_backend.require_library_loaded()
    return gpu_backend


def oracle_elliptic(T, g, bcx, bcy, d, jf, sigma, vol=None):
    mats = dict(leftx=T.derivative(0, g, T.inverse_bc(bcx), T.inverse_dir(d)),
                lefty=T.derivative(1, g, T.inverse_bc(bcy), T.inverse_dir(d)),
                rightx=T.derivative(0, g, bcx, d), righty=T.derivative(1, g, bcy, d),
                jumpx=T.jump(0, g, bcx), jumpy=T.jump(1, g, bcy))
    return orc.Elliptic2d(mats, sigma=sigma.copy(), vol=vol, jfactor=jf)


def walker(E, with_dot=False):
    E.set_kernel("walker")
    assert E.kernel(with_dot) == "walker"
    return E


CASES = [(n, N, bcx, bcy, d) for n in (2, 3) for N in GRIDS for (bcx, bcy) in BCS for d in (0, 1, 2)]


@pytest.mark.parametrize("n,N,bcx,bcy,d", CASES)
def test_walker_symv_vs_oracle(G, n, N, bcx, bcy, d):
    """plain and general epilogue (alpha, beta != 0), every bc family / direction, ragged grids; walker == tile == oracle"""
    from feltor_b200 import topology as T
    from feltor_b200.elliptic import Elliptic2d
    g = T.Grid([0, 0], [np.pi, 2 * np.pi], n, N, [bcx, bcy])
    r = rng(n * 1000 + N[0] * 10 + bcx * 3 + d)
    chi = 1. + 0.9 * r.uniform(0, 1, g.size)
    x, y0 = r.uniform(-1, 1, g.size), r.uniform(-1, 1, g.size)
    O = oracle_elliptic(T, g, bcx, bcy, d, 0.7, chi)
    E = walker(Elliptic2d(g, bcx, bcy, d, 0.7))
    E.set_chi(G.make(chi))
    for al, be in ((1., 0.), (-0.5, 2.)):
        yo = y0.copy()
        O.symv(al, x, be, yo)
        y = G.make(y0 if be != 0. else np.full(g.size, np.nan))
        E.symv(al, G.make(x), be, y)
        assert same_bits(G.get(y), yo), (n, N, bcx, bcy, d, al, be)
    # jfactor = 0 skips the jump terms altogether (elliptic.h:449)
    E.set_jfactor(0.)
    O0 = oracle_elliptic(T, g, bcx, bcy, d, 0., chi)
    yo = np.zeros(g.size)
    O0.symv(1., x, 0., yo)
    y = G.make(np.full(g.size, np.nan))
    E.symv(G.make(x), y)
    assert same_bits(G.get(y), yo)


@pytest.mark.parametrize("n,N,bcx,bcy,d", [c for c in CASES if c[1] != GRIDS[1]])
def test_walker_volume_form_vs_oracle(G, n, N, bcx, bcy, d):
    """curvilinear volume form: sigma = chi * vol, result divided by vol (elliptic.h:327,458), beta = 0 and beta != 0"""
    from feltor_b200 import topology as T
    from feltor_b200.elliptic import Elliptic2d
    g = T.Grid([0, 0], [np.pi, 2 * np.pi], n, N, [bcx, bcy])
    r = rng(n * 77 + N[1] + bcy * 5 + d)
    chi, vol = 1. + r.uniform(0, 1, g.size), 0.5 + r.uniform(0, 1, g.size)
    x, y0 = r.uniform(-1, 1, g.size), r.uniform(-1, 1, g.size)
    O = oracle_elliptic(T, g, bcx, bcy, d, 1.0, chi * vol, vol=vol)
    E = walker(Elliptic2d(g, bcx, bcy, d, 1.0))
    dvol = G.make(vol)
    E.set_vol(dvol)
    E.set_chi(G.make(chi))
    for al, be in ((1., 0.), (0.25, -3.)):
        yo = y0.copy()
        O.symv(al, x, be, yo)
        y = G.make(y0)
        E.symv(al, G.make(x), be, y)
        assert same_bits(G.get(y), yo), (n, N, bcx, bcy, d, al, be)


@pytest.mark.parametrize("n,N,bcx,bcy,d", [c for c in CASES if c[1] != GRIDS[2]])
def test_walker_helmholtz_vs_oracle(G, n, N, bcx, bcy, d):
    """GeneralHelmholtz epilogue (helmholtz.h:74-80): y = chi x - alpha (Elliptic x), chi = 1 and a chi field"""
    from feltor_b200 import topology as T
    from feltor_b200.elliptic import Elliptic2d
    from feltor_b200.toefl import Helmholtz
    g = T.Grid([0, 0], [3., 2.], n, N, [bcx, bcy])
    r = rng(n * 31 + N[0] + bcx + 7 * d)
    sig, hchi = 1. + 0.5 * r.uniform(0, 1, g.size), 1. + r.uniform(0, 1, g.size)
    x = r.uniform(-1, 1, g.size)
    O = oracle_elliptic(T, g, bcx, bcy, d, 1.0, sig)
    ex = np.zeros(g.size)
    O.symv(1., x, 0., ex)
    E = walker(Elliptic2d(g, bcx, bcy, d, 1.0))
    E.set_chi(G.make(sig))
    H = Helmholtz(-0.37, E)
    y = G.make(np.full(g.size, np.nan))
    H.symv(G.make(x), y)
    yo = ex.copy()
    orc.pointwiseDot(1., np.ones(g.size), x, 0.37, yo)
    assert same_bits(G.get(y), yo)
    H.set_chi(G.make(hchi))
    y = G.make(np.full(g.size, np.nan))
    H.symv(G.make(x), y)
    yo = ex.copy()
    orc.pointwiseDot(1., hchi, x, 0.37, yo)
    assert same_bits(G.get(y), yo)


PCG_CASES = [(3, [40, 24], 1, 0, 0, 1), (3, [33, 17], 4, 1, 2, 1), (2, [24, 40], 1, 1, 1, 3), (3, [37, 19], 0, 0, 0, 1),
             (2, [61, 33], 2, 3, 2, 1), (3, [33, 70], 3, 2, 1, 10), (3, [61, 33], 1, 0, 2, 1), (2, [37, 19], 0, 1, 0, 1)]


@pytest.mark.parametrize("n,N,bcx,bcy,d,tf", PCG_CASES)
def test_walker_pcg_vs_oracle(G, n, N, bcx, bcy, d, tf):
    """the fused-dot variant of the walker inside PCG: same iteration count and bit-identical solution"""
    from feltor_b200 import topology as T
    from feltor_b200.elliptic import Elliptic2d, PCG
    g = T.Grid([0, 0], [np.pi, 2 * np.pi], n, N, [bcx, bcy])
    chi = g.evaluate(lambda x, y: 1. + 0.9 * np.sin(x) * np.sin(y))
    b = g.evaluate(lambda x, y: np.sin(x) * np.sin(y) * (1 + np.cos(3 * y)))
    if bcx in (0, 4) and bcy in (0, 4):   # singular operator (constants in the kernel): make the right-hand side compatible
        b = g.evaluate(lambda x, y: np.cos(x) * np.sin(y))
    w = g.weights()
    O = oracle_elliptic(T, g, bcx, bcy, d, 1.0, chi)
    xo = np.zeros(g.size)
    ito = O.pcg_solve(xo, b, 1. / chi, w, 1e-9, 1.0, tf, max_iter=300)
    E = walker(Elliptic2d(g, bcx, bcy, d, 1.0), with_dot=True)
    E.set_chi(G.make(chi))
    x = G.make(np.zeros(g.size))
    pcg = PCG(g.size, 300)
    pcg.set_throw_on_fail(False)
    it = pcg.solve(E, x, G.make(b), E.precond(), E.weights(), 1e-9, 1.0, tf)
    assert it == ito
    assert same_bits(G.get(x), xo)


@pytest.mark.parametrize("n,N,d", [(3, [40, 24], 2), (2, [33, 40], 0)])
def test_walker_helmholtz_pcg_vs_unfused(G, n, N, d):
    """PCG on a Helmholtz plan (what toefl's gamma inversion runs): walker == tile == unfused bit for bit"""
    from feltor_b200 import topology as T
    from feltor_b200.elliptic import Elliptic2d, PCG
    from feltor_b200.toefl import Helmholtz
    g = T.Grid([0, 0], [3., 2.], n, N, [T.DIR, T.PER])
    r = rng(5)
    b = G.make(r.uniform(-1, 1, g.size))
    sols, its = [], []
    for kernel in ("walker", "tile", "unfused"):
        E = Elliptic2d(g, direction=d).set_kernel(kernel)
        H = Helmholtz(-0.5, E)
        x = G.make(np.zeros(g.size))
        pcg = PCG(g.size, 150)
        pcg.set_throw_on_fail(False)    # a fixed number of iterations is as good a comparison as a converged solve
        its.append(pcg.solve(H, x, b, H.precond(), H.weights(), 1e-10, 1.0, 1))
        sols.append(G.get(x))
        assert E.kernel(True) == kernel
    assert its[0] == its[1] == its[2] and its[0] > 0
    assert same_bits(sols[0], sols[1]) and same_bits(sols[0], sols[2])


@pytest.mark.parametrize("N,d", [(512, 2), (1024, 2), (1024, 0)])
def test_walker_full_size_vs_live_reference(G, ref, N, d):
    """benchmark-size grids, automatic kernel choice (= walker): plain and Helmholtz apply against the UNMODIFIED reference
    dg::Elliptic / pointwiseDot built in oracle/_ref (OpenMP backend), bit for bit"""
    if ref is None:
        pytest.skip("oracle/_ref/libdgref.so not present")
    import torch
    from feltor_b200 import topology as T
    from feltor_b200.elliptic import Elliptic2d
    from feltor_b200.toefl import Helmholtz
    g = T.Grid([0, 0], [np.pi, 2 * np.pi], 3, [N, N], [T.DIR, T.PER])
    r = rng(N + d)
    chi = 1. + 0.9 * r.uniform(0, 1, g.size)
    x = r.uniform(-1, 1, g.size)
    rg = ref.grid([0, 0], [np.pi, 2 * np.pi], 3, [N, N], [T.DIR, T.PER])
    R = ref.Elliptic2d(rg, T.DIR, T.PER, d, 1.0)
    R.set_chi(chi)
    yr = np.zeros(g.size)
    R.symv(1., x, 0., yr)
    E = Elliptic2d(g, T.DIR, T.PER, d, 1.0)
    assert E.kernel() == "walker" and E.kernel(True) == "walker"
    E.set_chi(G.make(chi))
    y = torch.full((g.size,), float("nan"), dtype=torch.float64, device="cuda")
    E.symv(G.make(x), y)
    assert same_bits(G.get(y), yr)
    hchi = 1. + r.uniform(0, 1, g.size)
    H = Helmholtz(-0.25, E)
    H.set_chi(G.make(hchi))
    H.symv(G.make(x), y)
    orc.pointwiseDot(1., hchi, x, 0.25, yr)   # helmholtz.h:79 on the reference's Elliptic result
    assert same_bits(G.get(y), yr)


def test_walker_slab_equals_global(G):
    """slab mode (multi-GPU partition) on the walker kernel at a small ragged size: every slab reproduces its rows"""
    import torch
    from feltor_b200 import topology as T
    from feltor_b200.elliptic import Elliptic2d
    from feltor_b200._lib import lib
    from feltor_b200._dev import ptr, stream
    import ctypes as C
    for d, bcy in ((0, 0), (2, 1), (1, 4)):
        g = T.Grid([0, 0], [np.pi, 2 * np.pi], 3, [37, 26], [T.DIR, bcy])
        r = rng(d)
        chi, x = 1. + r.uniform(0, 1, g.size), r.uniform(-1, 1, g.size)
        O = oracle_elliptic(T, g, T.DIR, bcy, d, 1.0, chi)
        yo = np.zeros(g.size)
        O.symv(1., x, 0., yo)
        row = 3 * 37 * 3                       # doubles per cell row
        ghost, Ny = 2, 26
        for (y0, rows) in ((0, 9), (9, 8), (17, 9)):
            E = walker(Elliptic2d(g, T.DIR, bcy, d, 1.0))
            lib().elliptic2d_set_slab(E.h, y0, rows, ghost)

            def padded(v):
                out = np.zeros((rows + 2 * ghost) * row)
                for k in range(-ghost, rows + ghost):
                    gy = y0 + k
                    if bcy == 0:
                        gy %= Ny
                    if 0 <= gy < Ny:
                        out[(k + ghost) * row:(k + ghost + 1) * row] = v[gy * row:(gy + 1) * row]
                return out
            xs, ss = G.make(padded(x)), G.make(padded(chi))
            lib().elliptic2d_set_sigma(E.h, C.c_void_p(ss.data_ptr() + ghost * row * 8))
            y = torch.full((rows * row,), float("nan"), dtype=torch.float64, device="cuda")
            lib().elliptic2d_symv(E.h, C.c_double(1.), C.c_void_p(xs.data_ptr() + ghost * row * 8), C.c_double(0.), ptr(y), stream())
            assert same_bits(G.get(y), yo[y0 * row:(y0 + rows) * row]), (d, bcy, y0)


@pytest.mark.parametrize("n,N,bcx,bcy,d", [(3, [130, 33], 1, 0, 0), (3, [126, 40], 1, 0, 2), (2, [97, 70], 1, 1, 1), (3, [420, 404], 1, 0, 0),
                                          (3, [404, 420], 4, 0, 2), (3, [33, 70], 2, 3, 1)])
def test_walker_relaxed_ordering_within_tolerance(G, n, N, bcx, bcy, d):
    """DGB_ORDER_RELAXED (opt-in): the interior rows accumulate every output in one FMA chain.  Not bit-identical by design;
    the north star's bound for symv is 1e-12 relative -- the kernel stays below 1e-13 of the result's norm, element-wise
    differences are a few ulp of the largest term, and PCG reaches the same solution."""
    from feltor_b200 import topology as T
    from feltor_b200.elliptic import Elliptic2d, PCG
    g = T.Grid([0, 0], [np.pi, 2 * np.pi], n, N, [bcx, bcy])
    r = rng(n + N[0] + d)
    chi, x = 1. + 0.9 * r.uniform(0, 1, g.size), r.uniform(-1, 1, g.size)
    O = oracle_elliptic(T, g, bcx, bcy, d, 0.7, chi)
    yo = np.zeros(g.size)
    O.symv(1., x, 0., yo)
    E = walker(Elliptic2d(g, bcx, bcy, d, 0.7)).set_ordering("relaxed")
    E.set_chi(G.make(chi))
    y = G.make(np.full(g.size, np.nan))
    E.symv(G.make(x), y)
    y = G.get(y)
    assert np.linalg.norm(y - yo) <= 1e-13 * np.linalg.norm(yo)
    assert np.abs(y - yo).max() <= 1e-12 * np.abs(yo).max()
    if N[0] >= 90:   # grids with at least one warp strip of interior columns take the relaxed path:
        assert not same_bits(y, yo)                       # it really is another rounding sequence
    E.set_ordering("reference")
    y = G.make(np.full(g.size, np.nan))
    E.symv(G.make(x), y)
    assert same_bits(G.get(y), yo)                        # and the default is restored bit for bit
    # PCG on the relaxed operator converges to the reference's solution
    b = g.evaluate(lambda X, Y: np.sin(X) * np.sin(Y) * (1 + np.cos(3 * Y)))
    if (bcx in (0, 4) and bcy in (0, 4)) or g.size > 60000:   # the single-threaded oracle solve stays within seconds
        return
    xo = np.zeros(g.size)
    O1 = oracle_elliptic(T, g, bcx, bcy, d, 0.7, chi)
    ito = O1.pcg_solve(xo, b, 1. / chi, g.weights(), 1e-7, 1.0, 1, max_iter=5000)
    E.set_ordering("relaxed")
    xs = G.make(np.zeros(g.size))
    pcg = PCG(g.size, 5000)
    it = pcg.solve(E, xs, G.make(b), E.precond(), E.weights(), 1e-7, 1.0, 1)
    assert 0 < ito < 5000 and abs(it - ito) <= max(5, ito // 20)
    assert np.linalg.norm(G.get(xs) - xo) <= 1e-5 * np.linalg.norm(xo)


FOLD_CASES = [(3, [40, 24], 1, 0, 0, 1, False), (3, [40, 33], 4, 1, 2, 1, False), (2, [61, 33], 2, 3, 1, 3, False),
              (3, [62, 45], 3, 2, 2, 1, True), (2, [36, 30], 1, 0, 0, 1, True), (3, [34, 70], 1, 4, 1, 10, False)]


@pytest.mark.parametrize("n,N,bcx,bcy,d,tf,helm", FOLD_CASES)
def test_walker_pcg_folded_direction_update(G, monkeypatch, n, N, bcx, bcy, d, tf, helm):
    """PCG with the direction update p = z + beta p folded into the loader of the walker kernel (two launches per iteration)
    against the three-kernel iteration (DGB_PCG_NO_FOLD=1) and the oracle: same iteration count, same bits, fewer launches"""
    from feltor_b200 import topology as T
    from feltor_b200.elliptic import Elliptic2d, PCG
    from feltor_b200.toefl import Helmholtz
    from feltor_b200._lib import lib
    g = T.Grid([0, 0], [np.pi, 2 * np.pi], n, N, [bcx, bcy])
    chi = g.evaluate(lambda x, y: 1. + 0.9 * np.sin(x) * np.sin(y))
    b = g.evaluate(lambda x, y: np.sin(x) * np.sin(y) * (1 + np.cos(3 * y)))
    if bcx == 4 and bcy in (0, 4):
        b = g.evaluate(lambda x, y: np.cos(x) * np.sin(y))
    res = {}
    for mode in ("fold", "nofold"):
        if mode == "nofold":
            monkeypatch.setenv("DGB_PCG_NO_FOLD", "1")
        else:
            monkeypatch.delenv("DGB_PCG_NO_FOLD", raising=False)
        E = walker(Elliptic2d(g, bcx, bcy, d, 1.0), with_dot=True)
        E.set_chi(G.make(chi))
        A = Helmholtz(-0.5, E) if helm else E
        x = G.make(np.zeros(g.size))
        pcg = PCG(g.size, 300)
        pcg.set_throw_on_fail(False)
        l0 = lib().raw["dgb_launch_count"]()
        it = pcg.solve(A, x, G.make(b), A.precond(), A.weights(), 1e-9, 1.0, tf)
        res[mode] = (it, G.get(x), lib().raw["dgb_launch_count"]() - l0)
    assert res["fold"][0] == res["nofold"][0] and res["fold"][0] > 3
    assert same_bits(res["fold"][1], res["nofold"][1])
    assert res["fold"][2] < res["nofold"][2] - (res["fold"][0] - 3), "the folded iteration did not run"
    if not helm:
        O = oracle_elliptic(T, g, bcx, bcy, d, 1.0, chi)
        xo = np.zeros(g.size)
        ito = O.pcg_solve(xo, b, 1. / chi, g.weights(), 1e-9, 1.0, tf, max_iter=300)
        assert ito == res["fold"][0] and same_bits(res["fold"][1], xo)


@pytest.mark.parametrize("N,bcy,d", [([40, 24], 0, 0), ([38, 31], 1, 2), ([44, 26], 0, 1)])
def test_walker_slab_pcg_folded(G, N, bcy, d):
    """the folded iteration in slab mode (communicator of size 1: ghost rows of z and of both direction buffers, ring-closing
    halo copy of z) against the plain single-GPU solve"""
    from feltor_b200 import topology as T
    from feltor_b200.elliptic import Elliptic2d, PCG
    from feltor_b200.dist import Comm, SlabElliptic2d, DistPCG
    from feltor_b200._lib import lib
    g = T.Grid([0, 0], [np.pi, 2 * np.pi], 3, N, [T.DIR, bcy])
    chi = g.evaluate(lambda x, y: 1. + 0.9 * np.sin(x) * np.sin(y))
    b = g.evaluate(lambda x, y: np.sin(x) * np.sin(y) * (1 + np.cos(3 * y)))
    E = Elliptic2d(g, T.DIR, bcy, d, 1.0).set_kernel("tile")
    E.set_chi(G.make(chi))
    x = G.make(np.zeros(g.size))
    it = PCG(g.size, g.size).solve(E, x, G.make(b), E.precond(), E.weights(), 1e-9, 1.0, 1)
    comm = Comm(0, 1)
    S = SlabElliptic2d(comm, g, T.DIR, bcy, d, 1.0)
    lib().elliptic2d_set_kernel(S.h, 2)
    S.set_chi(G.make(chi))
    xs = G.make(np.zeros(g.size))
    l0 = lib().raw["dgb_launch_count"]()
    its = DistPCG(comm, g.size, g.size).solve(S, xs, G.make(b), S.precond(), S.weights(), 1e-9, 1.0, 1)
    launches = lib().raw["dgb_launch_count"]() - l0
    assert its == it and it > 5
    assert same_bits(G.get(xs), G.get(x))
    assert launches < 5 * its + 30, "expected K1 + scalar + K2 + scalar per iteration (no direction kernel)"
