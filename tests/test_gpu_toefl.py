"""toefl (config 3 of BASELINE.json): the right-hand side toefl::Explicit, its building blocks and fixed-step
dg::ERKStep on libdgb200.so against the UNMODIFIED reference (src/toefl/toefl.h, OpenMP backend): committed golden
vectors (tests/golden/toefl_golden.npz, generator beside it) and, when oracle/_ref/libdgref_toefl.so travelled along,
the live reference on other sizes.  Everything is compared BITWISE: state, potentials, per-stage PCG iteration counts.
(CG amplifies a 1-ulp perturbation of its input to ~1e-8 of the solution within a few dozen iterations -- it is the bit
equality of every building block that makes the 1e-12 bar of the north star reachable at all.)"""
import ctypes as C
import os
import numpy as np
import pytest
from util import same_bits, rng

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def G():
    import gpu_backend
    gpu_backend.require_library_loaded()
    return gpu_backend


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(ROOT, "tests", "golden", "toefl_golden.npz"))


@pytest.fixture(scope="module")
def reft():
    from oracle import reftoefl
    return reftoefl if reftoefl.available() else None


def params(n, N, model):
    from oracle import reftoefl
    return reftoefl.default_params(n, N, N, model__type=model)


def run_steps(G, js, y0, y1, steps, dt=0.5):
    import torch
    from feltor_b200 import toefl as TF
    ex = TF.Explicit(TF.Parameters(js))
    u0 = [G.make(y0), G.make(y1)]
    u1 = [torch.zeros_like(u0[0]), torch.zeros_like(u0[0])]
    delta = [torch.zeros_like(u0[0]), torch.zeros_like(u0[0])]
    erk = TF.ERKStep("Bogacki-Shampine-4-2-3", u0)
    t = 0.
    for _ in range(steps):
        t = erk.step(ex, t, u0, u1, dt, delta)
        u0, u1 = u1, u0
    return ex, u0


@pytest.mark.parametrize("model", ["global", "local"])
def test_toefl_steps_vs_golden(G, gold, model):
    js = params(3, 24, model)
    ex, u = run_steps(G, js, gold[model + "_init0"], gold[model + "_init1"], 3)
    assert same_bits(G.get(ex.binv), gold[model + "_binv"])
    assert same_bits(G.get(u[0]), gold[model + "_y0"])
    assert same_bits(G.get(u[1]), gold[model + "_y1"])
    assert same_bits(G.get(ex.phi[0]), gold[model + "_phi0"])
    assert same_bits(G.get(ex.phi[1]), gold[model + "_phi1"])
    assert ex.ncalls == 10  # FSAL: 4 + 3 + 3 right-hand-side evaluations


@pytest.mark.parametrize("case,dt0,nsteps", [("adaptA", 1e-6, 8), ("adaptB", 60., 4)])
def test_toefl_adaptive_vs_golden(G, gold, case, dt0, nsteps):
    """dg::Adaptive<ERKStep> with pid_control / l2norm as src/toefl/toefl.cpp:88-91 drives it (adaptive.h:232-395):
    state, every proposed step size, the time and the number of rejected steps are bit-identical to the reference's"""
    from feltor_b200 import toefl as TF
    js = params(3, 24, "global")
    ex = TF.Explicit(TF.Parameters(js))
    u = [G.make(gold["global_init0"]), G.make(gold["global_init1"])]
    adapt = TF.Adaptive("Bogacki-Shampine-4-2-3", u)
    t, dt, dts = 0., dt0, []
    for _ in range(nsteps):
        t, dt = adapt.step(ex, t, u, u, dt, TF.pid_control, TF.l2norm, 1e-5, 1e-6)
        dts.append(dt)
    assert same_bits(np.array(dts), gold[case + "_dts"]), (dts, gold[case + "_dts"])
    assert t == gold[case + "_t_nfailed"][0] and adapt.nfailed == int(gold[case + "_t_nfailed"][1])
    assert same_bits(G.get(u[0]), gold[case + "_y0"]) and same_bits(G.get(u[1]), gold[case + "_y1"])


@pytest.mark.parametrize("case,tableau,dt,nsteps", [("msAB", "AB-3-3", 0.5, 6), ("msTVB", "TVB-3-3", 0.3, 5)])
def test_toefl_multistep_vs_golden(G, gold, case, tableau, dt, nsteps):
    """dg::ExplicitMultistep (multistep.h:59-100, :592-639) incl. its Shu-Osher start-up steps (runge_kutta.h:883-910):
    state, step times and the number of right-hand-side calls bit-identical to the reference's"""
    from feltor_b200 import toefl as TF
    ex = TF.Explicit(TF.Parameters(params(3, 24, "global")))
    u = [G.make(gold["global_init0"]), G.make(gold["global_init1"])]
    ms = TF.ExplicitMultistep(tableau, u)
    t, ts = 0., []
    ms.init(ex, t, u, dt)
    for _ in range(nsteps):
        t = ms.step(ex, t, u)
        ts.append(t)
    assert same_bits(np.array(ts), gold[case + "_ts"])
    assert ex.ncalls == int(gold[case + "_ncalls"][0])
    assert same_bits(G.get(u[0]), gold[case + "_y0"]) and same_bits(G.get(u[1]), gold[case + "_y1"])


@pytest.mark.parametrize("model", ["global", "local"])
def test_toefl_rhs_vs_golden(G, gold, model):
    import torch
    from feltor_b200 import toefl as TF
    ex = TF.Explicit(TF.Parameters(params(3, 24, model)))
    y = [G.make(gold[model + "_y0"]), G.make(gold[model + "_y1"])]
    yp = [torch.zeros_like(y[0]), torch.zeros_like(y[0])]  # Advection::upwind scales its output by beta = 0 (NaN would stay, as in the reference)
    ex(0., y, yp)
    assert same_bits(G.get(yp[0]), gold[model + "_rhs0"])
    assert same_bits(G.get(yp[1]), gold[model + "_rhs1"])
    assert same_bits(G.get(ex.phi[0]), gold[model + "_rhsphi0"])
    assert same_bits(G.get(ex.phi[1]), gold[model + "_rhsphi1"])


def test_toefl_building_blocks_vs_live_reference(G, reft):
    """Advection::upwind, Elliptic::variation, the Helmholtz and the chi-weighted polarisation multigrid solves"""
    if reft is None:
        pytest.skip("oracle/_ref/libdgref_toefl.so not present")
    import torch
    from feltor_b200 import toefl as TF
    from feltor_b200._lib import lib
    from feltor_b200._dev import ptr, stream
    js = params(3, 40, "global")
    ref = reft.RefToefl(js)
    ex = TF.Explicit(TF.Parameters(js))
    n, r = ref.size, rng(7)
    assert same_bits(G.get(ex.binv), ref.binv())
    f, vx, vy, res = (r.uniform(-1, 1, n) for _ in range(4))
    out = G.make(res)
    ex.adv.upwind(-1., G.make(vx), G.make(vy), G.make(f), 0.5, out)
    assert same_bits(G.get(out), ref.upwind(-1., vx, vy, f, 0.5, res))
    ar = TF.ArakawaX(ex.grid)
    out = G.make(res)
    ar(0.7, G.make(f), G.make(vx), -0.4, out)
    assert same_bits(G.get(out), ref.arakawa(0.7, f, vx, -0.4, res))
    phi = ex.grid.evaluate(lambda x, y: np.sin(0.05 * x) * np.cos(0.03 * y))
    u = torch.zeros(n, dtype=torch.float64, device="cuda")  # Axpby(alpha, 0) scales the old value: NaN would stay, as in the reference
    lib().elliptic2d_variation(ex.multi_pol[0].h, C.c_double(1.), None, ptr(G.make(phi)), C.c_double(0.), ptr(u), stream())
    assert same_bits(G.get(u), ref.variation(phi))
    b = ex.grid.evaluate(lambda x, y: np.exp(-((x - 60) ** 2 + (y - 100) ** 2) / 200.))
    x = G.make(np.zeros(n))
    num = ex.multigrid.solve(ex.multi_gamma1, x, G.make(b), ex.p.eps_gamma)
    xr, numr = ref.helmholtz_solve(np.zeros(n), b)
    assert num == numr and same_bits(G.get(x), xr)
    chi = 1. + b
    mc = ex.multigrid.project(G.make(chi))
    for k in range(3):
        ex.multi_pol[k].set_chi(mc[k])
    x = G.make(np.zeros(n))
    num = ex.multigrid.solve(ex.multi_pol, x, G.make(b), ex.p.eps_pol)
    xr, numr = ref.pol_solve(chi, np.zeros(n), b)
    assert num == numr and same_bits(G.get(x), xr)


@pytest.mark.parametrize("N,steps", [(48, 4), (64, 2), (416, 1)])
def test_toefl_steps_vs_live_reference(G, reft, N, steps):
    """N = 416: the finest multigrid stage is large enough for the automatic choice of the WALKER kernel, so its Helmholtz and
    centered polarisation solves run the folded two-kernel PCG iteration (direction update in the TMA loader) inside the
    nested iteration -- still bit for bit the unmodified reference"""
    if reft is None:
        pytest.skip("oracle/_ref/libdgref_toefl.so not present")
    js = params(3, N, "global")
    ref = reft.RefToefl(js)
    y0, y1 = ref.init()
    ra, rb, _ = ref.erk("Bogacki-Shampine-4-2-3", 0., 0.5, steps, y0, y1)
    ex, u = run_steps(G, js, y0, y1, steps)
    assert same_bits(G.get(u[0]), ra) and same_bits(G.get(u[1]), rb)
    assert same_bits(G.get(ex.phi[0]), ref.phi(0)) and same_bits(G.get(ex.phi[1]), ref.phi(1))
    # the initial condition built on the device agrees to rounding (host exp / contraction differ in the last bit)
    yi = ex.initial_condition()
    # (the gamma_inv image of that last bit grows with the resolution: the discrete Laplacian scales like (n N / lx)^2)
    assert np.abs(G.get(yi[0]) - y0).max() <= 4e-16 and np.abs(G.get(yi[1]) - y1).max() <= (2e-15 if N <= 64 else 1e-12)


def test_helmholtz_symv_fused_equals_composition(G):
    """the walker's Helmholtz epilogue == Elliptic symv followed by pointwiseDot(1, chi, x, -alpha, y), with and without chi"""
    import torch
    from feltor_b200 import blas1, topology as T
    from feltor_b200.elliptic import Elliptic2d
    from feltor_b200.toefl import Helmholtz
    g = T.Grid([0., 0.], [3., 2.], 3, [36, 28], [T.DIR, T.PER])
    r = rng(3)
    x, chi = G.make(r.uniform(-1, 1, g.size)), G.make(1. + r.uniform(0, 1, g.size))
    for direction in (T.FORWARD, T.CENTERED):
        E = Elliptic2d(g, direction=direction)
        ref_y = torch.empty_like(x)
        E.symv(x, ref_y)
        H = Helmholtz(-0.37, Elliptic2d(g, direction=direction))
        y = torch.full_like(x, float("nan"))
        H.symv(x, y)
        expect = ref_y.clone()
        blas1.axpby(1., x, 0.37, expect)
        assert same_bits(G.get(y), G.get(expect))
        H.set_chi(chi)
        H.symv(x, y)
        expect = ref_y.clone()
        blas1.pointwiseDot(1., chi, x, 0.37, expect)
        assert same_bits(G.get(y), G.get(expect))


@pytest.mark.parametrize("n,N,bcx,bcy", [(3, [37, 19], 1, 0), (2, [24, 40], 0, 1), (3, [33, 17], 4, 2), (4, [12, 15], 3, 0), (3, [5, 6], 1, 1),
                                         (3, [400, 300], 1, 0)])
def test_advection_upwind_fused(G, n, N, bcx, bcy):
    """dgb_advection_upwind (one kernel) == the reference's sequence of four Ell symv and two evaluate( Axpby, UpwindProduct)
    (advection.h:112-120) bit for bit: interior cells (constant-bank blocks), boundary block rows of every boundary condition,
    the periodic wrap, alpha / beta variants, velocities of both signs and exact zeros"""
    import torch
    from feltor_b200 import topology as T
    from feltor_b200 import toefl as TF
    g = T.Grid([0, 0], [3., 2.], n, N, [bcx, bcy])
    r = rng(n + N[0] + bcx)
    f, vx, vy, r0 = (r.uniform(-1, 1, g.size) for _ in range(4))
    vx[::7] = 0.
    vy[3::11] = 0.
    adv = TF.Advection(g)
    for alpha, beta in ((-1., 0.), (0.7, 1.), (1.3, -0.5)):
        a, b = G.make(r0), G.make(r0)
        adv.upwind(alpha, G.make(vx), G.make(vy), G.make(f), beta, a, fused=False)
        df, dvx, dvy = G.make(f), G.make(vx), G.make(vy)
        adv.upwind(alpha, dvx, dvy, df, beta, b, fused=True)
        assert adv._fused is True
        assert same_bits(G.get(a), G.get(b)), (alpha, beta)


@pytest.mark.parametrize("n,N,bcx,bcy", [(3, [37, 19], 1, 0), (2, [24, 40], 0, 1), (3, [33, 17], 4, 2), (4, [12, 15], 3, 0), (3, [5, 6], 1, 1),
                                         (3, [3, 3], 0, 0), (3, [400, 300], 1, 0)])
def test_arakawa_fused(G, n, N, bcx, bcy):
    """dgb_arakawa (two kernels) == the reference's ArakawaX::operator() sequence (arakawa.h:147-162: four symv, ArakawaFunctor, two
    symv with beta = 1, pointwiseDot) bit for bit: interior cells, boundary block rows of every boundary condition, the periodic
    wrap, alpha / beta variants, a non-trivial chi, the three-argument form and result aliasing lhs"""
    import torch
    from feltor_b200 import topology as T
    from feltor_b200 import toefl as TF
    g = T.Grid([0, 0], [3., 2.], n, N, [bcx, bcy])
    r = rng(2 * n + N[1] + bcy)
    lhs, rhs, r0 = (r.uniform(-1, 1, g.size) for _ in range(3))
    ar = TF.ArakawaX(g)
    ar.chi = G.make(r.uniform(0.5, 2., g.size))
    for alpha, beta in ((1., 0.), (1., 1.), (-0.7, 0.3)):
        a, b = G.make(r0), G.make(r0)
        ar(alpha, G.make(lhs), G.make(rhs), beta, a, fused=False)
        ar(alpha, G.make(lhs), G.make(rhs), beta, b, fused=True)
        assert ar._fused is True
        assert same_bits(G.get(a), G.get(b)), (alpha, beta)
    a, b = G.make(r0), G.make(r0)
    ar(G.make(lhs), G.make(rhs), a, fused=False)
    ar(G.make(lhs), G.make(rhs), b, fused=True)
    assert same_bits(G.get(a), G.get(b))
    la, lb = G.make(lhs), G.make(lhs)      # result aliases lhs (legal in the reference: lhs is last read by the functor)
    ar(1., la, G.make(rhs), 0.5, la, fused=False)
    ar(1., lb, G.make(rhs), 0.5, lb, fused=True)
    assert same_bits(G.get(la), G.get(lb))
