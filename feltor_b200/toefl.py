"""The toefl right-hand side on libdgb200.so: mirror of toefl::Explicit (src/toefl/toefl.h:8-310) and of the dg classes
it is built from -- dg::Helmholtz (inc/dg/helmholtz.h:27-82), dg::Advection (inc/dg/advection.h:60-120),
dg::Extrapolation (inc/dg/extrapolation.h:225-460), the explicit Runge-Kutta stepper dg::ERKStep
(inc/dg/runge_kutta.h:300-400) and dg::Adaptive with its controllers (inc/dg/adaptive.h:22-395) -- same member names, same call sequence, every arithmetic step a dgb_* call.
Like the rest of feltor_b200/*.py this is harness code over the C ABI (tests, benchmarks), not the product.

Models implemented: "global" (the default input of the reference) and "local".
"""
import ctypes as C
import math
import numpy as np
import torch
from ._lib import lib, DgbError
from ._dev import ptr, stream, dvec
from . import blas1, blas2, topology as T
from .elliptic import Elliptic2d, MultigridCG2d

d = C.c_double


def _zeros(n):
    return torch.zeros(n, dtype=torch.float64, device="cuda")


class Helmholtz:
    """dg::GeneralHelmholtz<dg::Elliptic2d> (helmholtz.h:27-82): chi x - alpha Elliptic x; usable wherever the solvers
    take an Elliptic2d (dgb_elliptic2d_set_helmholtz switches the plan's two-operand symv)."""

    def __init__(self, alpha, elliptic):
        self.E, self.alpha = elliptic, alpha
        self.h = elliptic.h
        self.size = elliptic.size
        self._chi = None
        lib().elliptic2d_set_helmholtz(self.h, 1, d(alpha), None)

    def weights(self):
        return self.E.weights()

    def precond(self):
        return self.E.precond()

    def set_chi(self, chi):
        self._chi = chi.clone()
        lib().elliptic2d_set_helmholtz(self.h, 1, d(self.alpha), ptr(self._chi))

    def symv(self, x, y):
        self.E.symv(x, y)


class Advection:
    """dg::Advection (advection.h:60-120): upwind discretisation of v . grad f"""

    def __init__(self, g, bcx=None, bcy=None):
        bcx = g.bc[0] if bcx is None else bcx
        bcy = g.bc[1] if bcy is None else bcy
        self.dxf = T.derivative(0, g, bcx, T.FORWARD)
        self.dyf = T.derivative(1, g, bcy, T.FORWARD)
        self.dxb = T.derivative(0, g, bcx, T.BACKWARD)
        self.dyb = T.derivative(1, g, bcy, T.BACKWARD)
        self.t0, self.t1 = _zeros(g.size), _zeros(g.size)
        self._fused = None

    def upwind(self, alpha, vx, vy, f, beta, result, fused=True):
        """one kernel (dgb_advection_upwind) when the library recognises the four matrices, else the reference's sequence"""
        n = f.numel()
        if fused and self._fused is not False:
            try:
                lib().advection_upwind(self.dxb.handle, self.dxf.handle, self.dyb.handle, self.dyf.handle, d(alpha), ptr(vx), ptr(vy),
                                       ptr(f), d(beta), ptr(result), stream())
                self._fused = True
                return
            except DgbError as e:
                if e.code != -2 or self._fused:   # DGB_ERR_UNSUPPORTED: fall back for good
                    raise
                self._fused = False
        self.dxb.symv(1., f, 0., self.t0)
        self.dxf.symv(1., f, 0., self.t1)
        lib().upwind_axpby(n, d(alpha), ptr(vx), ptr(self.t0), ptr(self.t1), d(beta), ptr(result), stream())
        self.dyb.symv(1., f, 0., self.t0)
        self.dyf.symv(1., f, 0., self.t1)
        lib().upwind_axpby(n, d(alpha), ptr(vy), ptr(self.t0), ptr(self.t1), d(1.), ptr(result), stream())


class ArakawaX:
    """dg::ArakawaX (arakawa.h:30-170): Poisson bracket {lhs, rhs} on a Cartesian grid (perp volume 1)"""

    def __init__(self, g, bcx=None, bcy=None):
        bcx = g.bc[0] if bcx is None else bcx
        bcy = g.bc[1] if bcy is None else bcy
        n = g.size
        self.work = _zeros(3 * n)   # m_dylhs, m_dxrhs, m_dyrhs of the class, contiguous for dgb_arakawa
        self.dylhs, self.dxrhs, self.dyrhs = self.work[:n], self.work[n:2 * n], self.work[2 * n:]
        self.dxlhs = _zeros(n)
        self.bdxf = T.derivative(0, g, bcx, T.CENTERED)
        self.bdyf = T.derivative(1, g, bcy, T.CENTERED)
        self.chi = torch.ones(n, dtype=torch.float64, device="cuda")  # 1 / perp_vol
        self._fused = None

    def __call__(self, *a, fused=True):
        """(lhs, rhs, result) | (alpha, lhs, rhs, beta, result): two kernels (dgb_arakawa) when the library recognises the
        matrices, else the reference's sequence of eight launches"""
        alpha, lhs, rhs, beta, result = (1., a[0], a[1], 0., a[2]) if len(a) == 3 else a
        if fused and self._fused is not False:
            try:
                lib().arakawa(self.bdxf.handle, self.bdyf.handle, d(alpha), ptr(lhs), ptr(rhs), ptr(self.chi), d(beta), ptr(result),
                              ptr(self.work), stream())
                self._fused = True
                return
            except DgbError as e:
                if e.code != -2 or self._fused:   # DGB_ERR_UNSUPPORTED: fall back for good
                    raise
                self._fused = False
        self.bdxf.symv(1., lhs, 0., self.dxlhs)
        self.bdyf.symv(1., lhs, 0., self.dylhs)
        self.bdxf.symv(1., rhs, 0., self.dxrhs)
        self.bdyf.symv(1., rhs, 0., self.dyrhs)
        lib().arakawa_functor(lhs.numel(), ptr(lhs), ptr(rhs), ptr(self.dxlhs), ptr(self.dylhs), ptr(self.dxrhs), ptr(self.dyrhs), stream())
        self.bdxf.symv(1., self.dylhs, 1., self.dyrhs)
        self.bdyf.symv(1., self.dxrhs, 1., self.dyrhs)
        blas1.pointwiseDot(alpha, self.chi, self.dyrhs, beta, result)


class Extrapolation:
    """dg::Extrapolation (extrapolation.h:225-460): polynomial through up to `max` past solutions"""

    def __init__(self, max_, copyable):
        self.max, self.counter = max_, 0
        self.x = [copyable.clone() for _ in range(max_)]
        self.t = [0.] * max_

    def extrapolate(self, t, new_x):
        c, tt, x = self.counter, self.t, self.x
        if c == 0:
            blas1.copy(0., new_x)
        elif c == 1:
            blas1.copy(x[0], new_x)
        elif c == 3:
            raise NotImplementedError("parabolic extrapolation is not used by toefl")
        else:
            f0 = (t - tt[1]) / (tt[0] - tt[1])
            f1 = (t - tt[0]) / (tt[1] - tt[0])
            blas1.axpby(f0, x[0], f1, x[1], new_x)

    def update(self, t_new, new_entry):
        if self.max == 0:
            return
        for i in range(self.counter):
            if abs(t_new - self.t[i]) < 1e-14:
                blas1.copy(new_entry, self.x[i])
                return
        if self.counter < self.max:
            self.counter += 1
        self.x = [self.x[-1]] + self.x[:-1]  # std::rotate by one towards the back
        self.t = [self.t[-1]] + self.t[:-1]
        self.t[0] = t_new
        blas1.copy(new_entry, self.x[0])


class Parameters:
    """toefl::Parameters (src/toefl/parameters.h) from the same JSON dictionary"""

    def __init__(self, js):
        g = js["grid"]
        self.n, self.Nx, self.Ny, self.lx, self.ly = g["n"], g["Nx"], g["Ny"], float(g["lx"]), float(g["ly"])
        e = js["elliptic"]
        self.num_stages = e["stages"]
        self.eps_pol = [float(e["eps_pol"][0])] + [float(v) * float(e["eps_pol"][0]) for v in e["eps_pol"][1:self.num_stages]]
        self.eps_gamma = [float(e["eps_gamma"][0])] + [float(v) * float(e["eps_gamma"][0]) for v in e["eps_gamma"][1:self.num_stages]]
        self.pol_dir = {"forward": T.FORWARD, "backward": T.BACKWARD, "centered": T.CENTERED}[e["direction"]]
        self.diff_dir = T.CENTERED
        i = js["init"]
        self.amp, self.sigma, self.posX, self.posY = float(i["amplitude"]), float(i["sigma"]), float(i["posX"]), float(i["posY"])
        self.flr = i.get("flr", "none")
        bcs = {"PER": T.PER, "DIR": T.DIR, "NEU": T.NEU, "DIR_NEU": T.DIR_NEU, "NEU_DIR": T.NEU_DIR}
        self.bcx, self.bcy = bcs[js["bc"][0]], bcs[js["bc"][1]]
        m = js["model"]
        self.model = m.get("type", "global")
        self.nu = float(m["nu"])
        self.boussinesq, self.tau, self.friction, self.kappa = False, 0., 0., 0.
        if self.model in ("local", "global"):
            self.kappa, self.tau = float(m["curvature"]), float(m["tau"])
            if self.model == "global":
                self.boussinesq = bool(m["boussinesq"])
        else:
            raise NotImplementedError("toefl model %r" % self.model)


class Explicit:
    """toefl::Explicit<CartesianGrid2d, DMatrix, DVec> (src/toefl/toefl.h)"""

    def __init__(self, p):
        self.p = p
        g = T.Grid([0., 0.], [p.lx, p.ly], p.n, [p.Nx, p.Ny], [p.bcx, p.bcy])
        self.grid, n = g, g.size
        self.chi, self.omega, self.uE2 = _zeros(n), _zeros(n), _zeros(n)
        # dg::LinearX (toefl.h:63, functors.h): a*x + b, which the reference's host compiler contracts to fma(a, x, b);
        # Fraction arithmetic is exact and float() rounds once, i.e. this IS the fused result
        from fractions import Fraction
        a, b = Fraction(p.kappa), Fraction(1. - p.kappa * p.posX * p.lx)
        line = np.array([float(a * Fraction(float(x)) + b) for x in g.abscissas(0)])
        self.binv = dvec(np.ascontiguousarray(np.broadcast_to(line, (g.shape(1), g.shape(0))).reshape(-1)))
        self.phi = [_zeros(n), _zeros(n)]
        self.dxphi, self.dyphi = [_zeros(n), _zeros(n)], [_zeros(n), _zeros(n)]
        self.ype, self.lapy, self.v = [_zeros(n), _zeros(n)], [_zeros(n), _zeros(n)], [_zeros(n), _zeros(n)]
        self.gamma_n = _zeros(n)
        self.laplaceM = Elliptic2d(g, direction=p.diff_dir)
        self.adv = Advection(g)
        self.multigrid = MultigridCG2d(g, p.num_stages)
        self.old_phi, self.old_psi, self.old_gammaN = (Extrapolation(2, self.chi) for _ in range(3))
        self.multi_chi = self.multigrid.project(self.chi)
        self.multi_pol = [Elliptic2d(self.multigrid.grid(u), direction=p.pol_dir, jfactor=1.) for u in range(p.num_stages)]
        self.multi_gamma1 = [Helmholtz(-0.5 * p.tau, Elliptic2d(self.multigrid.grid(u), direction=p.pol_dir))
                             for u in range(p.num_stages)]
        self.centered = [T.derivative(0, g, p.bcx, T.CENTERED), T.derivative(1, g, p.bcy, T.CENTERED)]
        self.ncalls = 0
        self.numbers = {}  # PCG iterations per stage of the last solves: "gammaN", "pol", "gammaPhi"

    def gamma_inv(self):
        return self.multi_gamma1[0]

    def initial_condition(self):
        """src/toefl/toefl.cpp:50-72"""
        p, g = self.p, self.grid
        x0, y0, s = p.posX * p.lx, p.posY * p.ly, p.sigma
        gauss = g.evaluate(lambda x, y: p.amp * np.exp(-((x - x0) * (x - x0) / 2. / s / s + (y - y0) * (y - y0) / 2. / s / s)))
        y = [dvec(gauss), dvec(gauss)]
        if p.tau != 0 and p.flr == "gamma_inv":
            self.gamma_inv().symv(y[0], y[1])
        return y

    def compute_psi(self, t):
        p = self.p
        if p.tau == 0.:
            blas1.axpby(1., self.phi[0], 0., self.phi[1])
        else:
            self.old_psi.extrapolate(t, self.phi[1])
            self.numbers["gammaPhi"] = self.multigrid.solve(self.multi_gamma1, self.phi[1], self.phi[0], p.eps_gamma)
            self.old_psi.update(t, self.phi[1])
        self._variation(self.phi[0], self.uE2)
        if p.model == "global":
            blas1.pointwiseDot(1., self.binv, self.binv, self.uE2, 0., self.uE2)
            blas1.axpby(-0.5, self.uE2, 1., self.phi[1])

    def _variation(self, phi, out):
        """Elliptic::variation of the finest polarisation operator (elliptic.h:497-502)"""
        lib().elliptic2d_variation(self.multi_pol[0].h, d(1.), None, ptr(phi), d(0.), ptr(out), stream())

    def polarisation(self, t, y):
        p = self.p
        if p.model == "global":
            # chi = (nt + 1) binv binv, same rounding sequence as the device lambda of toefl.h:116-119
            blas1.copy(y[1], self.chi)
            blas1.plus(self.chi, 1.)
            blas1.pointwiseDot(self.chi, self.binv, self.chi)
            blas1.pointwiseDot(self.chi, self.binv, self.chi)
            if not p.boussinesq:
                self.multi_chi = self.multigrid.project(self.chi)
                for u in range(len(self.multi_pol)):
                    self.multi_pol[u].set_chi(self.multi_chi[u])
        if p.tau == 0.:
            blas1.axpby(1., y[1], 0., self.gamma_n)
        else:
            self.old_gammaN.extrapolate(t, self.gamma_n)
            self.numbers["gammaN"] = self.multigrid.solve(self.multi_gamma1, self.gamma_n, y[1], p.eps_gamma)
            self.old_gammaN.update(t, self.gamma_n)
        blas1.axpby(-1., y[0], 1., self.gamma_n, self.omega)
        if p.model == "global" and p.boussinesq:
            blas1.pointwiseDivide(self.omega, self.chi, self.omega)
        self.old_phi.extrapolate(t, self.phi[0])
        self.numbers["pol"] = self.multigrid.solve(self.multi_pol, self.phi[0], self.omega, p.eps_pol)
        self.old_phi.update(t, self.phi[0])

    def __call__(self, t, y, yp):
        p = self.p
        self.ncalls += 1
        self.polarisation(t, y)
        self.compute_psi(t)
        tau = [-1., p.tau]
        for u in range(2):
            blas1.copy(y[u], self.ype[u])
            if p.model == "global":
                blas1.plus(self.ype[u], 1.)
        for u in range(2):
            self.centered[0].symv(1., self.phi[u], 0., self.dxphi[u])
            self.centered[1].symv(1., self.phi[u], 0., self.dyphi[u])
            if p.model == "global":
                blas1.pointwiseDot(-1., self.binv, self.dyphi[u], 0., self.v[0])
                blas1.pointwiseDot(+1., self.binv, self.dxphi[u], 0., self.v[1])
            else:
                blas1.axpby(-1., self.dyphi[u], 0., self.v[0])
                blas1.axpby(+1., self.dxphi[u], 0., self.v[1])
            blas1.plus(self.v[1], -tau[u] * p.kappa)
            self.adv.upwind(-1., self.v[0], self.v[1], y[u], 0., yp[u])
            if p.model == "global":
                blas1.pointwiseDot(p.kappa, self.ype[u], self.dyphi[u], 1., yp[u])
            else:
                blas1.axpby(p.kappa, self.dyphi[u], 1., yp[u])
        for u in range(2):
            self.laplaceM.symv(-1., y[u], 0., self.lapy[u])
            blas1.axpby(p.nu, self.lapy[u], 1., yp[u])


# Butcher tableaus of inc/dg/tableau.h used here: name -> (a, b, bt, c, fsal)
TABLEAUS = {
    "Bogacki-Shampine-4-2-3": ([[0, 0, 0, 0], [0.5, 0, 0, 0], [0, 0.75, 0, 0], [2. / 9., 1. / 3., 4. / 9., 0.]],
                               [2. / 9., 1. / 3., 4. / 9., 0.], [7. / 24., 1. / 4., 1. / 3., 1. / 8.], [0., 0.5, 3. / 4., 1.], True),
    "Runge-Kutta-4-4": ([[0, 0, 0, 0], [0.5, 0, 0, 0], [0, 0.5, 0, 0], [0, 0, 1., 0]], [1. / 6., 1. / 3., 1. / 3., 1. / 6.],
                        [1. / 6., 1. / 3., 1. / 3., 1. / 6.], [0, 0.5, 0.5, 1.], False),
}


def dense_gemv(alpha, ks, x, beta, y):
    """blas2::gemv(alpha, dg::asDenseMatrix(ks), x, beta, y) (blas2_densematrix.h:38-74): chunks of 8 / 4 / 2 / 1 columns"""
    size, n = len(x), y.numel()

    def pair_sum(cols, b):
        A = (C.c_double * len(cols))(*[x[j] for j in cols])
        X = (C.c_void_p * len(cols))(*[ks[j].data_ptr() for j in cols])
        lib().pair_sum_axpby(n, d(alpha), len(cols), A, X, d(b), ptr(y), stream())

    i = 0
    for i in range(size // 8):
        pair_sum(range(i * 8, i * 8 + 8), beta if i == 0 else 1.)
    i = size // 8
    l = 0
    if size % 8 >= 4:
        pair_sum(range(i * 8, i * 8 + 4), beta if size < 8 else 1.)
        l = 1
    k = 0
    if (size % 8) % 4 >= 2:
        pair_sum(range(i * 8 + l * 4, i * 8 + l * 4 + 2), beta if size < 4 else 1.)
        k = 1
    if ((size % 8) % 4) % 2 == 1:
        j = i * 8 + l * 4 + k * 2
        blas1.axpby(alpha * x[j], ks[j], beta if size < 2 else 1., y)


class ERKStep:
    """dg::ERKStep for std::array<DVec,2> (runge_kutta.h:300-400): same stage logic including FSAL, same kernels for the
    stage sums (dense gemv, EmbeddedPairSum) => bitwise the reference's arithmetic."""

    def __init__(self, tableau, copyable):
        self.a, self.b, self.bt, self.c, self.fsal = TABLEAUS[tableau]
        self.s = len(self.b)
        self.k = [[v.clone() for v in copyable] for _ in range(self.s)]
        self.rkd = [self.b[j] - self.bt[j] for j in range(self.s)]  # ButcherTableau::d (tableau.h:124)
        self.t1 = 1e300

    def step(self, rhs, t0, u0, u1, dt, delta):
        s = self.s
        if t0 != self.t1:
            rhs(t0, u0, self.k[0])
        for i in range(1, s):
            tu = _host_fma(dt, self.c[i], t0)  # DG_FMA on the host (runge_kutta.h:376): one rounding
            for q in range(2):
                blas1.copy(u0[q], delta[q])
                dense_gemv(dt, [self.k[l][q] for l in range(i)], self.a[i][:i], 1., delta[q])
            rhs(tu, delta, self.k[i])
        for q in range(2):
            blas1.copy(u0[q], u1[q])
            # detail::gemm({dt,dt}, k, {b, d}, {1., 0.}, {u1, delta}) (runge_kutta.h:24-64) for s = 4 columns
            assert s == 4, "only four-stage tableaus are wired up"
            blas1.embedded_pair_sum(u1[q], delta[q], 1., 0., [dt * v for v in self.b], [dt * v for v in self.rkd],
                                    [self.k[j][q] for j in range(s)])
        self.t1 = t1 = t0 + dt
        if not self.fsal:
            rhs(t1, u1, self.k[0])
        else:
            self.k[0], self.k[s - 1] = self.k[s - 1], self.k[0]
        return t1


def i_control(dt, eps, embedded_order, order):
    """adaptive.h:37-41"""
    return dt[0] * math.pow(eps[0], -1. / float(embedded_order))


def pi_control(dt, eps, embedded_order, order):
    """adaptive.h:43-53"""
    if dt[1] == 0:
        return i_control(dt, eps, embedded_order, order)
    factor = math.pow(eps[0], -0.8 / float(embedded_order)) * math.pow(eps[1], 0.31 / float(embedded_order))
    return dt[0] * factor


def pid_control(dt, eps, embedded_order, order):
    """adaptive.h:71-85"""
    if dt[1] == 0:
        return i_control(dt, eps, embedded_order, order)
    if dt[2] == 0:
        return pi_control(dt, eps, embedded_order, order)
    q = float(embedded_order)
    factor = math.pow(eps[0], -0.58 / q) * math.pow(eps[1], 0.21 / q) * math.pow(eps[2], -0.1 / q)
    return dt[0] * factor


def pair_dot(x, y):
    """blas1::dot of std::array<DVec,2> operands: the superaccumulators of the components are summed, normalised and
    rounded once (blas1_dispatch_vector.h:153-176)"""
    acc = np.zeros(blas2.BIN_COUNT, dtype=np.int64)
    for a, b in zip(x, y):
        part, _, st = blas2.superacc(a, b)
        if st != 0:
            raise FloatingPointError("dg::Error: dot product failed since one of the inputs contains NaN or Inf")
        acc += part
    lib().superacc_normalize_host(acc.ctypes.data_as(C.c_void_p), None)
    lib().superacc_round_host.restype = C.c_double
    return lib().superacc_round_host(acc.ctypes.data_as(C.c_void_p))


def l2norm(x):
    """adaptive.h:22"""
    return math.sqrt(pair_dot(x, x))


class Adaptive:
    """dg::Adaptive<dg::ERKStep<std::array<DVec,2>>> (adaptive.h:232-395): embedded step, error scaled by
    detail::Tolerance, controller with the step / error history, rejection with restart of the controller."""

    def __init__(self, tableau, copyable, embedded_order=2, order=3):
        self.stepper = ERKStep(tableau, copyable)
        self.next = [v.clone() for v in copyable]
        self.delta = [v.clone() for v in copyable]
        self.size = float(sum(v.numel() for v in copyable))   # dot(1, 1) over both components, exact
        self.embedded_order, self.order = embedded_order, order   # Bogacki-Shampine-4-2-3: q = 2, p = 3 (tableau.h)
        self.eps0 = self.eps1 = self.eps2 = 1.
        self.dt0 = self.dt1 = self.dt2 = 0.
        self.failed, self.nfailed, self.nsteps = False, 0, 0

    def step(self, rhs, t0, u0, u1, dt, control=pid_control, norm=l2norm, rtol=1e-5, atol=1e-6, reject_limit=2.):
        """returns (t1, dt_next); u1 may be the same list as u0 (AdaptiveTimeloop::do_integrate calls it so)"""
        t_next = self.stepper.step(rhs, t0, u0, self.next, dt, self.delta)
        self.nsteps += 1
        rs, as_ = rtol * math.sqrt(self.size), atol * math.sqrt(self.size)
        for q in range(2):
            lib().adaptive_tolerance(u0[q].numel(), d(rs), d(as_), ptr(u0[q]), ptr(self.delta[q]), stream())
        self.eps0 = norm(self.delta)
        self.dt0 = dt
        if self.eps0 > reject_limit or math.isnan(self.eps0):
            dt = control([self.dt0, 0., self.dt2], [self.eps0, self.eps1, self.eps2], self.embedded_order, self.order)
            if abs(dt) > 0.9 * abs(self.dt0):
                dt = 0.9 * self.dt0
            self.failed = True
            self.nfailed += 1
            if u1 is not u0:
                for q in range(2):
                    blas1.copy(u0[q], u1[q])
            return t0, dt
        if self.eps0 < 1e-30:
            dt = 1e14 * self.dt0
            self.eps0 = 1e-30
        else:
            dt = control([self.dt0, self.dt1, self.dt2], [self.eps0, self.eps1, self.eps2], self.embedded_order, self.order)
            if abs(dt) > 100 * abs(self.dt0):
                dt = 100 * self.dt0
        self.eps2, self.eps1 = self.eps1, self.eps0
        self.dt2, self.dt1 = self.dt1, self.dt0
        for q in range(2):
            blas1.copy(self.next[q], u1[q])
        self.failed = False
        return t_next, dt


def _host_fma(a, b, c):
    """a*b + c with ONE rounding: the reference's host compiler contracts such expressions (g++ -O2 -mfma), exact
    rational arithmetic reproduces that"""
    from fractions import Fraction
    return float(Fraction(a) * Fraction(b) + Fraction(c))


# Shu-Osher tableaus (tableau.h:1262-1286): lower triangles alpha(i,k), beta(i,k), k <= i
SHU_OSHER = {
    "SSPRK-2-2": (2, [[1.], [0.5, 0.5]], [[1.], [0., 0.5]]),
    "SSPRK-3-3": (3, [[1.], [3. / 4., 1. / 4.], [1. / 3., 0., 2. / 3.]], [[1.], [0., 1. / 4.], [0., 0., 2. / 3.]]),
}
# multistep tableaus (multistep_tableau.h:263-300): (order, a, b)
MULTISTEP = {
    "AB-1-1": (1, [1.], [1.]),
    "AB-2-2": (2, [1., 0.], [1.5, -0.5]),
    "AB-3-3": (3, [1., 0., 0.], [23. / 12., -4. / 3., 5. / 12.]),
    "TVB-2-2": (2, [4. / 3., -1. / 3.], [4. / 3., -2. / 3.]),
    "TVB-3-3": (3, [1.908535476882378, -1.334951446162515, 0.426415969280137],
                [1.502575553858997, -1.654746338401493, 0.670051276940255]),
}


class ShuOsher:
    """dg::ShuOsher with the identity limiter (runge_kutta.h:840-925) for std::array<DVec,2>"""

    def __init__(self, tableau, copyable):
        self.s, self.alpha, self.beta = SHU_OSHER[tableau]
        self.u = [[v.clone() for v in copyable] for _ in range(self.s)]
        self.k = [[v.clone() for v in copyable] for _ in range(self.s)]
        self.t1 = 1e300

    def step(self, rhs, t0, u0, u1, dt):
        s, al, be = self.s, self.alpha, self.beta
        ts = [t0] + [0.] * s
        for q in range(2):
            blas1.copy(u0[q], self.u[0][q])
        if t0 != self.t1:
            rhs(ts[0], self.u[0], self.k[0])
        for i in range(1, s + 1):
            out = u1 if i == s else self.u[i]
            for q in range(2):
                blas1.axpbypgz(al[i - 1][0], self.u[0][q], dt * be[i - 1][0], self.k[0][q], 0., out[q])
            ts[i] = _host_fma(al[i - 1][0], ts[0], dt * be[i - 1][0])
            for j in range(1, i):
                for q in range(2):
                    blas1.axpbypgz(al[i - 1][j], self.u[j][q], dt * be[i - 1][j], self.k[j][q], 1., out[q])
                ts[i] = ts[i] + _host_fma(al[i - 1][j], ts[j], dt * be[i - 1][j])
            if i != s:
                rhs(ts[i], self.u[i], self.k[i])
            else:
                rhs(ts[i], u1, self.k[0])
        self.t1 = ts[s]
        return ts[s]


class ExplicitMultistep:
    """dg::ExplicitMultistep (multistep.h:59-100; FilteredExplicitMultistep::init/step :592-639 with the identity filter):
    the first steps-1 steps are Shu-Osher Runge-Kutta steps of the same order, then
    u = sum_i a_i u_{n-i} + dt b_i f_{n-i}"""

    def __init__(self, tableau, copyable):
        self.order, self.a, self.b = MULTISTEP[tableau]
        self.steps = len(self.a)
        self.u = [[v.clone() for v in copyable] for _ in range(self.steps)]
        self.f = [[v.clone() for v in copyable] for _ in range(self.steps)]
        self.counter = 0

    def init(self, rhs, t0, u0, dt):
        self.tu, self.dt = t0, dt
        s = self.steps
        for q in range(2):
            blas1.copy(u0[q], self.u[s - 1][q])
        rhs(self.tu, self.u[s - 1], self.f[s - 1])
        self.counter = 0

    def step(self, rhs, t, u):
        """advances u in place, returns the new time"""
        s = self.steps
        if self.counter < s - 1:
            rk = ShuOsher({1: "SSPRK-2-2", 2: "SSPRK-2-2", 3: "SSPRK-3-3"}[self.order], u)
            t = rk.step(rhs, t, u, u, self.dt)
            self.counter += 1
            self.tu = t
            m = s - 1 - self.counter
            for q in range(2):
                blas1.copy(u[q], self.u[m][q])
            rhs(self.tu, self.u[m], self.f[m])
            return t
        t = self.tu = self.tu + self.dt
        for q in range(2):
            blas1.axpby(self.a[0], self.u[0][q], self.dt * self.b[0], self.f[0][q], u[q])
            for i in range(1, s):
                blas1.axpbypgz(self.a[i], self.u[i][q], self.dt * self.b[i], self.f[i][q], 1., u[q])
        self.f = [self.f[-1]] + self.f[:-1]   # std::rotate: the oldest slot becomes slot 0
        self.u = [self.u[-1]] + self.u[:-1]
        for q in range(2):
            blas1.copy(u[q], self.u[0][q])
        rhs(self.tu, self.u[0], self.f[0])
        return t
