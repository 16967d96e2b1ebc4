"""The reference's own known-answer tests, restated once and run against any backend
(oracle on CPU, libdgb200.so on the GPU):
  inc/dg/topology/evaluation_t.cpp:44-175  (exblas dot on 1d/2d/3d grids)
  inc/dg/topology/derivatives_t.cpp:54-133 (dx/dy/dz/jump symv + dot)
A backend supplies:  make(np)->vec, dot(x,y), dot3(x,w,y), symv(ell_like, alpha, x, beta, y), pdot(x,y,z)."""
import math
import numpy as np
from feltor_b200 import topology as T
from util import bits

PI = math.pi


def _shear(x, y):  # evaluation_t.cpp:22-29
    rho = 0.20943951023931953
    delta = 0.050000000000000003
    if y <= PI:
        return delta * math.cos(x) - 1. / rho / math.cosh((y - PI / 2.) / rho) / math.cosh((y - PI / 2.) / rho)
    return delta * math.cos(x) + 1. / rho / math.cosh((3. * PI / 2. - y) / rho) / math.cosh((3. * PI / 2. - y) / rho)


def evaluation_cases():
    """yields (name, kind, operands(np), golden int64)"""
    g1 = T.Grid([1.], [2.], 3, [12], [T.PER])
    f1 = g1.evaluate(math.exp, vectorized=False)
    w1 = g1.weights()
    yield "1d integral", "dot2", (w1, f1), 4616944842743393935
    yield "1d norm", "dot3", (f1, w1, f1), 4627337306989890294
    g2 = T.Grid([0., 0.], [6.2831853071795862] * 2, 3, [48, 48], [T.PER, T.PER])
    f2 = g2.evaluate(_shear, vectorized=False)
    w2 = g2.weights()
    yield "2d integral", "dot2", (w2, f2), -4823286950217646080
    yield "2d norm", "dot3", (f2, w2, f2), 4635333359953759707  # blas2::dot(w, f) = dot(f, w, f)
    g3 = T.Grid([1., 3., 5.], [2., 4., 6.], [3, 3, 1], [12, 28, 100], [T.PER] * 3)
    f3 = g3.evaluate(lambda x, y, z: math.exp(x) * math.exp(y) * math.exp(z), vectorized=False)
    w3 = g3.weights()
    yield "3d integral", "dot2", (w3, f3), 4675882723962622631
    yield "3d norm", "dot3", (f3, w3, f3), 4746764681002108278


def derivative_cases(three_d=True):
    """yields (name, ell, f, sol, w, golden, golden_gh, squared_first)"""
    n, Nx, Ny, Nz = 3, 24, 28, 100
    s, c = math.sin, math.cos
    g2 = T.Grid([0., 0.1], [PI, 2 * PI + 0.1], n, [Nx, Ny], [T.DIR, T.PER])
    w2 = g2.weights()
    f2 = g2.evaluate(lambda x, y: s(x) * s(y), vectorized=False)
    sols = [g2.evaluate(lambda x, y: c(x) * s(y), vectorized=False), g2.evaluate(lambda x, y: c(y) * s(x), vectorized=False),
            np.zeros(g2.size), np.zeros(g2.size)]
    mats = [T.derivative(0, g2, T.DIR, T.FORWARD), T.derivative(1, g2, T.PER, T.CENTERED), T.jump(0, g2, T.DIR),
            T.jump(1, g2, T.PER)]
    gold = [4562611930300281864, 4553674328256556132, 4567083257206218817, 4574111364446550002]
    gh = [4562611930300282861, 4553674328256673277, 4567083257206217158, 4574111364446550181]
    for i, nm in enumerate(("dx", "dy", "jx", "jy")):
        yield "2d " + nm, mats[i], f2, sols[i], w2, gold[i], gh[i], True
    if not three_d:
        return
    g3 = T.Grid([0., 0.1, PI / 2.], [PI, 2. * PI + 0.1, PI], [n, n, 1], [Nx, Ny, Nz], [T.DIR, T.PER, T.NEU_DIR])
    w3 = g3.weights()
    f3 = g3.evaluate(lambda x, y, z: s(x) * s(y) * s(z), vectorized=False)
    sols = [g3.evaluate(lambda x, y, z: c(x) * s(y) * s(z), vectorized=False),
            g3.evaluate(lambda x, y, z: c(y) * s(x) * s(z), vectorized=False),
            g3.evaluate(lambda x, y, z: c(z) * s(x) * s(y), vectorized=False)] + [np.zeros(g3.size)] * 3
    mats = [T.derivative(0, g3, T.DIR, T.FORWARD), T.derivative(1, g3, T.PER, T.CENTERED),
            T.derivative(2, g3, T.NEU_DIR, T.BACKWARD), T.jump(0, g3, T.DIR), T.jump(1, g3, T.PER), T.jump(2, g3, T.NEU_DIR)]
    gold = [4561946736820639666, 4553062895410573431, 4594213495911299616, 4566393134538626348, 4573262464593641240,
            4594304523193682043]
    gh = [4561946736820640320, 4553062895410783769, 4594213495911299616, 4566393134538622288, 4573262464593641524,
          4594304523193682043]
    for i, nm in enumerate(("dx", "dy", "dz", "jx", "jy", "jz")):
        yield "3d " + nm, mats[i], f3, sols[i], w3, gold[i], gh[i], False


def run_derivative_case(case, make, dot2, dot3, symv, pdot):
    name, m, f, sol, w, gold, gh, squared_first = case
    err = make(sol.copy())
    symv(m, -1., make(f), 1., err)
    if squared_first:  # 2d variant: pointwiseDot(error,error,error); sqrt(dot(w, error))
        pdot(err, err, err)
        norm = math.sqrt(dot2(make(w), err))
    else:              # 3d variant: sqrt(blas2::dot(error, w, error))
        norm = math.sqrt(dot3(err, make(w), err))
    return int(bits([norm])[0]), gold, gh
