#!/usr/bin/env python
"""Generates tests/golden/ref_golden.npz from the UNMODIFIED reference (oracle/_ref/libdgref.so, built by
`make -C oracle` from /root/reference).  Run in the build container:  python tests/golden/make_golden.py
Inputs are seeded; every array needed to replay a case (inputs AND reference outputs) is stored, so the tests
need neither /root/reference nor libdgref.so."""
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import refwrap as R  # noqa: E402
from util import wide  # noqa: E402
from backends import RefBlas1  # noqa: E402

out = {}
r = np.random.default_rng(20261017)

# ---- topology: matrices of every bc x direction on a small 1d grid, weights, abscissas
for n in (2, 3):
    g = R.grid([0.1], [2.3], n, [6], [0])
    out[f"topo/absc/n{n}"] = R.abscissas(g, 0)
    out[f"topo/w1d/n{n}"] = R.weights1d(g, 0)
    for bc in range(5):
        for d in range(3):
            m = R.ell_create(g, "derivative", 0, bc, d)
            out[f"topo/dx/n{n}/bc{bc}/dir{d}/data"] = m.data
            out[f"topo/dx/n{n}/bc{bc}/dir{d}/cols"] = m.cols_idx
            out[f"topo/dx/n{n}/bc{bc}/dir{d}/didx"] = m.data_idx
        m = R.ell_create(g, "jump", 0, bc)
        out[f"topo/jump/n{n}/bc{bc}/data"] = m.data
g = R.grid([0, 0], [1, 2], 3, [4, 6], [0, 1])
out["topo/w2d"] = R.weights(g)
for kind, a, b in (("fast_projection", 1, 2), ("fast_interpolation", 1, 2), ("fast_projection", 3, 1)):
    m = R.ell_create(g, kind, 1, a=a, b=b)
    out[f"topo/{kind}/{a}_{b}/data"] = m.data
    out[f"topo/{kind}/{a}_{b}/cols"] = m.cols_idx
    out[f"topo/{kind}/{a}_{b}/meta"] = m.meta()
for n in (3, 17):
    for w in range(4):
        out[f"topo/dlt/{w}/n{n}"] = R.dlt(w, n)

# ---- blas1 on random vectors (odd length: exercises the vector tail)
N1 = 1031
B = RefBlas1()
v = [r.uniform(-2, 2, N1) for _ in range(5)]
out["blas1/in"] = np.stack(v)
def run(name, f):
    w = [a.copy() for a in v]
    f(w)
    out["blas1/" + name] = np.stack(w)
run("axpby", lambda w: B.axpby(0.7, w[0], -1.3, w[1]))
run("axpbyz", lambda w: B.axpby(0.7, w[0], -1.3, w[1], w[2]))
run("axpbypgz", lambda w: B.axpbypgz(0.7, w[0], -1.3, w[1], 0.4, w[2]))
run("pdot", lambda w: B.pointwiseDot(0.7, w[0], w[1], -1.3, w[2]))
run("pdot_alias", lambda w: B.pointwiseDot(0.7, w[0], w[1], -1.3, w[1]))
run("pdot3", lambda w: B.pointwiseDot(0.7, w[0], w[1], w[2], -1.3, w[3]))
run("pdot2", lambda w: B.pointwiseDot(0.7, w[0], w[1], -1.3, w[2], w[3], 0.4, w[4]))
run("pdiv", lambda w: B.pointwiseDivide(0.7, w[0], w[1], -1.3, w[2]))
run("pdiv_alias", lambda w: B.pointwiseDivide(0.7, w[2], w[1], -1.3, w[2]))
run("tensor2d", lambda w: R.lib().ref_tensor_multiply2d(N1, R.dp(w[0]), R.dp(v[1]), R.dp(v[2]), R.dp(v[3]), R.dp(v[4]),
                                                        R.dp(w[1]), R.dp(w[2]), R.d(0.3), R.dp(w[3]), R.dp(w[4])))

# ---- exblas dot: wide dynamic range, cancellation, odd length
for name, n, lo, hi in (("small", 7, -3, 3), ("wide", 4099, -300, 300), ("mid", 10000, -40, 40)):
    x, w, y = wide(r, n, lo, hi), wide(r, n, lo // 3, hi // 3), wide(r, n, lo // 3, hi // 3)
    out[f"dot/{name}/x"], out[f"dot/{name}/w"], out[f"dot/{name}/y"] = x, w, y
    a2, s2 = R.dot2(x, y)
    a3, s3 = R.dot3(x, w, y)
    out[f"dot/{name}/val2"] = np.array([R.round_acc(a2)])
    out[f"dot/{name}/val3"] = np.array([R.round_acc(a3)])
    # the reference returns an un-normalised accumulator: store it normalised (exblas::cpu::Normalize is
    # applied by its callers, blas1.h:159) through Round's side effect-free twin in the oracle
    from oracle import orc
    out[f"dot/{name}/acc2"] = orc.normalize(a2)
    out[f"dot/{name}/acc3"] = orc.normalize(a3)
x = r.uniform(-1, 1, 100); x[:50] = 1e300; y = x.copy(); y[50:] = -1.  # overflow of single products -> status
a2, s2 = R.dot2(x, y)
out["dot/overflow/status"] = np.array([s2])

# ---- Ell symv: 2d and 3d, every coordinate, alpha/beta variants
g2 = R.grid([0, 0.1], [np.pi, 2 * np.pi + 0.1], 3, [8, 6], [R.DIR, R.PER])
x2 = r.uniform(-1, 1, R.grid_size(g2))
y2 = r.uniform(-1, 1, R.grid_size(g2))
out["ell/x2"], out["ell/y2"] = x2, y2
for coord, bc in ((0, R.DIR), (1, R.PER), (0, R.NEU_DIR), (1, R.NEU)):
    for d in range(3):
        m = R.ell_create(g2, "derivative", coord, bc, d)
        for al, be in ((1., 0.), (-1., 1.), (0.5, -2.)):
            y = y2.copy()
            m.symv(al, x2, be, y)
            out[f"ell/d{coord}/bc{bc}/dir{d}/a{al}b{be}"] = y
    m = R.ell_create(g2, "jump", coord, bc)
    y = y2.copy()
    m.symv(1., x2, 0., y)
    out[f"ell/j{coord}/bc{bc}"] = y
g3 = R.grid([0, 0.1, 1.], [np.pi, 2 * np.pi + 0.1, 2.], 3, [4, 3, 5], [R.DIR, R.PER, R.NEU_DIR])
x3 = r.uniform(-1, 1, R.grid_size(g3))
out["ell/x3"] = x3
for coord in range(3):
    m = R.ell_create(g3, "derivative", coord, g3.bc[coord], R.BACKWARD)
    y = np.zeros(R.grid_size(g3))
    m.symv(1., x3, 0., y)
    out[f"ell/3d/d{coord}"] = y

# ---- CSR spmv (random sparse rows incl. empty rows, unsorted columns)
nr, nc = 37, 29
counts = r.integers(0, 7, nr); counts[5] = 0
pos = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
idx = r.integers(0, nc, pos[-1]).astype(np.int32)
val = r.uniform(-1, 1, pos[-1])
xc, yc = r.uniform(-1, 1, nc), r.uniform(-1, 1, nr)
out["csr/pos"], out["csr/idx"], out["csr/val"], out["csr/x"], out["csr/y"] = pos, idx, val, xc, yc
for al, be in ((1., 0.), (0.5, 1.), (-2., 0.25)):
    y = yc.copy()
    R.csr_symv(nr, nc, pos, idx, val, al, xc, be, y)
    out[f"csr/a{al}b{be}"] = y

# ---- Elliptic2d + PCG (elliptic2d_b.cpp:22-38 problem at small size) and the cg2d_b-like NEU x PER case
for tag, bcx, bcy, d, jf in (("dirper_fwd", R.DIR, R.PER, R.FORWARD, 1.0), ("neu_cen", R.NEU, R.PER, R.CENTERED, 0.1),
                              ("dirneu_bwd", R.DIR_NEU, R.DIR, R.BACKWARD, 1.0)):
    g = R.grid([0, 0], [np.pi, 2 * np.pi], 3, [10, 8], [bcx, bcy])
    E = R.Elliptic2d(g, bcx, bcy, d, jf)
    chi = R.evaluate(g, "pol")
    E.set_chi(chi)
    x = r.uniform(-1, 1, E.size)
    y0 = r.uniform(-1, 1, E.size)
    out[f"elliptic/{tag}/chi"], out[f"elliptic/{tag}/x"], out[f"elliptic/{tag}/y0"] = chi, x, y0
    for al, be in ((1., 0.), (-0.5, 2.)):
        y = y0.copy()
        E.symv(al, x, be, y)
        out[f"elliptic/{tag}/a{al}b{be}"] = y
    out[f"elliptic/{tag}/weights"], out[f"elliptic/{tag}/precond"] = E.weights(), E.precond()
    if tag == "dirper_fwd":
        b = R.evaluate(g, "rhs")
        xs = np.zeros(E.size)
        it, _ = E.pcg_solve(xs, b, E.precond(), E.weights(), 1e-8, 1.0, 1)
        out[f"elliptic/{tag}/pcg_b"], out[f"elliptic/{tag}/pcg_x"], out[f"elliptic/{tag}/pcg_it"] = b, xs, np.array([it])
        sig = np.zeros(E.size)
        E.variation(1.0, chi, x, 0., sig)
        out[f"elliptic/{tag}/variation"] = sig

# ---- MultigridCG2d (multigrid.h:500-668): projection to stages + nested iteration solve
g = R.grid([0, 0], [np.pi, 2 * np.pi], 3, [16, 16], [R.DIR, R.PER])
M = R.Multigrid(g, 3)
chi = R.evaluate(g, "pol")
b = R.evaluate(g, "rhs")
proj = M.project(chi)
for u, p in enumerate(proj):
    out[f"multigrid/project{u}"] = p
M.set_chi(chi)
x = np.zeros(R.grid_size(g))
st, num, _ = M.solve(x, b, [1e-6, 1e-6 * 1.5, 1e-6 * 1.5 * 1.5])
out["multigrid/chi"], out["multigrid/b"], out["multigrid/x"], out["multigrid/num"] = chi, b, x, np.array(num)

path = os.path.join(ROOT, "tests", "golden", "ref_golden.npz")
np.savez_compressed(path, **out)
print("wrote", path, len(out), "arrays", os.path.getsize(path), "bytes")
