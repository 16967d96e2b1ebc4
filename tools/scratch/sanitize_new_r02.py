"""compute-sanitizer target for the kernels added late in round 2: spgemm (fast and large-table variants), the fused bracket and
upwind kernels, the one-pass projection / interpolation, the slope limiter, the gather / scatter-add / rank-sum kernels, and a few
PCG iterations with programmatic dependent launches (tile and walker kernel)."""
import ctypes as C
import os
import sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from feltor_b200 import topology as T, blas2, toefl as TF
from feltor_b200.elliptic import Elliptic2d, PCG, MultigridCG2d
from feltor_b200._lib import lib
from feltor_b200._dev import dvec, ptr, stream, hvec
from test_spgemm import random_pair

for seed, kw in ((0, {}), (5, dict(rows=12, mid=300, cols=6000, per_b=150, per_c=30))):
    shape, B, Cm = random_pair(seed, **kw)
    A = blas2.spgemm(*shape, tuple(dvec(a) for a in B), tuple(dvec(a) for a in Cm))
    print("spgemm", shape, int(A[1].numel()), flush=True)
g = T.Grid([0, 0], [3., 2.], 3, [37, 19], [1, 0])
r = np.random.default_rng(1)
f, vx, vy, r0 = (dvec(r.uniform(-1, 1, g.size)) for _ in range(4))
TF.Advection(g).upwind(-1., vx, vy, f, 0.5, r0)
TF.ArakawaX(g)(0.7, f, vx, 0.3, r0)
print("advection / arakawa", float(blas2.dot(r0, r0)), flush=True)
g2 = T.Grid([0, 0], [1., 2.], 3, [16, 24], [T.DIR, T.PER])
mg = MultigridCG2d(g2, 3)
pr = mg.project(dvec(r.uniform(-1, 1, g2.size)))
mg.interpolate(1, 1., pr[1], 1., pr[0])
print("projection / interpolation", float(blas2.dot(pr[0], pr[0])), flush=True)
pos, idx, val = T.limiter_stencil(g, 0)
y = torch.full((g.size,), float("nan"), dtype=torch.float64, device="cuda")
blas2.stencil("slope", dvec(pos), dvec(idx), dvec(val), f, y, alpha=0.)
n = 1000
ii = dvec(r.integers(0, g.size, n).astype(np.int32))
out = torch.empty(n, dtype=torch.float64, device="cuda")
lib().gather_indexed(n, ptr(ii), ptr(f), ptr(out), stream())
cp, ci, cv = dvec((np.arange(51) * 4).astype(np.int32)), dvec(r.integers(0, n, 200).astype(np.int32)), dvec(r.uniform(-1, 1, 200))
sc = dvec(np.arange(50, dtype=np.int32) * 3)
lib().csr_spmv_scatter_add(50, ptr(cp), ptr(ci), ptr(cv), ptr(out), ptr(sc), ptr(y), stream())
parts = dvec(r.uniform(-1, 1, 4 * 300))
ys = torch.empty(300, dtype=torch.float64, device="cuda")
lib().sum_ranks(4, 300, ptr(parts), ptr(ys), stream())
print("limiter / gather / scatter-add / rank sum", float(ys.sum()), flush=True)
for N in ([40, 24], [416, 404]):
    ge = T.Grid([0., 0.], [np.pi, 2 * np.pi], 3, N, [T.DIR, T.PER])
    E = Elliptic2d(ge, T.DIR, T.PER, T.CENTERED if N[0] < 100 else T.FORWARD, 1.0)
    E.set_chi(dvec(ge.evaluate(lambda x, y: 1. + 0.9 * np.sin(x) * np.sin(y))))
    b = dvec(ge.evaluate(lambda x, y: np.sin(x) * np.sin(y)))
    xs = torch.zeros_like(b)
    p = PCG(ge.size, 7)
    p.set_throw_on_fail(False)
    p.solve(E, xs, b, E.precond(), E.weights(), 1e-30, 1.0, 1)
    print("pcg with dependent launches", N, float(blas2.dot(xs, xs)), flush=True)
torch.cuda.synchronize()
print("done")
