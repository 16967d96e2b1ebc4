import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from feltor_b200 import toefl as TF, blas1
from feltor_b200._lib import lib
from feltor_b200._dev import dvec, hvec, ptr, stream
import ctypes as C
from oracle import reftoefl as R
N = int(sys.argv[1]) if len(sys.argv) > 1 else 48
js = R.default_params(3, N, N)
ref = R.RefToefl(js)
ex = TF.Explicit(TF.Parameters(js))
r = np.random.default_rng(1)
n = ref.size
def cmp(name, a, b, extra=""):
    print("%-12s rel %.3e  bitwise %s %s" % (name, float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)), np.array_equal(a.view(np.int64), b.view(np.int64)), extra), flush=True)
cmp("binv", hvec(ex.binv), ref.binv())
f, vx, vy, res = (r.uniform(-1, 1, n) for _ in range(4))
out = dvec(res)
ex.adv.upwind(-1., dvec(vx), dvec(vy), dvec(f), 0.5, out)
cmp("upwind", hvec(out), ref.upwind(-1., vx, vy, f, 0.5, res))
phi = ex.grid.evaluate(lambda x, y: np.sin(0.05 * x) * np.cos(0.03 * y))
u = torch.zeros(n, dtype=torch.float64, device="cuda")
lib().elliptic2d_variation(ex.multi_pol[0].h, C.c_double(1.), None, ptr(dvec(phi)), C.c_double(0.), ptr(u), stream())
cmp("variation", hvec(u), ref.variation(phi))
b = ex.grid.evaluate(lambda x, y: np.exp(-((x - 60) ** 2 + (y - 100) ** 2) / 200.))
x = dvec(np.zeros(n))
num = ex.multigrid.solve(ex.multi_gamma1, x, dvec(b), ex.p.eps_gamma)
xr, numr = ref.helmholtz_solve(np.zeros(n), b)
cmp("helmholtz", hvec(x), xr, "%s %s" % (num, numr))
chi = 1. + b
mc = ex.multigrid.project(dvec(chi))
for k in range(3):
    ex.multi_pol[k].set_chi(mc[k])
x = dvec(np.zeros(n))
num = ex.multigrid.solve(ex.multi_pol, x, dvec(b), ex.p.eps_pol)
xr, numr = ref.pol_solve(chi, np.zeros(n), b)
cmp("pol", hvec(x), xr, "%s %s" % (num, numr))
