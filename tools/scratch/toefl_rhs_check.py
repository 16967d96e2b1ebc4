import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from feltor_b200 import toefl as TF, blas1
from feltor_b200._dev import dvec, hvec
from oracle import reftoefl as R
N = int(sys.argv[1]) if len(sys.argv) > 1 else 48
model = sys.argv[2] if len(sys.argv) > 2 else "global"
js = R.default_params(3, N, N, model__type=model)
ref = R.RefToefl(js)
y0, y1 = ref.init()
ya, yb, _ = ref.erk("Bogacki-Shampine-4-2-3", 0., 0.5, 3, y0, y1)   # a state with a non-trivial potential
def cmp(name, a, b):
    print("%-10s rel %.3e  bitwise %s" % (name, float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)), np.array_equal(a.view(np.int64), b.view(np.int64))))
ref2 = R.RefToefl(js)   # fresh histories on both sides
rp0, rp1, _ = ref2.rhs(0., ya, yb)
ex = TF.Explicit(TF.Parameters(js))
cmp("binv", hvec(ex.binv), ref2.binv())
y = [dvec(ya), dvec(yb)]
yp = [torch.zeros_like(y[0]), torch.zeros_like(y[0])]
ex(0., y, yp)
print(ex.numbers)
cmp("phi0", hvec(ex.phi[0]), ref2.phi(0))
cmp("phi1", hvec(ex.phi[1]), ref2.phi(1))
cmp("yp0", hvec(yp[0]), rp0)
cmp("yp1", hvec(yp[1]), rp1)
