import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from feltor_b200 import topology as T, blas2
from feltor_b200.elliptic import Elliptic2d, PCG
from feltor_b200._dev import dvec
for N, bcx, bcy, d in (([420, 410], T.DIR, T.PER, T.FORWARD), ([404, 420], T.PER, T.PER, T.BACKWARD), ([401, 433], T.NEU, T.DIR, T.CENTERED), ([40, 24], T.DIR, T.PER, T.CENTERED)):
    g = T.Grid([0., 0.], [np.pi, 2 * np.pi], 3, N, [bcx, bcy])
    E = Elliptic2d(g, bcx, bcy, d, 1.0)
    E.set_chi(dvec(g.evaluate(lambda x, y: 1. + 0.9 * np.sin(x) * np.sin(y))))
    x = dvec(g.evaluate(lambda x, y: np.sin(x) * np.cos(3 * y) + 0.1 * x))
    y = torch.zeros_like(x)
    E.symv(x, y)
    E.symv(0.5, x, 0.25, y)
    b = dvec(g.evaluate(lambda x, y: np.sin(x) * np.sin(y)))
    xs = torch.zeros_like(x)
    p = PCG(g.size, 6)
    p.set_throw_on_fail(False)
    p.solve(E, xs, b, E.precond(), E.weights(), 1e-30, 1.0, 1)
    print(N, d, float(blas2.dot(y, y)), float(blas2.dot(xs, xs)), flush=True)
torch.cuda.synchronize()
print("done")
