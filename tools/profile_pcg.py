#!/usr/bin/env python
"""Small driver for ncu: a few Elliptic applies, exact dots and a short fixed-count PCG solve at n=3, N x N cells."""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from feltor_b200 import blas1, blas2, topology as T  # noqa: E402
from feltor_b200.elliptic import Elliptic2d, PCG  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 12
g = T.Grid([0., 0.], [np.pi, 2 * np.pi], 3, [N, N], [T.DIR, T.PER])
E = Elliptic2d(g, T.DIR, T.PER, T.FORWARD, 1.0)
chi = torch.from_numpy(g.evaluate(lambda x, y: 1. + 0.9 * np.sin(x) * np.sin(y)).copy()).cuda()
E.set_chi(chi)
b = torch.from_numpy(g.evaluate(lambda x, y: np.sin(x) * np.sin(y)).copy()).cuda()
x = torch.zeros(g.size, dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
for _ in range(3):
    E.symv(b, y)
    blas2.dot(b, E.weights(), y)
    blas2.dot(b, y)
    blas1.axpby(1.0, b, 0.5, y)
pcg = PCG(g.size, iters + 1)
pcg.set_throw_on_fail(False)
it = pcg.solve(E, x, b, E.precond(), E.weights(), 1e-30, 1.0, 1)
torch.cuda.synchronize()
print("iterations", it)
