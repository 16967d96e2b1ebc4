#!/usr/bin/env python
"""Short driver for ncu captures: exdot2/exdot3 and the Elliptic apply (plain, and 12 PCG iterations) at n=3 1024^2.
   ncu --set full -k regex:'exdot|walker' ... python tools/prof_kernels.py [dot|k1|pcg]"""
import os, sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feltor_b200 as fb
from feltor_b200 import blas2, topology as T
from feltor_b200._dev import ptr, stream
from feltor_b200.elliptic import Elliptic2d, PCG
what = sys.argv[1] if len(sys.argv) > 1 else "dot"
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
L = fb.lib()
g = T.Grid([0., 0.], [np.pi, 2 * np.pi], 3, [N, N], [T.DIR, T.PER])
n = g.size
v = [torch.rand(n, dtype=torch.float64, device="cuda") + 0.5 for _ in range(4)]
if what == "dot":
    ws = blas2.DotWorkspace()
    res = torch.zeros(41, dtype=torch.int64, device="cuda")
    for _ in range(3):
        L.exdot2(ws.h, n, ptr(v[0]), C.c_double(0), ptr(v[1]), C.c_double(0), ptr(res), stream())
        L.exdot3(ws.h, n, ptr(v[0]), C.c_double(0), ptr(v[1]), C.c_double(0), ptr(v[2]), C.c_double(0), ptr(res), stream())
elif what == "k1":
    for d in (T.FORWARD, T.CENTERED):
        E = Elliptic2d(g, T.DIR, T.PER, d, 1.0)
        E.set_chi(v[2])
        for _ in range(3):
            E.symv(v[0], v[1])
else:
    E = Elliptic2d(g, T.DIR, T.PER, T.FORWARD, 1.0)
    E.set_chi(v[2])
    pcg = PCG(n, 8)
    pcg.set_throw_on_fail(False)
    x = torch.zeros(n, dtype=torch.float64, device="cuda")
    pcg.solve(E, x, v[0], E.precond(), E.weights(), 1e-30, 1.0, 1)
torch.cuda.synchronize()
