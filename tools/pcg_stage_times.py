#!/usr/bin/env python
"""Per-kernel times of one PCG iteration (CUDA events inside the solver, dgb_pcg_set_profile) for the grid sizes of toefl's
multigrid stages.  python tools/pcg_stage_times.py [sizes...]"""
import os, sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feltor_b200 as fb
from feltor_b200 import topology as T
from feltor_b200.elliptic import Elliptic2d, PCG
L = fb.lib()
sizes = [int(a) for a in sys.argv[1:]] or [128, 256, 512, 1024]
for N in sizes:
    g = T.Grid([0., 0.], [np.pi, 2 * np.pi], 3, [N, N], [T.DIR, T.PER])
    n = g.size
    chi = torch.from_numpy(g.evaluate(lambda x, y: 1. + 0.9 * np.sin(x) * np.sin(y)).copy()).cuda()
    b = torch.from_numpy(g.evaluate(lambda x, y: np.sin(x) * np.sin(y)).copy()).cuda()
    for dname, d in (("forward", T.FORWARD), ("centered", T.CENTERED)):
        for kern in ("auto", "tile", "walker"):
            E = Elliptic2d(g, T.DIR, T.PER, d, 1.0).set_kernel(kern)
            E.set_chi(chi)
            pcg = PCG(n, 101)
            pcg.set_throw_on_fail(False)
            x = torch.zeros(n, dtype=torch.float64, device="cuda")
            for _ in range(2):
                x.zero_(); pcg.solve(E, x, b, E.precond(), E.weights(), 1e-30, 1.0, 1)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            x.zero_()
            e0.record(); it = pcg.solve(E, x, b, E.precond(), E.weights(), 1e-30, 1.0, 1); e1.record(); torch.cuda.synchronize()
            tot = e0.elapsed_time(e1) * 1e3 / max(it, 1)
            L.pcg_set_profile(pcg.h, 1)
            x.zero_(); pcg.solve(E, x, b, E.precond(), E.weights(), 1e-30, 1.0, 1)
            p = [C.c_double(), C.c_double(), C.c_double()]; pn = C.c_longlong()
            L.pcg_get_profile(pcg.h, C.byref(p[0]), C.byref(p[1]), C.byref(p[2]), C.byref(pn))
            L.pcg_set_profile(pcg.h, 0)
            k = [v.value / max(pn.value, 1) * 1e3 for v in p]
            print(f"N={N:5d} {dname:8s} {kern:6s}->{E.kernel(True):6s} it/s {1e6/tot:9.0f}  us/iter {tot:7.1f} | K1 {k[0]:6.1f} K2 {k[1]:6.1f} K3 {k[2]:6.1f} | roofline(128 B/dof) {128*n/6540.5e3:6.1f} us", flush=True)
