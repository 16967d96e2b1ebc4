"""Elliptic3d (compute-in-2d) at the feltor grid of BASELINE config 5: n=3, 192 x 192 x 64; us per apply and GB/s (24 B/dof)"""
import sys, os, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from feltor_b200 import topology as T
from feltor_b200.elliptic import Elliptic3d
for N in ([192, 192, 64], [96, 96, 64]):
    g = T.Grid([3., -1., 0.], [5., 1., 2 * np.pi], [3, 3, 1], N, [T.DIR, T.DIR, T.PER])
    op = Elliptic3d(g, direction=T.CENTERED, jfactor=1., cylindrical=True)
    n = g.size
    x = torch.rand(n, dtype=torch.float64, device="cuda"); y = torch.zeros_like(x)
    op.set_chi(torch.rand(n, dtype=torch.float64, device="cuda") + 0.5)
    for _ in range(3):
        op.symv(x, y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        op.symv(x, y)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print("N", N, "loop" if os.environ.get("DGB_ELLIPTIC_PLANES_LOOP") else "one launch", "%.1f us" % us, "%.0f GB/s (24 B/dof + vol)" % (n * 24 / us / 1e3))
