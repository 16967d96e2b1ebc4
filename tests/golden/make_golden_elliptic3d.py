#!/usr/bin/env python
"""Generates tests/golden/elliptic3d_golden.npz from the UNMODIFIED reference classes dg::Elliptic3d (compute-in-2d mode, on a
CartesianGrid3d and a CylindricalGrid3d) and dg::Elliptic1d (inc/dg/elliptic.h:65-200,557-797) through oracle/_ref/libdgref.so.
    python tests/golden/make_golden_elliptic3d.py
Inputs are seeded and stored with the outputs, so the tests need neither /root/reference nor the .so."""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refwrap as R  # noqa: E402

out = {}
r = np.random.default_rng(20261019)
N, n = [9, 7, 4], 9 * 7 * 4 * 9
out["x"], out["y0"], out["chi"] = r.uniform(-1, 1, n), r.uniform(-1, 1, n), r.uniform(0.5, 2., n)
for cyl in (0, 1):
    x0, x1 = ([3., -1., 0.], [5., 1., 2 * np.pi]) if cyl else ([0., 0., 0.], [1., 2., 3.])
    bc = [1, 4 if cyl else 0, 0]   # DIR, NEU / PER, PER
    rg = R.grid(x0, x1, 3, N, bc)
    for direction in (0, 2):
        for cwj in (0, 1):
            y, w, p = R.elliptic3d_symv(rg, cyl, direction, 0.7, cwj, out["chi"], -0.5, out["x"], 0.3, out["y0"])
            out[f"e3d/cyl{cyl}/dir{direction}/cwj{cwj}/y"] = y
    out[f"e3d/cyl{cyl}/weights"], out[f"e3d/cyl{cyl}/precond"] = w, p
n1 = 21 * 3
out["x1d"], out["y1d"], out["chi1d"] = r.uniform(-1, 1, n1), r.uniform(-1, 1, n1), r.uniform(0.5, 2., n1)
for bcx in range(5):
    rg = R.grid([0.3], [2.1], 3, [21], [bcx])
    for direction in (0, 1, 2):
        y, w, p = R.elliptic1d_symv(rg, bcx, direction, 0.7, out["chi1d"], -0.5, out["x1d"], 0.3, out["y1d"])
        out[f"e1d/bc{bcx}/dir{direction}/y"] = y
# full 3-d mode (compute_in_2d = 0: z derivative + 3-d tensor product, elliptic.h:688-697) and the restricted one, same call
Nf, nf = [9, 7, 5], 9 * 7 * 5 * 9
out["x_full"], out["y0_full"], out["chi_full"] = r.uniform(-1, 1, nf), r.uniform(-1, 1, nf), r.uniform(0.5, 2., nf)
for (cyl, direction, cwj, in2d) in [(0, 0, 0, 0), (1, 2, 0, 0), (0, 1, 1, 0), (1, 0, 1, 0), (1, 2, 1, 1), (0, 2, 0, 1)]:
    x0, x1 = ([3., -1., 0.], [5., 1., 2 * np.pi]) if cyl else ([0., 0., 0.], [1., 2., 3.])
    rg = R.grid(x0, x1, 3, Nf, [1, 4 if cyl else 0, 0])
    out[f"e3dfull/cyl{cyl}/dir{direction}/cwj{cwj}/in2d{in2d}/y"] = R.elliptic3d_symv_mode(
        rg, cyl, direction, 0.7, cwj, in2d, out["chi_full"], -0.5, out["x_full"], 0.3, out["y0_full"])
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "elliptic3d_golden.npz"), **out)
print("wrote", len(out), "arrays")
