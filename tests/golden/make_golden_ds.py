#!/usr/bin/env python
"""Generates tests/golden/ds_golden.npz from the UNMODIFIED reference free functions dg::geo::ds_* / dss_centered /
dssd_centered / ds_div* / ds_average (inc/geometries/ds.h:743-1016) and dg::TensorMultiply3d (multiply.h:34-58), wrapped
by oracle/ref_ds.cpp (oracle/_ref/libdgref_ds.so, built by `make -C oracle`).  Run in the build container:
    python tests/golden/make_golden_ds.py
Inputs are seeded and stored with the outputs, so the tests need neither /root/reference nor the .so."""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libdgref_ds.so"))
c_dp = C.POINTER(C.c_double)
dp = lambda a: a.ctypes.data_as(c_dp)
d = C.c_double

r = np.random.default_rng(20261018)
n = 1031
out = {}
names = ["f", "fm", "fp", "fmm", "fpp", "g0"]
for k in names:
    out["ds/" + k] = r.uniform(-1, 1, n)
for k in ("Gm", "G0", "Gp", "bm", "b0", "bp"):
    out["ds/" + k] = r.uniform(0.5, 1.5, n)
out["ds/delta"] = np.array([2 * np.pi / 7])
out["ds/alpha_beta"] = np.array([0.7, -0.3])
# operands (a, b, c) per kind in the order of dgb_ds_apply / dgb_ds_apply_vol
ARGS = {0: ("f", "fp", None), 1: ("f", "fm", None), 2: ("fm", "fp", None), 3: ("f", "fp", "fpp"), 4: ("f", "fm", "fmm"),
        5: ("fm", "f", "fp"), 6: ("fm", "f", "fp"), 7: ("fm", "f", None), 8: ("f", "fp", None), 9: ("fm", "fp", None),
        10: ("fm", "fp", None)}
for kind, (a, b, c) in ARGS.items():
    for beta in (-0.3, 0.0):
        g = out["ds/g0"].copy()
        L.ref_ds_apply(kind, n, d(0.7), dp(out["ds/" + a]), dp(out["ds/" + b]), dp(out["ds/" + c]) if c else None,
                       dp(out["ds/Gm"]), dp(out["ds/G0"]), dp(out["ds/Gp"]), dp(out["ds/bm"]), dp(out["ds/b0"]), dp(out["ds/bp"]),
                       d(out["ds/delta"][0]), d(beta), dp(g))
        out[f"ds/kind{kind}/beta{int(beta != 0)}"] = g

t = [r.uniform(-2, 2, n) for _ in range(9)]
ins = [r.uniform(-2, 2, n) for _ in range(3)]
outs = [r.uniform(-2, 2, n) for _ in range(3)]
lam = r.uniform(-2, 2, n)
out["t3d/t"], out["t3d/in"], out["t3d/out0"], out["t3d/lambda"] = np.stack(t), np.stack(ins), np.stack(outs), lam
o = [a.copy() for a in outs]
L.ref_tensor_multiply3d(n, dp(lam), (c_dp * 9)(*[dp(a) for a in t]), (c_dp * 3)(*[dp(a) for a in ins]), d(0.3),
                        (c_dp * 3)(*[dp(a) for a in o]))
out["t3d/out"] = np.stack(o)
o = [a.copy() for a in ins]   # in place (out aliases in), mu = 0
L.ref_tensor_multiply3d(n, dp(lam), (c_dp * 9)(*[dp(a) for a in t]), (c_dp * 3)(*[dp(a) for a in o]), d(0.),
                        (c_dp * 3)(*[dp(a) for a in o]))
out["t3d/out_alias"] = np.stack(o)
# ---- blas2::stencil with the CSR filters on window-like stencils of 3, 4, 5, 9 and 12..25 points (ties included)
L.ref_csr_stencil.argtypes = None
nr = 500
xs = np.round(r.uniform(-1, 1, nr), 1) + 0.0    # coarse values: many ties (+0.0: no negative zeros)
xs[::7] = r.uniform(-1, 1, len(xs[::7]))
counts = r.choice([3, 4, 5, 9, 12, 25], nr)
pos = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
idx = np.concatenate([np.clip(i + r.integers(-12, 13, c), 0, nr - 1) for i, c in enumerate(counts)]).astype(np.int32)
val = r.uniform(-1, 1, pos[-1])
out["stencil/x"], out["stencil/pos"], out["stencil/idx"], out["stencil/val"] = xs, pos, idx, val
c_ip = C.POINTER(C.c_int)
for kind in range(4):
    y = np.zeros(nr)
    L.ref_csr_stencil(kind, nr, nr, pos.ctypes.data_as(c_ip), idx.ctypes.data_as(c_ip), dp(val), d(1.5), dp(xs), dp(y))
    out[f"stencil/kind{kind}"] = y
# ---- assign_bc_along_field_2nd / _1st (ds.h:169-296): masks bbm/bbo/bbp in {0,1} (at most one set), wall distances in (0, delta)
nb = 777
bc_in = {k: r.uniform(-1, 1, nb) for k in ("fm", "f", "fp")}
delta = 2 * np.pi / 7
bc_in["hbm"], bc_in["hbp"] = r.uniform(0.05, 0.95, nb) * delta, r.uniform(0.05, 0.95, nb) * delta
which = r.integers(0, 4, nb)
for j, k in enumerate(("bbm", "bbo", "bbp")):
    bc_in[k] = (which == j + 1).astype(np.float64)
for k, v in bc_in.items():
    out["bc/" + k] = v
for order in (1, 2):
    for bound in (4, 1):   # NEU, DIR
        g0, g1 = np.zeros(nb), np.zeros(nb)
        L.ref_assign_bc_along_field(order, bound, nb, d(delta), dp(bc_in["fm"]), dp(bc_in["f"]), dp(bc_in["fp"]), dp(bc_in["hbm"]),
                                    dp(bc_in["hbp"]), dp(bc_in["bbm"]), dp(bc_in["bbo"]), dp(bc_in["bbp"]), d(0.3), d(-0.2), dp(g0), dp(g1))
        out[f"bc/order{order}/bound{bound}/fmg"], out[f"bc/order{order}/bound{bound}/fpg"] = g0, g1
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "ds_golden.npz"), **out)
print("wrote ds_golden.npz with", len(out), "arrays")
