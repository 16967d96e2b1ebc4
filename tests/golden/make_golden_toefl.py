#!/usr/bin/env python
"""Generates tests/golden/toefl_golden.npz from the UNMODIFIED reference toefl::Explicit + dg::ERKStep
(oracle/_ref/libdgref_toefl.so, built by oracle/Makefile from /root/reference/src/toefl/toefl.h).
Cases: default input of src/toefl/input/default.json on a 24 x 24 grid (n = 3), models "global" and "local";
stored: initial condition, the state and both potentials after 3 fixed Bogacki-Shampine-4-2-3 steps of dt = 0.5, and one
right-hand-side evaluation (fresh object) at that state; plus two dg::Adaptive<ERKStep> runs (see below).   python tests/golden/make_golden_toefl.py"""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import reftoefl as R  # noqa: E402

out = {}
for model in ("global", "local"):
    js = R.default_params(3, 24, 24, model__type=model)
    ref = R.RefToefl(js)
    y0, y1 = ref.init()
    a, b, _ = ref.erk("Bogacki-Shampine-4-2-3", 0., 0.5, 3, y0, y1)
    out[model + "_init0"], out[model + "_init1"] = y0, y1
    out[model + "_y0"], out[model + "_y1"] = a, b
    out[model + "_phi0"], out[model + "_phi1"] = ref.phi(0), ref.phi(1)
    fresh = R.RefToefl(js)
    p0, p1, _ = fresh.rhs(0., a, b)
    out[model + "_rhs0"], out[model + "_rhs1"] = p0, p1
    out[model + "_rhsphi0"], out[model + "_rhsphi1"] = fresh.phi(0), fresh.phi(1)
    out[model + "_binv"] = ref.binv()
# dg::Adaptive<ERKStep> with pid_control / l2norm as src/toefl/toefl.cpp:88-91 drives it: (a) 8 steps from the timeloop's
# initial guess dt = 1e-6 (growth limited to x100 per step), (b) 4 steps from dt = 60 (first step rejected, controller restart)
js = R.default_params(3, 24, 24, model__type="global")
for name, dt0, nsteps in (("adaptA", 1e-6, 8), ("adaptB", 60., 4)):
    ref = R.RefToefl(js)
    y0, y1 = ref.init()
    a, b, t, dts, nf = ref.adaptive("Bogacki-Shampine-4-2-3", 0., dt0, nsteps, 1e-5, 1e-6, y0, y1)
    out[name + "_y0"], out[name + "_y1"], out[name + "_dts"] = a, b, dts
    out[name + "_t_nfailed"] = np.array([t, nf])
    print(name, "t =", t, "dts =", dts, "failed", nf)
# dg::ExplicitMultistep (multistep.h:59-100): init + steps; the first two steps are SSPRK-3-3 Shu-Osher steps
for name, tab, dt, nsteps in (("msAB", "AB-3-3", 0.5, 6), ("msTVB", "TVB-3-3", 0.3, 5)):
    ref = R.RefToefl(js)
    y0, y1 = ref.init()
    a, b, ts = ref.multistep(tab, 0., dt, nsteps, y0, y1)
    out[name + "_y0"], out[name + "_y1"], out[name + "_ts"] = a, b, ts
    out[name + "_ncalls"] = np.array([ref.ncalls()])
    print(name, "ts =", ts, "rhs calls", ref.ncalls())
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "toefl_golden.npz"), **out)
print("wrote", len(out), "arrays")
