#!/usr/bin/env python
"""Generates tests/golden/toefl_golden.npz from the UNMODIFIED reference toefl::Explicit + dg::ERKStep
(oracle/_ref/libdgref_toefl.so, built by oracle/Makefile from /root/reference/src/toefl/toefl.h).
Cases: default input of src/toefl/input/default.json on a 24 x 24 grid (n = 3), models "global" and "local";
stored: initial condition, the state and both potentials after 3 fixed Bogacki-Shampine-4-2-3 steps of dt = 0.5, and one
right-hand-side evaluation (fresh object) at that state.   python tests/golden/make_golden_toefl.py"""
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
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "toefl_golden.npz"), **out)
print("wrote", len(out), "arrays")
