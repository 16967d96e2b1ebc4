#!/usr/bin/env python
"""toefl (config 3 of BASELINE.json): 2D interchange blob, n=3, Nx=Ny=N cells (default 1024), model "global",
nested MultigridCG2d (3 stages) for the two Helmholtz and the polarisation solve, fixed-step Bogacki-Shampine-4-2-3 (the
tableau the shipped toefl.cpp uses; FSAL: 3 right-hand sides per step).  Prints steps/s and RHS/s on one GPU and, with
--reference, the unmodified reference (OpenMP, oracle/_ref/libdgref_toefl.so) on a bounded sample of the same run.
  python tools/toefl_bench.py [--cells 1024] [--steps 6] [--warmup 2] [--dt 0.5] [--reference]"""
import argparse
import json
import os
import sys
import time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feltor_b200 as fb  # noqa: E402
from feltor_b200 import toefl as TF  # noqa: E402
from feltor_b200._dev import hvec  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=1024)
ap.add_argument("--steps", type=int, default=6)
ap.add_argument("--warmup", type=int, default=2)
ap.add_argument("--dt", type=float, default=0.5)
ap.add_argument("--reference", action="store_true")
args = ap.parse_args()
N = args.cells
js = {"grid": {"n": 3, "Nx": N, "Ny": N, "lx": 200, "ly": 200},
      "init": {"amplitude": 1.0, "sigma": 10, "posX": 0.3, "posY": 0.5, "flr": "gamma_inv"},
      "bc": ["DIR", "PER"],
      "elliptic": {"stages": 3, "eps_pol": [1e-6, 1, 1], "eps_gamma": [1e-7, 1, 1], "direction": "centered"},
      "model": {"type": "global", "boussinesq": False, "curvature": 0.00015, "tau": 1, "nu": 1e-6}}
ex = TF.Explicit(TF.Parameters(js))
u0 = ex.initial_condition()
y_init = [hvec(u0[0]).copy(), hvec(u0[1]).copy()]
u1 = [torch.zeros_like(u0[0]), torch.zeros_like(u0[0])]
delta = [torch.zeros_like(u0[0]), torch.zeros_like(u0[0])]
erk = TF.ERKStep("Bogacki-Shampine-4-2-3", u0)
t = 0.
its = {"gammaN": [], "pol": [], "gammaPhi": []}


def step():
    global t, u0, u1
    t = erk.step(ex, t, u0, u1, args.dt, delta)
    u0, u1 = u1, u0
    for k in its:
        its[k].append(ex.numbers[k])


for _ in range(args.warmup):
    step()
torch.cuda.synchronize()
for k in its:
    its[k].clear()
calls0, launches0 = ex.ncalls, fb.lib().raw["dgb_launch_count"]()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.steps):
    step()
e1.record()
torch.cuda.synchronize()
sec = e0.elapsed_time(e1) * 1e-3
calls = ex.ncalls - calls0
out = {"workload": "toefl global n=3 %dx%d, 3-stage MultigridCG2d, Bogacki-Shampine-4-2-3 fixed dt=%g" % (N, N, args.dt),
       "steps_per_s": args.steps / sec, "rhs_per_s": calls / sec, "ms_per_step": sec / args.steps * 1e3,
       "rhs_calls": calls, "kernel_launches": int(fb.lib().raw["dgb_launch_count"]() - launches0),
       "mean_pcg_iterations_per_solve(stage0,1,2)": {k: [float(np.mean([v[s] for v in vals])) for s in range(3)] for k, vals in its.items()},
       "dof": ex.grid.size}
if args.reference:
    from oracle import reftoefl as R
    if R.available():
        ref = R.RefToefl(js)
        nref = max(1, min(2, args.steps))
        a, b, rsec = ref.erk("Bogacki-Shampine-4-2-3", 0., args.dt, nref, y_init[0], y_init[1])
        out["reference_cpu"] = {"steps_per_s": nref / rsec, "sample": "%d steps from the same initial state" % nref,
                                "cores": os.cpu_count(), "kind": "reference"}
print(json.dumps(out), flush=True)
