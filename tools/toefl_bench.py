#!/usr/bin/env python
"""toefl (config 3 of BASELINE.json): 2D interchange blob, n=3, Nx=Ny=N cells (default 1024), model "global",
nested MultigridCG2d (3 stages) for the two Helmholtz and the polarisation solve, fixed-step Bogacki-Shampine-4-2-3 (the
tableau the shipped toefl.cpp uses; FSAL: 3 right-hand sides per step).  Prints steps/s and RHS/s on one GPU.
`python bench.py --workload toefl` runs this and times the unmodified reference (OpenMP) beside it.
  python tools/toefl_bench.py [--cells 1024] [--steps 6] [--warmup 2] [--dt 0.5] [--stepper erk|multistep] [--tableau NAME]"""
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


def params(N):
    return {"grid": {"n": 3, "Nx": N, "Ny": N, "lx": 200, "ly": 200},
            "init": {"amplitude": 1.0, "sigma": 10, "posX": 0.3, "posY": 0.5, "flr": "gamma_inv"},
            "bc": ["DIR", "PER"],
            "elliptic": {"stages": 3, "eps_pol": [1e-6, 1, 1], "eps_gamma": [1e-7, 1, 1], "direction": "centered"},
            "model": {"type": "global", "boussinesq": False, "curvature": 0.00015, "tau": 1, "nu": 1e-6}}


def run(cells=1024, steps=6, warmup=2, dt=0.5, comm=None, stepper="erk", tableau=None):
    """returns (result dict, initial state as two numpy arrays); comm: feltor_b200.dist.Comm -> the grid is cut into y-slabs
    (one per rank, feltor_b200/dist_toefl.py) and the time is the maximum over the ranks.
    stepper "erk": fixed-step dg::ERKStep (default tableau Bogacki-Shampine-4-2-3, 3 right-hand sides per step);
    stepper "multistep": dg::ExplicitMultistep (default tableau TVB-3-3, ONE right-hand side per step; its first two steps are
    Shu-Osher start-up steps, so warmup must be >= 2) -- the stepper BASELINE.json's config 3 names"""
    N = cells
    if comm is not None and comm.size > 1:
        from feltor_b200.dist_toefl import DistExplicit
        ex = DistExplicit(comm, TF.Parameters(params(N)))
    else:
        comm = None
        ex = TF.Explicit(TF.Parameters(params(N)))
    u0 = ex.initial_condition()
    y_init = [hvec(u0[0]).copy(), hvec(u0[1]).copy()]
    u1 = [torch.zeros_like(u0[0]), torch.zeros_like(u0[0])]
    delta = [torch.zeros_like(u0[0]), torch.zeros_like(u0[0])]
    multistep = stepper == "multistep"
    tableau = tableau or ("TVB-3-3" if multistep else "Bogacki-Shampine-4-2-3")
    state = {"t": 0., "u0": u0, "u1": u1}
    its = {"gammaN": [], "pol": [], "gammaPhi": []}
    if multistep:
        if warmup < 2:
            raise ValueError("the multistep record needs warmup >= 2 (Shu-Osher start-up steps)")
        ms = TF.ExplicitMultistep(tableau, u0)
        ms.init(ex, 0., u0, dt)
    else:
        erk = TF.ERKStep(tableau, u0)

    def step():
        if multistep:
            state["t"] = ms.step(ex, state["t"], state["u0"])
        else:
            state["t"] = erk.step(ex, state["t"], state["u0"], state["u1"], dt, delta)
            state["u0"], state["u1"] = state["u1"], state["u0"]
        for k in its:
            its[k].append(ex.numbers[k])

    def sync():
        torch.cuda.synchronize()
        if comm is not None:
            import torch.distributed as dist
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(warmup):
        step()
    sync()
    for k in its:
        its[k].clear()
    calls0, launches0 = ex.ncalls, fb.lib().raw["dgb_launch_count"]()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    sync()
    sec = e0.elapsed_time(e1) * 1e-3
    if comm is not None:
        import torch.distributed as dist
        tmax = torch.tensor([sec], dtype=torch.float64, device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        sec = tmax.item()
    calls = ex.ncalls - calls0
    out = {"workload": "toefl global n=3 %dx%d, 3-stage MultigridCG2d, %s %s fixed dt=%g" % (N, N, "dg::ExplicitMultistep" if multistep else "dg::ERKStep", tableau, dt),
           "steps_per_s": steps / sec, "rhs_per_s": calls / sec, "ms_per_step": sec / steps * 1e3, "steps": steps,
           "rhs_calls": calls, "kernel_launches": int(fb.lib().raw["dgb_launch_count"]() - launches0),
           "mean_pcg_iterations_per_solve(stage0,1,2)": {k: [float(np.mean([v[s] for v in vals])) for s in range(3)] for k, vals in its.items()},
           "dof": ex.grid.size, "n_gpus": 1 if comm is None else comm.size}
    if comm is not None:
        out["workload"] += "; y-slabs x%d (strong scaling of the fixed grid), ghost rows by dgb_comm_halo_rows, distributed PCG" % comm.size
    return out, y_init


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--dt", type=float, default=0.5)
    ap.add_argument("--stepper", default="erk", choices=["erk", "multistep"])
    ap.add_argument("--tableau", default=None)
    a = ap.parse_args()
    print(json.dumps(run(a.cells, a.steps, a.warmup, a.dt, stepper=a.stepper, tableau=a.tableau)[0]), flush=True)
