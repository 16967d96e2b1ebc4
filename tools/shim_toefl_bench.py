#!/usr/bin/env python
"""toefl (config 3) through the REFERENCE's own API: the unmodified src/toefl/toefl.h + dg::ERKStep compiled against the
libdgb200 binding (integration/_build/libdgshim_toefl.so).  Prints steps/s with the fused hooks on and off.
  python tools/shim_toefl_bench.py [--cells 1024] [--steps 9]"""
import argparse, importlib.util, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def load():
    path = os.path.join(ROOT, "integration", "_build", "libdgshim_toefl.so")
    if not os.path.exists(path):
        return None
    spec = importlib.util.spec_from_file_location("shimtoefl", os.path.join(ROOT, "oracle", "reftoefl.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    m._PATH = path
    return m


def run(cells=1024, steps=9, fusion=1):
    """`steps` Bogacki-Shampine steps of ONE dg::ERKStep from the application's initial state on a fresh application object
    (a new stepper re-evaluates the right-hand side at its start time, which re-solves converged systems: not part of a time
    loop, so nothing is split off as warm-up here; the one-time costs -- plans, first touches -- are paid by a throw-away
    object first)."""
    m = load()
    if m is None:
        return None
    import feltor_b200  # noqa: F401  (loads libdgb200.so from the tree)
    m.lib().ref_set_fusion(fusion)
    saved, devnull = os.dup(1), os.open(os.devnull, os.O_WRONLY)
    sys.stdout.flush()
    os.dup2(devnull, 1)  # toefl.h prints its solver statistics (set_benchmark(true))
    try:
        for nsteps in (1, steps):
            T = m.RefToefl(m.default_params(3, cells, cells))
            y0, y1 = T.init()
            _, _, sec = T.erk("Bogacki-Shampine-4-2-3", 0., 0.5, nsteps, y0, y1)
            del T
    finally:
        os.dup2(saved, 1)
        os.close(devnull)
        os.close(saved)
    m.lib().ref_set_fusion(1)
    return {"steps_per_s": steps / sec, "ms_per_step": sec / steps * 1e3, "steps": steps, "fusion": bool(fusion), "cells": cells,
            "what": "unmodified toefl::Explicit + dg::ERKStep (reference headers) on the libdgb200 binding, steps 1..%d from the initial state" % steps}


def run_harness(cells=1024, steps=9):
    """the same steps from the same initial state (the reference's, bit for bit) through the C-ABI harness feltor_b200/toefl.py"""
    m = load()
    import time
    import torch
    from feltor_b200 import toefl as TF
    from feltor_b200._dev import dvec
    saved, devnull = os.dup(1), os.open(os.devnull, os.O_WRONLY)
    sys.stdout.flush()
    os.dup2(devnull, 1)
    try:
        T = m.RefToefl(m.default_params(3, cells, cells))
        a, b = T.init()
        del T
    finally:
        os.dup2(saved, 1)
        os.close(devnull)
        os.close(saved)
    for nsteps in (1, steps):
        ex = TF.Explicit(TF.Parameters(m.default_params(3, cells, cells)))
        u0 = [dvec(a), dvec(b)]
        u1 = [torch.zeros_like(u0[0]), torch.zeros_like(u0[0])]
        delta = [torch.zeros_like(u0[0]), torch.zeros_like(u0[0])]
        erk = TF.ERKStep("Bogacki-Shampine-4-2-3", u0)
        stage0 = []

        def rhs(t, y, yp):
            ex(t, y, yp)
            stage0.append(ex.numbers["pol"][0])
        torch.cuda.synchronize()
        t0, t = time.time(), 0.
        for _ in range(nsteps):
            t = erk.step(rhs, t, u0, u1, 0.5, delta)
            u0, u1 = u1, u0
        torch.cuda.synchronize()
        sec = time.time() - t0
    return {"steps_per_s": steps / sec, "ms_per_step": sec / steps * 1e3, "steps": steps, "cells": cells,
            "polarisation_iterations_on_the_fine_stage_per_rhs": stage0,
            "what": "C-ABI harness (feltor_b200/toefl.py), the same steps from the same initial state"}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=9)
    a = ap.parse_args()
    for f in (1, 0):
        print(json.dumps(run(a.cells, a.steps, f)), flush=True)
    print(json.dumps(run_harness(a.cells, a.steps)), flush=True)
