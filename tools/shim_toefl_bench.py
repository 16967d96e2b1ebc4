#!/usr/bin/env python
"""toefl (config 3) through the REFERENCE's own API: the unmodified src/toefl/toefl.h + dg::ERKStep compiled against the
libdgb200 binding (integration/_build/libdgshim_toefl.so).  Prints steps/s with the fused hooks on and off.
  python tools/shim_toefl_bench.py [--cells 1024] [--steps 4]"""
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


def run(cells=1024, steps=4, fusion=1):
    m = load()
    if m is None:
        return None
    import feltor_b200  # noqa: F401  (loads libdgb200.so from the tree)
    m.lib().ref_set_fusion(fusion)
    saved, devnull = os.dup(1), os.open(os.devnull, os.O_WRONLY)
    sys.stdout.flush()
    os.dup2(devnull, 1)  # toefl.h prints its solver statistics (set_benchmark(true))
    try:
        T = m.RefToefl(m.default_params(3, cells, cells))
        y0, y1 = T.init()
        a, b, _ = T.erk("Bogacki-Shampine-4-2-3", 0., 0.5, 2, y0, y1)      # warm-up (plans, extrapolation history)
        a, b, sec = T.erk("Bogacki-Shampine-4-2-3", 1.0, 0.5, steps, a, b)
    finally:
        os.dup2(saved, 1)
        os.close(devnull)
        os.close(saved)
    m.lib().ref_set_fusion(1)
    return {"steps_per_s": steps / sec, "ms_per_step": sec / steps * 1e3, "steps": steps, "fusion": bool(fusion), "cells": cells,
            "what": "unmodified toefl::Explicit + dg::ERKStep (reference headers) on the libdgb200 binding"}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=4)
    a = ap.parse_args()
    for f in (1, 0):
        print(json.dumps(run(a.cells, a.steps, f)), flush=True)
