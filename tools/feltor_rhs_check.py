#!/usr/bin/env python
"""The UNMODIFIED 3-d application class feltor::Explicit (src/feltor/feltor.h) on the libdgb200 binding
(integration/_build/libdgshim_feltor.so) against the same wrapper on the reference's OpenMP backend (oracle/_ref/libdgref_feltor.so):
right-hand sides and potentials of a few evaluations, time per evaluation on both.
    python tools/feltor_rhs_check.py [Nx Ny Nz [mx my]]"""
import importlib.util
import json
import os
import sys
import time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def clone(path):
    spec = importlib.util.spec_from_file_location("feltor_" + os.path.basename(path).replace(".", "_"), os.path.join(ROOT, "oracle", "reffeltor.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    m._PATH = path
    return m


def run(dims=None, evaluations=3, quiet=True):
    import feltor_b200  # noqa: F401  (libdgb200.so first)
    js = json.load(open(os.path.join(ROOT, "tests", "golden", "feltor_input.json")))
    if dims:
        js["grid"].update({"Nx": dims[0], "Ny": dims[1], "Nz": dims[2]})
        if len(dims) >= 5:
            js["FCI"]["refine"] = [dims[3], dims[4]]
    text = json.dumps(js)
    dev = clone(os.path.join(ROOT, "integration", "_build", "libdgshim_feltor.so"))
    ref = clone(os.path.join(ROOT, "oracle", "_ref", "libdgref_feltor.so"))
    assert dev.lib().ref_feltor_is_device() == 1 and ref.lib().ref_feltor_is_device() == 0
    saved, devnull = os.dup(1), os.open(os.devnull, os.O_WRONLY)
    if quiet:
        sys.stdout.flush()
        os.dup2(devnull, 1)
    try:
        rec = {"config": "feltor::Explicit n=%d %dx%dx%d refine %s, circular field, stages %d" % (js["grid"]["n"], js["grid"]["Nx"], js["grid"]["Ny"],
                                                                                               js["grid"]["Nz"], js["FCI"]["refine"], js["elliptic"]["stages"])}
        objs = []
        for name, m in (("device", dev), ("openmp", ref)):
            t0 = time.time()
            objs.append(m.RefFeltor(text))
            rec[name + "_construct_s"] = time.time() - t0
        D, H = objs
        rec["dof"] = D.size
        worst, times = 0., {"device": [], "openmp": []}
        for k in range(evaluations):
            outs = {}
            for name, o in (("device", D), ("openmp", H)):
                t0 = time.time()
                outs[name] = o.rhs(0.01 * k) + [o.potential(0), o.potential(1)]
                times[name].append(time.time() - t0)
                o.euler(1e-3)
            for a, b in zip(outs["device"], outs["openmp"]):
                worst = max(worst, float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300)))
        rec["max_rel_diff"] = worst
        rec["device_ms_per_rhs"] = [t * 1e3 for t in times["device"]]
        rec["openmp_ms_per_rhs"] = [t * 1e3 for t in times["openmp"]]
        rec["openmp_threads"] = os.cpu_count()
    finally:
        if quiet:
            os.dup2(saved, 1)
        os.close(devnull)
        os.close(saved)
    return rec


if __name__ == "__main__":
    dims = [int(v) for v in sys.argv[1:]] or None
    print(json.dumps(run(dims)))
