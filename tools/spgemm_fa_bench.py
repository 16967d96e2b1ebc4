#!/usr/bin/env python
"""Construction time of the reference's dg::geo::Fieldaligned (n = 3, 96 x 96, mx = my = 10, "dg": config 4 of BASELINE.json) compiled
on the binding (integration/_build/libdgshim_fa.so): the product projection * interpolation (fieldaligned.h:645-658) through the
device spgemm of libdgb200.so against the reference's host kernel (DGB_SHIM_NO_FUSION=1), same process image, matrices compared.
    python tools/spgemm_fa_bench.py [Nx Ny mx my]"""
import ctypes as C
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "integration", "_build", "libdgshim_fa.so")


def child(Nx, Ny, mx, my, out):
    import numpy as np
    sys.path.insert(0, ROOT)
    import feltor_b200  # noqa: F401  (loads libdgb200.so first so that the binding resolves against it)
    L = C.CDLL(SO)
    L.ref_fa_create.restype = C.c_void_p
    t0 = time.time()
    h = L.ref_fa_create(3, Nx, Ny, 4, mx, my, b"dg")
    sec = time.time() - t0
    assert h
    h = C.c_void_p(h)
    res = {"construct_s": sec}
    arrays = {}
    for which, name in ((0, "plus"), (1, "minus")):
        rows, nnz = L.ref_fa_plane_size(h), L.ref_fa_nnz(h, which)
        pos, idx, val = np.empty(rows + 1, dtype=np.int32), np.empty(nnz, dtype=np.int32), np.empty(nnz)
        L.ref_fa_csr(h, which, pos.ctypes.data_as(C.c_void_p), idx.ctypes.data_as(C.c_void_p), val.ctypes.data_as(C.c_void_p))
        arrays[name + "_pos"], arrays[name + "_idx"], arrays[name + "_val"] = pos, idx, val
        res[name + "_nnz"] = int(nnz)
    np.savez(out, **arrays)
    print("RESULT " + json.dumps(res), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--child":
        child(*[int(v) for v in sys.argv[2:6]], sys.argv[6])
        sys.exit(0)
    import numpy as np
    dims = [int(v) for v in sys.argv[1:5]] if len(sys.argv) >= 5 else [96, 96, 10, 10]
    record = {"config": "dg::geo::Fieldaligned n=3 %dx%d mx=%d my=%d method dg (reference class on the binding)" % tuple(dims)}
    for label, env in (("device_spgemm", {}), ("host_spgemm", {"DGB_SHIM_NO_FUSION": "1"})):
        out = "/tmp/fa_%s.npz" % label
        p = subprocess.run([sys.executable, __file__, "--child"] + [str(v) for v in dims] + [out], env=dict(os.environ, **env),
                           capture_output=True, text=True)
        if p.returncode != 0:
            print(p.stdout[-2000:], p.stderr[-2000:])
            sys.exit(1)
        lines = p.stdout.splitlines()
        res = json.loads([l for l in lines if l.startswith("RESULT ")][0][7:])
        for l in lines:
            if "Multiplication PI" in l:
                res["multiplication_PI_s"] = float(l.split(":")[-1])
            if "Computing all points" in l:
                res["fieldline_integration_s"] = float(l.split(":")[-1])
        record[label] = res
    a, b = np.load("/tmp/fa_device_spgemm.npz"), np.load("/tmp/fa_host_spgemm.npz")
    record["matrices_bitwise_equal"] = bool(all(np.array_equal(a[k].view(np.int64) if a[k].dtype == np.float64 else a[k],
                                                               b[k].view(np.int64) if b[k].dtype == np.float64 else b[k]) for k in a.files))
    record["speedup_multiplication_PI"] = record["host_spgemm"]["multiplication_PI_s"] / record["device_spgemm"]["multiplication_PI_s"]
    print(json.dumps(record))
