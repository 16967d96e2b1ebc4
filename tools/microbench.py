#!/usr/bin/env python
"""Kernel micro-benchmark (the blas_b.cpp table of the reference, inc/dg/blas_b.cpp:122-341) on one GPU:
prints GB/s (algorithmic bytes) per kernel.  python tools/microbench.py [N=1024] [reps=50]"""
import json
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feltor_b200 as fb  # noqa: E402
from feltor_b200 import blas1, blas2, topology as T  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 50
g = T.Grid([0., 0.], [2 * np.pi] * 2, 3, [N, N], [T.PER, T.PER])
n = g.size
gen = torch.Generator(device="cuda").manual_seed(0)
vec = [torch.rand(n, dtype=torch.float64, device="cuda", generator=gen) + 0.5 for _ in range(8)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > L2 (126 MB)
PEAK = 6541.5
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass


def timeit(f, bytes_per_call, name, flush_l2=True):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush_l2:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        f()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    t = float(np.median(ts))
    gbs = bytes_per_call / t / 1e9
    print(f"{name:34s} {t*1e6:9.1f} us  {gbs:8.1f} GB/s  {gbs/PEAK*100:5.1f}% of {PEAK:.0f}", flush=True)
    return t


B = 8 * n
print(f"# n=3 {N}x{N}: {n} doubles per vector ({B/1e6:.1f} MB), reps={reps}, L2 flushed between calls")
timeit(lambda: vec[1].copy_(vec[0]), 2 * B, "torch copy_ (comparison)")
timeit(lambda: blas1.copy(vec[0], vec[1]), 2 * B, "copy")
timeit(lambda: blas1.scal(vec[1], 1.0000001), 2 * B, "scal")
timeit(lambda: blas1.axpby(1.0000001, vec[0], 0.9999999, vec[1]), 3 * B, "axpby")
timeit(lambda: blas1.axpby(1.0000001, vec[0], 0.9999999, vec[1], vec[2]), 3 * B, "axpbyz")
timeit(lambda: blas1.axpbypgz(1.0000001, vec[0], 0.9999999, vec[1], 0.5, vec[2]), 4 * B, "axpbypgz")
timeit(lambda: blas1.pointwiseDot(1.0000001, vec[0], vec[1], 0.5, vec[2]), 4 * B, "pointwiseDot a x1 x2 + b y")
timeit(lambda: blas1.pointwiseDot(1.0, vec[0], vec[1], 2.0, vec[2], vec[3], 0.0, vec[4]), 6 * B, "pointwiseDot 6 operands")
timeit(lambda: blas1.pointwiseDivide(1.0, vec[0], vec[1], 0.5, vec[2]), 4 * B, "pointwiseDivide")
timeit(lambda: blas1.tensor_multiply2d(vec[0], (vec[1], vec[2], vec[3], vec[4]), vec[5], vec[6], 0., vec[5], vec[6]),
       9 * B, "tensor_multiply2d (7 in, 2 out)")
ws = blas2.DotWorkspace()
res = torch.zeros(41, dtype=torch.int64, device="cuda")
from feltor_b200._dev import ptr, stream  # noqa: E402
import ctypes as C  # noqa: E402
L = fb.lib()
timeit(lambda: L.exdot2(ws.h, n, ptr(vec[0]), C.c_double(0), ptr(vec[1]), C.c_double(0), ptr(res), stream()), 2 * B,
       "exdot2 (device result)")
timeit(lambda: L.exdot3(ws.h, n, ptr(vec[0]), C.c_double(0), ptr(vec[1]), C.c_double(0), ptr(vec[2]), C.c_double(0),
                        ptr(res), stream()), 3 * B, "exdot3 (device result)")
timeit(lambda: blas2.dot(vec[0], vec[1], vec[2]), 3 * B, "dot3 incl. D2H + sync")
timeit(lambda: torch.dot(vec[0], vec[1]), 2 * B, "torch.dot (comparison, not reproducible)")
for name, m in (("dx forward", T.derivative(0, g, T.PER, T.FORWARD)), ("dy forward", T.derivative(1, g, T.PER, T.FORWARD)),
                ("dx centered", T.derivative(0, g, T.PER, T.CENTERED)), ("dy centered", T.derivative(1, g, T.PER, T.CENTERED)),
                ("jump x", T.jump(0, g, T.PER)), ("jump y", T.jump(1, g, T.PER)),
                ("dx forward DIR", T.derivative(0, g, T.DIR, T.FORWARD))):
    m.handle
    timeit(lambda: m.symv(1.0, vec[0], 0.0, vec[1]), 2 * B, f"ell {name} beta=0")
    timeit(lambda: m.symv(1.0, vec[0], 1.0, vec[1]), 3 * B, f"ell {name} beta=1")
    timeit(lambda: m.symv(1.0, vec[0], 0.0, vec[1], generic=True), 2 * B, f"ell {name} GENERIC beta=0")
# ---- Elliptic2d apply (config 2 operator) and PCG iteration rate
from feltor_b200.elliptic import Elliptic2d, PCG  # noqa: E402
ge = T.Grid([0., 0.], [np.pi, 2 * np.pi], 3, [N, N], [T.DIR, T.PER])
for dname, dd in (("forward", T.FORWARD), ("centered", T.CENTERED)):
    E = Elliptic2d(ge, T.DIR, T.PER, dd, 1.0)
    E.set_chi(vec[2])
    timeit(lambda: E.symv(vec[0], vec[1]), 3 * B, f"Elliptic2d {dname} FUSED (24 B/dof)")
    timeit(lambda: E.symv(0.5, vec[0], 2.0, vec[1]), 4 * B, f"Elliptic2d {dname} FUSED beta!=0")
    timeit(lambda: E.symv(vec[0], vec[1], unfused=True), 3 * B, f"Elliptic2d {dname} unfused (24 B/dof)")
    E.set_ordering("relaxed")
    timeit(lambda: E.symv(vec[0], vec[1]), 3 * B, f"Elliptic2d {dname} FUSED relaxed ordering")
    E.set_ordering("reference")
E = Elliptic2d(ge, T.DIR, T.PER, T.FORWARD, 1.0)
chi = torch.from_numpy(ge.evaluate(lambda x, y: 1. + 0.9 * np.sin(x) * np.sin(y))).cuda()
E.set_chi(chi)
amp = 0.9
b = torch.from_numpy(ge.evaluate(lambda x, y: 2. * np.sin(x) * np.sin(y) * (amp * np.sin(x) * np.sin(y) + 1)
                                 - amp * np.sin(x) ** 2 * np.cos(y) ** 2 - amp * np.cos(x) ** 2 * np.sin(y) ** 2)).cuda()
pcg = PCG(ge.size, 10000)
pcg.set_throw_on_fail(False)
for K in (100, 400):
    pcg.set_max(K)
    xs = torch.zeros(ge.size, dtype=torch.float64, device="cuda")
    pcg.solve(E, xs, b, E.precond(), E.weights(), 1e-30, 1.0, 1)
    xs.zero_()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    it = pcg.solve(E, xs, b, E.precond(), E.weights(), 1e-30, 1.0, 1)
    e1.record()
    torch.cuda.synchronize()
    t = e0.elapsed_time(e1) * 1e-3
    print(f"PCG fixed {it} iterations: {t*1e3:.2f} ms  {it/t:.1f} it/s  {128*ge.size*it/t/1e9:.1f} GB/s (128 B/dof/it)", flush=True)
pcg.set_max(10000)
xs = torch.zeros(ge.size, dtype=torch.float64, device="cuda")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
it = pcg.solve(E, xs, b, E.precond(), E.weights(), 1e-8, 1.0, 1)
e1.record()
torch.cuda.synchronize()
print(f"PCG to eps=1e-8: {it} iterations in {e0.elapsed_time(e1):.2f} ms", flush=True)
print("launches:", L.raw["dgb_launch_count"]())
