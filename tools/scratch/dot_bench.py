import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from feltor_b200 import blas1, blas2
n = 9437184
gen = torch.Generator(device="cuda").manual_seed(0)
v = [torch.rand(n, dtype=torch.float64, device="cuda", generator=gen) + 0.5 for _ in range(3)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timeit(f, nbytes, name):
    for _ in range(3): f()
    ts = []
    for _ in range(30):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    t = float(np.median(ts))
    print("variant %s %-8s %7.1f us %7.1f GB/s" % (os.environ.get("DGB_DOT_VARIANT"), name, t * 1e6, nbytes / t / 1e9), flush=True)
import ctypes as C
import feltor_b200 as fb
from feltor_b200._dev import ptr, stream
ws = blas2.DotWorkspace()
res = torch.zeros(41, dtype=torch.int64, device="cuda")
L = fb.lib()
timeit(lambda: L.exdot2(ws.h, n, ptr(v[0]), C.c_double(0), ptr(v[1]), C.c_double(0), ptr(res), stream()), 16 * n, "dot2")
timeit(lambda: L.exdot3(ws.h, n, ptr(v[0]), C.c_double(0), ptr(v[1]), C.c_double(0), ptr(v[2]), C.c_double(0), ptr(res), stream()), 24 * n, "dot3")
for mult in (2, 4, 8):
    nn = n * mult
    a = torch.rand(nn, dtype=torch.float64, device="cuda", generator=gen) + 0.5
    b = torch.rand(nn, dtype=torch.float64, device="cuda", generator=gen) + 0.5
    timeit(lambda: L.exdot2(ws.h, nn, ptr(a), C.c_double(0), ptr(b), C.c_double(0), ptr(res), stream()), 16 * nn, "dot2 x%d" % mult)
    timeit(lambda: torch.dot(a, b), 16 * nn, "torch x%d" % mult)
    timeit(lambda: b.copy_(a), 16 * nn, "copy x%d" % mult)
    del a, b
