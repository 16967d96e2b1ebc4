import os, sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import feltor_b200 as fb
from feltor_b200 import blas2
from feltor_b200._dev import ptr, stream
L = fb.lib()
ws = blas2.DotWorkspace()
res = torch.zeros(41, dtype=torch.int64, device="cuda")
for n in (1024, 65536, 1 << 20, 9437184):
    a = torch.rand(n, dtype=torch.float64, device="cuda"); b = torch.rand(n, dtype=torch.float64, device="cuda")
    for _ in range(5): L.exdot2(ws.h, n, ptr(a), C.c_double(0), ptr(b), C.c_double(0), ptr(res), stream())
    torch.cuda.synchronize()
    # back-to-back (throughput of the launch pipeline) and isolated (latency)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): L.exdot2(ws.h, n, ptr(a), C.c_double(0), ptr(b), C.c_double(0), ptr(res), stream())
    e1.record(); torch.cuda.synchronize()
    t_b2b = e0.elapsed_time(e1) / 50 * 1e3
    ts = []
    for _ in range(20):
        e0.record(); L.exdot2(ws.h, n, ptr(a), C.c_double(0), ptr(b), C.c_double(0), ptr(res), stream()); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    t2 = []
    c = torch.empty_like(a)
    for _ in range(20):
        e0.record(); c.copy_(a); e1.record(); torch.cuda.synchronize()
        t2.append(e0.elapsed_time(e1) * 1e3)
    print("n=%9d exdot2 back-to-back %.1f us  isolated %.1f us   (torch copy isolated %.1f us)" % (n, t_b2b, np.median(ts), np.median(t2)), flush=True)
