#!/usr/bin/env python
"""exdot timing against vector size (fixed cost vs marginal rate): python tools/dot_sweep.py"""
import os, sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feltor_b200 as fb
from feltor_b200 import blas2
from feltor_b200._dev import ptr, stream
L = fb.lib()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
ws = blas2.DotWorkspace()
res = torch.zeros(41, dtype=torch.int64, device="cuda")
reps = 30

def t_of(f, fl=True):
    for _ in range(3): f()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if fl: flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return float(np.median(ts))

print("empty-kernel event floor:", t_of(lambda: L.fill(1, C.c_double(0.), ptr(res.view(torch.float64)), stream()), False), "us")
for n in (128*128*9, 256*256*9, 512*512*9, 1024*1024*9, 2*1024*1024*9, 4*1024*1024*9, 8*1024*1024*9):
    v = [torch.rand(n, dtype=torch.float64, device="cuda") + 0.5 for _ in range(3)]
    t2 = t_of(lambda: L.exdot2(ws.h, n, ptr(v[0]), C.c_double(0), ptr(v[1]), C.c_double(0), ptr(res), stream()))
    t3 = t_of(lambda: L.exdot3(ws.h, n, ptr(v[0]), C.c_double(0), ptr(v[1]), C.c_double(0), ptr(v[2]), C.c_double(0), ptr(res), stream()))
    tc = t_of(lambda: v[1].copy_(v[0]))
    ts = t_of(lambda: torch.sum(v[0]))
    td = t_of(lambda: torch.dot(v[0], v[1]))
    t2w = t_of(lambda: L.exdot2(ws.h, n, ptr(v[0]), C.c_double(0), ptr(v[1]), C.c_double(0), ptr(res), stream()), False)
    print(f"n={n:9d} ({8*n/1e6:7.1f} MB)  exdot2 {t2:7.1f} us {16*n/t2/1e3:7.0f} GB/s | exdot3 {t3:7.1f} us {24*n/t3/1e3:7.0f} GB/s | copy {tc:7.1f} us {16*n/tc/1e3:7.0f} | torch.sum {ts:7.1f} us {8*n/ts/1e3:7.0f} | torch.dot {td:7.1f} {16*n/td/1e3:7.0f} | exdot2 warm-L2 {t2w:7.1f}", flush=True)
    del v
