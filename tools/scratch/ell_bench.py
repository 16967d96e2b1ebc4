#!/usr/bin/env python
"""experiment: time the fused Elliptic apply (plain and PCG-dot variant) and check it bitwise against the unfused path"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from feltor_b200 import topology as T
from feltor_b200.elliptic import Elliptic2d, PCG
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for bcx, bcy, name in ((T.DIR, T.PER, "DIRxPER"), (T.PER, T.PER, "PERxPER")):
    for direction, dname in ((T.FORWARD, "fwd"), (T.CENTERED, "cen")):
        g = T.Grid([0., 0.], [np.pi, 2 * np.pi], 3, [N, N], [bcx, bcy])
        E = Elliptic2d(g, bcx, bcy, direction, 1.0)
        E.set_chi(torch.from_numpy(g.evaluate(lambda x, y: 1. + 0.9 * np.sin(x) * np.sin(y)).copy()).cuda())
        x = torch.from_numpy(g.evaluate(lambda x, y: np.sin(x) * np.cos(3 * y) + 0.1 * x).copy()).cuda()
        y = torch.empty_like(x); y2 = torch.empty_like(x)
        E.symv(x, y)
        E.symv(x, y2, unfused=True)
        same = bool((y.view(torch.int64) == y2.view(torch.int64)).all())
        ts = []
        for _ in range(20):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); E.symv(x, y); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        t = float(np.median(ts))
        print(f"{name} {dname}: fused {t:7.1f} us  {24*g.size/t/1e3:7.1f} GB/s  bitwise==unfused: {same}", flush=True)
g = T.Grid([0., 0.], [np.pi, 2 * np.pi], 3, [N, N], [T.DIR, T.PER])
E = Elliptic2d(g, T.DIR, T.PER, T.CENTERED if os.environ.get("DIRN") == "cen" else T.FORWARD, 1.0)
E.set_chi(torch.from_numpy(g.evaluate(lambda x, y: 1. + 0.9 * np.sin(x) * np.sin(y)).copy()).cuda())
b = torch.from_numpy(g.evaluate(lambda x, y: np.sin(x) * np.sin(y)).copy()).cuda()
x = torch.zeros_like(b)
pcg = PCG(g.size, 201); pcg.set_throw_on_fail(False)
for rep in range(2):
    x.zero_(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); it = pcg.solve(E, x, b, E.precond(), E.weights(), 1e-30, 1.0, 1); e1.record(); torch.cuda.synchronize()
print(f"PCG 200 it: {e0.elapsed_time(e1):.2f} ms  {200/e0.elapsed_time(e1)*1e3:.0f} it/s")
