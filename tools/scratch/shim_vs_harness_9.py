import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch
import shim_toefl_bench as s
m = s.load()
import feltor_b200
from feltor_b200 import toefl as TF
from feltor_b200._dev import dvec, hvec
m.lib().ref_set_fusion(1)
T = m.RefToefl(m.default_params(3, 1024, 1024))
a, b = T.init()
print("=== shim 9 steps in one call", flush=True)
a9, b9, sec = T.erk("Bogacki-Shampine-4-2-3", 0., 0.5, 9, a, b)
print("=== shim seconds", sec, flush=True)
ex = TF.Explicit(TF.Parameters(m.default_params(3, 1024, 1024)))
u0 = [dvec(a), dvec(b)]
u1 = [torch.zeros_like(u0[0]), torch.zeros_like(u0[0])]
delta = [torch.zeros_like(u0[0]), torch.zeros_like(u0[0])]
erk = TF.ERKStep("Bogacki-Shampine-4-2-3", u0)
calls = []
def rhs(t, y, yp):
    ex(t, y, yp)
    calls.append((t, dict(ex.numbers)))
t = 0.
for k in range(9):
    torch.cuda.synchronize(); t0 = time.time()
    n0 = len(calls)
    t = erk.step(rhs, t, u0, u1, 0.5, delta)
    torch.cuda.synchronize()
    u0, u1 = u1, u0
    print("=== harness step %d: %.1f ms, rhs calls %d" % (k + 1, (time.time() - t0) * 1e3, len(calls) - n0), [c[1]["pol"] for c in calls[n0:]], flush=True)
print("=== max rel diff", np.abs(hvec(u0[0]) - a9).max() / np.abs(a9).max())
