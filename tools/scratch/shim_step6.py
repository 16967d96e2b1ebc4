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
# harness from the same initial state
ex = TF.Explicit(TF.Parameters(m.default_params(3, 1024, 1024)))
u0 = [dvec(a), dvec(b)]
u1 = [torch.zeros_like(u0[0]), torch.zeros_like(u0[0])]
delta = [torch.zeros_like(u0[0]), torch.zeros_like(u0[0])]
t = 0.
for k in range(8):
    print("=== shim step", k + 1, flush=True)
    a, b, sec = T.erk("Bogacki-Shampine-4-2-3", t, 0.5, 1, a, b)
    erk = TF.ERKStep("Bogacki-Shampine-4-2-3", u0)   # fresh stepper like the wrapper's: 4 right-hand sides
    torch.cuda.synchronize(); t0 = time.time()
    erk.step(ex, t, u0, u1, 0.5, delta)
    torch.cuda.synchronize(); hs = time.time() - t0
    u0, u1 = u1, u0
    t += 0.5
    d0 = np.abs(hvec(u0[0]) - a).max() / np.abs(a).max()
    print("=== step %d: shim %.1f ms, harness %.1f ms, harness last numbers %s, max rel diff of the states %.2e" % (k + 1, sec * 1e3, hs * 1e3, ex.numbers, d0), flush=True)
