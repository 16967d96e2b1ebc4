import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from feltor_b200 import toefl as TF
from feltor_b200._dev import dvec, hvec
from oracle import reftoefl as R
N = int(sys.argv[1]) if len(sys.argv) > 1 else 48
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
model = sys.argv[3] if len(sys.argv) > 3 else "global"
js = R.default_params(3, N, N, model__type=model)
ref = R.RefToefl(js)
y0, y1 = ref.init()
ex = TF.Explicit(TF.Parameters(js))
yi = ex.initial_condition()
print("init rel diff", np.abs(hvec(yi[0]) - y0).max(), np.abs(hvec(yi[1]) - y1).max())
def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
# k fixed ERK steps from the reference's initial condition
dt = 0.5
ra, rb, rsec = ref.erk("Bogacki-Shampine-4-2-3", 0., dt, steps, y0, y1)
u0 = [dvec(y0), dvec(y1)]
u1 = [torch.zeros_like(u0[0]), torch.zeros_like(u0[0])]
delta = [torch.zeros_like(u0[0]), torch.zeros_like(u0[0])]
erk = TF.ERKStep("Bogacki-Shampine-4-2-3", u0)
t = 0.
torch.cuda.synchronize(); t0 = time.time()
for k in range(steps):
    t = erk.step(ex, t, u0, u1, dt, delta)
    u0, u1 = u1, u0
    print("step", k, ex.numbers)
torch.cuda.synchronize(); sec = time.time() - t0
a, b = hvec(u0[0]), hvec(u0[1])
print("after %d steps: rel diff y0 %.3e y1 %.3e phi0 %.3e phi1 %.3e  bitwise %s %s" % (steps, rel(a, ra), rel(b, rb), rel(hvec(ex.phi[0]), ref.phi(0)),
      rel(hvec(ex.phi[1]), ref.phi(1)), np.array_equal(a.view(np.int64), ra.view(np.int64)), np.array_equal(b.view(np.int64), rb.view(np.int64))))
print("seconds: ours %.3f reference %.3f (%d rhs calls)" % (sec, rsec, ex.ncalls))
