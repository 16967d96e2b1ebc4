import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
from oracle import reftoefl
from feltor_b200 import toefl as TF
from feltor_b200._dev import dvec, hvec
N = int(sys.argv[1]) if len(sys.argv) > 1 else 416
js = reftoefl.default_params(3, N, N, model__type="global")
ref = reftoefl.RefToefl(js)
y0, y1 = ref.init()
ra, rb, _ = ref.erk("Bogacki-Shampine-4-2-3", 0., 0.5, 1, y0, y1)
ex = TF.Explicit(TF.Parameters(js))
u0 = [dvec(y0), dvec(y1)]
u1 = [torch.zeros_like(u0[0]), torch.zeros_like(u0[0])]
delta = [torch.zeros_like(u0[0]), torch.zeros_like(u0[0])]
erk = TF.ERKStep("Bogacki-Shampine-4-2-3", u0)
erk.step(ex, 0., u0, u1, 0.5, delta)
for name, a, b in (("y0", hvec(u1[0]), ra), ("y1", hvec(u1[1]), rb), ("phi0", hvec(ex.phi[0]), ref.phi(0)), ("phi1", hvec(ex.phi[1]), ref.phi(1))):
    print(os.environ.get("DGB_PCG_NO_FOLD"), os.environ.get("DGB_ELLIPTIC_TILE"), name, "bitwise", np.array_equal(a.view(np.int64), b.view(np.int64)), "maxrel", np.abs(a - b).max() / np.abs(b).max(), flush=True)
print("numbers", ex.numbers)
