import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import shim_toefl_bench as s
m = s.load()
import feltor_b200
m.lib().ref_set_fusion(1)
T = m.RefToefl(m.default_params(3, 1024, 1024))
y0, y1 = T.init()
a, b, _ = T.erk("Bogacki-Shampine-4-2-3", 0., 0.5, 2, y0, y1)
print("=== timed step", flush=True)
t0 = time.time()
a, b, sec = T.erk("Bogacki-Shampine-4-2-3", 1.0, 0.5, 1, a, b)
print("=== one step: erk seconds", sec, "wall incl. copies", time.time() - t0, flush=True)
