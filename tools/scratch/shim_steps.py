import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import shim_toefl_bench as s
m = s.load()
import feltor_b200
m.lib().ref_set_fusion(1)
saved, devnull = os.dup(1), os.open(os.devnull, os.O_WRONLY)
sys.stdout.flush(); os.dup2(devnull, 1)
T = m.RefToefl(m.default_params(3, 1024, 1024))
a, b = T.init()
out = []
t = 0.
for k in range(8):
    a, b, sec = T.erk("Bogacki-Shampine-4-2-3", t, 0.5, 1, a, b)
    t += 0.5
    out.append(sec)
a2, b2, sec4 = T.erk("Bogacki-Shampine-4-2-3", t, 0.5, 4, a, b)
os.dup2(saved, 1)
print("single steps (4 RHS each):", ["%.1f ms" % (x * 1e3) for x in out])
print("4 steps in one call: %.1f ms per step" % (sec4 / 4 * 1e3))
