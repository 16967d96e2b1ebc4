import sys, os
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import numpy as np, ctypes as C
import kat
from oracle import orc, refwrap
import test_gpu_shim as t
shim = t._clone("shimwrap", "refwrap.py", "LIB_PATH", t.SHIM)
for name, kind, ops, gold in kat.evaluation_cases():
    if kind == "dot2":
        acc_s, st = shim.dot2(ops[0], ops[1]); acc_o, _ = orc.exdot2(ops[0], ops[1])
    else:
        acc_s, st = shim.dot3(*ops); acc_o, _ = orc.exdot3(*ops)
    no = orc.normalize(acc_o)
    print(name, ops[0].size, "acc equal", np.array_equal(acc_s, no), "round orc", np.float64(orc.round_acc(no)).view(np.int64), "round ref(shimacc)", np.float64(refwrap.round_acc(acc_s)).view(np.int64), "gold", gold)
    out = C.c_double()
    if kind == "dot2":
        shim.lib().ref_blas1_dot(ops[0].size, shim.dp(ops[0]), shim.dp(ops[1]), C.byref(out))
        print("   blas1_dot", np.float64(out.value).view(np.int64))
        # shim-side round
        print("   shim round", np.float64(shim.round_acc(acc_s)).view(np.int64))
