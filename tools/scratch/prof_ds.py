import os, sys, ctypes as C
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tools"))
import ds_bench
out = ds_bench.run_real(96, 64, reps=2, methods=("dg",), cpu_reps=1)
print(out[0]["celltile_us"])
