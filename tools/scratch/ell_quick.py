import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from feltor_b200 import topology as T
from feltor_b200.elliptic import Elliptic2d
N = 1024
BC = os.environ.get("BC", "DIR")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
g = T.Grid([0., 0.], [np.pi, 2 * np.pi], 3, [N, N], [getattr(T, BC), T.PER])
E = Elliptic2d(g, getattr(T, BC), T.PER, T.FORWARD, 1.0)
E.set_chi(torch.from_numpy(g.evaluate(lambda x, y: 1. + 0.9 * np.sin(x) * np.sin(y)).copy()).cuda())
x = torch.from_numpy(g.evaluate(lambda x, y: np.sin(x) * np.cos(3 * y) + 0.1 * x).copy()).cuda()
y = torch.empty_like(x)
E.symv(x, y)
ts = []
for _ in range(30):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); E.symv(x, y); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
print(os.environ.get("DGB_WALK_SLOW_WEIGHT"), os.environ.get("DGB_WALK_MANUAL_WEIGHT"), BC, "fused %.1f us (min %.1f)" % (float(np.median(ts)), min(ts)), flush=True)
