import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from feltor_b200 import topology as T
from feltor_b200.elliptic import Elliptic2d
N = int(sys.argv[1]) if len(sys.argv) > 1 else 64
g = T.Grid([0., 0.], [np.pi, 2 * np.pi], 3, [N, N], [T.DIR, T.DIR])
E = Elliptic2d(g, T.DIR, T.DIR, T.FORWARD, 1.0)
x = torch.rand(g.size, dtype=torch.float64, device="cuda")
y = torch.empty_like(x)
y2 = torch.empty_like(x)
E.symv(x, y)
torch.cuda.synchronize()
E.symv(x, y2, unfused=True)
torch.cuda.synchronize()
print("equal", torch.equal(y.view(torch.int64), y2.view(torch.int64)), float((y - y2).abs().max()))
