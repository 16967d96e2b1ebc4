"""GPU side of the parity tests: everything goes through the C ABI of libdgb200.so (feltor_b200 harness)."""
import numpy as np
import torch
import feltor_b200 as fb
from feltor_b200 import blas1, blas2
from feltor_b200._dev import dvec, hvec


def make(a):
    return dvec(np.asarray(a, dtype=np.float64))


def get(t):
    return hvec(t)


def dot2(x, y):
    return blas2.dot(x, y)


def dot3(x, w, y):
    return blas2.dot(x, w, y)


_cache = {}


def ell(m):
    """device plan of a host Ell (cached per object)"""
    key = id(m)
    if key not in _cache:
        _cache[key] = (m, blas2.Ell.from_like(m))
    return _cache[key][1]


def symv(m, alpha, x, beta, y, generic=False):
    ell(m).symv(alpha, x, beta, y, generic=generic)


def require_library_loaded():
    """fail loudly if the native library is not the thing that runs"""
    L = fb.lib()
    assert L.cdll is not None
    n0 = L.raw["dgb_launch_count"]()
    t = torch.zeros(8, dtype=torch.float64, device="cuda")
    blas1.plus(t, 1.0)
    torch.cuda.synchronize()
    assert L.raw["dgb_launch_count"]() == n0 + 1
    assert float(t.sum()) == 8.0
