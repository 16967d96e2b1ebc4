"""GPU parity of the CSR spmv (single vector and all-planes variant) against the oracle / reference fixtures."""
import ctypes as C
import numpy as np
import pytest
from oracle import orc
from util import same_bits, rng

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import gpu_backend
    gpu_backend.require_library_loaded()
    return gpu_backend


def spmv(G, pos, idx, val, alpha, x, beta, y, nplanes=None, shift=0):
    import torch
    from feltor_b200 import lib
    from feltor_b200._dev import ptr, stream, dvec
    dpos, didx, dval = dvec(pos), dvec(idx), dvec(val)
    nr = len(pos) - 1
    if nplanes is None:
        nc = x.numel()
        lib().csr_spmv(nr, nc, ptr(dpos), ptr(didx), ptr(dval), C.c_double(alpha), ptr(x), C.c_double(beta), ptr(y), stream())
    else:
        nc = x.numel() // nplanes
        lib().csr_spmv_planes(nr, nc, ptr(dpos), ptr(didx), ptr(dval), C.c_double(alpha), ptr(x), C.c_double(beta),
                              ptr(y), nplanes, shift, stream())
    torch.cuda.synchronize()


def test_csr_fixtures(G, golden):
    pos, idx, val, x, y0 = (golden["csr/" + k] for k in ("pos", "idx", "val", "x", "y"))
    for al, be in ((1., 0.), (0.5, 1.), (-2., 0.25)):
        y = G.make(y0 if be != 0. else np.full_like(y0, np.nan))
        spmv(G, pos, idx, val, al, G.make(x), be, y)
        assert same_bits(G.get(y), golden[f"csr/a{al}b{be}"]), (al, be)


def random_csr(r, nr, nc, maxlen):
    counts = r.integers(0, maxlen + 1, nr)
    pos = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    idx = r.integers(0, nc, pos[-1]).astype(np.int32)
    val = r.uniform(-1, 1, pos[-1])
    return pos, idx, val


@pytest.mark.parametrize("nr,nc,maxlen,nplanes,shift", [(1, 1, 1, 1, 0), (300, 257, 40, 1, 0), (300, 300, 80, 5, 1),
                                                        (1000, 1000, 35, 8, -1), (513, 513, 12, 3, 1), (64, 64, 0, 4, 0)])
def test_csr_planes_vs_oracle(G, nr, nc, maxlen, nplanes, shift):
    """y[p] = alpha A x[(p+shift) mod nplanes] + beta y[p] (Fieldaligned::ePlus/eMinus, fieldaligned.h:850-912)"""
    r = rng(nr + nplanes)
    pos, idx, val = random_csr(r, nr, nc, maxlen)
    x = r.uniform(-1, 1, nc * nplanes)
    y0 = r.uniform(-1, 1, nr * nplanes)
    for al, be in ((1., 0.), (0.5, 1.), (-2., 0.25)):
        yo = y0.copy()
        for p in range(nplanes):
            src = (p + shift) % nplanes
            yp = yo[p * nr:(p + 1) * nr]
            orc.csr_spmv(pos, idx, val, al, np.ascontiguousarray(x[src * nc:(src + 1) * nc]), be, yp)
        y = G.make(y0)
        spmv(G, pos, idx, val, al, G.make(x), be, y, nplanes, shift)
        assert same_bits(G.get(y), yo), (al, be)
