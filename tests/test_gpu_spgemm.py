"""GPU: dgb_csr_spgemm (A = B C on the device) against the oracle's restatement of the reference's host kernel
dg::detail::spgemm_cpu_kernel (sparsematrix_cpu.h:19-95; oracle pinned on the live reference in tests/test_spgemm.py) and against
the committed golden product of the unmodified reference: row offsets, sorted columns and values bit for bit."""
import os
import numpy as np
import pytest
from util import same_bits

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import gpu_backend
    gpu_backend.require_library_loaded()
    return gpu_backend


def gpu_product(shape, B, Cm):
    from feltor_b200 import blas2
    from feltor_b200._dev import dvec, hvec
    dB, dC = tuple(dvec(a) for a in B), tuple(dvec(a) for a in Cm)
    return tuple(hvec(t) for t in blas2.spgemm(shape[0], shape[1], shape[2], dB, dC))


def check(got, want):
    assert np.array_equal(got[0], want[0]), "row offsets"
    assert np.array_equal(got[1], want[1]), "columns"
    assert same_bits(got[2], want[2]), "values"


def test_spgemm_golden(G):
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "spgemm_golden.npz"))
    got = gpu_product(tuple(int(v) for v in g["shape"]), (g["Bpos"], g["Bidx"], g["Bval"]), (g["Cpos"], g["Cidx"], g["Cval"]))
    check(got, (g["Apos"], g["Aidx"], g["Aval"]))


@pytest.mark.parametrize("seed,kw", [(0, {}), (1, dict(sort=True)), (2, dict(rows=700, mid=900, cols=800, per_b=40, per_c=30)),
                                     (3, dict(rows=33, mid=5, cols=7, per_b=64, per_c=7)),          # many duplicates per column
                                     (4, dict(rows=1, mid=1, cols=1, per_b=1, per_c=1)),
                                     (5, dict(rows=40, mid=400, cols=6000, per_b=150, per_c=30)),   # rows beyond the fast table
                                     (6, dict(rows=3000, mid=2000, cols=3000, per_b=100, per_c=9))])
def test_spgemm_vs_oracle(G, seed, kw):
    from oracle import orc
    from test_spgemm import random_pair
    shape, B, Cm = random_pair(seed, **kw)
    want = orc.spgemm(shape[0], shape[2], B, Cm)
    check(gpu_product(shape, B, Cm), want)
    if seed == 5:
        assert np.diff(want[0]).max() > 512, "the case does not reach the large-table variant"


def test_spgemm_projection_times_interpolation(G):
    """the product dg::geo::Fieldaligned forms (fieldaligned.h:549-735): a fine-to-coarse projection (every coarse row reads the
    (n mx) x (n my) fine points of its cell) times an interpolation with n^2 entries per fine row into neighbouring coarse cells"""
    from oracle import orc
    n, Nx, Ny, mx, my = 3, 10, 8, 4, 3
    r = np.random.default_rng(3)
    fx, fy = n * Nx * mx, n * Ny * my
    coarse, fine = n * n * Nx * Ny, fx * fy
    Bpos, Bidx = [0], []
    for cy in range(Ny):
        for ky in range(n):
            for cx in range(Nx):
                for kx in range(n):
                    cols = [(cy * n * my + j) * fx + cx * n * mx + i for j in range(n * my) for i in range(n * mx)]
                    Bidx += cols
                    Bpos.append(len(Bidx))
    Bpos, Bidx = np.array(Bpos, dtype=np.int32), np.array(Bidx, dtype=np.int32)
    Bval = r.uniform(-1, 1, Bidx.size)
    Cpos = (np.arange(fine + 1) * n * n).astype(np.int32)
    tx = np.clip((np.arange(fine) % fx) // (n * mx) + r.integers(-1, 2, fine), 0, Nx - 1)      # target cell of the field line
    ty = np.clip((np.arange(fine) // fx) // (n * my) + r.integers(-1, 2, fine), 0, Ny - 1)
    Cidx = np.concatenate([[((ty[f] * n + j) * Nx + tx[f]) * n + i for j in range(n) for i in range(n)] for f in range(fine)]).astype(np.int32)
    Cval = r.uniform(-1, 1, Cidx.size)
    shape = (coarse, fine, coarse)
    want = orc.spgemm(coarse, coarse, (Bpos, Bidx, Bval), (Cpos, Cidx, Cval))
    check(gpu_product(shape, (Bpos, Bidx, Bval), (Cpos, Cidx, Cval)), want)
    from feltor_b200 import blas2
    check(blas2.spgemm_host(*shape, (Bpos, Bidx, Bval), (Cpos, Cidx, Cval)), want)


def test_spgemm_too_many_columns(G):
    from feltor_b200 import DgbError
    r = np.random.default_rng(0)
    mid, cols = 300, 20000
    Bpos, Bidx, Bval = np.array([0, mid], dtype=np.int32), np.arange(mid, dtype=np.int32), r.uniform(-1, 1, mid)
    Cpos = (np.arange(mid + 1) * 20).astype(np.int32)
    Cidx, Cval = np.arange(mid * 20, dtype=np.int32), r.uniform(-1, 1, mid * 20)          # 6000 distinct columns in one row
    with pytest.raises(DgbError) as e:
        gpu_product((1, mid, cols), (Bpos, Bidx, Bval), (Cpos, Cidx, Cval))
    assert e.value.code == -2
