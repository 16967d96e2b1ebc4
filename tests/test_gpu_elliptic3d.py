"""GPU: dg::Elliptic1d and dg::Elliptic3d in its compute-in-2d mode (inc/dg/elliptic.h:557-797, the mode src/feltor/feltor.h runs) through
dgb_elliptic2d_symv_planes, against the UNMODIFIED reference class (oracle/_ref/libdgref.so, live) on Cartesian and
cylindrical 3-d grids.  Tolerance of the north star for symv: 1e-12 relative (measured: bit-identical)."""
import numpy as np
import pytest
from util import same_bits, rng

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import gpu_backend
    gpu_backend.require_library_loaded()
    return gpu_backend


@pytest.mark.parametrize("cyl", [False, True])
@pytest.mark.parametrize("direction", [0, 2])
@pytest.mark.parametrize("chi_weight_jump,N", [(False, [12, 10, 5]), (True, [12, 10, 5]), (False, [70, 66, 4]), (False, [33, 40, 3])])
def test_elliptic3d_compute_in_2d(G, cyl, direction, chi_weight_jump, N):
    """chi_weight_jump = False: all planes in one launch of the tile kernel (interior tiles by TMA from the stacked planes,
    tiles at a plane's y boundary by the boundary-aware loader); True: per-plane unfused path"""
    from oracle import refwrap as R
    if not R.available():
        pytest.skip("oracle/_ref/libdgref.so not present")
    from feltor_b200 import topology as T
    from feltor_b200.elliptic import Elliptic3d
    x0, x1 = ([3., -1., 0.], [5., 1., 2 * np.pi]) if cyl else ([0., 0., 0.], [1., 2., 3.])
    bc = [T.DIR, T.NEU if cyl else T.PER, T.PER]
    g = T.Grid(x0, x1, [3, 3, 1], N, bc)
    rg = R.grid(x0, x1, 3, N, bc)
    r = rng(7 + 2 * cyl + direction)
    n = g.size
    x, y0, chi = r.uniform(-1, 1, n), r.uniform(-1, 1, n), r.uniform(0.5, 2., n)
    op = Elliptic3d(g, direction=direction, jfactor=0.7, chi_weight_jump=chi_weight_jump, cylindrical=cyl)
    for use_chi, (alpha, beta) in ((False, (1., 0.)), (True, (1., 0.)), (True, (-0.5, 0.3))):
        want, w, p = R.elliptic3d_symv(rg, cyl, direction, 0.7, chi_weight_jump, chi if use_chi else None, alpha, x, beta, y0)
        if use_chi:
            op.set_chi(G.make(chi))
        y = G.make(y0 if beta != 0. else np.full(n, np.nan))
        op.symv(alpha, G.make(x), beta, y)
        got = G.get(y)
        scale = np.abs(want).max()
        assert np.abs(got - want).max() <= 1e-12 * scale, (cyl, direction, use_chi, np.abs(got - want).max() / scale)
        assert same_bits(got, want), "within tolerance but not bit-identical"
        assert same_bits(G.get(op.weights()), w) and same_bits(G.get(op.precond()), p)


def test_elliptic3d_rejects_wrong_size(G):
    from feltor_b200 import topology as T
    from feltor_b200.elliptic import Elliptic3d
    g = T.Grid([0., 0., 0.], [1., 1., 1.], [3, 3, 1], [4, 4, 3], [T.PER, T.PER, T.PER])
    op = Elliptic3d(g)
    with pytest.raises(ValueError):
        op.symv(G.make(np.zeros(7)), G.make(np.zeros(g.size)))


def test_helmholtz_planes_equal_per_plane_calls(G):
    """a Helmholtz plan (chi x - alpha Elliptic x, chi and sigma 3-d) on stacked planes == the 2-d calls plane by plane"""
    import ctypes as C
    import torch
    from feltor_b200 import topology as T, lib
    from feltor_b200._dev import ptr, stream
    from feltor_b200.elliptic import Elliptic2d
    g = T.Grid([0., 0.], [1., 2.], 3, [40, 36], [T.DIR, T.PER])
    nz, n2 = 6, g.size
    r = rng(77)
    x, sig, chi = (G.make(r.uniform(0.5, 1.5, n2 * nz)) for _ in range(3))
    op = Elliptic2d(g, direction=T.CENTERED, jfactor=1.)
    want = torch.zeros(n2 * nz, dtype=torch.float64, device="cuda")
    for k in range(nz):
        lib().elliptic2d_set_sigma(op.h, ptr(sig[k * n2:(k + 1) * n2]))
        lib().elliptic2d_set_helmholtz(op.h, 1, C.c_double(-0.5), ptr(chi[k * n2:(k + 1) * n2]))
        lib().elliptic2d_symv(op.h, C.c_double(1.), ptr(x[k * n2:(k + 1) * n2]), C.c_double(0.), ptr(want[k * n2:(k + 1) * n2]), stream())
    lib().elliptic2d_set_helmholtz(op.h, 1, C.c_double(-0.5), ptr(chi))
    got = torch.full((n2 * nz,), float("nan"), dtype=torch.float64, device="cuda")
    lib().elliptic2d_symv_planes(op.h, nz, ptr(sig), C.c_double(1.), ptr(x), C.c_double(0.), ptr(got), stream())
    assert same_bits(G.get(got), G.get(want))
    lib().elliptic2d_set_helmholtz(op.h, 0, C.c_double(0.), None)
    lib().elliptic2d_set_sigma(op.h, ptr(op._sigma))


@pytest.mark.parametrize("bcx", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("direction", [0, 1, 2])
def test_elliptic1d(G, bcx, direction):
    """dg::Elliptic1d (elliptic.h:65-200) against the unmodified reference class, bit for bit"""
    from oracle import refwrap as R
    if not R.available():
        pytest.skip("oracle/_ref/libdgref.so not present")
    from feltor_b200 import topology as T
    from feltor_b200.elliptic import Elliptic1d
    g = T.Grid([0.3], [2.1], 3, [37], [bcx])
    rg = R.grid([0.3], [2.1], 3, [37], [bcx])
    r = rng(11 + bcx)
    n = g.size
    x, y0, chi = r.uniform(-1, 1, n), r.uniform(-1, 1, n), r.uniform(0.5, 2., n)
    op = Elliptic1d(g, direction=direction, jfactor=0.7)
    for use_chi, (alpha, beta) in ((False, (1., 0.)), (True, (1., 0.)), (True, (-0.5, 0.3))):
        want, w, p = R.elliptic1d_symv(rg, bcx, direction, 0.7, chi if use_chi else None, alpha, x, beta, y0)
        if use_chi:
            op.set_chi(G.make(chi))
        y = G.make(y0 if beta != 0. else np.full(n, np.nan))
        op.symv(alpha, G.make(x), beta, y)
        assert same_bits(G.get(y), want), (bcx, direction, use_chi, alpha, beta)
        assert same_bits(G.get(op.weights()), w) and same_bits(G.get(op.precond()), p)


class _GpuBackend:
    """dg-shaped primitives of the C ABI on device vectors: the twin of tests/test_elliptic3d_oracle._OrcBackend"""

    def __init__(self, G):
        from feltor_b200 import blas1
        self.G, self.b = G, blas1

    def make(self, a): return self.G.make(a)
    def symv(self, m, a, x, b, y): self.G.symv(m, a, x, b, y)
    def tensor_multiply3d(self, lam, t, ins, mu, outs): self.b.tensor_multiply3d(lam, t, list(ins), mu, list(outs))
    def tensor_multiply2d(self, lam, t, i0, i1, mu, o0, o1): self.b.tensor_multiply2d(lam, t, i0, i1, mu, o0, o1)
    def pointwiseDot(self, *a): self.b.pointwiseDot(*a)
    def pointwiseDivide(self, *a): self.b.pointwiseDivide(*a)
    def axpbypgz(self, *a): self.b.axpbypgz(*a)
    def scal(self, x, a): self.b.scal(x, a)


@pytest.mark.parametrize("cyl,direction,cwj,in2d", [(0, 0, 0, 0), (1, 2, 0, 0), (0, 1, 1, 0), (1, 0, 1, 0), (1, 2, 1, 1), (0, 2, 0, 1)])
def test_elliptic3d_full_3d_mode(G, cyl, direction, cwj, in2d):
    """dg::Elliptic3d::symv in its FULL 3-d mode (z derivative through the 3-d Ell matrices, TensorMultiply3d, elliptic.h:688-697)
    composed call by call from the C ABI (dgb_ell_symv in x / y / z, dgb_tensor_multiply3d, blas1): bitwise equal to the
    reference's outputs committed in tests/golden/elliptic3d_golden.npz (and to the oracle composition the CPU suite pins)"""
    import os
    from feltor_b200 import topology as T
    from util import elliptic3d_full_symv
    gold3 = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "elliptic3d_golden.npz"))
    x0, x1 = ([3., -1., 0.], [5., 1., 2 * np.pi]) if cyl else ([0., 0., 0.], [1., 2., 3.])
    N, bc = [9, 7, 5], [1, 4 if cyl else 0, 0]
    x, y0, chi = gold3["x_full"], gold3["y0_full"], gold3["chi_full"]
    y = G.make(y0)
    elliptic3d_full_symv(_GpuBackend(G), T, x0, x1, N, bc, direction, 0.7, bool(cwj), bool(cyl), G.make(chi), -0.5, G.make(x), 0.3, y, bool(in2d))
    assert same_bits(G.get(y), gold3[f"e3dfull/cyl{cyl}/dir{direction}/cwj{cwj}/in2d{in2d}/y"])
