"""The reference's OWN headers and application on libdgb200.so.

integration/Makefile compiles oracle/ref_wrap.cpp and oracle/ref_toefl.cpp -- the very wrappers that oracle/Makefile compiles
on the reference's OpenMP backend, the second one around the UNMODIFIED src/toefl/toefl.h -- with nvcc against a copy of the
reference's inc/ tree in which only the dg::CudaTag backend files are replaced by integration/dgb_shim/ (INTEGRATION.md).
These tests drive both builds through the same C calls and compare them:
  * dg::blas1 functions, dg::blas1::dot / dg::blas2::dot, EllSparseBlockMat / CSR symv, dg::Elliptic2d, dg::PCG,
    dg::MultigridCG2d: BITWISE equal to the OpenMP backend (and to the reference's own golden bit patterns) -- they only touch
    library functors, which dispatch into libdgb200.so;
  * toefl::Explicit + dg::ERKStep / dg::Adaptive / dg::ExplicitMultistep: <= 1e-12 relative through the application's device
    lambdas (compiled by nvcc instead of g++), iteration counts equal;
  * the dispatch counters prove the calls landed in the library (dgb_launch_count grows, almost nothing takes a generic kernel).
"""
import ctypes as C
import importlib.util
import os
import sys
import numpy as np
import pytest
from util import same_bits, bits, rng, blas1_sequence, BLAS1_GOLDEN

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "integration", "_build", "libdgshim.so")
SHIM_TOEFL = os.path.join(ROOT, "integration", "_build", "libdgshim_toefl.so")


def _clone(modname, filename, attr, libpath):
    spec = importlib.util.spec_from_file_location(modname, os.path.join(ROOT, "oracle", filename))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    setattr(m, attr, libpath)
    return m


@pytest.fixture(scope="module")
def shim():
    """oracle/refwrap.py bound to the device build of the same wrapper"""
    if not os.path.exists(SHIM):
        pytest.skip("integration/_build/libdgshim.so not built (needs the reference tree at build time)")
    import gpu_backend
    gpu_backend.require_library_loaded()
    m = _clone("shimwrap", "refwrap.py", "LIB_PATH", SHIM)
    assert m.lib().ref_backend_is_device() == 1
    return m


def counters(lib, name="ref_dispatch_counters"):
    a, b = C.c_longlong(), C.c_longlong()
    getattr(lib, name)(C.byref(a), C.byref(b))
    return a.value, b.value


def launches():
    import feltor_b200 as fb
    fn = fb.lib().raw["dgb_launch_count"]
    fn.restype = C.c_longlong
    return fn()


def test_shim_blas1_goldens(shim):
    """inc/dg/blas1_t.cpp:102-184 through the reference's dg::blas1 templates -> binding -> libdgb200: every golden exact"""
    from backends import RefBlas1, np_make, np_get
    l0, lib0 = launches(), counters(shim.lib())
    out = blas1_sequence(RefBlas1(shim.lib()), np_make, np_get)
    for name, gold in BLAS1_GOLDEN:
        assert int(out[name][0]) == gold, name
    lib1 = counters(shim.lib())
    assert lib1[0] - lib0[0] >= len(BLAS1_GOLDEN) and lib1[1] == lib0[1]      # all in the library, none generic
    assert launches() - l0 >= len(BLAS1_GOLDEN)


def test_shim_blas1_random_vs_openmp(shim, ref):
    if ref is None:
        pytest.skip("oracle/_ref/libdgref.so not present")
    from backends import RefBlas1
    A, B = RefBlas1(shim.lib()), RefBlas1(ref.lib())
    r = rng(11)
    n = 10007
    x1, x2, x3, x4 = (r.uniform(-2, 2, n) for _ in range(4))

    def both(f):
        ya, yb = r.uniform(-1, 1, n), None
        yb = ya.copy()
        f(A, ya)
        f(B, yb)
        assert same_bits(ya, yb)
    both(lambda L, y: L.axpby(0.3, x1, -1.7, y))
    both(lambda L, y: L.axpby(0.3, x1, -1.7, x2, y))
    both(lambda L, y: L.axpbypgz(0.3, x1, -1.7, x2, 0.9, y))
    both(lambda L, y: L.pointwiseDot(0.3, x1, x2, -1.7, y))
    both(lambda L, y: L.pointwiseDot(x1, x2, y))
    both(lambda L, y: L.pointwiseDot(0.3, x1, x2, x3, 1.1, y))
    both(lambda L, y: L.pointwiseDot(0.3, x1, x2, 0.7, x3, x4, 1.1, y))
    both(lambda L, y: L.pointwiseDivide(0.3, x1, x2, 1.1, y))
    both(lambda L, y: L.pointwiseDivide(x1, x2, y))
    # transform( x, y, dg::EXP()): exp is the device's vs glibc's -- last-bit differences are the math library's
    ya, yb = np.zeros(n), np.zeros(n)
    A.transform(x1, ya, "exp")
    B.transform(x1, yb, "exp")
    assert np.abs(ya - yb).max() <= 2e-16 * np.abs(yb).max()
    both(lambda L, y: L.scal(y, 1.3))
    both(lambda L, y: L.plus(y, 1.3))


def test_shim_dots_goldens_and_openmp(shim, ref):
    """evaluation_t.cpp goldens through dg::blas1::dot / dg::blas2::dot of the reference templates on the binding"""
    import kat
    from oracle import orc
    out = C.c_double()
    for name, kind, ops, gold in kat.evaluation_cases():
        if kind == "dot2":
            assert shim.lib().ref_blas1_dot(ops[0].size, shim.dp(ops[0]), shim.dp(ops[1]), C.byref(out)) == 0
        else:
            assert shim.lib().ref_blas2_dot(ops[0].size, shim.dp(ops[0]), shim.dp(ops[1]), shim.dp(ops[2]), C.byref(out)) == 0
        oval = orc.dot2(*ops)[0] if kind == "dot2" else orc.dot3(*ops)[0]
        assert same_bits([out.value], [oval]), name                 # same inputs -> the same bits as the oracle
        assert abs(int(bits([out.value])[0]) - gold) < 2, name      # the golden itself within the reference's own tolerance
                                                                    # (blas1_t.cpp:41: the host's exp() differs between CPUs)
    # superaccumulator words: the binding returns them normalised; equal to the OpenMP backend's after Normalize
    r = rng(3)
    from util import wide
    for n in (1, 33, 4097, 100003):
        x, y, w = wide(r, n), wide(r, n), wide(r, n, -5, 5)
        acc_s, st = shim.dot2(x, y)
        acc_o, _ = orc.exdot2(x, y)
        assert st == 0 and np.array_equal(acc_s, orc.normalize(acc_o))
        acc_s, st = shim.dot3(x, w, y)
        acc_o, _ = orc.exdot3(x, w, y)
        assert st == 0 and np.array_equal(acc_s, orc.normalize(acc_o))
    x = r.uniform(-1, 1, 100)
    x[17] = np.nan
    assert shim.lib().ref_blas1_dot(100, shim.dp(x), shim.dp(x), C.byref(out)) == 1     # dg::Error, blas1.h:161


def test_shim_derivative_goldens(shim):
    """derivatives_t.cpp:54-133: dx / dy / dz / jump symv of the reference's EllSparseBlockMat on the binding + dot"""
    import kat
    from oracle import orc

    def symv(m, alpha, x, beta, y):
        meta = np.ascontiguousarray(m.meta(), dtype=np.int32)
        e = shim.ell_from_arrays(meta, m.data, m.cols_idx, m.data_idx)
        e.symv(alpha, x, beta, y)

    def dot2(x, y):
        out = C.c_double()
        shim.lib().ref_blas1_dot(x.size, shim.dp(x), shim.dp(y), C.byref(out))
        return out.value

    def dot3(x, w, y):
        out = C.c_double()
        shim.lib().ref_blas2_dot(x.size, shim.dp(x), shim.dp(w), shim.dp(y), C.byref(out))
        return out.value

    def pdot(x, y, z):
        shim.lib().ref_pointwiseDot_xy(x.size, shim.dp(x), shim.dp(y), shim.dp(z))
    l0 = counters(shim.lib())
    for case in kat.derivative_cases():
        got, gold, gh = kat.run_derivative_case(case, np.array, dot2, dot3, symv, pdot)
        ogot, _, _ = kat.run_derivative_case(case, np.array, lambda x, y: orc.dot2(x, y)[0], lambda x, w, y: orc.dot3(x, w, y)[0],
                                             lambda m, a, x, b, y: orc.ell_symv(m, a, x, b, y), orc.pointwiseDot_xy)
        assert got == ogot, case[0]
        assert abs(got - gold) < 2 or abs(got - gh) < 2, case[0]
    l1 = counters(shim.lib())
    assert l1[0] > l0[0] and l1[1] == l0[1]


@pytest.mark.parametrize("fusion", [1, 0], ids=["fused", "backend-only"])
@pytest.mark.parametrize("N,bcx,bcy,d", [([37, 19], 1, 0, 0), ([24, 40], 4, 1, 2), ([33, 17], 2, 3, 1), ([420, 404], 1, 0, 2)])
def test_shim_elliptic_and_pcg_vs_openmp(shim, ref, N, bcx, bcy, d, fusion):
    """the reference's dg::Elliptic2d / dg::PCG templates: device build on the binding == OpenMP build, bit for bit -- with
    the fused-kernel hooks of dgb_fused.h (one kernel per apply, three per PCG iteration) and with the backend dispatch alone"""
    if ref is None:
        pytest.skip("oracle/_ref/libdgref.so not present")
    shim.lib().ref_set_fusion(fusion)
    outs = []
    l0 = launches()
    for L in (shim, ref):
        g = L.grid([0, 0], [np.pi, 2 * np.pi], 3, N, [bcx, bcy])
        n = L.grid_size(g)
        rr = rng(N[0] + d)
        chi, x, y0 = 1. + rr.uniform(0, 1, n), rr.uniform(-1, 1, n), rr.uniform(-1, 1, n)
        E = L.Elliptic2d(g, bcx, bcy, d, 0.7)
        E.set_chi(chi)
        y = y0.copy()
        E.symv(-0.5, x, 2., y)
        sig = y0.copy()
        E.variation(0.3, chi, x, 1.5, sig)
        xs = np.zeros(n)
        it, _ = E.pcg_solve(xs, x, 1. / chi, E.weights(), 1e-6, 1.0, 1, max_iter=60)
        outs.append((y, sig, xs, it, E.weights(), E.precond()))
        if L is shim:
            used = launches() - l0
    shim.lib().ref_set_fusion(1)
    its = outs[0][3]
    # fused: ~3 launches per PCG iteration; backend-only: the reference's loop (8 per apply + blas1 + 3 dots per iteration)
    assert used < 6 * its + 60 if fusion else used > 12 * its
    for a, b in zip(outs[0], outs[1]):
        if isinstance(a, np.ndarray):
            assert same_bits(a, b)
        else:
            assert a == b


@pytest.mark.parametrize("cyl,d,cwj,in2d", [(0, 0, 0, 0), (1, 2, 0, 0), (0, 1, 1, 0), (1, 0, 1, 0), (1, 2, 0, 1)])
def test_shim_elliptic3d_full_3d_vs_openmp(shim, ref, cyl, d, cwj, in2d):
    """the reference's dg::Elliptic3d in its FULL 3-d mode (z derivative, 3-d tensor product, z jump: elliptic.h:688-697,727-746)
    and its variation, on Cartesian and cylindrical grids: the device build on the binding (Ell symv in x, y and z, blas1 and
    tensor functors -- all library kernels) equals the OpenMP build bit for bit"""
    if ref is None:
        pytest.skip("oracle/_ref/libdgref.so not present")
    outs = []
    g0 = counters(shim.lib())
    l0 = launches()
    for L in (shim, ref):
        g = L.grid([1., -1., 0.], [2., 1., 2 * np.pi], 3, [11, 9, 6], [1, 4, 0])
        n = L.grid_size(g)
        rr = rng(7 + d + 3 * cyl)
        chi, x, y0 = 1. + rr.uniform(0, 1, n), rr.uniform(-1, 1, n), rr.uniform(-1, 1, n)
        y, var = L.elliptic3d_symv_mode(g, cyl, d, 0.7, cwj, in2d, chi, -0.5, x, 2., y0, variation=True)
        outs.append((y, var))
    assert same_bits(outs[0][0], outs[1][0]) and same_bits(outs[0][1], outs[1][1])
    g1 = counters(shim.lib())
    # Ell symv in three directions, blas1 and TensorMultiply2d land in the library; TensorMultiply3d / TensorDot3d run through the
    # binding's generic subroutine template
    assert launches() > l0 and g1[0] > g0[0]


def test_shim_vdot_and_reduce_vs_openmp(shim, ref):
    """dg::blas1::vdot (FPE accumulation, exblas/fpedot_cuda.cuh:66-183 seam) and dg::blas1::reduce (blas1_cuda.cuh:96-102 seam)
    through the binding's single-launch kernel templates, library functor and user functors, against the OpenMP backend:
    vdot is an extended-precision sum rounded at the end (not binary reproducible by the reference's own contract, blas1.h:74)
    -> within 2 ulp; max / min are order independent -> equal; the sum of squares is an ordinary floating-point reduction
    (different association) -> 1e-14"""
    if ref is None:
        pytest.skip("oracle/_ref/libdgref.so not present")
    r = rng(21)
    for n in (1, 7, 1000, 40000, 300001):
        x, y = r.uniform(-1, 1, n), r.uniform(-1, 1, n) * 10. ** r.integers(-3, 4, n)
        for kind in (0, 1):
            a, b = shim.vdot(kind, x, y), ref.vdot(kind, x, y)
            assert abs(a - b) <= 4e-16 * abs(b), (n, kind, a, b)   # blas1.h:74: extended precision, "does not guarantee binary reproducible results"
        for kind in (1, 2):
            assert shim.reduce(kind, x) == ref.reduce(kind, x), (n, kind)
        a, b = shim.reduce(0, x), ref.reduce(0, x)
        assert abs(a - b) <= 1e-14 * abs(b), (n, a, b)


@pytest.mark.parametrize("fusion", [1, 0], ids=["fused", "backend-only"])
def test_shim_multigrid_vs_openmp(shim, ref, fusion):
    """dg::MultigridCG2d (nested iterations, fast projection / interpolation = MultiMatrix of Ell matrices) on the binding"""
    if ref is None:
        pytest.skip("oracle/_ref/libdgref.so not present")
    shim.lib().ref_set_fusion(fusion)
    res = []
    for L in (shim, ref):
        g = L.grid([0, 0], [np.pi, 2 * np.pi], 3, [48, 40], [1, 0])
        n = L.grid_size(g)
        rr = rng(8)
        chi, b = 1. + rr.uniform(0, 1, n), rr.uniform(-1, 1, n)
        M = L.Multigrid(g, 3, 2, 1.0)
        proj = M.project(chi)
        M.set_chi(chi)
        x = np.zeros(n)
        st, num, _ = M.solve(x, b, [1e-8, 1e-7, 1e-7])
        assert st == 0
        res.append((proj, x, num))
    shim.lib().ref_set_fusion(1)
    for a, b in zip(res[0][0], res[1][0]):
        assert same_bits(a, b)
    assert res[0][2] == res[1][2]
    assert same_bits(res[0][1], res[1][1])


def test_shim_csr_vs_openmp(shim, ref):
    """dg::SparseMatrix (CSR) symv: the binding replaces cuSPARSE by the library's row-ordered kernel == OpenMP kernel"""
    if ref is None:
        pytest.skip("oracle/_ref/libdgref.so not present")
    r = rng(4)
    nr, nc = 1500, 1700
    cnt = r.integers(0, 9, nr)
    pos = np.zeros(nr + 1, dtype=np.int32)
    pos[1:] = np.cumsum(cnt)
    idx = r.integers(0, nc, pos[-1]).astype(np.int32)
    val = r.uniform(-1, 1, pos[-1])
    x = r.uniform(-1, 1, nc)
    for al, be in ((1., 0.), (0.5, 1.), (-2., 0.3)):
        ya, yb = r.uniform(-1, 1, nr), None
        yb = ya.copy()
        shim.csr_symv(nr, nc, pos, idx, val, al, x, be, ya)
        ref.csr_symv(nr, nc, pos, idx, val, al, x, be, yb)
        assert same_bits(ya, yb), (al, be)


# ------------------------------------------------------------------------------------------------ toefl
@pytest.fixture(scope="module")
def toefl_pair():
    from oracle import reftoefl
    if not os.path.exists(SHIM_TOEFL):
        pytest.skip("integration/_build/libdgshim_toefl.so not built")
    if not reftoefl.available():
        pytest.skip("oracle/_ref/libdgref_toefl.so not present")
    import gpu_backend
    gpu_backend.require_library_loaded()
    dev = _clone("shimtoefl", "reftoefl.py", "_PATH", SHIM_TOEFL)
    assert dev.lib().ref_toefl_backend_is_device() == 1 and reftoefl.lib().ref_toefl_backend_is_device() == 0
    return dev, reftoefl


def rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)


@pytest.mark.parametrize("fusion", [1, 0], ids=["fused", "backend-only"])
@pytest.mark.parametrize("model,N", [("global", 48), ("local", 40)])
def test_shim_toefl_rhs_and_steps(toefl_pair, model, N, fusion):
    """UNMODIFIED toefl::Explicit + dg::ERKStep compiled on the binding vs the OpenMP build: right-hand side, potentials, PCG
    iteration numbers and the state after fixed Bogacki-Shampine steps within 1e-12 (the device lambdas of toefl.h are
    compiled by nvcc instead of g++; everything else is bit-identical)"""
    dev, omp = toefl_pair
    dev.lib().ref_set_fusion(fusion)
    js = omp.default_params(3, N, N, model__type=model)
    D, O = dev.RefToefl(js), omp.RefToefl(js)
    y0, y1 = O.init()
    d0, d1 = D.init()
    assert rel(d0, y0) < 1e-13 and rel(d1, y1) < 1e-12
    l0, c0 = launches(), counters(dev.lib())
    ra = O.rhs(0., y0, y1)
    da = D.rhs(0., y0, y1)
    assert rel(da[0], ra[0]) < 1e-12 and rel(da[1], ra[1]) < 1e-12
    assert rel(D.phi(0), O.phi(0)) < 1e-12 and rel(D.phi(1), O.phi(1)) < 1e-12
    c1 = counters(dev.lib())
    assert launches() - l0 > 100                       # the right-hand side ran inside libdgb200.so ...
    assert c1[0] - c0[0] > 50 and (c1[1] - c0[1]) * 5 < (c1[0] - c0[0])     # ... and all but the app's own lambdas went there
    oa, ob, _ = O.erk("Bogacki-Shampine-4-2-3", 0., 0.5, 3, y0, y1)
    xa, xb, _ = D.erk("Bogacki-Shampine-4-2-3", 0., 0.5, 3, y0, y1)
    dev.lib().ref_set_fusion(1)
    assert rel(xa, oa) < 1e-12 and rel(xb, ob) < 1e-12
    assert D.ncalls() == O.ncalls()


def test_shim_toefl_adaptive_and_multistep(toefl_pair):
    """dg::Adaptive<dg::ERKStep> with pid_control / l2norm as src/toefl/toefl.cpp:88-91 and dg::ExplicitMultistep (config 3's
    "multistep") -- the reference's own stepper templates on the binding: same accepted steps, states within 1e-12"""
    dev, omp = toefl_pair
    js = omp.default_params(3, 48, 48)
    D, O = dev.RefToefl(js), omp.RefToefl(js)
    y0, y1 = O.init()
    o = O.adaptive("Bogacki-Shampine-4-2-3", 0., 1e-3, 4, 1e-5, 1e-6, y0, y1)
    d = D.adaptive("Bogacki-Shampine-4-2-3", 0., 1e-3, 4, 1e-5, 1e-6, y0, y1)
    assert rel(d[0], o[0]) < 1e-12 and rel(d[1], o[1]) < 1e-12
    assert abs(d[2] - o[2]) <= 1e-12 * abs(o[2]) and np.allclose(d[3], o[3], rtol=1e-9, atol=0) and d[4] == o[4]
    o = O.multistep("TVB-3-3", 0., 0.2, 5, y0, y1)
    d = D.multistep("TVB-3-3", 0., 0.2, 5, y0, y1)
    assert rel(d[0], o[0]) < 1e-12 and rel(d[1], o[1]) < 1e-12 and np.array_equal(d[2], o[2])


def test_shim_toefl_operators(toefl_pair):
    """Advection::upwind / ArakawaX / Elliptic::variation / Helmholtz + polarisation multigrid solves of the reference classes
    on the binding: library functors only -> BITWISE equal to the OpenMP backend"""
    dev, omp = toefl_pair
    js = omp.default_params(3, 40, 36)
    D, O = dev.RefToefl(js), omp.RefToefl(js)
    n = O.size
    r = rng(9)
    vx, vy, f, res = (r.uniform(-1, 1, n) for _ in range(4))
    assert same_bits(D.upwind(0.7, vx, vy, f, -0.3, res), O.upwind(0.7, vx, vy, f, -0.3, res))
    assert same_bits(D.arakawa(1.3, vx, f, 0.5, res), O.arakawa(1.3, vx, f, 0.5, res))
    assert same_bits(D.variation(f), O.variation(f))
    xd, nd = D.helmholtz_solve(np.zeros(n), f)
    xo, no = O.helmholtz_solve(np.zeros(n), f)
    assert nd == no and same_bits(xd, xo)
    chi = 1. + r.uniform(0, 1, n)
    xd, nd = D.pol_solve(chi, np.zeros(n), f)
    xo, no = O.pol_solve(chi, np.zeros(n), f)
    assert nd == no and same_bits(xd, xo)


SHIM_FA = os.path.join(ROOT, "integration", "_build", "libdgshim_fa.so")


def test_shim_fieldaligned_and_ds_vs_openmp():
    """the reference's dg::geo::Fieldaligned and dg::geo::DS templates instantiated on DEVICE containers (dg::IDMatrix, dg::DVec)
    and compiled on the binding -- ePlus / eMinus through dgb_csr_spmv, DS::centered / forward / backward / dss / divCentered
    through the class's own device lambdas -- against the same wrapper on the OpenMP backend: identical matrices and fields
    (host construction), bitwise field-line shifts (row-ordered CSR kernel), DS members within 1e-12"""
    from oracle import reffa
    if not os.path.exists(SHIM_FA) or not reffa.available():
        pytest.skip("integration/_build/libdgshim_fa.so or oracle/_ref/libdgref_fa.so not built")
    import gpu_backend
    gpu_backend.require_library_loaded()
    dev = _clone("shimfa", "reffa.py", "_PATH", SHIM_FA)
    assert dev.lib().ref_fa_is_device() == 1 and reffa.lib().ref_fa_is_device() == 0
    l0 = launches()
    objs = [m.RefFieldaligned(3, 14, 12, 8, 5, 5, "dg") for m in (dev, reffa)]
    D, H = objs
    for which in ("plus", "minus"):
        for a, b in zip(D.csr(which), H.csr(which)):
            assert np.array_equal(a, b)
    for name in ("bphi", "bphiM", "bphiP", "sqrtG", "hbm", "hbp"):
        assert same_bits(D.field(name), H.field(name))
    f = H.testfunction()
    assert same_bits(D.testfunction(), f)
    r = rng(3)
    g0 = r.uniform(-1, 1, f.size)
    for which in ("plus", "minus"):
        assert same_bits(D.shift(which, f), H.shift(which, f)), which
    for kind in ("centered", "forward", "backward", "dss", "divCentered"):
        for alpha, beta in ((1., 0.), (-0.5, 0.3)):
            gd, _ = D.ds(kind, alpha, f, beta, g0)
            gh, _ = H.ds(kind, alpha, f, beta, g0)
            scale = np.max(np.abs(gh))
            assert np.max(np.abs(gd - gh)) <= 1e-12 * scale, (kind, alpha, beta, np.max(np.abs(gd - gh)) / scale)
    assert launches() > l0, "the device build did not reach libdgb200.so"


SHIM_FELTOR = os.path.join(ROOT, "integration", "_build", "libdgshim_feltor.so")


@pytest.mark.parametrize("dims", [None, (24, 30, 8, 2, 2)])
def test_shim_feltor_explicit_vs_openmp(dims):
    """the UNMODIFIED 3-d application class feltor::Explicit (src/feltor/feltor.h: Elliptic3d + Helmholtz multigrid solves, staggered
    Fieldaligned / DS parallel derivatives, perpendicular advection and diffusion on device containers) compiled on the binding
    against the same wrapper (oracle/ref_feltor.cpp) on the reference's OpenMP backend: the four right-hand sides and both
    potentials of three consecutive evaluations (explicit Euler steps in between) agree to 1e-10 (measured: 5e-14 .. 3e-13);
    the work runs inside libdgb200.so (launch counter)"""
    from oracle import reffeltor
    if not os.path.exists(SHIM_FELTOR) or not reffeltor.available():
        pytest.skip("integration/_build/libdgshim_feltor.so or oracle/_ref/libdgref_feltor.so not built")
    import gpu_backend
    gpu_backend.require_library_loaded()
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import feltor_rhs_check
    l0 = launches()
    rec = feltor_rhs_check.run(list(dims) if dims else None, evaluations=3)
    assert rec["max_rel_diff"] <= 1e-10, rec
    assert launches() - l0 > 1000, "the device build did not reach libdgb200.so"
