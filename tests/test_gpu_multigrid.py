"""GPU parity of NestedGrids::project and MultigridCG2d::solve against fixtures produced by the unmodified reference
(tests/golden/make_golden.py: 16 x 16 cells, n=3, 3 stages, DIR x PER) and against the live reference when present."""
import numpy as np
import pytest
from util import same_bits

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import gpu_backend
    gpu_backend.require_library_loaded()
    return gpu_backend


def build(G, T, g, stages, chi, direction=0, jfactor=1.0):
    from feltor_b200.elliptic import Elliptic2d, MultigridCG2d
    mg = MultigridCG2d(g, stages)
    ops = [Elliptic2d(mg.grid(u), g.bc[0], g.bc[1], direction, jfactor) for u in range(stages)]
    proj = mg.project(G.make(chi))
    for u in range(stages):
        ops[u].set_chi(proj[u])
    return mg, ops, proj


def test_multigrid_fixture(G, golden):
    from feltor_b200 import topology as T
    g = T.Grid([0, 0], [np.pi, 2 * np.pi], 3, [16, 16], [T.DIR, T.PER])
    mg, ops, proj = build(G, T, g, 3, golden["multigrid/chi"])
    assert [gr.N for gr in mg.grids] == [[16, 16], [8, 8], [4, 4]]
    for u in range(3):
        assert same_bits(G.get(proj[u]), golden[f"multigrid/project{u}"]), u
    x = G.make(np.zeros(g.size))
    num = mg.solve(ops, x, G.make(golden["multigrid/b"]), [1e-6, 1e-6 * 1.5, 1e-6 * 1.5 * 1.5])
    assert num == [int(v) for v in golden["multigrid/num"]]
    assert same_bits(G.get(x), golden["multigrid/x"])


@pytest.mark.parametrize("N,stages,bcx,bcy", [([32, 48], 3, 1, 0), ([64, 64], 4, 4, 1), ([40, 24], 2, 0, 0)])
def test_multigrid_vs_live_reference(G, ref, N, stages, bcx, bcy):
    """iteration numbers per stage and the solution are identical to the reference's MultigridCG2d"""
    if ref is None:
        pytest.skip("oracle/_ref/libdgref.so not built")
    if bcx == 0 and bcy == 0:
        pytest.skip("pure periodic Poisson problem is singular")
    from feltor_b200 import topology as T
    rg = ref.grid([0, 0], [np.pi, 2 * np.pi], 3, N, [bcx, bcy])
    g = T.Grid([0, 0], [np.pi, 2 * np.pi], 3, N, [bcx, bcy])
    chi, b = ref.evaluate(rg, "pol"), ref.evaluate(rg, "rhs")
    M = ref.Multigrid(rg, stages)
    M.set_chi(chi)
    xr = np.zeros(g.size)
    eps = [1e-7 * 1.5 ** u for u in range(stages)]
    st, numr, _ = M.solve(xr, b, eps)
    assert st == 0
    mg, ops, proj = build(G, T, g, stages, chi)
    pr = M.project(chi)
    for u in range(stages):
        assert same_bits(G.get(proj[u]), pr[u])
    x = G.make(np.zeros(g.size))
    num = mg.solve(ops, x, G.make(b), eps)
    assert num == numr
    assert same_bits(G.get(x), xr)


def test_multigrid_full_size(G):
    """config 3 size (n=3, 1024^2, 3 stages): the nested-iteration solve converges and needs far fewer fine-grid
    iterations than plain PCG (~16 000); the result satisfies the stopping criterion"""
    import torch
    from feltor_b200 import topology as T, blas1, blas2
    g = T.Grid([0, 0], [np.pi, 2 * np.pi], 3, [1024, 1024], [T.DIR, T.PER])
    chi = g.evaluate(lambda x, y: 1. + 0.9 * np.sin(x) * np.sin(y))
    mg, ops, _ = build(G, T, g, 3, chi)
    amp = 0.9
    b = G.make(g.evaluate(lambda x, y: 2. * np.sin(x) * np.sin(y) * (amp * np.sin(x) * np.sin(y) + 1)
                          - amp * np.sin(x) ** 2 * np.cos(y) ** 2 - amp * np.cos(x) ** 2 * np.sin(y) ** 2))
    x = torch.zeros(g.size, dtype=torch.float64, device="cuda")
    num = mg.solve(ops, x, b, [1e-8, 1.5e-8, 2.25e-8])
    assert 0 < num[0] < 16000
    r = torch.empty_like(x)
    ops[0].symv(x, r)
    blas1.axpby(1., b, -1., r)
    res = np.sqrt(blas2.dot(r, ops[0].weights(), r))
    nrmb = np.sqrt(blas2.dot(b, ops[0].weights(), b))
    assert res < 1e-8 * (nrmb + 1.0) * 5


@pytest.mark.parametrize("n,N,stages", [(3, [16, 24], 3), (2, [20, 12], 3), (4, [8, 12], 2), (3, [2, 2], 2), (3, [256, 192], 3)])
def test_multimatrix_fused(G, n, N, stages):
    """the one-pass factor-2 projection / interpolation kernels == dg::MultiMatrix::symv as the reference runs it (x-matrix into a
    temporary, y-matrix from there, fast_interpolation.h:71-84) bit for bit, for alpha / beta variants"""
    from feltor_b200 import topology as T
    from feltor_b200.elliptic import MultigridCG2d
    from feltor_b200._lib import lib
    g = T.Grid([0, 0], [1., 2.], n, N, [T.DIR, T.PER])
    mg = MultigridCG2d(g, stages)
    r = np.random.default_rng(n + N[0])
    src = r.uniform(-1, 1, g.size)
    try:
        lib().multigrid2d_set_two_pass(1)
        two = [G.get(v) for v in mg.project(G.make(src))]
        lib().multigrid2d_set_two_pass(0)
        one = [G.get(v) for v in mg.project(G.make(src))]
        for u in range(stages):
            assert same_bits(one[u], two[u]), u
        for u in range(1, stages):
            xc, xf0 = r.uniform(-1, 1, mg.sizes[u]), r.uniform(-1, 1, mg.sizes[u - 1])
            for alpha, beta in ((1., 1.), (1., 0.), (-0.6, 0.3)):
                out = []
                for two_pass in (1, 0):
                    lib().multigrid2d_set_two_pass(two_pass)
                    xf = G.make(xf0 if beta != 0. else np.full(mg.sizes[u - 1], np.nan))
                    mg.interpolate(u, alpha, G.make(xc), beta, xf)
                    out.append(G.get(xf))
                assert same_bits(out[0], out[1]), (u, alpha, beta)
    finally:
        lib().multigrid2d_set_two_pass(0)
