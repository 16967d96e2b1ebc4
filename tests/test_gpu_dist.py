"""Slab (multi-GPU) path on ONE device: a communicator of size 1 exercises the padded operands, the ring-closing halo
copy, the slab tables of the fused kernel and the allreduce+scalar-kernel completion of the dots.  Results must be
bit-identical to the plain single-GPU objects.  (tools/dist_check.py runs the same comparison on N GPUs under torchrun.)"""
import numpy as np
import pytest
from util import same_bits, rng

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def G():
    import gpu_backend
    gpu_backend.require_library_loaded()
    return gpu_backend


@pytest.mark.parametrize("N,bcx,bcy,d", [([40, 24], 1, 0, 0), ([37, 19], 4, 0, 2), ([33, 40], 1, 1, 1), ([64, 16], 0, 0, 0),
                                        ([416, 420], 1, 0, 0), ([420, 416], 0, 0, 1)])  # the last two: large enough for the walker kernel
def test_slab_symv_equals_global(G, N, bcx, bcy, d):
    from feltor_b200 import topology as T
    from feltor_b200.elliptic import Elliptic2d
    from feltor_b200.dist import Comm, SlabElliptic2d
    g = T.Grid([0, 0], [np.pi, 2 * np.pi], 3, N, [bcx, bcy])
    r = rng(5)
    chi = 1. + r.uniform(0, 1, g.size)
    x = r.uniform(-1, 1, g.size)
    E = Elliptic2d(g, bcx, bcy, d, 0.7)
    E.set_chi(G.make(chi))
    y = G.make(np.zeros(g.size))
    E.symv(G.make(x), y)
    comm = Comm(0, 1)
    S = SlabElliptic2d(comm, g, bcx, bcy, d, 0.7)
    S.set_chi(G.make(S.local(chi)))
    ys = G.make(np.full(g.size, np.nan))
    S.symv(G.make(S.local(x)), ys)
    assert same_bits(G.get(ys), G.get(y))


def test_slab_pcg_equals_global(G):
    from feltor_b200 import topology as T
    from feltor_b200.elliptic import Elliptic2d, PCG
    from feltor_b200.dist import Comm, SlabElliptic2d, DistPCG
    g = T.Grid([0, 0], [np.pi, 2 * np.pi], 3, [40, 24], [T.DIR, T.PER])
    chi = g.evaluate(lambda x, y: 1. + 0.9 * np.sin(x) * np.sin(y))
    b = g.evaluate(lambda x, y: np.sin(x) * np.sin(y) * (1 + np.cos(3 * y)))
    E = Elliptic2d(g, T.DIR, T.PER, T.FORWARD, 1.0)
    E.set_chi(G.make(chi))
    x = G.make(np.zeros(g.size))
    it = PCG(g.size, g.size).solve(E, x, G.make(b), E.precond(), E.weights(), 1e-9, 1.0, 1)
    comm = Comm(0, 1)
    S = SlabElliptic2d(comm, g, T.DIR, T.PER, T.FORWARD, 1.0)
    S.set_chi(G.make(chi))
    xs = G.make(np.zeros(g.size))
    its = DistPCG(comm, g.size, g.size).solve(S, xs, G.make(b), S.precond(), S.weights(), 1e-9, 1.0, 1)
    assert its == it
    assert same_bits(G.get(xs), G.get(x))


def _toefl_params(N, model="global"):
    return {"grid": {"n": 3, "Nx": N, "Ny": N, "lx": 200, "ly": 200},
            "init": {"amplitude": 1.0, "sigma": 10, "posX": 0.3, "posY": 0.5, "flr": "gamma_inv"},
            "bc": ["DIR", "PER"],
            "elliptic": {"stages": 3, "eps_pol": [1e-6, 1, 1], "eps_gamma": [1e-7, 1, 1], "direction": "centered"},
            "model": {"type": model, "boussinesq": False, "curvature": 0.00015, "tau": 1, "nu": 1e-6}}


@pytest.mark.parametrize("N,model", [(32, "global"), (40, "local")])
def test_dist_toefl_equals_single_gpu(G, N, model):
    """toefl::Explicit + Bogacki-Shampine steps on the slab harness (feltor_b200/dist_toefl.py: row-sliced block matrices with
    remapped columns, ghost-row exchanges, slab Helmholtz / polarisation plans, distributed PCG inside the restated nested
    iteration) with a communicator of size 1 against the plain single-GPU harness: state, potentials and every per-stage
    iteration count bit for bit (tools/dist_check.py repeats it on N GPUs)"""
    import torch
    from feltor_b200 import toefl as TF
    from feltor_b200.dist import Comm
    from feltor_b200.dist_toefl import DistExplicit
    js = _toefl_params(N, model)
    results = []
    for dist in (False, True):
        ex = DistExplicit(Comm(0, 1), TF.Parameters(js)) if dist else TF.Explicit(TF.Parameters(js))
        u0 = ex.initial_condition()
        u1 = [torch.zeros_like(u0[0]), torch.zeros_like(u0[0])]
        delta = [torch.zeros_like(u0[0]), torch.zeros_like(u0[0])]
        erk = TF.ERKStep("Bogacki-Shampine-4-2-3", u0)
        t, nums = 0., []
        for _ in range(2):
            t = erk.step(ex, t, u0, u1, 0.5, delta)
            u0, u1 = u1, u0
            nums.append(dict(ex.numbers))
        results.append((G.get(u0[0]), G.get(u0[1]), G.get(ex.phi[0]), G.get(ex.phi[1]), nums))
    a, b = results
    assert a[4] == b[4], (a[4], b[4])
    for k in range(4):
        assert same_bits(a[k], b[k]), k


def test_dist_ds_centered_z_slab_equals_global(G):
    """DS::centered with the planes owned in blocks (z decomposition, ghost planes by dgb_comm_halo_rows) on a size-1
    communicator against the global cell-tiled call: bit for bit (tools/dist_check.py repeats it on N GPUs)"""
    import ctypes as C
    import torch
    from feltor_b200._lib import lib
    from feltor_b200._dev import ptr, stream
    from feltor_b200.dist import Comm
    from feltor_b200.dist_ds import DistDSCentered
    from test_gpu_ds import _fieldaligned_like_matrix
    n, Nx, Ny, Nz = 3, 40, 12, 10
    r = rng(4)
    P, M = _fieldaligned_like_matrix(r, n, Nx, Ny), _fieldaligned_like_matrix(r, n, Nx, Ny, 2)
    rows = n * n * Nx * Ny
    f, bphi, g0 = r.uniform(-1, 1, rows * Nz), r.uniform(0.5, 1.5, rows * Nz), r.uniform(-1, 1, rows * Nz)
    dP = [torch.from_numpy(a).cuda() for a in P]
    dM = [torch.from_numpy(a).cuda() for a in M]
    hp, hm = C.c_void_p(), C.c_void_p()
    lib().celltile_plan_create(C.byref(hp), n, Nx, Ny, ptr(dP[0]), ptr(dP[1]), ptr(dP[2]), stream())
    lib().celltile_plan_create(C.byref(hm), n, Nx, Ny, ptr(dM[0]), ptr(dM[1]), ptr(dM[2]), stream())
    D = DistDSCentered(Comm(0, 1), n, Nx, Ny, Nz, dP, dM, G.make(bphi), 0.1)
    df, dbphi = G.make(f), G.make(bphi)
    for alpha, beta in ((0.7, 0.), (-1.3, 0.5)):
        a, b = G.make(g0), G.make(g0)
        lib().celltile_ds_centered(hp, hm, Nz, C.c_double(alpha), ptr(df), ptr(dbphi), C.c_double(0.1), C.c_double(beta), ptr(a), stream())
        D.centered(alpha, df, beta, b)
        assert same_bits(G.get(a), G.get(b)), (alpha, beta)
    lib().celltile_plan_destroy(hp)
    lib().celltile_plan_destroy(hm)
