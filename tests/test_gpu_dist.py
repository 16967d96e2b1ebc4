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


@pytest.mark.parametrize("size,band,nrows,ncols", [(1, None, 500, 640), (3, None, 3001, 2500), (4, 40, 5000, 5000), (2, 5, 64, 64)])
def test_dist_csr_equals_global(G, size, band, nrows, ncols):
    """dg::MPIDistMat::symv (feltor_b200/dist_csr.py: pack, exchange, inner product, fused outer product + scatter) on `size`
    ranks emulated in this process -- the kernels are the ones N processes run, the exchange between the emulated ranks is a
    copy -- equals dgb_csr_spmv of the global matrix bit for bit; with one rank the exchange is the library's dgb_comm_gather"""
    import ctypes as C
    import torch
    from feltor_b200 import blas2
    from feltor_b200._lib import lib
    from feltor_b200._dev import dvec, ptr, stream
    from feltor_b200.dist import Comm, partition
    from feltor_b200.dist_csr import DistCsr, DistCsrPlan, contiguous_owner
    from test_dist_cpu import random_csr
    r = rng(size + nrows)
    pos, idx, val = random_csr(r, nrows, ncols, 30, band)
    x = r.uniform(-1, 1, ncols)
    want = G.make(np.full(nrows, np.nan))
    keep = [dvec(pos), dvec(idx), dvec(val), G.make(x)]   # alive until the call is enqueued
    lib().csr_spmv(nrows, ncols, *[ptr(a) for a in keep[:3]], C.c_double(1.), ptr(keep[3]), C.c_double(0.), ptr(want), stream())
    want = G.get(want)
    row_part, col_part = partition(nrows, size), partition(ncols, size)
    g2l = contiguous_owner(col_part)
    local = lambda rank: (pos[row_part[rank][0]:row_part[rank][0] + row_part[rank][1] + 1] - pos[row_part[rank][0]],
                          idx[pos[row_part[rank][0]]:pos[row_part[rank][0] + row_part[rank][1]]],
                          val[pos[row_part[rank][0]]:pos[row_part[rank][0] + row_part[rank][1]]])
    plans = [DistCsrPlan(rank, size, *local(rank), g2l, col_part[rank][1]) for rank in range(size)]
    comm = Comm(0, 1)
    mats = []
    for rank in range(size):
        m = DistCsr.__new__(DistCsr)
        comm_r = comm if size == 1 else type("EmulatedRank", (), {"rank": rank, "size": size, "h": None})()
        DistCsr.__init__(m, comm_r, *local(rank), g2l, col_part[rank][1], asked=[plans[p].requests[rank] for p in range(size)])
        mats.append(m)
    xs = [G.make(x[o:o + c]) for o, c in col_part]
    ys = [G.make(np.full(c, np.nan)) for _, c in row_part]
    if size == 1:
        mats[0].symv(xs[0], ys[0])
        assert mats[0].plan.buffer_size == 0
    else:
        for rank in range(size):
            mats[rank].pack(xs[rank])
        for rank in range(size):                        # the exchange: segment of p's send buffer meant for `rank` -> rank's buffer
            ro = 0
            for p in range(size):
                so, cnt = int(mats[p].plan.send_counts[:rank].sum()), int(mats[p].plan.send_counts[rank])
                assert cnt == int(mats[rank].plan.recv_counts[p])
                mats[rank].recv_buf[ro:ro + cnt].copy_(mats[p].send_buf[so:so + cnt])
                ro += cnt
        for rank in range(size):
            mats[rank].apply_inner(xs[rank], ys[rank])
            mats[rank].apply_outer(ys[rank])
        assert sum(m.plan.scatter.size for m in mats) > 0
    got = np.concatenate([G.get(y) for y in ys])
    assert same_bits(got, want)


def test_dist_csr_self_exchange(G):
    """a size-1 communicator whose column map sends every second element through the gather buffer: exercises dgb_comm_gather's
    self message, the side stream and the fused outer kernel inside DistCsr.symv itself"""
    import ctypes as C
    from feltor_b200._lib import lib
    from feltor_b200._dev import dvec, ptr, stream
    from feltor_b200.dist import Comm
    from feltor_b200.dist_csr import DistCsr
    from test_dist_cpu import random_csr
    r = rng(91)
    nrows = ncols = 4000
    pos, idx, val = random_csr(r, nrows, ncols, 25, 30)
    x = r.uniform(-1, 1, ncols)
    want = G.make(np.full(nrows, np.nan))
    keep = [dvec(pos), dvec(idx), dvec(val), G.make(x)]   # alive until the call is enqueued
    lib().csr_spmv(nrows, ncols, *[ptr(a) for a in keep[:3]], C.c_double(1.), ptr(keep[3]), C.c_double(0.), ptr(want), stream())

    comm = Comm(0, 1)
    g2l_true = lambda gi: (np.zeros(len(gi), dtype=np.int64), np.asarray(gi, dtype=np.int64))
    from feltor_b200.dist_csr import DistCsrPlan as plan_cls
    # build the matrix normally (everything inner), then move every row that has an odd column into the outer matrix by hand and
    # let the rank ask itself for those columns
    p = plan_cls(0, 1, pos, idx, val, g2l_true, ncols)
    assert p.buffer_size == 0
    counts = np.diff(pos)
    row_of = np.repeat(np.arange(nrows), counts)
    outer_row = np.zeros(nrows, dtype=bool)
    outer_row[row_of[idx % 2 == 1]] = True
    eo = outer_row[row_of]
    m = DistCsr(comm, pos, idx, val, g2l_true, ncols)
    uniq, inv = np.unique(idx[eo], return_inverse=True)
    m.plan.inner_pos = np.concatenate([[0], np.cumsum(np.where(outer_row, 0, counts))]).astype(np.int32)
    m.plan.scatter = np.nonzero(outer_row)[0].astype(np.int32)
    m.plan.buffer_size = int(uniq.size)
    m.plan.send_idx = uniq.astype(np.int32)
    m.inner = tuple(dvec(a) for a in (m.plan.inner_pos, idx[~eo].astype(np.int32), val[~eo]))
    m.outer = tuple(dvec(a) for a in (np.concatenate([[0], np.cumsum(counts[m.plan.scatter])]).astype(np.int32), inv.astype(np.int32), val[eo]))
    m.scatter, m.send_idx = dvec(m.plan.scatter), dvec(m.plan.send_idx)
    import torch
    m.send_buf = torch.empty(uniq.size, dtype=torch.float64, device="cuda")
    m.recv_buf = torch.full((uniq.size,), float("nan"), dtype=torch.float64, device="cuda")
    m._sc = (C.c_int * 1)(int(uniq.size))
    m._rc = (C.c_int * 1)(int(uniq.size))
    y = G.make(np.full(nrows, np.nan))
    for _ in range(3):
        m.symv(G.make(x), y)
    assert same_bits(G.get(y), G.get(want))


def _emulated_dist_ell(G, m, size, row_part, col_part, left, right, alpha, beta, xs, ys):
    """run DistEll.symv step by step for `size` ranks in this process (the exchange between them is a copy)"""
    from feltor_b200.dist_ell import DistEll, DistEllPlan
    plans = [DistEllPlan(r, size, m, row_part[r][0], row_part[r][1], col_part, left, right) for r in range(size)]
    mats = []
    for r in range(size):
        comm_r = type("EmulatedRank", (), {"rank": r, "size": size, "h": None})()
        mats.append(DistEll(comm_r, m, row_part[r][0], row_part[r][1], col_part, left, right, asked=[plans[p].requests[r] for p in range(size)]))
    for r in range(size):
        mats[r].pack(xs[r])
    for r in range(size):
        ro = 0
        for p in range(size):
            ch = mats[p].plan.chunk
            so, cnt = int(mats[p].plan.send_blocks[:r].sum()) * ch, int(mats[p].plan.send_blocks[r]) * ch
            assert cnt == int(mats[r].plan.recv_blocks[p]) * ch
            mats[r].recv_buf[ro:ro + cnt].copy_(mats[p].send_buf[so:so + cnt])
            ro += cnt
    for r in range(size):
        mats[r].apply_inner(alpha, xs[r], beta, ys[r])
        mats[r].apply_outer(alpha, ys[r])
    return mats


@pytest.mark.parametrize("coord,size,bc,direction,n,N", [(0, 3, 1, 0, 3, [24, 10]), (0, 4, 0, 2, 3, [37, 6]), (1, 3, 0, 1, 3, [9, 31]), (1, 2, 4, 2, 2, [8, 12]),
                                                         (0, 5, 2, "jump", 4, [25, 3]), (1, 8, 0, "jump", 3, [5, 16])])
def test_dist_ell_equals_global(G, coord, size, bc, direction, n, N):
    """dg::MPISparseBlockMat::symv (feltor_b200/dist_ell.py: inner Ell + outer Coo from make_mpi_sparseblockmat's row split, packed
    block columns, exchange, mpi_matrix.h:183-330) for derivatives / jumps distributed along THEIR OWN axis -- x- and y-
    decompositions, periodic wrap across the first and last rank, uneven partitions -- equals dgb_ell_symv of the global matrix
    on the global vector bit for bit"""
    from feltor_b200 import topology as T
    from feltor_b200.dist import partition
    bcs = [bc, bc]
    g = T.Grid([0, 0], [1., 2.], n, N, bcs)
    m = T.jump(coord, g, bc) if direction == "jump" else T.derivative(coord, g, bc, direction)
    r = rng(coord * 10 + size)
    x, y0 = r.uniform(-1, 1, g.size), r.uniform(-1, 1, g.size)
    nx, ny = n * N[0], n * N[1]
    part = partition(N[coord], size)
    if coord == 0:
        left, right = ny, 1
        cut = lambda v, o, c: np.ascontiguousarray(v.reshape(ny, nx)[:, o * n:(o + c) * n]).reshape(-1)
    else:
        left, right = 1, nx
        cut = lambda v, o, c: np.ascontiguousarray(v.reshape(ny, nx)[o * n:(o + c) * n]).reshape(-1)
    for alpha, beta in ((1., 0.), (-0.7, 0.4)):
        want = G.make(y0)
        m.symv(alpha, G.make(x), beta, want)
        want = G.get(want)
        xs = [G.make(cut(x, o, c)) for o, c in part]
        ys = [G.make(cut(y0, o, c) if beta != 0. else np.full(c * n * (ny if coord == 0 else nx), np.nan)) for o, c in part]
        mats = _emulated_dist_ell(G, m, size, part, part, left, right, alpha, beta, xs, ys)
        for (o, c), yl in zip(part, ys):
            assert same_bits(G.get(yl), cut(want, o, c)), (alpha, beta, o)
        assert sum(mt.plan.coo_rows.size for mt in mats) > 0


def test_dist_ell_projection(G):
    """a rectangular block matrix (fast projection: half as many block rows as columns) with different row and column
    distributions"""
    from feltor_b200 import topology as T
    from feltor_b200.dist import partition
    n, N = 3, [12, 16]
    g = T.Grid([0, 0], [1., 2.], n, N, [1, 0])
    m = T.fast_projection(1, g, 1, 2)           # along y: 8 block rows, 16 block columns, right = n Nx
    r = rng(2)
    x = r.uniform(-1, 1, g.size)
    nx = n * N[0]
    want = G.make(np.full(m.total_rows, np.nan))
    m.symv(1., G.make(x), 0., want)
    want = G.get(want)
    size = 3
    row_part, col_part = partition(N[1] // 2, size), [(0, 5), (5, 4), (9, 7)]      # rows and columns cut at unrelated places
    xs = [G.make(x.reshape(-1, nx)[o * n:(o + c) * n].reshape(-1).copy()) for o, c in col_part]
    ys = [G.make(np.full(c * n * nx, np.nan)) for _, c in row_part]
    _emulated_dist_ell(G, m, size, row_part, col_part, 1, nx, 1., 0., xs, ys)
    for (o, c), yl in zip(row_part, ys):
        assert same_bits(G.get(yl), want.reshape(-1, nx)[o * n:(o + c) * n].reshape(-1))


@pytest.mark.timeout(600)
def test_dist_check_on_all_visible_gpus(G):
    """tools/dist_check.py under torch.distributed.run on every visible GPU (needs >= 2): slab Elliptic / PCG / exact dot / toefl /
    DS on z-slabs / distributed CSR and block matrices, each rank's part bitwise equal to the single-GPU result, with the real
    NCCL / peer-memory exchanges instead of the size-1 communicator and the emulated ranks of the tests above"""
    import os
    import subprocess
    import sys
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("one visible GPU: the N > 1 exchanges are covered by tools/dist_check.py runs (profiles/dist_check_r02_n*.log)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    n = 2 if ngpu < 4 else 4
    port = 29600 + os.getpid() % 300
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(n), "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(root, "tools", "dist_check.py")], capture_output=True, text=True, timeout=580)
    assert p.returncode == 0 and "DIST_CHECK PASS" in p.stdout, p.stdout[-3000:] + p.stderr[-2000:]


@pytest.mark.parametrize("size", [1, 3, 8])
def test_dist_csr_allreduce_mode(G, size):
    """MPIDistMat in allreduce mode (dg::Average over a distributed axis, feltor_b200/dist_csr.py::DistCsrAllreduce): column blocks
    of an averaging matrix on `size` emulated ranks, partial results summed in rank order by dgb_sum_ranks -- bitwise the
    rank-ordered sum of the single-rank partial products, every rank the same bits, 1e-14 from the undistributed product"""
    import ctypes as C
    from feltor_b200._lib import lib
    from feltor_b200._dev import dvec, ptr, stream
    from feltor_b200.dist import partition
    from feltor_b200.dist_csr import DistCsrAllreduce
    r = rng(size)
    nx, ny = 90, 64                                   # average over y: rows = nx, every row sums ny weighted entries
    w = r.uniform(0.5, 1.5, ny)
    pos = (np.arange(nx + 1) * ny).astype(np.int32)
    idx = (np.arange(ny)[None, :] * nx + np.arange(nx)[:, None]).reshape(-1).astype(np.int32)
    val = np.tile(w, nx)
    x = r.uniform(-1, 1, nx * ny)
    keep = [dvec(pos), dvec(idx), dvec(val), G.make(x)]
    full = G.make(np.full(nx, np.nan))
    lib().csr_spmv(nx, nx * ny, *[ptr(a) for a in keep[:3]], C.c_double(1.), ptr(keep[3]), C.c_double(0.), ptr(full), stream())
    part = partition(ny, size)                        # ranks own blocks of y rows = contiguous pieces of the vector
    mats, partials = [], []
    for rank, (o, c) in enumerate(part):
        lp = (np.arange(nx + 1) * c).astype(np.int32)
        li = (np.arange(c)[None, :] * nx + np.arange(nx)[:, None]).reshape(-1).astype(np.int32)
        lv = np.tile(w[o:o + c], nx)
        comm_r = type("EmulatedRank", (), {"rank": rank, "size": size, "h": None})()
        m = DistCsrAllreduce(comm_r, nx, c * nx, lp, li, lv)
        m.apply_local(G.make(x[o * nx:(o + c) * nx]))
        mats.append(m)
        partials.append(G.get(m.partial)[:nx])
    want = partials[0].copy()
    for p_ in partials[1:]:
        want = want + p_                               # rank order, one rounding per addition
    for m in mats:
        for rank in range(size):
            m.parts[rank * nx:(rank + 1) * nx].copy_(mats[rank].partial[:nx])       # the exchange, emulated
        y = G.make(np.full(nx, np.nan))
        m.reduce(y)
        assert same_bits(G.get(y), want)
    assert np.max(np.abs(want - G.get(full))) <= 1e-14 * np.max(np.abs(G.get(full))) * ny
    if size == 1:
        assert same_bits(want, G.get(full))
