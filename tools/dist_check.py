#!/usr/bin/env python
"""N-GPU parity check (run under torchrun): the slab-decomposed Elliptic apply and PCG solve give, on every rank, exactly
the rows of the single-GPU result (bitwise) and the same iteration count; the global dot is bit-reproducible.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/dist_check.py"""
import os
import sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from feltor_b200 import topology as T, blas2  # noqa: E402
from feltor_b200.elliptic import Elliptic2d, PCG  # noqa: E402
from feltor_b200.dist import Comm, SlabElliptic2d, DistPCG  # noqa: E402
from feltor_b200._dev import dvec, hvec, ptr, stream  # noqa: E402
import feltor_b200 as fb  # noqa: E402
import ctypes as C  # noqa: E402

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
comm = Comm.from_torch_distributed()
rank, size = comm.rank, comm.size
ok = True
for N, bcx, bcy, d in (([40, 24], T.DIR, T.PER, T.FORWARD), ([37, 21], T.NEU, T.DIR, T.CENTERED), ([96, 64], T.DIR, T.PER, T.BACKWARD)):
    g = T.Grid([0, 0], [np.pi, 2 * np.pi], 3, N, [bcx, bcy])
    r = np.random.default_rng(3)
    chi = 1. + r.uniform(0, 1, g.size)
    x = r.uniform(-1, 1, g.size)
    b = g.evaluate(lambda xx, yy: np.sin(xx) * np.sin(yy) * (1 + np.cos(3 * yy)))
    # single-GPU reference on every rank
    E = Elliptic2d(g, bcx, bcy, d, 0.7)
    E.set_chi(dvec(chi))
    y = torch.zeros(g.size, dtype=torch.float64, device="cuda")
    E.symv(dvec(x), y)
    xs1 = torch.zeros(g.size, dtype=torch.float64, device="cuda")
    it1 = PCG(g.size, g.size).solve(E, xs1, dvec(b), E.precond(), E.weights(), 1e-9, 1.0, 1)
    # decomposed
    S = SlabElliptic2d(comm, g, bcx, bcy, d, 0.7)
    S.set_chi(dvec(S.local(chi)))
    ys = torch.full((S.size,), float("nan"), dtype=torch.float64, device="cuda")
    S.symv(dvec(S.local(x)), ys)
    same_symv = np.array_equal(hvec(ys).view(np.int64), S.local(hvec(y)).view(np.int64))
    xs = torch.zeros(S.size, dtype=torch.float64, device="cuda")
    its = DistPCG(comm, S.size, g.size).solve(S, xs, dvec(S.local(b)), S.precond(), S.weights(), 1e-9, 1.0, 1)
    same_pcg = its == it1 and np.array_equal(hvec(xs).view(np.int64), S.local(hvec(xs1)).view(np.int64))
    # global exact dot through the integer allreduce
    ws = blas2.DotWorkspace()
    rec = torch.zeros(41, dtype=torch.int64, device="cuda")
    xl, wl = dvec(S.local(x)), S.weights()
    fb.lib().exdot3(ws.h, S.size, ptr(xl), C.c_double(0), ptr(wl), C.c_double(0), ptr(xl), C.c_double(0), ptr(rec), stream())
    comm.allreduce_dot(rec)
    val = rec[39:40].view(torch.float64).item()
    same_dot = val == blas2.dot(dvec(x), E.weights(), dvec(x))
    print(f"rank {rank}/{size} N={N} bc=({bcx},{bcy}) dir={d}: symv {same_symv} pcg {same_pcg} (it {its} vs {it1}) dot {same_dot}", flush=True)
    ok = ok and same_symv and same_pcg and same_dot
# a grid large enough that every rank's slab takes the walker kernel (>= 400^2 cells per rank) and the peer-memory halo
# path (equal slabs): symv and a fixed number of PCG iterations, bitwise against the single-GPU run
N, bcx, bcy, d = [416, 416 * size], T.DIR, T.PER, T.FORWARD
g = T.Grid([0, 0], [np.pi, 2 * np.pi], 3, N, [bcx, bcy])
r = np.random.default_rng(11)
chi = 1. + r.uniform(0, 1, g.size)
x = r.uniform(-1, 1, g.size)
b = g.evaluate(lambda xx, yy: np.sin(xx) * np.sin(yy) * (1 + np.cos(3 * yy)))
E = Elliptic2d(g, bcx, bcy, d, 1.0)
E.set_chi(dvec(chi))
y = torch.zeros(g.size, dtype=torch.float64, device="cuda")
E.symv(dvec(x), y)
p1 = PCG(g.size, 41)
p1.set_throw_on_fail(False)
xs1 = torch.zeros(g.size, dtype=torch.float64, device="cuda")
it1 = p1.solve(E, xs1, dvec(b), E.precond(), E.weights(), 1e-30, 1.0, 1)
S = SlabElliptic2d(comm, g, bcx, bcy, d, 1.0)
S.set_chi(dvec(S.local(chi)))
ys = torch.full((S.size,), float("nan"), dtype=torch.float64, device="cuda")
S.symv(dvec(S.local(x)), ys)
same_symv = np.array_equal(hvec(ys).view(np.int64), S.local(hvec(y)).view(np.int64))
p2 = DistPCG(comm, S.size, 41)
p2.throw_on_fail = False
xs = torch.zeros(S.size, dtype=torch.float64, device="cuda")
its = p2.solve(S, xs, dvec(S.local(b)), S.precond(), S.weights(), 1e-30, 1.0, 1)
same_pcg = its == it1 and np.array_equal(hvec(xs).view(np.int64), S.local(hvec(xs1)).view(np.int64))
print(f"rank {rank}/{size} N={N} (walker slabs): symv {same_symv} pcg-40-iterations {same_pcg}", flush=True)
ok = ok and same_symv and same_pcg
# toefl (config 3) on slabs: two Bogacki-Shampine steps of the right-hand side on N ranks against the single-GPU harness on every
# rank: state rows, potentials and every per-stage PCG iteration count bit for bit
from feltor_b200 import toefl as TF  # noqa: E402
from feltor_b200.dist_toefl import DistExplicit  # noqa: E402
Nt = 32 * size if size <= 4 else 16 * size
js = {"grid": {"n": 3, "Nx": 48, "Ny": Nt, "lx": 200, "ly": 200}, "init": {"amplitude": 1.0, "sigma": 10, "posX": 0.3, "posY": 0.5, "flr": "gamma_inv"},
      "bc": ["DIR", "PER"], "elliptic": {"stages": 3, "eps_pol": [1e-6, 1, 1], "eps_gamma": [1e-7, 1, 1], "direction": "centered"},
      "model": {"type": "global", "boussinesq": False, "curvature": 0.00015, "tau": 1, "nu": 1e-6}}
res = []
for exd in (TF.Explicit(TF.Parameters(js)), DistExplicit(comm, TF.Parameters(js))):
    u0 = exd.initial_condition()
    u1 = [torch.zeros_like(u0[0]), torch.zeros_like(u0[0])]
    delta = [torch.zeros_like(u0[0]), torch.zeros_like(u0[0])]
    erk = TF.ERKStep("Bogacki-Shampine-4-2-3", u0)
    tt, nums = 0., []
    for _ in range(2):
        tt = erk.step(exd, tt, u0, u1, 0.5, delta)
        u0, u1 = u1, u0
        nums.append(dict(exd.numbers))
    res.append((hvec(u0[0]), hvec(u0[1]), hvec(exd.phi[0]), nums, exd))
slab = res[1][4].slab
same_toefl = res[0][3] == res[1][3] and all(np.array_equal(slab.local(res[0][k]).view(np.int64), res[1][k].view(np.int64)) for k in range(3))
print(f"rank {rank}/{size} toefl 48x{Nt} on slabs: {same_toefl} (iterations {res[1][3][-1]})", flush=True)
ok = ok and same_toefl
# DS::centered on a z-decomposition (ghost planes by the halo exchange): every rank's planes against the global call
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_ds import _fieldaligned_like_matrix  # noqa: E402
from feltor_b200.dist_ds import DistDSCentered  # noqa: E402
nd, Nxd, Nyd, Nzd = 3, 40, 12, 4 * size
rd = np.random.default_rng(4)
Pm, Mm = _fieldaligned_like_matrix(rd, nd, Nxd, Nyd), _fieldaligned_like_matrix(rd, nd, Nxd, Nyd, 2)
rowsd = nd * nd * Nxd * Nyd
fd, bd, gd0 = rd.uniform(-1, 1, rowsd * Nzd), rd.uniform(0.5, 1.5, rowsd * Nzd), rd.uniform(-1, 1, rowsd * Nzd)
dPm = [torch.from_numpy(a).cuda() for a in Pm]
dMm = [torch.from_numpy(a).cuda() for a in Mm]
hp, hm = C.c_void_p(), C.c_void_p()
fb.lib().celltile_plan_create(C.byref(hp), nd, Nxd, Nyd, ptr(dPm[0]), ptr(dPm[1]), ptr(dPm[2]), stream())
fb.lib().celltile_plan_create(C.byref(hm), nd, Nxd, Nyd, ptr(dMm[0]), ptr(dMm[1]), ptr(dMm[2]), stream())
ga, dfd, dbd = dvec(gd0), dvec(fd), dvec(bd)
fb.lib().celltile_ds_centered(hp, hm, Nzd, C.c_double(-1.3), ptr(dfd), ptr(dbd), C.c_double(0.1), C.c_double(0.5), ptr(ga), stream())
DD = DistDSCentered(comm, nd, Nxd, Nyd, Nzd, dPm, dMm, dvec(bd[(Nzd // size) * rank * rowsd:(Nzd // size) * (rank + 1) * rowsd]), 0.1)
gb = dvec(DD.local(gd0).copy())
DD.centered(-1.3, dvec(DD.local(fd).copy()), 0.5, gb)
same_ds = np.array_equal(hvec(gb).view(np.int64), DD.local(hvec(ga)).view(np.int64))
print(f"rank {rank}/{size} DS::centered on z-slabs ({Nzd} planes): {same_ds}", flush=True)
ok = ok and same_ds
# row-distributed CSR product (dg::MPIDistMat, feltor_b200/dist_csr.py) on N ranks: every rank's rows against the global product
from test_dist_cpu import random_csr  # noqa: E402
from feltor_b200.dist import partition  # noqa: E402
from feltor_b200.dist_csr import DistCsr, contiguous_owner  # noqa: E402
for band in (None, 60):
    rc = np.random.default_rng(21)
    nr_, nc_ = 20000, 18000
    cpos, cidx, cval = random_csr(rc, nr_, nc_, 40, band)
    cx = rc.uniform(-1, 1, nc_)
    cy = torch.full((nr_,), float("nan"), dtype=torch.float64, device="cuda")
    keep = [dvec(cpos), dvec(cidx), dvec(cval), dvec(cx)]
    fb.lib().csr_spmv(nr_, nc_, *[ptr(a) for a in keep[:3]], C.c_double(1.), ptr(keep[3]), C.c_double(0.), ptr(cy), stream())
    rp, cp = partition(nr_, size), partition(nc_, size)
    r0, nrl = rp[rank]
    M = DistCsr(comm, cpos[r0:r0 + nrl + 1] - cpos[r0], cidx[cpos[r0]:cpos[r0 + nrl]], cval[cpos[r0]:cpos[r0 + nrl]], contiguous_owner(cp), cp[rank][1])
    yl = torch.full((nrl,), float("nan"), dtype=torch.float64, device="cuda")
    xl = dvec(cx[cp[rank][0]:cp[rank][0] + cp[rank][1]].copy())
    for _ in range(3):
        M.symv(xl, yl)
    same_csr = np.array_equal(hvec(yl).view(np.int64), hvec(cy)[r0:r0 + nrl].view(np.int64))
    print(f"rank {rank}/{size} DistCsr band={band}: {same_csr} (outer rows {M.plan.scatter.size}, buffer {M.plan.buffer_size}, sends {M.plan.send_idx.size})", flush=True)
    ok = ok and same_csr
# allreduce mode of MPIDistMat (dg::Average over a distributed axis): column blocks, rank-ordered sum, same bits on every rank
from feltor_b200.dist_csr import DistCsrAllreduce  # noqa: E402
nxa, nya = 120, 16 * size
ra = np.random.default_rng(31)
wa, xa = ra.uniform(0.5, 1.5, nya), ra.uniform(-1, 1, nxa * nya)
oa, ca = partition(nya, size)[rank]
lpa = (np.arange(nxa + 1) * ca).astype(np.int32)
lia = (np.arange(ca)[None, :] * nxa + np.arange(nxa)[:, None]).reshape(-1).astype(np.int32)
Ma = DistCsrAllreduce(comm, nxa, ca * nxa, lpa, lia, np.tile(wa[oa:oa + ca], nxa))
ya = torch.full((nxa,), float("nan"), dtype=torch.float64, device="cuda")
Ma.symv(dvec(xa[oa * nxa:(oa + ca) * nxa].copy()), ya)
exact = (xa.reshape(nya, nxa) * wa[:, None]).sum(axis=0)
everyone = [torch.empty_like(ya) for _ in range(size)]
dist.all_gather(everyone, ya)
same_avg = all(torch.equal(everyone[0].view(torch.int64), e.view(torch.int64)) for e in everyone) and \
    float(np.max(np.abs(hvec(ya) - exact))) <= 1e-13 * float(np.max(np.abs(exact)))
print(f"rank {rank}/{size} DistCsrAllreduce (average over the distributed axis): {same_avg}", flush=True)
ok = ok and same_avg
# block matrices distributed along their own axis (dg::MPISparseBlockMat, feltor_b200/dist_ell.py): an x-decomposition of dx and
# a y-decomposition of the jump matrix, periodic across the ranks, against the global product
from feltor_b200.dist_ell import DistEll  # noqa: E402
for coord, what in ((0, "dx centered"), (1, "jump y")):
    ge = T.Grid([0, 0], [1., 2.], 3, [16 * size + 3, 12 * size + 1], [T.PER, T.PER])
    me = T.derivative(0, ge, T.PER, T.CENTERED) if coord == 0 else T.jump(1, ge, T.PER)
    re_ = np.random.default_rng(8)
    xe, ye0 = re_.uniform(-1, 1, ge.size), re_.uniform(-1, 1, ge.size)
    nxe, nye = 3 * ge.N[0], 3 * ge.N[1]
    ywant = dvec(ye0)
    me.symv(-0.6, dvec(xe), 0.3, ywant)
    parte = partition(ge.N[coord], size)
    o_, c_ = parte[rank]
    if coord == 0:
        cut = lambda v: np.ascontiguousarray(v.reshape(nye, nxe)[:, o_ * 3:(o_ + c_) * 3]).reshape(-1)
        De = DistEll(comm, me, o_, c_, parte, nye, 1)
    else:
        cut = lambda v: np.ascontiguousarray(v.reshape(nye, nxe)[o_ * 3:(o_ + c_) * 3]).reshape(-1)
        De = DistEll(comm, me, o_, c_, parte, 1, nxe)
    yl, xl = dvec(cut(ye0)), dvec(cut(xe))
    De.symv(-0.6, xl, 0.3, yl)
    same_ell = np.array_equal(hvec(yl).view(np.int64), cut(hvec(ywant)).view(np.int64))
    print(f"rank {rank}/{size} DistEll {what}: {same_ell} (outer entries {De.plan.coo_rows.size}, chunks {De.plan.num_chunks})", flush=True)
    ok = ok and same_ell
t = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("DIST_CHECK", "PASS" if t.item() == 1 else "FAIL", flush=True)
dist.destroy_process_group()
sys.exit(0 if t.item() == 1 else 1)
