"""CPU (gloo, world_size 2): host-side logic of the N > 1 path -- the slab partition, the reproducible global dot
(normalised superaccumulators summed as integers, exblas/mpi_accumulate.h:94-125) and the halo-exchange posting order
used by feltor_b200/csrc/comm.cu (which must also be right when lower == upper neighbour, i.e. two ranks in a ring)."""
import os
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import orc
    from feltor_b200.dist import partition
    from util import wide
    ok = True
    # ---- reproducible global dot
    r = np.random.default_rng(11)
    n_rows, row_len = 37, 50
    x = wide(r, n_rows * row_len, -300, 300)
    w = wide(r, n_rows * row_len, -20, 20)
    off, rows = partition(n_rows, world)[rank]
    sl = slice(off * row_len, (off + rows) * row_len)
    acc, st = orc.exdot3(np.ascontiguousarray(x[sl]), np.ascontiguousarray(w[sl]), np.ascontiguousarray(x[sl]))
    t = torch.from_numpy(acc.copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)          # integer sum of normalised words: associative
    total = orc.normalize(t.numpy())
    full, _ = orc.exdot3(x, w, x)
    ok = ok and np.array_equal(total, full) and orc.round_acc(total) == orc.round_acc(full)
    # ---- halo exchange in a periodic ring of slabs, same posting order as comm_halo_rows()
    ghost = 2
    glob = np.arange(n_rows * row_len, dtype=np.float64).reshape(n_rows, row_len)
    pad = np.full((rows + 2 * ghost, row_len), -1.0)
    pad[ghost:ghost + rows] = glob[off:off + rows]
    lower, upper = (rank - 1) % world, (rank + 1) % world
    bottom = torch.from_numpy(pad[ghost:2 * ghost].copy())
    top = torch.from_numpy(pad[rows:rows + ghost].copy())
    up_ghost, lo_ghost = torch.empty(ghost, row_len, dtype=torch.float64), torch.empty(ghost, row_len, dtype=torch.float64)
    ops = [dist.P2POp(dist.isend, bottom, lower), dist.P2POp(dist.isend, top, upper),
           dist.P2POp(dist.irecv, up_ghost, upper), dist.P2POp(dist.irecv, lo_ghost, lower)]
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    want_up = glob[[(off + rows + k) % n_rows for k in range(ghost)]]
    want_lo = glob[[(off - ghost + k) % n_rows for k in range(ghost)]]
    ok = ok and np.array_equal(up_ghost.numpy(), want_up) and np.array_equal(lo_ghost.numpy(), want_lo)
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


def test_partition():
    import sys
    sys.path.insert(0, ROOT)
    from feltor_b200.dist import partition
    for n, s in ((1024, 8), (10, 3), (7, 7), (5, 1)):
        p = partition(n, s)
        assert p[0][0] == 0 and sum(r for _, r in p) == n
        assert all(p[k][0] + p[k][1] == p[k + 1][0] for k in range(s - 1))
        assert max(r for _, r in p) - min(r for _, r in p) <= 1


@pytest.mark.timeout(120)
def test_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in range(2)]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(0, True), (1, True)]


# ---------------------------------------------------------------------------------------------------------------------------
# dg::make_mpi_matrix / dg::MPIGather host logic of feltor_b200/dist_csr.py (mpi_projection.h:50-124, mpi_gather.h:476-520)
def random_csr(r, nrows, ncols, max_per_row, band=None):
    """unsorted columns with duplicates, empty rows -- everything a CSR stencil / interpolation matrix may contain"""
    counts = r.integers(0, max_per_row + 1, nrows)
    pos = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
    if band is None:
        idx = r.integers(0, ncols, pos[-1])
    else:
        centre = np.repeat((np.arange(nrows) * ncols) // max(nrows, 1), counts)
        idx = (centre + r.integers(-band, band + 1, pos[-1])) % ncols
    return pos, idx.astype(np.int32), r.uniform(-1, 1, pos[-1])


def check_plan(plan, pos, idx, val, col_part, rank):
    """every row of the plan lists exactly the (global column, value) pairs of the original row, in the original order; a row is
    in the outer matrix iff it touches another rank"""
    recv_pid = np.repeat(np.arange(plan.size), plan.recv_counts)
    recv_lidx = np.concatenate(plan.requests) if plan.buffer_size else np.zeros(0, dtype=np.int64)
    offs = np.array([o for o, _ in col_part])
    buffer_global = offs[recv_pid] + recv_lidx
    assert np.all(np.diff(recv_pid * 10 ** 9 + recv_lidx) > 0)              # unique, ascending (rank, index)
    outer_of = {int(row): k for k, row in enumerate(plan.scatter)}
    for i in range(plan.num_rows):
        cols, vals = idx[pos[i]:pos[i + 1]], val[pos[i]:pos[i + 1]]
        lo, hi = col_part[rank]
        remote = np.any((cols < lo) | (cols >= lo + hi))
        a, b = plan.inner_pos[i], plan.inner_pos[i + 1]
        if remote:
            assert a == b and i in outer_of
            k = outer_of[i]
            c, d = plan.outer_pos[k], plan.outer_pos[k + 1]
            assert np.array_equal(buffer_global[plan.outer_idx[c:d]], cols) and np.array_equal(plan.outer_val[c:d], vals)
        else:
            assert i not in outer_of
            assert np.array_equal(plan.inner_idx[a:b] + lo, cols) and np.array_equal(plan.inner_val[a:b], vals)


@pytest.mark.parametrize("size,band", [(1, None), (3, None), (4, 6), (2, 3)])
def test_dist_csr_plan(size, band):
    import sys
    sys.path.insert(0, ROOT)
    from feltor_b200.dist import partition
    from feltor_b200.dist_csr import DistCsrPlan, contiguous_owner
    r = np.random.default_rng(size)
    nrows, ncols = 57, 64
    pos, idx, val = random_csr(r, nrows, ncols, 9, band)
    row_part, col_part = partition(nrows, size), partition(ncols, size)
    g2l = contiguous_owner(col_part)
    plans = []
    for rank in range(size):
        r0, nr = row_part[rank]
        lp = pos[r0:r0 + nr + 1] - pos[r0]
        li, lv = idx[pos[r0]:pos[r0 + nr]], val[pos[r0]:pos[r0 + nr]]
        plan = DistCsrPlan(rank, size, lp, li, lv, g2l, col_part[rank][1])
        check_plan(plan, lp, li, lv, col_part, rank)
        plans.append(plan)
    x = r.uniform(-1, 1, ncols)
    for rank in range(size):                                                  # the exchange, emulated: pack -> route -> buffer
        plans[rank].set_sends([plans[p].requests[rank] for p in range(size)])
    for rank in range(size):
        segs = []
        for p in range(size):
            o = plans[p].send_counts[:rank].sum()
            xp = x[col_part[p][0]:col_part[p][0] + col_part[p][1]]
            segs.append(xp[plans[p].send_idx[o:o + plans[p].send_counts[rank]]])
        buf = np.concatenate(segs) if segs else np.zeros(0)
        offs = np.array([o for o, _ in col_part])
        want = x[(offs[np.repeat(np.arange(size), plans[rank].recv_counts)] + np.concatenate(plans[rank].requests)).astype(int)] \
            if plans[rank].buffer_size else np.zeros(0)
        assert np.array_equal(buf, want)
    with pytest.raises(ValueError):
        contiguous_owner(col_part)(np.array([ncols]))


def _csr_worker(rank, world, port, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from feltor_b200.dist import partition
    from feltor_b200.dist_csr import DistCsrPlan, contiguous_owner, exchange_requests
    r = np.random.default_rng(5)                                             # same matrix on both ranks
    nrows, ncols = 40, 33
    pos, idx, val = random_csr(r, nrows, ncols, 7)
    row_part, col_part = partition(nrows, world), partition(ncols, world)
    r0, nr = row_part[rank]
    plan = DistCsrPlan(rank, world, pos[r0:r0 + nr + 1] - pos[r0], idx[pos[r0]:pos[r0 + nr]], val[pos[r0]:pos[r0 + nr]],
                       contiguous_owner(col_part), col_part[rank][1])
    asked = exchange_requests(plan.requests, rank, world)
    plan.set_sends(asked)
    everyone = [None] * world
    dist.all_gather_object(everyone, [q_.tolist() for q_ in plan.requests])
    ok = all(np.array_equal(asked[p], np.asarray(everyone[p][rank], dtype=np.int32)) for p in range(world))
    ok = ok and int(plan.send_counts.sum()) == sum(len(everyone[p][rank]) for p in range(world))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_dist_csr_requests_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31000 + os.getpid() % 2000
    procs = [ctx.Process(target=_csr_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=100) for _ in range(2)]
    for p in procs:
        p.join(timeout=30)
    assert sorted(res) == [(0, True), (1, True)]


@pytest.mark.parametrize("size,periodic", [(1, True), (3, True), (4, False)])
def test_dist_ell_plan(size, periodic):
    """dg::make_mpi_sparseblockmat host logic (feltor_b200/dist_ell.py, mpi_matrix.h:242-330) on a three-block stencil matrix: a
    block row is in the outer (Coo) matrix iff it touches another rank, the inner matrix keeps it empty, every row lists its
    blocks in slot order with the right global column, and the requests of all ranks match the sends"""
    import sys
    sys.path.insert(0, ROOT)
    from feltor_b200.dist import partition
    from feltor_b200.dist_ell import DistEllPlan
    N, n, bpl = 13, 2, 3

    class M:
        pass
    m = M()
    m.num_rows = m.num_cols = N
    m.bpl, m.n = bpl, n
    cols = np.array([[i - 1, i, i + 1] for i in range(N)])
    cols = cols % N if periodic else np.where((cols < 0) | (cols >= N), -1, cols)
    m.cols_idx = cols.reshape(-1).astype(np.int32)
    m.data_idx = np.tile(np.arange(bpl), N).astype(np.int32)
    m.data = np.arange(bpl * n * n, dtype=np.float64)
    part = partition(N, size)
    plans = [DistEllPlan(r, size, m, part[r][0], part[r][1], part, 4, 5) for r in range(size)]
    offs = np.array([o for o, _ in part])
    for r, p in enumerate(plans):
        o, c = part[r]
        recv_pid = np.repeat(np.arange(size), p.recv_blocks)
        chunk_global = offs[recv_pid] + (np.concatenate(p.requests) if p.num_chunks else np.zeros(0, dtype=np.int64))
        inner = p.inner_cols.reshape(c, bpl)
        for i in range(c):
            g = cols[o + i]
            remote = np.any((g >= 0) & ((g < o) | (g >= o + c)))
            mine = np.nonzero(p.coo_rows == i)[0]
            if remote:
                assert np.all(inner[i] == -1)
                assert np.array_equal(chunk_global[p.coo_cols[mine]], g[g >= 0])          # slot order kept
                assert np.array_equal(p.coo_didx[mine], m.data_idx.reshape(N, bpl)[o + i][g >= 0])
            else:
                assert mine.size == 0
                assert np.array_equal(np.where(inner[i] >= 0, inner[i] + o, -1), g)
        p.set_sends([plans[q].requests[r] for q in range(size)])
        assert p.send_idx.size == int(p.send_blocks.sum()) * p.chunk and p.chunk == n * 4 * 5
    if size == 1:
        assert plans[0].num_chunks == 0 and plans[0].coo_rows.size == 0
