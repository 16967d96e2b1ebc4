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
