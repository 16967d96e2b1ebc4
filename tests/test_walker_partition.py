"""CPU: the host-side work partition of the warp-walker Elliptic kernel (feltor_b200/csrc/elliptic_walker.cu) covers every
(warp column, cell row) exactly once, keeps pieces at least as long as the ring depth, uses no more warps than there are
and balances their cost -- for odd grid sizes, boundary / periodic columns and both stencil widths."""
import ctypes as C
import numpy as np
import pytest
import feltor_b200 as fb


def partition(Nx, Ny, centered, nwarps, fx_lo, fx_hi, wrapx, tma=1, dot=0):
    L = fb.lib()
    maxt = 8 * nwarps + 64
    tasks = (C.c_int * (3 * maxt))()
    tb = (C.c_int * (nwarps + 1))()
    nt = C.c_int()
    L.debug_walker_partition(Nx, Ny, centered, nwarps, fx_lo, fx_hi, wrapx, tma, dot, tasks, maxt, C.byref(nt), tb)
    t = np.array(tasks[:3 * nt.value]).reshape(-1, 3)
    return t, np.array(tb[:])


@pytest.mark.parametrize("Nx,Ny,centered,wrapx", [(1024, 1024, 0, 0), (1024, 1024, 1, 0), (416, 420, 0, 1), (401, 433, 1, 0),
                                                  (37, 19, 1, 0), (5, 5, 0, 0), (3000, 7, 0, 1), (64, 4096, 0, 0), (1024, 128, 1, 1)])
@pytest.mark.parametrize("nwarps", [148 * 12, 148 * 8, 7])
def test_partition_covers_grid_once(Nx, Ny, centered, wrapx, nwarps):
    HL = 2 if centered else 1
    UL = 32 - 2 * HL
    ncols = (Nx + UL - 1) // UL
    fx_lo, fx_hi = (1, Nx - 1) if not wrapx else (1, Nx - 1)
    t, tb = partition(Nx, Ny, centered, nwarps, fx_lo, fx_hi, wrapx)
    cover = np.zeros((ncols, Ny), dtype=np.int32)
    for c, a, b in t:
        assert 0 <= c < ncols and 0 <= a < b <= Ny
        cover[c, a:b] += 1
    assert (cover == 1).all()
    assert tb[0] == 0 and tb[-1] == len(t) and (np.diff(tb) >= 0).all()
    min_rows = (2 if centered else 1) + 2 + 1
    short = [(c, a, b) for c, a, b in t if b - a < min(min_rows, Ny)]
    assert not short, short


def test_partition_balances_cost():
    # n = 3, 1024^2, forward, Dirichlet in x: interior columns cost 1 per row, the two boundary columns 2.3 (default weight)
    nw = 148 * 12
    t, tb = partition(1024, 1024, 0, nw, 1, 1023, 0)
    ncols = 35
    w = np.ones(ncols)
    w[0] = w[-1] = 2.3
    cost = np.array([sum((b - a) * w[c] + 2.0 for c, a, b in t[tb[g]:tb[g + 1]]) for g in range(nw)])
    used = cost[cost > 0]
    assert len(used) >= 0.95 * nw
    assert used.max() <= 1.12 * used.mean()
