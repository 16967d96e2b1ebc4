"""Golden vectors of dg::create::limiter_stencil + dg::CSRSlopeLimiter (inc/dg/topology/stencil.h:89-256, filter.h:288-336) from the
UNMODIFIED reference (oracle/_ref/libdgref_ds.so, built by oracle/Makefile from /root/reference).  Run in the build container:
    python tests/golden/make_golden_limiter.py          -> tests/golden/limiter_golden.npz"""
import ctypes as C
import os
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
L = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libdgref_ds.so"))
vp = lambda a: a.ctypes.data_as(C.c_void_p)


def ref_limiter(x0, x1, n, N, bc, direction, bound):
    ndim = len(N)
    rows = int(np.prod([n * v for v in N]))
    pos, idx, val = np.empty(rows + 1, dtype=np.int32), np.empty(3 * rows, dtype=np.int32), np.empty(3 * rows)
    arr = lambda t, v: (t * ndim)(*v)
    nnz = L.ref_limiter_stencil(ndim, arr(C.c_double, x0), arr(C.c_double, x1), n, arr(C.c_int, N), arr(C.c_int, bc), direction, bound,
                                vp(pos), vp(idx), vp(val))
    assert nnz == 3 * rows
    return pos, idx, val


def ref_apply(pos, idx, val, mod, x):
    y = np.full(x.size, np.nan)
    L.ref_csr_stencil.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
    L.ref_csr_stencil(4, x.size, x.size, vp(pos), vp(idx), vp(val), mod, vp(x), vp(y))
    return y


def field(rows, seed):
    """smooth part + jumps + noise: some cells are limited, some are not"""
    r = np.random.default_rng(seed)
    t = np.linspace(0, 1, rows)
    return np.sin(7 * t) + (t > 0.4) * 1.5 - (t > 0.8) * 2.2 + 0.05 * r.uniform(-1, 1, rows)


CASES = [("1d_n3_bc%d" % bc, [0.], [1.], 3, [17], [bc], 0, bc) for bc in range(5)] + \
        [("1d_n2_per", [0.], [2.], 2, [9], [0], 0, 0), ("1d_n4_dir", [0.], [2.], 4, [6], [1], 0, 1),
         ("2d_x_dir", [0., 0.], [1., 1.], 3, [7, 5], [1, 0], 0, 1), ("2d_y_neu", [0., 0.], [1., 1.], 3, [7, 5], [1, 2], 1, 2),
         ("2d_y_per", [0., 0.], [1., 1.], 2, [4, 6], [3, 0], 1, 0)]

if __name__ == "__main__":
    out = {}
    for k, (name, x0, x1, n, N, bc, direction, bound) in enumerate(CASES):
        pos, idx, val = ref_limiter(x0, x1, n, N, bc, direction, bound)
        x = field(pos.size - 1, k)
        out[name + "/pos"], out[name + "/idx"], out[name + "/val"], out[name + "/x"] = pos, idx, val, x
        for mod in (0., 0.3):
            out[name + "/y_mod%g" % mod] = ref_apply(pos, idx, val, mod, x)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "limiter_golden.npz"), **out)
    print("wrote limiter_golden.npz with", len(out), "arrays")
