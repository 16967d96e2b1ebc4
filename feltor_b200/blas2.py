"""dg::blas1::dot / dg::blas2::dot / dg::blas2::symv on device tensors (inc/dg/blas1.h:152, inc/dg/blas2.h:94,325)."""
import ctypes as C
import numpy as np
import torch
from ._lib import lib, DgbError
from ._dev import ptr, stream, hptr

d = C.c_double
BIN_COUNT = 39


class DotResult(C.Structure):
    """mirror of dgb_dot_result (include/dgb200.h)"""
    _fields_ = [("acc", C.c_int64 * BIN_COUNT), ("value", C.c_double), ("status", C.c_int32), ("pad", C.c_int32)]


class DotWorkspace:
    """Device scratch of the superaccumulator reduction (dgb_dot_ws)."""

    def __init__(self):
        self.h = C.c_void_p()
        lib().dot_ws_create(C.byref(self.h))

    def __del__(self):
        try:
            lib().dot_ws_destroy(self.h)
        except Exception:
            pass


_default_ws = None


def _ws():
    global _default_ws
    if _default_ws is None:
        _default_ws = DotWorkspace()
    return _default_ws


def _operand(v):
    """(pointer, scalar) pair: dot(1., v) style scalar operands are allowed (blas1_cuda.cuh:30)."""
    if isinstance(v, (int, float)):
        return None, d(float(v))
    return ptr(v), d(0.0)


def superacc(*ops, ws=None):
    """doDot_superacc: returns (acc[39] normalised int64 numpy, status).  2 or 3 operands."""
    ws = ws or _ws()
    sizes = {o.numel() for o in ops if not isinstance(o, (int, float))}
    if len(sizes) != 1:
        raise ValueError("dg::Error: dot: operand sizes differ %s" % sorted(sizes))   # blas1.h: the reference throws as well
    n = sizes.pop()
    acc = np.zeros(BIN_COUNT, dtype=np.int64)
    val = C.c_double()
    st = C.c_int()
    tensors = all(not isinstance(o, (int, float)) for o in ops)
    try:
        if tensors and len(ops) == 2:
            lib().dot2(ws.h, n, ptr(ops[0]), ptr(ops[1]), hptr(acc), C.byref(val), C.byref(st), stream())
        elif tensors and len(ops) == 3:
            lib().dot3(ws.h, n, ptr(ops[0]), ptr(ops[1]), ptr(ops[2]), hptr(acc), C.byref(val), C.byref(st), stream())
        else:
            res = torch.zeros(ctypes_sizeof_result() // 8, dtype=torch.int64, device="cuda")
            pr = [_operand(o) for o in ops]
            if len(ops) == 2:
                lib().exdot2(ws.h, n, pr[0][0], pr[0][1], pr[1][0], pr[1][1], ptr(res), stream())
            else:
                lib().exdot3(ws.h, n, pr[0][0], pr[0][1], pr[1][0], pr[1][1], pr[2][0], pr[2][1], ptr(res), stream())
            r = DotResult.from_buffer_copy(res.cpu().numpy().tobytes())
            return np.array(r.acc[:], dtype=np.int64), r.value, r.status
    except DgbError as e:
        if e.code != -3:
            raise
    return acc, val.value, st.value


def ctypes_sizeof_result():
    return C.sizeof(DotResult)


def dot(*ops, ws=None):
    """blas1::dot(x,y) / blas2::dot(x,W,y) / blas2::dot(W,x) = dot(x,W,x).  Raises like blas1.h:161 on NaN/Inf."""
    if len(ops) == 2 and getattr(ops[0], "_is_weights", False):
        ops = (ops[1], ops[0], ops[1])
    acc, val, st = superacc(*ops, ws=ws)
    if st != 0:
        raise FloatingPointError("dg::Error: dot product failed since one of the inputs contains NaN or Inf")
    return val


class Ell:
    """EllSparseBlockMat (inc/dg/backend/sparseblockmat.h:44-188): host arrays + device-resident plan."""

    def __init__(self, num_rows, num_cols, bpl, n, left_size, right_size, data, cols_idx, data_idx, right_range=None):
        self.num_rows, self.num_cols, self.bpl, self.n = int(num_rows), int(num_cols), int(bpl), int(n)
        self.left_size, self.right_size = int(left_size), int(right_size)
        self.data = np.ascontiguousarray(data, dtype=np.float64)
        self.cols_idx = np.ascontiguousarray(cols_idx, dtype=np.int32)
        self.data_idx = np.ascontiguousarray(data_idx, dtype=np.int32)
        self.nblocks = self.data.size // (self.n * self.n)
        self.right_range = tuple(right_range) if right_range is not None else (0, self.right_size)
        self._h = None

    @classmethod
    def from_like(cls, m):
        """from any object that carries the EllSparseBlockMat fields (num_rows, ..., data, cols_idx, data_idx)"""
        return cls(m.num_rows, m.num_cols, m.bpl, m.n, m.left_size, m.right_size, m.data, m.cols_idx, m.data_idx,
                   m.right_range)

    def meta(self):
        return np.array([self.num_rows, self.num_cols, self.bpl, self.n, self.left_size, self.right_size,
                         self.nblocks, self.right_range[0], self.right_range[1], 0], dtype=np.int32)

    @property
    def total_rows(self):
        return self.num_rows * self.n * self.left_size * self.right_size

    @property
    def total_cols(self):
        return self.num_cols * self.n * self.left_size * self.right_size

    class _Host(C.Structure):
        _fields_ = [("num_rows", C.c_int), ("num_cols", C.c_int), ("blocks_per_line", C.c_int), ("n", C.c_int),
                    ("left_size", C.c_int), ("right_size", C.c_int), ("num_blocks", C.c_int),
                    ("right_range", C.c_int * 2), ("data", C.c_void_p), ("cols_idx", C.c_void_p),
                    ("data_idx", C.c_void_p)]

    def host_struct(self):
        h = Ell._Host(self.num_rows, self.num_cols, self.bpl, self.n, self.left_size, self.right_size, self.nblocks,
                      (C.c_int * 2)(*self.right_range), self.data.ctypes.data, self.cols_idx.ctypes.data,
                      self.data_idx.ctypes.data)
        return h

    @property
    def handle(self):
        if self._h is None:
            self._h = C.c_void_p()
            hs = self.host_struct()
            lib().ell_create(C.byref(self._h), C.byref(hs))
        return self._h

    def set_left_size(self, v):
        self.left_size = int(v)
        if self._h is not None:
            lib().ell_set_left_size(self._h, int(v))

    def set_right_size(self, v):
        self.right_size = int(v)
        self.right_range = (0, int(v))
        if self._h is not None:
            lib().ell_set_right_size(self._h, int(v))

    def set_right_range(self, a, b):
        self.right_range = (int(a), int(b))
        if self._h is not None:
            lib().ell_set_right_range(self._h, int(a), int(b))

    def symv(self, alpha, x, beta, y, generic=False):
        """blas2::symv(alpha, M, x, beta, y); size mismatch raises like blas2_sparseblockmat.h:45-50"""
        if x.numel() != self.total_cols or y.numel() != self.total_rows:
            raise ValueError("dg::Error: x/y size does not match matrix size")
        fn = lib().ell_symv_generic if generic else lib().ell_symv
        fn(self.handle, d(alpha), ptr(x), d(beta), ptr(y), stream())

    def __del__(self):
        try:
            if self._h is not None:
                lib().ell_destroy(self._h)
        except Exception:
            pass


def symv(*a):
    """blas2::symv(M,x,y) | symv(alpha,M,x,beta,y) for Ell matrices and vector preconditioners (blas2.h:139-330)."""
    if len(a) == 3:
        M, x, y = a
        alpha, beta = 1.0, 0.0
    else:
        alpha, M, x, beta, y = a
    if isinstance(M, torch.Tensor):  # diagonal matrix: blas2_dispatch_shared.h:116-135 -> pointwiseDot
        from . import blas1
        if len(a) == 3:
            return blas1.pointwiseDot(M, x, y)
        return blas1.pointwiseDot(alpha, M, x, beta, y)
    return M.symv(alpha, x, beta, y)


STENCILS = {"median": 0, "swm": 1, "average": 2, "symv": 3, "slope": 4}


def stencil(kind, pos, idx, val, x, y, alpha=0.0):
    """blas2::stencil(f, M, x, y) (blas2.h:454) with f = CSRMedianFilter / CSRSWMFilter(alpha) / CSRAverageFilter /
    CSRSymvFilter / CSRSlopeLimiter(alpha) (topology/filter.h:174-336); pos, idx int32 device tensors, val float64 or None."""
    lib().csr_stencil(STENCILS[kind], pos.numel() - 1, ptr(pos), ptr(idx), ptr(val), d(alpha), ptr(x), ptr(y), stream())


def spgemm(B_rows, B_cols, C_cols, B, Cm):
    """A = B Cm for CSR matrices on the device (pos, idx int32; val float64), bitwise the reference's host product
    dg::SparseMatrix::operator* (sparsematrix.h:549-566, sparsematrix_cpu.h:19-95): returns (pos, idx, val) device tensors"""
    import torch
    h, nnz = C.c_void_p(), C.c_longlong()
    lib().csr_spgemm_symbolic(C.byref(h), B_rows, B_cols, C_cols, ptr(B[0]), ptr(B[1]), ptr(Cm[0]), ptr(Cm[1]), C.byref(nnz), stream())
    try:
        pos = torch.empty(B_rows + 1, dtype=torch.int32, device="cuda")
        idx = torch.empty(max(nnz.value, 1), dtype=torch.int32, device="cuda")
        val = torch.empty(max(nnz.value, 1), dtype=torch.float64, device="cuda")
        lib().csr_spgemm_numeric(h, ptr(B[0]), ptr(B[1]), ptr(B[2]), ptr(Cm[0]), ptr(Cm[1]), ptr(Cm[2]), ptr(pos), ptr(idx), ptr(val), stream())
    finally:
        lib().csr_spgemm_destroy(h)
    return pos, idx[:nnz.value], val[:nnz.value]


def spgemm_host(B_rows, B_cols, C_cols, B, Cm):
    """the same for host (numpy) arrays, as the binding of SparseMatrix::operator* uses it"""
    import numpy as np
    B = (np.ascontiguousarray(B[0], dtype=np.int32), np.ascontiguousarray(B[1], dtype=np.int32), np.ascontiguousarray(B[2], dtype=np.float64))
    Cm = (np.ascontiguousarray(Cm[0], dtype=np.int32), np.ascontiguousarray(Cm[1], dtype=np.int32), np.ascontiguousarray(Cm[2], dtype=np.float64))
    h, nnz = C.c_void_p(), C.c_longlong()
    a = lambda x: C.c_void_p(x.ctypes.data)
    lib().csr_spgemm_host_begin(C.byref(h), B_rows, B_cols, C_cols, a(B[0]), a(B[1]), a(B[2]), a(Cm[0]), a(Cm[1]), a(Cm[2]), C.byref(nnz))
    pos, idx, val = np.empty(B_rows + 1, dtype=np.int32), np.empty(max(nnz.value, 1), dtype=np.int32), np.empty(max(nnz.value, 1))
    lib().csr_spgemm_host_finish(h, a(pos), a(idx), a(val))
    return pos, idx[:nnz.value], val[:nnz.value]

