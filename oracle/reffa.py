"""TEST INFRASTRUCTURE ONLY: ctypes binding of oracle/_ref/libdgref_fa.so (the unmodified reference dg::geo::Fieldaligned / dg::geo::DS on
the circular field of inc/geometries/ds_b.cpp, wrapped by oracle/ref_fa.cpp)."""
import ctypes as C
import os
import numpy as np

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libdgref_fa.so")
_lib = None


def available():
    return os.path.exists(_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(_PATH)
        _lib.ref_fa_create.restype = C.c_void_p
        _lib.ref_fa_create.argtypes = [C.c_int] * 6 + [C.c_char_p]
        for name in ("ref_fa_free", "ref_fa_plane_size", "ref_fa_size"):
            getattr(_lib, name).argtypes = [C.c_void_p]
        _lib.ref_fa_delta_phi.restype = C.c_double
        _lib.ref_fa_delta_phi.argtypes = [C.c_void_p]
        _lib.ref_fa_nnz.argtypes = [C.c_void_p, C.c_int]
        _lib.ref_fa_csr.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.ref_fa_field.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib.ref_fa_testfunction.argtypes = [C.c_void_p, C.c_void_p]
        _lib.ref_fa_shift.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _lib.ref_fa_ds.restype = C.c_double
        _lib.ref_fa_ds.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_double, C.c_void_p, C.c_int]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


FIELDS = {"bphi": 0, "bphiM": 1, "bphiP": 2, "sqrtG": 3, "sqrtGm": 4, "sqrtGp": 5, "hbm": 6, "hbp": 7}
KINDS = {"centered": 0, "forward": 1, "backward": 2, "dss": 3, "divCentered": 4}


class RefFieldaligned:
    def __init__(self, n, Nx, Ny, Nz, mx=10, my=10, method="dg"):
        self.h = lib().ref_fa_create(n, Nx, Ny, Nz, mx, my, method.encode())
        if not self.h:
            raise RuntimeError("ref_fa_create failed")
        self.plane, self.size = lib().ref_fa_plane_size(self.h), lib().ref_fa_size(self.h)
        self.Nz = self.size // self.plane
        self.delta_phi = lib().ref_fa_delta_phi(self.h)

    def csr(self, which):
        """which: "plus" | "minus" -> (row_offsets, column_indices, values) of the 2-d interpolation matrix"""
        w = 0 if which == "plus" else 1
        nnz = lib().ref_fa_nnz(self.h, w)
        pos, idx, val = np.empty(self.plane + 1, dtype=np.int32), np.empty(nnz, dtype=np.int32), np.empty(nnz)
        lib().ref_fa_csr(self.h, w, _p(pos), _p(idx), _p(val))
        return pos, idx, val

    def field(self, name):
        out = np.empty(self.size)
        lib().ref_fa_field(self.h, FIELDS[name], _p(out))
        return out

    def testfunction(self):
        out = np.empty(self.size)
        lib().ref_fa_testfunction(self.h, _p(out))
        return out

    def shift(self, which, f):
        out = np.empty(self.size)
        lib().ref_fa_shift(self.h, 0 if which == "plus" else 1, _p(np.ascontiguousarray(f)), _p(out))
        return out

    def ds(self, kind, alpha, f, beta, g, reps=1):
        """returns (result, seconds per application)"""
        out = np.array(g, dtype=np.float64, copy=True)
        sec = lib().ref_fa_ds(self.h, KINDS[kind], alpha, _p(np.ascontiguousarray(f)), beta, _p(out), reps)
        return out, sec

    def threads(self):
        return lib().ref_fa_threads()

    def __del__(self):
        try:
            lib().ref_fa_free(self.h)
        except Exception:
            pass
