"""CPU: the oracle's restatement of dg::detail::spgemm_cpu_kernel (inc/dg/backend/sparsematrix_cpu.h:19-95) against the UNMODIFIED
reference (dg::SparseMatrix::operator* through oracle/_ref/libdgref_ds.so) and against a golden product committed under
tests/golden/spgemm_golden.npz (made by this file: python tests/test_spgemm.py)."""
import ctypes as C
import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import orc
from util import same_bits

REF = os.path.join(ROOT, "oracle", "_ref", "libdgref_ds.so")
GOLD = os.path.join(ROOT, "tests", "golden", "spgemm_golden.npz")


def random_pair(seed, rows=60, mid=45, cols=70, per_b=8, per_c=6, sort=False):
    """unsorted rows with duplicate columns, empty rows, explicit zeros, values of mixed magnitude (rounding order matters)"""
    r = np.random.default_rng(seed)

    def mat(nr, nc, per):
        counts = r.integers(0, per + 1, nr)
        pos = np.concatenate([[0], np.cumsum(counts)]).astype(np.int32)
        idx = r.integers(0, nc, pos[-1]).astype(np.int32)
        val = r.uniform(-1, 1, pos[-1]) * 10. ** r.integers(-6, 7, pos[-1])
        val[r.uniform(size=val.size) < 0.05] = 0.
        if sort:
            for i in range(nr):
                o = np.argsort(idx[pos[i]:pos[i + 1]], kind="stable")
                idx[pos[i]:pos[i + 1]], val[pos[i]:pos[i + 1]] = idx[pos[i]:pos[i + 1]][o], val[pos[i]:pos[i + 1]][o]
        return pos, idx, val
    return (rows, mid, cols), mat(rows, mid, per_b), mat(mid, cols, per_c)


def ref_spgemm(shape, B, Cm):
    L = C.CDLL(REF)
    L.ref_spgemm.restype = C.c_longlong
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    pos = np.zeros(shape[0] + 1, dtype=np.int32)
    args = [shape[0], shape[1], shape[2], vp(B[0]), vp(B[1]), vp(B[2]), vp(Cm[0]), vp(Cm[1]), vp(Cm[2]), vp(pos)]
    nnz = L.ref_spgemm(*args, None, None)
    idx, val = np.zeros(max(nnz, 1), dtype=np.int32), np.zeros(max(nnz, 1))
    L.ref_spgemm(*args, vp(idx), vp(val))
    return pos, idx[:nnz], val[:nnz]


@pytest.mark.parametrize("seed", range(6))
def test_oracle_spgemm_vs_live_reference(seed):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/libdgref_ds.so not built (needs /root/reference)")
    shape, B, Cm = random_pair(seed, sort=seed % 2 == 0)
    want, got = ref_spgemm(shape, B, Cm), orc.spgemm(shape[0], shape[2], B, Cm)
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]) and same_bits(got[2], want[2])


def test_oracle_spgemm_vs_golden():
    g = np.load(GOLD)
    B, Cm = (g["Bpos"], g["Bidx"], g["Bval"]), (g["Cpos"], g["Cidx"], g["Cval"])
    got = orc.spgemm(int(g["shape"][0]), int(g["shape"][2]), B, Cm)
    assert np.array_equal(got[0], g["Apos"]) and np.array_equal(got[1], g["Aidx"]) and same_bits(got[2], g["Aval"])
    assert np.all(np.diff(g["Apos"]) >= 0)
    for i in range(int(g["shape"][0])):                                  # sorted, distinct columns in every row
        assert np.all(np.diff(g["Aidx"][g["Apos"][i]:g["Apos"][i + 1]]) > 0)


if __name__ == "__main__":
    shape, B, Cm = random_pair(1234, rows=90, mid=80, cols=64, per_b=12, per_c=9)
    A = ref_spgemm(shape, B, Cm)
    np.savez_compressed(GOLD, shape=np.array(shape), Bpos=B[0], Bidx=B[1], Bval=B[2], Cpos=Cm[0], Cidx=Cm[1], Cval=Cm[2], Apos=A[0], Aidx=A[1], Aval=A[2])
    print("wrote", GOLD, "nnz", A[1].size)
