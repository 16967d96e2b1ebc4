#!/usr/bin/env python
"""A = B C at the size dg::geo::Fieldaligned multiplies for config 4 (n = 3, 96 x 96 cells, mx = my = 10): B = fine-to-coarse
projection (900 entries per row, 74.6 M), C = interpolation with n^2 entries per fine row (74.6 M).  Times dgb_csr_spgemm with
device-resident operands (CUDA events), the host-array entry points the binding uses (wall clock, copies included) and the
reference's host kernel (oracle/_ref/libdgref_ds.so: dg::SparseMatrix::operator*), and compares the three products bit for bit.
    python tools/spgemm_bench.py [Nx Ny mx my]"""
import ctypes as C
import json
import os
import sys
import time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def operands(n, Nx, Ny, mx, my, seed=3):
    r = np.random.default_rng(seed)
    fx, fy = n * Nx * mx, n * Ny * my
    coarse, fine = n * n * Nx * Ny, fx * fy
    cy, ky, cx, kx = np.meshgrid(np.arange(Ny), np.arange(n), np.arange(Nx), np.arange(n), indexing="ij")
    j, i = np.meshgrid(np.arange(n * my), np.arange(n * mx), indexing="ij")
    Bidx = ((cy.reshape(-1, 1, 1) * n * my + j[None]) * fx + cx.reshape(-1, 1, 1) * n * mx + i[None]).reshape(-1).astype(np.int32)
    Bpos = (np.arange(coarse + 1, dtype=np.int64) * (n * my * n * mx)).astype(np.int32)
    Bval = r.uniform(-1, 1, Bidx.size)
    f = np.arange(fine)
    tx = np.clip((f % fx) // (n * mx) + r.integers(-1, 2, fine), 0, Nx - 1)
    ty = np.clip((f // fx) // (n * my) + r.integers(-1, 2, fine), 0, Ny - 1)
    jj, ii = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    Cidx = (((ty[:, None, None] * n + jj[None]) * Nx + tx[:, None, None]) * n + ii[None]).reshape(-1).astype(np.int32)
    Cpos = (np.arange(fine + 1, dtype=np.int64) * n * n).astype(np.int32)
    Cval = r.uniform(-1, 1, Cidx.size)
    return (coarse, fine, coarse), (Bpos, Bidx, Bval), (Cpos, Cidx, Cval)


if __name__ == "__main__":
    import torch
    from feltor_b200 import blas2
    from feltor_b200._dev import dvec, hvec
    dims = [int(v) for v in sys.argv[1:5]] if len(sys.argv) >= 5 else [96, 96, 10, 10]
    shape, B, Cm = operands(3, *dims)
    rec = {"config": "projection x interpolation, n=3 %dx%d mx=%d my=%d" % tuple(dims), "B_nnz": int(B[1].size), "C_nnz": int(Cm[1].size)}
    dB, dC = tuple(dvec(a) for a in B), tuple(dvec(a) for a in Cm)
    blas2.spgemm(*shape, dB, dC)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    times = []
    for _ in range(3):
        torch.cuda.synchronize()
        ev[0].record()
        A = blas2.spgemm(*shape, dB, dC)
        ev[1].record()
        torch.cuda.synchronize()
        times.append(ev[0].elapsed_time(ev[1]))
    rec["device_ms"] = float(np.median(times))
    rec["A_nnz"] = int(A[1].numel())
    rec["candidates"] = int(B[1].size) * 9
    t0 = time.time()
    Ah = blas2.spgemm_host(*shape, B, Cm)
    rec["host_arrays_ms"] = (time.time() - t0) * 1e3
    Ad = tuple(hvec(t) for t in A)
    same = all(np.array_equal(a, b) for a, b in zip(Ad[:2], Ah[:2])) and np.array_equal(Ad[2].view(np.int64), Ah[2].view(np.int64))
    rec["device_equals_host_entry"] = bool(same)
    ref = os.path.join(ROOT, "oracle", "_ref", "libdgref_ds.so")
    if os.path.exists(ref):
        from test_spgemm import ref_spgemm
        t0 = time.time()
        Ar = ref_spgemm(shape, B, Cm)
        rec["reference_host_ms"] = (time.time() - t0) * 1e3
        rec["bitwise_equal_reference"] = bool(np.array_equal(Ar[0], Ad[0]) and np.array_equal(Ar[1], Ad[1]) and np.array_equal(Ar[2].view(np.int64), Ad[2].view(np.int64)))
        rec["speedup_device_vs_reference"] = rec["reference_host_ms"] / rec["device_ms"]
        rec["speedup_host_entry_vs_reference"] = rec["reference_host_ms"] / rec["host_arrays_ms"]
    print(json.dumps(rec))
