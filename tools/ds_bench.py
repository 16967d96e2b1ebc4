#!/usr/bin/env python
"""DS benchmark (config 4 of BASELINE.json): ds.centered(f, g) on a 3-D grid n=3, Nx=Ny=96, Nz=64 with
field-line interpolation matrices I+ / I- of the reference's structure (one CSR matrix per direction on the perpendicular
plane, applied to every plane; "dg": the target point's 2x2 cells x n^2 nodes = 36 entries per row, "cubic": 3x3 cells = 81).
The matrices are synthetic (the field-line tracing that builds them is host set-up code outside the hot path): a smooth
displacement field of up to two cells, positive weights summing to one.  Prints time and GB/s for the fused kernel and for the
three-launch composition (ePlus, eMinus, ds_centered); algorithmic bytes = 24 B per 3-D element + both matrices once.
  python tools/ds_bench.py [--cells 96] [--planes 64] [--reps 30]"""
import argparse
import ctypes as C
import json
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feltor_b200 as fb  # noqa: E402
from feltor_b200._dev import dvec, ptr, stream  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=96)
ap.add_argument("--planes", type=int, default=64)
ap.add_argument("--reps", type=int, default=30)
ap.add_argument("--reference", action="store_true", help="also time the reference's OpenMP CSR kernel (oracle/_ref) on the host")
args = ap.parse_args()
n, N, Nz = 3, args.cells, args.planes
L = fb.lib()
try:
    PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0


def interpolation_matrix(rng, cells_per_dim):
    """rows = (n N)^2 perpendicular points; row of point (cx,kx,cy,ky) couples to all nodes of a cells_per_dim^2 block of cells"""
    S = n * N
    rows = S * S
    iy, ix = np.divmod(np.arange(rows), S)
    cx, cy = ix // n, iy // n
    # field lines displace neighbouring points coherently: a smooth displacement field of up to two cells
    xs, ys = (ix + 0.5) / S, (iy + 0.5) / S
    dx = np.rint(2. * np.sin(2 * np.pi * ys) * np.cos(np.pi * xs) + rng.uniform(-0.3, 0.3, rows)).astype(np.int64)
    dy = np.rint(2. * np.cos(2 * np.pi * xs) * np.sin(np.pi * ys) + rng.uniform(-0.3, 0.3, rows)).astype(np.int64)
    tx = np.clip(cx + dx, 0, N - cells_per_dim)  # first cell of the target block
    ty = np.clip(cy + dy, 0, N - cells_per_dim)
    w = cells_per_dim * n
    off = np.arange(w)
    cols = ((ty * n)[:, None, None] + off[None, :, None]) * S + (tx * n)[:, None, None] + off[None, None, :]
    idx = cols.reshape(rows, -1).astype(np.int32)
    val = rng.uniform(0.1, 1., idx.shape)
    val /= val.sum(axis=1, keepdims=True)
    pos = (np.arange(rows + 1) * idx.shape[1]).astype(np.int32)
    return pos, idx.reshape(-1), val.reshape(-1)


def timeit(f, reps):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        f()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    return float(np.median(ts))


rng = np.random.default_rng(0)
rows = (n * N) ** 2
size = rows * Nz
f = dvec(rng.uniform(-1, 1, size))
bphi = dvec(rng.uniform(0.5, 1.5, size))
g = torch.zeros(size, dtype=torch.float64, device="cuda")
fp, fm = torch.zeros_like(g), torch.zeros_like(g)
ghost = torch.zeros(rows, dtype=torch.float64, device="cuda")
dphi = 2 * np.pi / Nz
print(f"# DS centered, n={n} {N}x{N}x{Nz}: {size} elements ({size*8/1e6:.1f} MB per 3-D vector), L2 flushed between calls")
for name, cpd in (("dg (36 per row)", 2), ("cubic (81 per row)", 3)):
    P = [dvec(a) for a in interpolation_matrix(rng, cpd)]
    M = [dvec(a) for a in interpolation_matrix(rng, cpd)]
    nnz = P[1].numel()
    alg = 24 * size + 2 * nnz * 12
    t = timeit(lambda: L.ds_centered_fused(rows, Nz, ptr(P[0]), ptr(P[1]), ptr(P[2]), ptr(M[0]), ptr(M[1]), ptr(M[2]), C.c_double(1.),
                                           ptr(f), ptr(bphi), C.c_double(dphi), C.c_double(0.), ptr(g), stream()), args.reps)
    print(f"{name:20s} fused DS::centered      {t*1e6:9.1f} us  {alg/t/1e9:8.1f} GB/s  {alg/t/1e9/PEAK*100:5.1f}% of {PEAK:.0f}  ({alg/size:.1f} B/element)", flush=True)

    hp, hm = C.c_void_p(), C.c_void_p()
    L.gather_plan_create(C.byref(hp), rows, rows, ptr(P[0]), ptr(P[1]), ptr(P[2]), stream())
    L.gather_plan_create(C.byref(hm), rows, rows, ptr(M[0]), ptr(M[1]), ptr(M[2]), stream())
    g2 = torch.zeros_like(g)
    t3 = timeit(lambda: L.gather_ds_centered(hp, hm, Nz, C.c_double(1.), ptr(f), ptr(bphi), C.c_double(dphi), C.c_double(0.), ptr(g2), stream()), args.reps)
    same = bool((g.view(torch.int64) == g2.view(torch.int64)).all())
    print(f"{name:20s} gather plan (sliced ELL) {t3*1e6:9.1f} us  {alg/t3/1e9:8.1f} GB/s  {alg/t3/1e9/PEAK*100:5.1f}%  bitwise == CSR kernel: {same}", flush=True)
    L.gather_plan_destroy(hp); L.gather_plan_destroy(hm)

    def composed():
        L.fa_shift(1, rows, Nz, ptr(P[0]), ptr(P[1]), ptr(P[2]), ptr(f), ptr(fp), 0, None, None, ptr(ghost), C.c_double(dphi), stream())
        L.fa_shift(0, rows, Nz, ptr(M[0]), ptr(M[1]), ptr(M[2]), ptr(f), ptr(fm), 0, None, None, ptr(ghost), C.c_double(dphi), stream())
        L.ds_apply(2, size, C.c_double(1.), ptr(fm), ptr(fp), None, None, ptr(bphi), None, C.c_double(dphi), C.c_double(0.), ptr(g), stream())
    if args.reference:
        from oracle import refwrap as R
        if R.available():
            import time
            hP, hM = [a.cpu().numpy() for a in P], [a.cpu().numpy() for a in M]
            hf = f.cpu().numpy()
            hb = bphi.cpu().numpy()
            tp, tm, hg = np.zeros(size), np.zeros(size), np.zeros(size)
            rP, rM = R.Csr(rows, rows, hP[0], hP[1], hP[2]), R.Csr(rows, rows, hM[0], hM[1], hM[2])
            t0 = time.time()
            for k in range(Nz):  # Fieldaligned::ePlus / eMinus: one symv per plane (fieldaligned.h:850-912), then the formula
                rP.symv(1., hf[((k + 1) % Nz) * rows:((k + 1) % Nz + 1) * rows], 0., tp[k * rows:(k + 1) * rows])
                rM.symv(1., hf[((k - 1) % Nz) * rows:((k - 1) % Nz + 1) * rows], 0., tm[k * rows:(k + 1) * rows])
            hg[:] = 1. * hb * (tp - tm) / 2. / dphi
            tr = time.time() - t0
            print(f"{name:20s} reference OpenMP CSR x {2*Nz} planes + numpy formula {tr*1e6:9.1f} us  {alg/tr/1e9:8.2f} GB/s  ({R.lib().ref_get_max_threads()} host threads)", flush=True)
    t2 = timeit(composed, args.reps)
    print(f"{name:20s} ePlus + eMinus + formula {t2*1e6:9.1f} us  {alg/t2/1e9:8.1f} GB/s  {alg/t2/1e9/PEAK*100:5.1f}% (same algorithmic bytes)", flush=True)
