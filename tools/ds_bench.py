#!/usr/bin/env python
"""DS benchmark (config 4 of BASELINE.json): ds.centered(f, g) on a 3-D grid n=3, Nx=Ny=96, Nz=64 with
field-line interpolation matrices I+ / I- of the reference's structure (one CSR matrix per direction on the perpendicular
plane, applied to every plane; "dg": the target point's 2x2 cells x n^2 nodes = 36 entries per row, "cubic": 3x3 cells = 81).
The matrices are synthetic (the field-line tracing that builds them is host set-up code outside the hot path): a smooth
displacement field of up to two cells, positive weights summing to one.  Prints time and GB/s for the gather-plan kernel
(sliced ELL), the CSR fused kernel and the three-launch composition (ePlus, eMinus, ds_centered); algorithmic bytes = 24 B
per 3-D element + both matrices once.  `python bench.py --workload ds` runs this with the reference's OpenMP CSR kernel
beside it.
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

n = 3


def peak():
    try:
        return json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        return 6650.0


def interpolation_matrix(rng, N, cells_per_dim):
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


def run(cells=96, planes=64, reps=30, kinds=(("dg", 2), ("cubic", 3))):
    """returns a list of result dicts (one per interpolation method)"""
    L = fb.lib()
    N, Nz = cells, planes
    rng = np.random.default_rng(0)
    rows = (n * N) ** 2
    size = rows * Nz
    f = dvec(rng.uniform(-1, 1, size))
    bphi = dvec(rng.uniform(0.5, 1.5, size))
    g = torch.zeros(size, dtype=torch.float64, device="cuda")
    g2 = torch.zeros_like(g)
    fp, fm = torch.zeros_like(g), torch.zeros_like(g)
    ghost = torch.zeros(rows, dtype=torch.float64, device="cuda")
    dphi = 2 * np.pi / Nz
    out = []
    for name, cpd in kinds:
        P = [dvec(a) for a in interpolation_matrix(rng, N, cpd)]
        M = [dvec(a) for a in interpolation_matrix(rng, N, cpd)]
        nnz = P[1].numel()
        alg = 24 * size + 2 * nnz * 12
        t_csr = timeit(lambda: L.ds_centered_fused(rows, Nz, ptr(P[0]), ptr(P[1]), ptr(P[2]), ptr(M[0]), ptr(M[1]), ptr(M[2]), C.c_double(1.),
                                                   ptr(f), ptr(bphi), C.c_double(dphi), C.c_double(0.), ptr(g), stream()), reps)
        hp, hm = C.c_void_p(), C.c_void_p()
        L.gather_plan_create(C.byref(hp), rows, rows, ptr(P[0]), ptr(P[1]), ptr(P[2]), stream())
        L.gather_plan_create(C.byref(hm), rows, rows, ptr(M[0]), ptr(M[1]), ptr(M[2]), stream())
        t_plan = timeit(lambda: L.gather_ds_centered(hp, hm, Nz, C.c_double(1.), ptr(f), ptr(bphi), C.c_double(dphi), C.c_double(0.), ptr(g2), stream()), reps)
        same = bool((g.view(torch.int64) == g2.view(torch.int64)).all())
        L.gather_plan_destroy(hp); L.gather_plan_destroy(hm)

        def composed():
            L.fa_shift(1, rows, Nz, ptr(P[0]), ptr(P[1]), ptr(P[2]), ptr(f), ptr(fp), 0, None, None, ptr(ghost), C.c_double(dphi), stream())
            L.fa_shift(0, rows, Nz, ptr(M[0]), ptr(M[1]), ptr(M[2]), ptr(f), ptr(fm), 0, None, None, ptr(ghost), C.c_double(dphi), stream())
            L.ds_apply(2, size, C.c_double(1.), ptr(fm), ptr(fp), None, None, ptr(bphi), None, C.c_double(dphi), C.c_double(0.), ptr(g), stream())
        t_comp = timeit(composed, reps)
        out.append({"method": name, "entries_per_row": nnz // rows, "elements": size, "algorithmic_bytes": alg,
                    "gather_plan_us": t_plan * 1e6, "csr_fused_us": t_csr * 1e6, "composition_us": t_comp * 1e6,
                    "gather_plan_gbs": alg / t_plan / 1e9, "bitwise_plan_equals_csr": same, "cells": N, "planes": Nz})
    return out


def run_real(cells=96, planes=64, reps=30, methods=("dg",), mx=10, my=10, cpu_reps=2):
    """config 4 on the REAL matrices: dg::geo::Fieldaligned of the unmodified reference (circular field of ds_b.cpp:70-84) is
    built on the host through oracle/_ref/libdgref_fa.so (test infrastructure; about a minute at 96 x 96), its I+ / I- and bphi
    feed the library's three layouts; the reference's own ds.centered (OpenMP) is timed beside them and checks the result."""
    from oracle import reffa
    if not reffa.available():
        return None
    L = fb.lib()
    out = []
    for method in methods:
        F = reffa.RefFieldaligned(n, cells, cells, planes, mx, my, method)
        rows, size, Nz = F.plane, F.size, planes
        fh = F.testfunction()
        gref, cpu_sec = F.ds("centered", 1., fh, 0., np.zeros(size), reps=cpu_reps)
        P = [torch.from_numpy(a).cuda() for a in F.csr("plus")]
        M = [torch.from_numpy(a).cuda() for a in F.csr("minus")]
        f, bphi = dvec(fh), dvec(F.field("bphi"))
        g, g2, g3 = (torch.zeros(size, dtype=torch.float64, device="cuda") for _ in range(3))
        nnz = P[1].numel()
        alg = 24 * size + 2 * nnz * 12
        dphi = F.delta_phi
        t_csr = timeit(lambda: L.ds_centered_fused(rows, Nz, ptr(P[0]), ptr(P[1]), ptr(P[2]), ptr(M[0]), ptr(M[1]), ptr(M[2]), C.c_double(1.),
                                                   ptr(f), ptr(bphi), C.c_double(dphi), C.c_double(0.), ptr(g), stream()), reps)
        hp, hm = C.c_void_p(), C.c_void_p()
        L.gather_plan_create(C.byref(hp), rows, rows, ptr(P[0]), ptr(P[1]), ptr(P[2]), stream())
        L.gather_plan_create(C.byref(hm), rows, rows, ptr(M[0]), ptr(M[1]), ptr(M[2]), stream())
        t_plan = timeit(lambda: L.gather_ds_centered(hp, hm, Nz, C.c_double(1.), ptr(f), ptr(bphi), C.c_double(dphi), C.c_double(0.), ptr(g2), stream()), reps)
        L.gather_plan_destroy(hp); L.gather_plan_destroy(hm)
        rec = {"method": method, "matrices": "dg::geo::Fieldaligned of the reference, circular field R0=10 I0=20 (ds_b.cpp:70-84), mx=my=%d" % mx,
               "entries_per_row": nnz / rows, "elements": size, "algorithmic_bytes": alg, "cells": cells, "planes": Nz,
               "csr_fused_us": t_csr * 1e6, "gather_plan_us": t_plan * 1e6, "gather_plan_gbs": alg / t_plan / 1e9,
               "bitwise_plan_equals_csr": bool((g.view(torch.int64) == g2.view(torch.int64)).all()),
               "reference_openmp_ms": cpu_sec * 1e3, "reference_threads": F.threads()}
        cp, cm = C.c_void_p(), C.c_void_p()
        try:
            L.celltile_plan_create(C.byref(cp), n, cells, cells, ptr(P[0]), ptr(P[1]), ptr(P[2]), stream())
            L.celltile_plan_create(C.byref(cm), n, cells, cells, ptr(M[0]), ptr(M[1]), ptr(M[2]), stream())
            t_ct = timeit(lambda: L.celltile_ds_centered(cp, cm, Nz, C.c_double(1.), ptr(f), ptr(bphi), C.c_double(dphi), C.c_double(0.), ptr(g3), stream()), reps)
            info = [C.c_int(), C.c_int(), C.c_int(), C.c_longlong()]
            L.celltile_plan_info(cp, *[C.byref(v) for v in info])
            rec.update({"celltile_us": t_ct * 1e6, "celltile_gbs": alg / t_ct / 1e9, "celltile_frac_of_peak": alg / t_ct / 1e9 / peak(),
                        "bitwise_celltile_equals_csr": bool((g.view(torch.int64) == g3.view(torch.int64)).all()),
                        "celltile_tiles": info[0].value, "celltile_max_source_cells": info[1].value, "celltile_planes_per_cta": info[2].value})
            L.celltile_plan_destroy(cp); L.celltile_plan_destroy(cm)
        except fb.DgbError as e:
            rec["celltile"] = "unsupported: %s" % e
        gh = g2.cpu().numpy()
        rec["max_rel_diff_vs_reference_ds_centered"] = float(np.abs(gh - gref).max() / np.abs(gref).max())
        out.append(rec)
        del F
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cells", type=int, default=96)
    ap.add_argument("--planes", type=int, default=64)
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--real", action="store_true", help="the reference's own Fieldaligned matrices (oracle/_ref/libdgref_fa.so)")
    ap.add_argument("--methods", default="dg")
    a = ap.parse_args()
    if a.real:
        for rec in run_real(a.cells, a.planes, a.reps, tuple(a.methods.split(","))) or [{"real": "oracle/_ref/libdgref_fa.so not present"}]:
            print(json.dumps(rec), flush=True)
        sys.exit(0)
    PEAK = peak()
    rows = run(a.cells, a.planes, a.reps)
    print(f"# DS centered, n=3 {a.cells}x{a.cells}x{a.planes}: {rows[0]['elements']} elements, L2 flushed between calls")
    for r in rows:
        alg = r["algorithmic_bytes"]
        for key, label in (("gather_plan_us", "gather plan (sliced ELL)"), ("csr_fused_us", "CSR fused DS::centered"), ("composition_us", "ePlus + eMinus + formula")):
            t = r[key] * 1e-6
            print(f"{r['method']:6s} ({r['entries_per_row']} per row) {label:26s} {r[key]:9.1f} us  {alg/t/1e9:8.1f} GB/s  {alg/t/1e9/PEAK*100:5.1f}% of {PEAK:.0f}  "
                  f"({alg/r['elements']:.1f} B/element)" + ("  bitwise == CSR kernel: %s" % r["bitwise_plan_equals_csr"] if key == "gather_plan_us" else ""), flush=True)
