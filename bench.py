#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json on the B200-native dg core.

Workload (config 2 of BASELINE.json): 2-D dg::Elliptic + dg::PCG Poisson problem, n=3, Nx=Ny=1024
(9 437 184 dof), Dirichlet x periodic, chi = 1 + 0.9 sin x sin y, forward discretisation, jfactor 1, P = 1/chi,
W = weights, eps = 1e-8, x0 = 0 (inc/dg/elliptic2d_b.cpp:22-38).  One STEP = one call of PCG::solve limited to a
fixed number of iterations (set_max(k), set_throw_on_fail(false); the full solve needs ~16 000 iterations at this
size).  metric = PCG iterations per second (whole job).

  value   device-resident inputs, timed with CUDA events on the launching stream
  e2e     the same solve through the C-ABI entry point with HOST buffers: b and x0 copied host->device from
          pinned memory and x copied back inside the timed region, every step
  roofline  dominant kernel = the fused Elliptic apply + dot(p,W,Ap) kernel (K1): algorithmic 32 B/dof
          (read p, sigma, W; write Ap) / its live CUDA-event duration, against MEASURED_PEAKS.json
  cpu_baseline  the reference's own OpenMP implementation (oracle/_ref/libdgref.so) on the host cores, bounded sample

  python bench.py [--gpus N] [--steps K] [--warmup W] [--iters M] [--cells 1024] [--impl reference]
N > 1 (torchrun): weak scaling of ONE global problem n=3, Nx=1024, Ny=1024*N cut into N slabs of cell rows (one per
GPU): NCCL halo exchange of the search direction and an integer allreduce of the three exact dots per iteration
(bit-identical to the single-GPU arithmetic).  value = N * iterations / max-over-ranks time (dof-iterations are
what scales; every rank performs the same iteration count).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

AMP = 0.9


def problem_functions():
    chi = lambda x, y: 1. + AMP * np.sin(x) * np.sin(y)  # noqa: E731  elliptic2d_b.cpp:33
    rhs = lambda x, y: (2. * np.sin(x) * np.sin(y) * (AMP * np.sin(x) * np.sin(y) + 1)  # noqa: E731  :38
                        - AMP * np.sin(x) ** 2 * np.cos(y) ** 2 - AMP * np.cos(x) ** 2 * np.sin(y) ** 2)
    return chi, rhs


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for k, nm in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


def reference_arm(args, rank, world):
    """the reference's own CPU implementation of the path (oracle/_ref) on the host cores; rank 0 only"""
    if rank != 0:
        return
    from oracle import refwrap as R
    cells = args.cells
    kind = "reference"
    m = args.ref_iters
    if R.available():
        g = R.grid([0, 0], [np.pi, 2 * np.pi], 3, [cells, cells], [R.DIR, R.PER])
        E = R.Elliptic2d(g, R.DIR, R.PER, R.FORWARD, 1.0)
        chi = R.evaluate(g, "pol")
        E.set_chi(chi)
        b = R.evaluate(g, "rhs")
        P, W = E.precond(), E.weights()
        cores = R.lib().ref_get_max_threads()

        def step():
            x = np.zeros(E.size)
            it, sec = E.pcg_solve(x, b, P, W, 1e-8, 1.0, 1, max_iter=m + 1)
            return min(it, m), sec
    else:  # the oracle port (single thread)
        from oracle import orc
        from feltor_b200 import topology as T
        kind, cores = "port", 1
        g = T.Grid([0, 0], [np.pi, 2 * np.pi], 3, [cells, cells], [T.DIR, T.PER])
        fchi, frhs = problem_functions()
        chi, b, W = g.evaluate(fchi), g.evaluate(frhs), g.weights()
        mats = dict(leftx=T.derivative(0, g, T.NEU, T.BACKWARD), lefty=T.derivative(1, g, T.PER, T.BACKWARD),
                    rightx=T.derivative(0, g, T.DIR, T.FORWARD), righty=T.derivative(1, g, T.PER, T.FORWARD),
                    jumpx=T.jump(0, g, T.DIR), jumpy=T.jump(1, g, T.PER))
        E = orc.Elliptic2d(mats, sigma=chi.copy(), jfactor=1.0)
        P = 1. / chi

        def step():
            x = np.zeros(g.size)
            t0 = time.time()
            it = E.pcg_solve(x, b, P, W, 1e-8, 1.0, 1, max_iter=m + 1)
            return min(it, m), time.time() - t0
    for _ in range(args.warmup if args.warmup < 2 else 1):
        step()
    its, secs = 0, 0.
    for _ in range(args.steps):
        i, s = step()
        its += i
        secs += s
    v = its / secs
    sample = "%d PCG iterations per step of the n=3 %dx%d problem, %d steps" % (m, cells, cells, args.steps)
    out = {"metric": "pcg_iterations_per_second", "value": v, "unit": "iterations/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
           "config": {"workload": "2D dg::Elliptic+dg::PCG Poisson n=3 Nx=Ny=%d eps=1e-8 DIRxPER (config 2)" % cells,
                      "iterations_per_step": m},
           "cpu_baseline": {"value": v, "unit": "iterations/s", "cores": cores, "kind": kind, "sample": sample},
           "e2e": {"value": v, "unit": "iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def secondary_workload(args, rank):
    """configs 3 (toefl) and 4 (DS) of BASELINE.json on one GPU: GPU leg from tools/{toefl,ds}_bench.py, CPU baseline = the
    unmodified reference (oracle/_ref) on a bounded sample; one JSON line"""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    if args.workload == "toefl":
        import toefl_bench
        out, y_init = toefl_bench.run(args.cells, args.steps, max(args.warmup, 2), 0.5)
        line = {"metric": "toefl_steps_per_second", "value": out["steps_per_s"], "unit": "steps/s", "n_gpus": 1, "steps": args.steps,
                "warmup": max(args.warmup, 2), "ms_per_step": out["ms_per_step"], "higher_is_better": True, "dtype": "f64",
                "data": "synthetic", "config": {"workload": out["workload"]}, "detail": out, "gpu_launches": out["kernel_launches"]}
        if not args.no_cpu_baseline:
            from oracle import reftoefl as R
            if R.available():
                ref = R.RefToefl(toefl_bench.params(args.cells))
                nref = 2
                sys.stdout.flush()
                saved, devnull = os.dup(1), os.open(os.devnull, os.O_WRONLY)
                os.dup2(devnull, 1)  # the reference prints its solver statistics to stdout (toefl.h: set_benchmark(true))
                try:
                    _, _, rsec = ref.erk("Bogacki-Shampine-4-2-3", 0., 0.5, nref, y_init[0], y_init[1])
                finally:
                    os.dup2(saved, 1)
                    os.close(devnull)
                    os.close(saved)
                line["cpu_baseline"] = {"value": nref / rsec, "unit": "steps/s", "cores": os.cpu_count(), "kind": "reference",
                                        "sample": "%d steps from the same initial state (toefl::Explicit + dg::ERKStep, OpenMP)" % nref}
        print(json.dumps(line), flush=True)
        return
    import ds_bench
    import numpy as np_
    rows = ds_bench.run(96, 64, 20)
    line = {"metric": "ds_centered_gbs", "value": rows[0]["gather_plan_gbs"], "unit": "GB/s", "n_gpus": 1, "higher_is_better": True,
            "dtype": "f64", "data": "synthetic", "config": {"workload": "DS::centered n=3 96x96x64, synthetic field-line matrices (dg: 36 per row)"},
            "detail": rows, "roofline": {"bound": "hbm", "achieved": rows[0]["gather_plan_gbs"], "peak": peaks()[0], "unit": "GB/s",
                                         "frac": rows[0]["gather_plan_gbs"] / peaks()[0], "traffic": None}}
    if not args.no_cpu_baseline:
        from oracle import refwrap as R
        if R.available():
            rng = np_.random.default_rng(0)
            nrows, Nz = (3 * 96) ** 2, 64
            size = nrows * Nz
            hf = rng.uniform(-1, 1, size)
            rng.uniform(0.5, 1.5, size)
            P, M = ds_bench.interpolation_matrix(rng, 96, 2), ds_bench.interpolation_matrix(rng, 96, 2)
            rP, rM = R.Csr(nrows, nrows, *P), R.Csr(nrows, nrows, *M)
            tp, tm = np_.zeros(size), np_.zeros(size)
            t0 = time.time()
            for k in range(Nz):  # Fieldaligned::ePlus / eMinus: one symv per plane (fieldaligned.h:850-912)
                rP.symv(1., hf[((k + 1) % Nz) * nrows:((k + 1) % Nz + 1) * nrows], 0., tp[k * nrows:(k + 1) * nrows])
                rM.symv(1., hf[((k - 1) % Nz) * nrows:((k - 1) % Nz + 1) * nrows], 0., tm[k * nrows:(k + 1) * nrows])
            sec = time.time() - t0
            line["cpu_baseline"] = {"value": rows[0]["algorithmic_bytes"] / sec / 1e9, "unit": "GB/s", "cores": R.lib().ref_get_max_threads(),
                                    "kind": "reference", "sample": "the 128 plane-wise CSR symv of one DS::centered (reference OpenMP kernel), formula excluded"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--iters", type=int, default=200, help="PCG iterations per step (our arm)")
    ap.add_argument("--ref-iters", type=int, default=10, help="PCG iterations per step of the CPU reference arm")
    ap.add_argument("--cells", type=int, default=1024)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="pcg", choices=["pcg", "toefl", "ds"],
                    help="pcg: the headline (config 2); toefl: config 3; ds: config 4 -- the two secondary workloads print their own "
                         "JSON line (single GPU) with the reference's CPU path timed beside them")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload != "pcg":
        return secondary_workload(args, rank)
    if args.impl == "reference":
        return reference_arm(args, rank, world)

    import torch
    import torch.distributed as dist
    import feltor_b200 as fb
    from feltor_b200 import topology as T
    from feltor_b200.elliptic import Elliptic2d, PCG
    from feltor_b200._dev import ptr, stream
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = fb.lib()
    cells, M = args.cells, args.iters
    fchi, frhs = problem_functions()
    if world == 1:
        g = T.Grid([0., 0.], [np.pi, 2 * np.pi], 3, [cells, cells], [T.DIR, T.PER])
        E = Elliptic2d(g, T.DIR, T.PER, T.FORWARD, 1.0)
        E.set_chi(torch.from_numpy(g.evaluate(fchi).copy()).cuda())
        ndof = g.size
        b_np = g.evaluate(frhs).copy()
        pcg = PCG(ndof, M + 1)
        pcg.set_throw_on_fail(False)
    else:
        from feltor_b200.dist import Comm, SlabElliptic2d, DistPCG
        comm = Comm.from_torch_distributed()
        g = T.Grid([0., 0.], [np.pi, 2 * np.pi * world], 3, [cells, cells * world], [T.DIR, T.PER])
        E = SlabElliptic2d(comm, g, T.DIR, T.PER, T.FORWARD, 1.0)
        E.set_chi(torch.from_numpy(E.evaluate(fchi)).cuda())
        ndof = E.size
        b_np = E.evaluate(frhs)
        pcg = DistPCG(comm, ndof, M + 1)
        pcg.throw_on_fail = False
    b_host = torch.from_numpy(b_np).pin_memory()
    x0_host = torch.zeros(ndof, dtype=torch.float64).pin_memory()
    xout_host = torch.empty(ndof, dtype=torch.float64).pin_memory()
    b = b_host.cuda()
    x = torch.zeros(ndof, dtype=torch.float64, device="cuda")
    P, W = E.precond(), E.weights()

    def solve_device():
        x.zero_()
        return min(pcg.solve(E, x, b, P, W, 1e-8, 1.0, 1), M)

    def solve_e2e():
        b.copy_(b_host, non_blocking=True)
        x.copy_(x0_host, non_blocking=True)
        it = min(pcg.solve(E, x, b, P, W, 1e-8, 1.0, 1), M)
        xout_host.copy_(x, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return it

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        its = 0
        for _ in range(steps):
            its += fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms, float(its)], dtype=torch.float64, device="cuda")
            tmax = t.clone()
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return tmax[0].item(), int(t[1].item())
        return ms, its

    for _ in range(max(args.warmup, 3)):
        solve_device()
    # timed region: exactly K steps, no instrumentation inside
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = L.raw["dgb_launch_count"]()
    ms, its = timed(solve_device, args.steps)
    launches = L.raw["dgb_launch_count"]() - launches0
    clocks = sampler.stop() if rank == 0 else None
    # per-kernel durations for the roofline: a separate pass with CUDA events on the launching stream around K1/K2/K3
    # (the events cost a few percent, so they stay out of the timed region above)
    L.pcg_set_profile(pcg.h, 1)
    timed(solve_device, min(args.steps, 2))
    prof = [C.c_double(), C.c_double(), C.c_double()]
    pn = C.c_longlong()
    L.pcg_get_profile(pcg.h, C.byref(prof[0]), C.byref(prof[1]), C.byref(prof[2]), C.byref(pn))
    L.pcg_set_profile(pcg.h, 0)
    solve_e2e()
    ms_e2e, its_e2e = timed(solve_e2e, args.steps)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_kind = peaks()
    k1_ms = prof[0].value / max(pn.value, 1)
    k2_ms = prof[1].value / max(pn.value, 1)
    # algorithmic bytes per launch: K1 = fused Elliptic apply + dot: read p, sigma, W, write Ap = 32 B/dof;
    # K2 = update + two dots: read p, Ap, x, r, P, W, write x, r, z = 72 B/dof
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
    except Exception:
        tr = {}
    kernels = [
        {"kernel": "elliptic2d_walker_kernel<3,fwd,dot> (Elliptic apply + dot(p,W,Ap))", "bytes_per_launch": 32 * ndof, "ms_per_launch": k1_ms,
         "traffic": tr.get("elliptic2d_fused_dot_bytes_per_launch")},
        {"kernel": "pcg_update_kernel (x, r, z = P r updates + dot(r,W,r), dot(z,W,r))", "bytes_per_launch": 72 * ndof, "ms_per_launch": k2_ms,
         "traffic": tr.get("pcg_update_bytes_per_launch")},
    ]
    for k in kernels:
        k["achieved"] = k["bytes_per_launch"] / (k["ms_per_launch"] * 1e-3) / 1e9 if k["ms_per_launch"] > 0 else None
        k["frac"] = k["achieved"] / peak if k["achieved"] else None
    dom = max(kernels, key=lambda k: k["ms_per_launch"])  # the dominant kernel of the step = the one with the largest share of it
    value = its / (ms * 1e-3)
    out = {
        "metric": "pcg_iterations_per_second", "value": value, "unit": "iterations/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "2D dg::Elliptic+dg::PCG Poisson n=3 Nx=Ny=%d eps=1e-8 DIRxPER (config 2)" % cells,
                   "dof_per_gpu": ndof, "iterations_per_step": M, "l2": "working set 8 vectors x %.0f MB > 126 MB L2"
                   % (ndof * 8 / 1e6), "parallelism": ("y-slabs x%d, NCCL halo + int64 superacc allreduce, global grid %dx%d cells" % (world, cells, cells * world))
                   if world > 1 else "single"},
        "e2e": {"value": its_e2e / (ms_e2e * 1e-3), "unit": "iterations/s", "h2d_bytes_per_step": 2 * ndof * 8,
                "d2h_bytes_per_step": ndof * 8},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved"], "peak": peak, "peak_kind": peak_kind,
                     "unit": "GB/s", "frac": dom["frac"], "traffic": dom["traffic"], "bytes_per_launch": dom["bytes_per_launch"],
                     "ms_per_launch": dom["ms_per_launch"]},
        "roofline_other_kernels": [{kk: k[kk] for kk in ("kernel", "achieved", "frac", "traffic", "bytes_per_launch", "ms_per_launch")}
                                   for k in kernels if k is not dom],
        "kernels_ms_per_iteration": {"apply_dot": k1_ms, "update_dots": prof[1].value / max(pn.value, 1),
                                     "direction": prof[2].value / max(pn.value, 1)},
        "pcg_gbs_at_128B_per_dof": 128 * ndof * value / world / 1e9,
    }
    if not args.no_cpu_baseline and world == 1:
        try:
            from oracle import refwrap as R
            if R.available():
                m = args.ref_iters * 4
                gr = R.grid([0, 0], [np.pi, 2 * np.pi], 3, [cells, cells], [R.DIR, R.PER])
                Er = R.Elliptic2d(gr, R.DIR, R.PER, R.FORWARD, 1.0)
                Er.set_chi(R.evaluate(gr, "pol"))
                br = R.evaluate(gr, "rhs")
                xr = np.zeros(Er.size)
                it, sec = Er.pcg_solve(xr, br, Er.precond(), Er.weights(), 1e-8, 1.0, 1, max_iter=m + 1)
                out["cpu_baseline"] = {"value": min(it, m) / sec, "unit": "iterations/s",
                                       "cores": R.lib().ref_get_max_threads(), "kind": "reference",
                                       "sample": "%d PCG iterations of the same n=3 %dx%d problem (reference OpenMP "
                                       "backend, oracle/_ref)" % (m, cells, cells)}
            else:
                out["cpu_baseline"] = {"value": None, "unit": "iterations/s", "cores": 0, "kind": "reference",
                                       "sample": "oracle/_ref/libdgref.so not present"}
        except Exception as e:  # never lose the GPU number over the baseline leg
            out["cpu_baseline"] = {"value": None, "unit": "iterations/s", "cores": 0, "kind": "reference",
                                   "sample": "failed: %r" % (e,)}
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
