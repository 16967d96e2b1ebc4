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
  roofline  the kernel the north star names = the fused Elliptic apply + dot(p,W,Ap) kernel (K1).  With the PCG direction
          update folded into its loader (the default on the walker kernel, dgb_pcg_last_folded) it reads z, the old
          direction, sigma, W and writes the new direction and Ap: algorithmic 48 B/dof (32 B/dof for the classic
          three-kernel iteration) / its live CUDA-event duration, against MEASURED_PEAKS.json; the streaming update
          kernel K2 (72 B/dof) is listed beside it under roofline_other_kernels
  micro     config 1 (and the same kernels at the benchmark size): axpby, pointwiseDot, dot, dx/dy, Elliptic apply in GB/s
  toefl     config 3 on one GPU: steps/s of the toefl right-hand side + Bogacki-Shampine step (N = 1 only)
  cpu_baseline  the reference's own OpenMP implementation (oracle/_ref/libdgref.so) on the host cores, bounded sample

  python bench.py [--gpus N] [--steps K] [--warmup W] [--iters M] [--cells 1024] [--impl reference]
N > 1 (torchrun): weak scaling of ONE global problem n=3, Nx=1024, Ny=1024*N cut into N slabs of cell rows (one per
GPU): the halo rows of the search direction and the three exact dots per iteration travel through CUDA-IPC peer memory
over NVLink (NCCL if peer memory is unavailable), bit-identical to the single-GPU arithmetic.  value = N * iterations /
max-over-ranks time (dof-iterations are what scales; every rank performs the same iteration count).  A "strong_scaling"
record (the FIXED 1024^2 grid of config 2 cut into N slabs) is measured after the timed region and attached to the line.
The reference arm at N > 1 runs the SAME global problem on all host cores (rank 0 only) and reports N * iterations / s.
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
    # torchrun exports OMP_NUM_THREADS=1 to its children: the CPU arm uses every host core, set explicitly and stated
    ncores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(ncores)
    from oracle import refwrap as R
    cells = args.cells
    kind = "reference"
    m = args.ref_iters
    ny = cells * world  # N > 1: the same global problem our arm cuts into N slabs
    if R.available():
        R.lib().ref_set_num_threads(ncores)
        g = R.grid([0, 0], [np.pi, 2 * np.pi * world], 3, [cells, ny], [R.DIR, R.PER])
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
        g = T.Grid([0, 0], [np.pi, 2 * np.pi * world], 3, [cells, ny], [T.DIR, T.PER])
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
    nwarm = min(args.warmup, 1)  # every step re-solves from x = 0: one untimed step warms caches and the thread pool
    for _ in range(nwarm):
        step()
    its, secs = 0, 0.
    for _ in range(args.steps):
        i, s = step()
        its += i
        secs += s
    v = world * its / secs  # in units of the per-GPU (1024^2) problem, like our arm's weak-scaling value
    sample = "%d PCG iterations per step of the n=3 %dx%d problem, %d steps, %d OpenMP threads" % (m, cells, ny, args.steps, cores)
    out = {"metric": "pcg_iterations_per_second", "value": v, "unit": "iterations/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": nwarm, "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "impl": "reference",
           "config": {"workload": "2D dg::Elliptic+dg::PCG Poisson n=3 Nx=Ny=%d eps=1e-8 DIRxPER (config 2)" % cells,
                      "iterations_per_step": m, "global_grid_cells": [cells, ny]},
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
    # config 4 on the REAL matrices: dg::geo::Fieldaligned of the unmodified reference (oracle/_ref/libdgref_fa.so, test
    # infrastructure) builds I+ / I- on the host (about a minute); synthetic matrices of the same structure otherwise
    real = ds_bench.run_real(96, 64, 20, ("dg",)) if not args.no_cpu_baseline else None
    if real:
        r = real[0]
        best = min(r.get("celltile_us", 1e30), r["gather_plan_us"])
        gbs = r["algorithmic_bytes"] / (best * 1e-6) / 1e9
        line = {"metric": "ds_centered_gbs", "value": gbs, "unit": "GB/s", "n_gpus": 1, "higher_is_better": True, "dtype": "f64", "data": "synthetic field, reference matrices",
                "config": {"workload": "DS::centered n=3 96x96x64, I+/I- from the reference's dg::geo::Fieldaligned (circular field of ds_b.cpp:70-84, mx=my=10, method dg: %.1f entries per row)" % r["entries_per_row"]},
                "detail": real, "us_per_call": best, "kernel": "celltile_kernel" if best == r.get("celltile_us") else "gather_ds_centered_kernel",
                "roofline": {"bound": "hbm", "achieved": gbs, "peak": peaks()[0], "unit": "GB/s", "frac": gbs / peaks()[0], "traffic": None,
                             "bytes_per_launch": r["algorithmic_bytes"]},
                "cpu_baseline": {"value": r["algorithmic_bytes"] / (r["reference_openmp_ms"] * 1e-3) / 1e9, "unit": "GB/s", "cores": r["reference_threads"], "kind": "reference",
                                 "sample": "ds.centered(f, g) of the unmodified dg::geo::DS on the same Fieldaligned object, mean of 2 calls (%.1f ms)" % r["reference_openmp_ms"]}}
        print(json.dumps(line), flush=True)
        return
    rows = ds_bench.run(96, 64, 20)
    line = {"metric": "ds_centered_gbs", "value": rows[0]["gather_plan_gbs"], "unit": "GB/s", "n_gpus": 1, "higher_is_better": True,
            "dtype": "f64", "data": "synthetic", "config": {"workload": "DS::centered n=3 96x96x64, synthetic field-line matrices (dg: 36 per row)"},
            "detail": rows, "roofline": {"bound": "hbm", "achieved": rows[0]["gather_plan_gbs"], "peak": peaks()[0], "unit": "GB/s",
                                         "frac": rows[0]["gather_plan_gbs"] / peaks()[0], "traffic": None}}
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
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = 1024 x 1024*N cells (default), strong = the fixed 1024^2 grid cut into N slabs")
    ap.add_argument("--no-strong", action="store_true", help="N > 1: skip the attached strong-scaling record")
    ap.add_argument("--no-micro", action="store_true", help="skip the config-1 kernel table")
    ap.add_argument("--no-toefl", action="store_true", help="skip the config-3 (toefl) record")
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
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = fb.lib()
    cells, M = args.cells, args.iters
    fchi, frhs = problem_functions()
    comm = None
    if world > 1:
        from feltor_b200.dist import Comm, SlabElliptic2d, DistPCG
        comm = Comm.from_torch_distributed()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    class Problem:
        """config 2 on `world` GPUs: a global grid of cells x ny_global cells, cut into slabs of cell rows when world > 1"""

        def __init__(self, ny_global):
            self.ny = ny_global
            ly = 2 * np.pi * ny_global / cells
            if world == 1:
                g = T.Grid([0., 0.], [np.pi, ly], 3, [cells, ny_global], [T.DIR, T.PER])
                self.E = Elliptic2d(g, T.DIR, T.PER, T.FORWARD, 1.0)
                self.E.set_chi(torch.from_numpy(g.evaluate(fchi).copy()).cuda())
                self.ndof = g.size
                b_np = g.evaluate(frhs).copy()
                self.pcg = PCG(self.ndof, M + 1)
                self.pcg.set_throw_on_fail(False)
            else:
                g = T.Grid([0., 0.], [np.pi, ly], 3, [cells, ny_global], [T.DIR, T.PER])
                self.E = SlabElliptic2d(comm, g, T.DIR, T.PER, T.FORWARD, 1.0)
                self.E.set_chi(torch.from_numpy(self.E.evaluate(fchi)).cuda())
                self.ndof = self.E.size
                b_np = self.E.evaluate(frhs)
                self.pcg = DistPCG(comm, self.ndof, M + 1)
                self.pcg.throw_on_fail = False
            self.b_host = torch.from_numpy(b_np).pin_memory()
            self.x0_host = torch.zeros(self.ndof, dtype=torch.float64).pin_memory()
            self.xout_host = torch.empty(self.ndof, dtype=torch.float64).pin_memory()
            self.b = self.b_host.cuda()
            self.x = torch.zeros(self.ndof, dtype=torch.float64, device="cuda")
            self.P, self.W = self.E.precond(), self.E.weights()

        def solve_device(self):
            self.x.zero_()
            return min(self.pcg.solve(self.E, self.x, self.b, self.P, self.W, 1e-8, 1.0, 1), M)

        def solve_e2e(self):
            self.b.copy_(self.b_host, non_blocking=True)
            self.x.copy_(self.x0_host, non_blocking=True)
            it = min(self.pcg.solve(self.E, self.x, self.b, self.P, self.W, 1e-8, 1.0, 1), M)
            self.xout_host.copy_(self.x, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return it

        def run_e2e_pipelined(self, steps):
            """`steps` end-to-end solves, every one with its own host -> device copy of b and x0 (pinned memory) and device ->
            host copy of its solution, the way a driver loop over independent right-hand sides would issue them: the inputs of
            solve k+1 go up on a copy stream while solve k runs (two sets of device buffers), the solution of solve k comes down
            on a third stream while solve k+1 runs.  Every byte still crosses PCIe inside the timed region."""
            if not hasattr(self, "b2"):
                self.b2, self.x2 = torch.empty_like(self.b), torch.empty_like(self.x)
                self.xout2_host = torch.empty(self.ndof, dtype=torch.float64).pin_memory()
                self.h2d, self.d2h = torch.cuda.Stream(), torch.cuda.Stream()
            cs = torch.cuda.current_stream()
            bufs = [(self.b, self.x, self.xout_host), (self.b2, self.x2, self.xout2_host)]
            up = [torch.cuda.Event(), torch.cuda.Event()]
            down = [torch.cuda.Event(), torch.cuda.Event()]
            start = torch.cuda.Event()
            start.record(cs)

            def upload(k):
                b, x, _ = bufs[k % 2]
                with torch.cuda.stream(self.h2d):
                    self.h2d.wait_event(start if k < 2 else down[k % 2])   # the buffers' previous solution has left
                    b.copy_(self.b_host, non_blocking=True)
                    x.copy_(self.x0_host, non_blocking=True)
                    up[k % 2].record(self.h2d)
            upload(0)
            its = 0
            for k in range(steps):
                if k + 1 < steps:
                    upload(k + 1)
                b, x, xo = bufs[k % 2]
                cs.wait_event(up[k % 2])
                its += min(self.pcg.solve(self.E, x, b, self.P, self.W, 1e-8, 1.0, 1), M)
                fin = torch.cuda.Event()
                fin.record(cs)
                with torch.cuda.stream(self.d2h):
                    self.d2h.wait_event(fin)
                    xo.copy_(x, non_blocking=True)
                    down[k % 2].record(self.d2h)
            for k in range(max(0, steps - 2), steps):
                cs.wait_event(down[k % 2])
            return its

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

    strong = args.scaling == "strong"
    if strong and cells % world:
        raise SystemExit("bench.py: --scaling strong needs --cells divisible by the GPU count")
    pr = Problem(cells if strong else cells * world)
    nwarm = max(args.warmup, 3)
    for _ in range(nwarm):
        pr.solve_device()
    # timed region: exactly K steps, no instrumentation inside
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = L.raw["dgb_launch_count"]()
    ms, its = timed(pr.solve_device, args.steps)
    launches = L.raw["dgb_launch_count"]() - launches0
    clocks = sampler.stop() if rank == 0 else None
    # per-kernel durations for the roofline: a separate pass with CUDA events on the launching stream around K1/K2/K3
    # (the events cost a few percent, so they stay out of the timed region above)
    L.pcg_set_profile(pr.pcg.h, 1)
    timed(pr.solve_device, min(args.steps, 2))
    prof = [C.c_double(), C.c_double(), C.c_double()]
    pn = C.c_longlong()
    L.pcg_get_profile(pr.pcg.h, C.byref(prof[0]), C.byref(prof[1]), C.byref(prof[2]), C.byref(pn))
    L.pcg_set_profile(pr.pcg.h, 0)
    folded = C.c_int(0)
    L.pcg_last_folded(pr.pcg.h, C.byref(folded))
    folded = bool(folded.value)
    pr.solve_e2e()
    ms_e2e_serial, its_e2e_serial = timed(pr.solve_e2e, args.steps)
    pr.run_e2e_pipelined(2)
    ms_e2e, its_e2e = timed(lambda: pr.run_e2e_pipelined(args.steps), 1)
    ndof = pr.ndof
    # strong scaling of the named 1024^2 grid (N > 1, default weak run): the same solver on cells x cells cells cut into N slabs
    strong_rec = None
    if world > 1 and not strong and cells % world == 0 and not args.no_strong:
        del pr
        torch.cuda.empty_cache()
        ps = Problem(cells)
        for _ in range(3):
            ps.solve_device()
        ms_s, its_s = timed(ps.solve_device, args.steps)
        strong_rec = {"metric": "pcg_iterations_per_second", "value": (its_s / world) / (ms_s * 1e-3), "unit": "iterations/s",
                      "scaling": "strong", "n_gpus": world, "global_grid_cells": [cells, cells], "dof_per_gpu": ps.ndof,
                      "ms_per_step": ms_s / args.steps, "steps": args.steps, "warmup": 3,
                      "note": "iterations/s of the FIXED n=3 %dx%d problem; divide by the N=1 value for the strong-scaling speed-up" % (cells, cells)}
        del ps

    # toefl (config 3) on the fixed grid cut into N slabs: strong scaling of the step rate (N > 1; N = 1 runs further down)
    toefl_dist = None
    if world > 1 and not args.no_toefl and cells % (4 * world) == 0:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import toefl_bench
            torch.cuda.empty_cache()
            t_out, _ = toefl_bench.run(cells, 4, 2, 0.5, comm=comm)
            toefl_dist = {"metric": "toefl_steps_per_second", "value": t_out["steps_per_s"], "unit": "steps/s", "n_gpus": world,
                          "scaling": "strong", "rhs_per_s": t_out["rhs_per_s"], "config": t_out["workload"],
                          "mean_pcg_iterations_per_solve": t_out["mean_pcg_iterations_per_solve(stage0,1,2)"]}
        except Exception as e:
            toefl_dist = {"failed": repr(e)}
        try:  # the multistep stepper config 3 names: one right-hand side per step
            torch.cuda.empty_cache()
            m_out, _ = toefl_bench.run(cells, 12, 4, 0.5, comm=comm, stepper="multistep")
            toefl_dist["multistep"] = {"value": m_out["steps_per_s"], "unit": "steps/s", "config": m_out["workload"], "steps": m_out["steps"]}
        except Exception as e:
            toefl_dist["multistep"] = {"failed": repr(e)}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_kind = peaks()
    k1_ms = prof[0].value / max(pn.value, 1)
    k2_ms = prof[1].value / max(pn.value, 1)
    # algorithmic bytes per launch: K1 = fused Elliptic apply + dot: read p, sigma, W, write Ap = 32 B/dof,
    # with the folded direction update read z, p_old, sigma, W, write p_new, Ap = 48 B/dof;
    # K2 = update + two dots: read p, Ap, x, r, P, W, write x, r, z = 72 B/dof; K3 (classic iteration only) 24 B/dof
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
    except Exception:
        tr = {}
    kernels = [
        {"kernel": ("elliptic2d_walker_kernel<3,fwd,dot,fold> (p = z + beta p in the loader, Elliptic apply, dot(p,W,Ap))" if folded else
                    "elliptic2d_walker_kernel<3,fwd,dot> (Elliptic apply + dot(p,W,Ap))"),
         "bytes_per_launch": (48 if folded else 32) * ndof, "ms_per_launch": k1_ms,
         "traffic": tr.get("elliptic2d_fold_dot_bytes_per_launch" if folded else "elliptic2d_fused_dot_bytes_per_launch")},
        {"kernel": "pcg_update_kernel (x, r, z = P r updates + dot(r,W,r), dot(z,W,r))", "bytes_per_launch": 72 * ndof, "ms_per_launch": k2_ms,
         "traffic": tr.get("pcg_update_bytes_per_launch")},
    ]
    for k in kernels:
        k["achieved"] = k["bytes_per_launch"] / (k["ms_per_launch"] * 1e-3) / 1e9 if k["ms_per_launch"] > 0 else None
        k["frac"] = k["achieved"] / peak if k["achieved"] else None
        k["share_of_iteration"] = k["ms_per_launch"] / max(prof[0].value + prof[1].value + prof[2].value, 1e-30) * max(pn.value, 1)
    named = kernels[0]  # the kernel the north star names (the fused Elliptic apply); K2 has the larger time share and is listed beside it
    value = (its / world if strong else its) / (ms * 1e-3)
    if world == 1:
        par = "single GPU"
    else:
        plane = ("CUDA-IPC peer memory over NVLink: the update kernel stores the boundary rows of z = P r into the neighbours' ghost rows, the 39-word int64 dot "
                 "records are exchanged by peer stores + polling (NCCL only ships the IPC handles)") if comm.peer_memory else \
                "NCCL: grouped ncclSend/ncclRecv of the ghost rows, ncclAllReduce(int64) of the dot records"
        par = "y-slabs x%d of a global grid of %dx%d cells; %s" % (world, cells, cells if strong else cells * world, plane)
    out = {
        "metric": "pcg_iterations_per_second", "value": value, "unit": "iterations/s", "n_gpus": world,
        "steps": args.steps, "warmup": nwarm, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "2D dg::Elliptic+dg::PCG Poisson n=3 Nx=Ny=%d eps=1e-8 DIRxPER (config 2)" % cells,
                   "dof_per_gpu": ndof, "iterations_per_step": M, "l2": "working set 8 vectors x %.0f MB > 126 MB L2"
                   % (ndof * 8 / 1e6), "parallelism": par},
        "e2e": {"value": (its_e2e / world if strong else its_e2e) / (ms_e2e * 1e-3), "unit": "iterations/s", "h2d_bytes_per_step": 2 * ndof * 8,
                "d2h_bytes_per_step": ndof * 8,
                "how": "every step copies b and x0 host->device and its solution device->host (pinned memory) inside the timed region; "
                       "the copies of neighbouring steps run on copy streams beside the solve (double-buffered device operands)",
                "serial_value": (its_e2e_serial / world if strong else its_e2e_serial) / (ms_e2e_serial * 1e-3),
                "serial_how": "copy in -> solve -> copy out -> synchronize, one step after the other"},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": named["kernel"], "achieved": named["achieved"], "peak": peak, "peak_kind": peak_kind,
                     "unit": "GB/s", "frac": named["frac"], "traffic": named["traffic"], "bytes_per_launch": named["bytes_per_launch"],
                     "ms_per_launch": named["ms_per_launch"], "share_of_iteration": named["share_of_iteration"]},
        "roofline_other_kernels": [{kk: k[kk] for kk in ("kernel", "achieved", "frac", "traffic", "bytes_per_launch", "ms_per_launch",
                                                         "share_of_iteration")} for k in kernels if k is not named],
        "kernels_ms_per_iteration": {"apply_dot": k1_ms, "update_dots": prof[1].value / max(pn.value, 1),
                                     "direction": prof[2].value / max(pn.value, 1)},
        "pcg_iteration": {"launches": 2 if folded else 3, "algorithmic_bytes_per_dof": 120 if folded else 128,
                          "gbs": (120 if folded else 128) * ndof * (its / (ms * 1e-3)) / world / 1e9},
    }
    if strong_rec is not None:
        out["strong_scaling"] = strong_rec
    if toefl_dist is not None:
        out["toefl"] = toefl_dist
    if world == 1 and not args.no_micro:
        try:
            out["micro"] = micro_table(cells, peak)
        except Exception as e:
            out["micro"] = {"failed": repr(e)}
    if world == 1 and not args.no_toefl:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import toefl_bench
            t_out, _ = toefl_bench.run(cells, 4, 2, 0.5)
            out["toefl"] = {"metric": "toefl_steps_per_second", "value": t_out["steps_per_s"], "unit": "steps/s", "rhs_per_s": t_out["rhs_per_s"],
                            "config": t_out["workload"], "kernel_launches_per_step": t_out["kernel_launches"] / t_out["steps"],
                            "mean_pcg_iterations_per_solve": t_out["mean_pcg_iterations_per_solve(stage0,1,2)"]}
        except Exception as e:
            out["toefl"] = {"failed": repr(e)}
        try:  # the multistep stepper config 3 names (dg::ExplicitMultistep TVB-3-3: one right-hand side per step)
            m_out, _ = toefl_bench.run(cells, 20, 5, 0.5, stepper="multistep")
            out["toefl"]["multistep"] = {"metric": "toefl_steps_per_second", "value": m_out["steps_per_s"], "unit": "steps/s", "config": m_out["workload"],
                                         "steps": m_out["steps"], "kernel_launches_per_step": m_out["kernel_launches"] / m_out["steps"],
                                         "mean_pcg_iterations_per_solve": m_out["mean_pcg_iterations_per_solve(stage0,1,2)"]}
        except Exception as e:
            out["toefl"]["multistep"] = {"failed": repr(e)}
    if not args.no_cpu_baseline and world == 1:
        try:
            from oracle import refwrap as R
            if R.available():
                R.lib().ref_set_num_threads(os.cpu_count() or 1)
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


def micro_table(cells, peak):
    """config 1 of BASELINE.json (n=3, 128^2: 1.2 MB vectors, launch/latency bound) and the same kernels at the benchmark
    size: algorithmic GB/s from CUDA events, median of 20, L2 flushed between calls"""
    import torch
    import feltor_b200 as fb
    from feltor_b200 import blas1, blas2, topology as T
    from feltor_b200.elliptic import Elliptic2d
    from feltor_b200._dev import ptr, stream
    L = fb.lib()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ws = blas2.DotWorkspace()
    res = torch.zeros(41, dtype=torch.int64, device="cuda")
    table = {"unit": "GB/s algorithmic (us per call)", "peak_gbs": peak, "method": "CUDA events, median of 20, 256 MB written between calls (L2 flush)"}

    def timeit(f):
        for _ in range(3):
            f()
        torch.cuda.synchronize()
        ts = []
        for _ in range(20):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            f()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e-3)
        return float(np.median(ts))

    for N in sorted({128, cells}):
        g = T.Grid([0., 0.], [2 * np.pi] * 2, 3, [N, N], [T.PER, T.PER])
        n = g.size
        gen = torch.Generator(device="cuda").manual_seed(0)
        v = [torch.rand(n, dtype=torch.float64, device="cuda", generator=gen) + 0.5 for _ in range(4)]
        B = 8 * n
        rows = {}

        def rec(name, nbytes, f):
            t = timeit(f)
            rows[name] = {"gbs": nbytes / t / 1e9, "us": t * 1e6, "frac": nbytes / t / 1e9 / peak}
        rec("axpby", 3 * B, lambda: blas1.axpby(1.0000001, v[0], 0.9999999, v[1]))
        rec("pointwiseDot", 4 * B, lambda: blas1.pointwiseDot(1.0000001, v[0], v[1], 0.5, v[2]))
        rec("dot2", 2 * B, lambda: L.exdot2(ws.h, n, ptr(v[0]), C.c_double(0), ptr(v[1]), C.c_double(0), ptr(res), stream()))
        rec("dot3", 3 * B, lambda: L.exdot3(ws.h, n, ptr(v[0]), C.c_double(0), ptr(v[1]), C.c_double(0), ptr(v[2]), C.c_double(0),
                                            ptr(res), stream()))
        for name, m in (("dx_forward", T.derivative(0, g, T.PER, T.FORWARD)), ("dy_forward", T.derivative(1, g, T.PER, T.FORWARD)),
                        ("dx_centered", T.derivative(0, g, T.PER, T.CENTERED)), ("dy_centered", T.derivative(1, g, T.PER, T.CENTERED))):
            m.handle
            rec(name, 2 * B, lambda: m.symv(1.0, v[0], 0.0, v[1]))
        ge = T.Grid([0., 0.], [np.pi, 2 * np.pi], 3, [N, N], [T.DIR, T.PER])
        for dname, dd in (("elliptic_forward", T.FORWARD), ("elliptic_centered", T.CENTERED)):
            E = Elliptic2d(ge, T.DIR, T.PER, dd, 1.0)
            E.set_chi(v[2])
            rec(dname, 3 * B, lambda: E.symv(v[0], v[1]))
            rows[dname]["kernel"] = E.kernel()
        table["n3_%dx%d" % (N, N)] = rows
    return table


if __name__ == "__main__":
    main()
