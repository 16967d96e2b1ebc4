"""The C++ host layer (include/dg_b200.hpp) compiles against the C ABI and runs: host-only part on CPU, the Poisson
multigrid + PCG demo (elliptic2d_b.cpp problem) on the GPU with the same iteration numbers as the Python harness."""
import os
import re
import subprocess
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "poisson_demo")


def build():
    src = os.path.join(ROOT, "tests", "cpp", "poisson_demo.cpp")
    hdr = os.path.join(ROOT, "include", "dg_b200.hpp")
    if os.path.exists(EXE) and os.path.getmtime(EXE) > max(os.path.getmtime(src), os.path.getmtime(hdr)):
        return
    subprocess.check_call(["g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"), src,
                           "-L" + os.path.join(ROOT, "feltor_b200"), "-ldgb200",
                           "-Wl,-rpath," + os.path.join(ROOT, "feltor_b200"), "-Wl,-rpath,$ORIGIN/../../feltor_b200", "-o", EXE])


def test_cpp_host_only():
    build()
    out = subprocess.run([EXE, "--host-only"], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "caught: dgb200:" in out.stdout


@pytest.mark.gpu
def test_cpp_poisson_demo_matches_harness():
    import torch
    import gpu_backend as G
    from feltor_b200 import topology as T
    from feltor_b200.elliptic import Elliptic2d, MultigridCG2d, PCG
    if not os.path.exists(EXE):
        build()
    out = subprocess.run([EXE, "48", "32", "3"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    num = [int(v) for v in re.search(r"multigrid iterations:((?: \d+)+)", out.stdout).group(1).split()]
    it = int(re.search(r"pcg iterations: (\d+)", out.stdout).group(1))
    g = T.Grid([0, 0], [np.pi, 2 * np.pi], 3, [48, 32], [T.DIR, T.PER])
    amp = 0.9
    import math
    chi = g.evaluate(lambda x, y: 1. + amp * math.sin(x) * math.sin(y), vectorized=False)
    b = g.evaluate(lambda x, y: 2. * math.sin(x) * math.sin(y) * (amp * math.sin(x) * math.sin(y) + 1)
                   - amp * math.sin(x) * math.sin(x) * math.cos(y) * math.cos(y)
                   - amp * math.cos(x) * math.cos(x) * math.sin(y) * math.sin(y), vectorized=False)
    mg = MultigridCG2d(g, 3)
    ops = [Elliptic2d(mg.grid(u), T.DIR, T.PER, T.FORWARD, 1.0) for u in range(3)]
    for u, p in enumerate(mg.project(G.make(chi))):
        ops[u].set_chi(p)
    x = G.make(np.zeros(g.size))
    assert mg.solve(ops, x, G.make(b), 1e-6) == num
    y = G.make(np.zeros(g.size))
    assert PCG(g.size, 100000).solve(ops[0], y, G.make(b), ops[0].precond(), ops[0].weights(), 1e-6) == it
