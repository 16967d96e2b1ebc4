#!/usr/bin/env python3
"""Build the include tree an application compiles against to run Feltor's dg library on libdgb200.so.

    python integration/make_tree.py [--reference /root/reference] [--out integration/_build/inc]

What it does -- this is the complete change a Feltor maintainer would make to inc/dg/backend:
  1. copies the reference's inc/ tree (nothing under /root/reference is modified);
  2. REPLACES the four files that hold the dg::CudaTag overloads with the ones in integration/dgb_shim/dg/backend/
         blas1_cuda.cuh  sparseblockmat_gpu_kernels.cuh  sparsematrix_gpu.cuh  exblas/exdot_cuda.cuh  exblas/fpedot_cuda.cuh
     and adds dgb_shim.h / dgb_parallel_for.cuh next to them;
  3. applies two small edits to files that stay the reference's:
       backend/sparseblockmat.h   EllSparseBlockMat gets a launch-plan cache member (as SparseMatrix has CSRCache_gpu,
                                  sparsematrix.h:620-628); set_default_range / set_right_size / set_left_size drop it
       backend/blas2_stencil.h    the CUDA section (stencil_kernel + doParallelFor_dispatch( CudaTag,...)) is replaced
                                  by #include "dgb_parallel_for.cuh"
Applications (src/toefl/toefl.h, ...) and every other header compile UNCHANGED:
    nvcc -x cu -std=c++17 --extended-lambda -arch=sm_100a -I integration/_build/inc -I include  app.cpp -L feltor_b200 -ldgb200
"""
import argparse
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def edit(path, old, new, count=1):
    s = open(path).read()
    if s.count(old) < 1:
        sys.exit("make_tree.py: anchor not found in %s:\n%s" % (path, old))
    s = s.replace(old, new, count)
    open(path, "w").write(s)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(HERE, "_build", "inc"))
    a = ap.parse_args()
    src = os.path.join(a.reference, "inc")
    if not os.path.isdir(src):
        sys.exit("make_tree.py: %s not found" % src)
    if os.path.isdir(a.out):
        shutil.rmtree(a.out)
    shutil.copytree(src, a.out)
    # 2. overlay
    shim = os.path.join(HERE, "dgb_shim")
    for root, _, files in os.walk(shim):
        for f in files:
            rel = os.path.relpath(os.path.join(root, f), shim)
            dst = os.path.join(a.out, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(os.path.join(root, f), dst)
    # 3a. launch-plan cache of the Ell block matrix
    sbm = os.path.join(a.out, "dg", "backend", "sparseblockmat.h")
    edit(sbm, '#include "sparsematrix.h"', '#include "sparsematrix.h"\n#include "dgb_shim.h" // libdgb200 binding: dgb::shim::EllCache')
    edit(sbm, "    int right_size; //!< size of the right Kronecker delta (is e.g 1 for a x - derivative)\n    private:\n",
         "    int right_size; //!< size of the right Kronecker delta (is e.g 1 for a x - derivative)\n"
         "    mutable dgb::shim::EllCache m_dgb_cache; //!< libdgb200 launch plan, built on first symv (copies start without)\n"
         "    private:\n")
    edit(sbm, "    void set_default_range(){\n", "    void set_default_range(){\n        m_dgb_cache.forget();\n")
    edit(sbm, "    void set_left_size( int new_left_size ){\n", "    void set_left_size( int new_left_size ){\n        m_dgb_cache.forget();\n")
    # 3b. parallel_for / stencil dispatch
    st = os.path.join(a.out, "dg", "backend", "blas2_stencil.h")
    s = open(st).read()
    beg = s.index("#if THRUST_DEVICE_SYSTEM==THRUST_DEVICE_SYSTEM_CUDA")
    end = s.index("#elif THRUST_DEVICE_SYSTEM==THRUST_DEVICE_SYSTEM_OMP")
    ns_close = "}//namespace detail\n}//namespace blas2\n}//namespace dg\n"
    ns_open = "namespace dg\n{\nnamespace blas2\n{\nnamespace detail\n{\n"
    s = (s[:beg] + "#if THRUST_DEVICE_SYSTEM==THRUST_DEVICE_SYSTEM_CUDA\n" + ns_close +
         '#include "dgb_parallel_for.cuh" // libdgb200 binding: doParallelFor_dispatch( CudaTag, ...)\n' + ns_open + s[end:])
    open(st, "w").write(s)
    print("make_tree.py: wrote", a.out)


if __name__ == "__main__":
    main()
