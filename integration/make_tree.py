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
       backend/sparsematrix.h     SparseMatrix::operator* of two host matrices (the reference has no device spgemm) sends
                                  large products to dgb_csr_spgemm_host_* (bit-identical, seconds -> milliseconds in
                                  dg::geo::Fieldaligned's constructor)
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


def fusion_hooks(out):
    """4. second level (integration/dgb_shim/dg/backend/dgb_fused.h): dg::Elliptic2d::symv and dg::PCG::solve reach the fused
    kernels.  Each hook is a prologue of a few lines guarded by `if constexpr` on device double vectors; when the operator is
    outside the fused kernels' scope the hook returns false and the reference's own code runs, so every other instantiation
    (host vectors, MPI vectors, other value types, curvilinear grids) compiles and behaves as before."""
    ell = os.path.join(out, "dg", "elliptic.h")
    edit(ell, '#include "topology/geometry.h"', '#include "topology/geometry.h"\n#include "backend/dgb_fused.h" // libdgb200 binding: fused Elliptic2d')
    # (a) Elliptic2d: plan cache member + accessor (the FIRST "m_chi_weight_jump;\n};" after "class Elliptic2d" closes that class)
    s = open(ell).read()
    k = s.index("class Elliptic2d")
    member = "    value_type m_jfactor;\n    bool m_chi_weight_jump;\n};"
    j = s.index(member, k)
    s = (s[:j] + "    value_type m_jfactor;\n    bool m_chi_weight_jump;\n"
         "    mutable dgb::shim::EllipticPlanCache m_dgb; //!< libdgb200 fused-kernel plan (built on first use, dropped on copy)\n"
         "    public:\n"
         "    /// libdgb200 binding: the fused-kernel plan of this operator or nullptr if it is outside the kernels' scope\n"
         "    dgb_elliptic2d* dgb_plan() const {\n"
         "        if constexpr( dgb::shim::is_device_dvec<Container>::value && std::is_same_v<dg::get_value_type<Matrix>, double>)\n"
         "            return dgb::shim::elliptic2d_plan( m_dgb, m_leftx, m_lefty, m_rightx, m_righty, m_jumpX, m_jumpY, m_sigma, m_vol, m_chi, m_jfactor, m_chi_weight_jump);\n"
         "        else return nullptr;\n"
         "    }\n"
         "};" + s[j + len(member):])
    # (b) symv prologue of Elliptic2d (first occurrence of the 4-argument symv after the class head)
    head = ("    void symv( value_type alpha, const ContainerType0& x, value_type beta, ContainerType1& y)\n    {\n"
            "        //compute gradient\n        dg::blas2::gemv( m_rightx, x, m_tempx); //R_x*f\n        dg::blas2::gemv( m_righty, x, m_tempy); //R_y*f\n")
    j = s.index(head, k)
    hook = ("    void symv( value_type alpha, const ContainerType0& x, value_type beta, ContainerType1& y)\n    {\n"
            "        if constexpr( dgb::shim::all_device_dvec<Container, ContainerType0, ContainerType1>::value)\n"
            "        {   // libdgb200 binding: the whole operator in one kernel\n"
            "            if( dgb_elliptic2d* plan = dgb_plan())\n"
            "            {\n"
            "                dgb::shim::check( dgb_elliptic2d_symv( plan, alpha, dgb::shim::cptr(x), beta, dgb::shim::mptr(y), nullptr), \"dg::Elliptic2d::symv\");\n"
            "                dgb::shim::note_library();\n"
            "                return;\n"
            "            }\n"
            "        }\n"
            "        //compute gradient\n        dg::blas2::gemv( m_rightx, x, m_tempx); //R_x*f\n        dg::blas2::gemv( m_righty, x, m_tempy); //R_y*f\n")
    s = s[:j] + hook + s[j + len(head):]
    # (c) a new tensor invalidates the plan (the assignment may reuse the old storage)
    old = "        m_chi = SparseTensor<Container>(tau);\n"
    j = s.index(old, k)
    s = s[:j] + old + "        m_dgb.forget();\n" + s[j + len(old):]
    open(ell, "w").write(s)
    # (d) GeneralHelmholtz: the plan of the wrapped operator in Helmholtz mode (chi x - alpha A x in the kernel epilogue)
    hh = os.path.join(out, "dg", "helmholtz.h")
    edit(hh, "    const Container& chi() const{return m_chi;}\n    private:\n    value_type m_alpha;\n    Matrix m_matrix;\n    Container m_chi;\n",
         "    const Container& chi() const{return m_chi;}\n"
         "    /// libdgb200 binding: plan of the wrapped operator switched to y = chi x - alpha A x, or nullptr\n"
         "    dgb_elliptic2d* dgb_plan() const {\n"
         "        if constexpr( dgb::shim::has_dgb_plan<const Matrix>::value && dgb::shim::is_device_dvec<Container>::value) {\n"
         "            if( m_alpha == 0) return nullptr;\n"
         "            dgb_elliptic2d* plan = m_matrix.dgb_plan();\n"
         "            if( plan) dgb::shim::check( dgb_elliptic2d_set_helmholtz( plan, 1, m_alpha, dgb::shim::cptr( m_chi)), \"dgb_elliptic2d_set_helmholtz\");\n"
         "            return plan;\n"
         "        }\n"
         "        else return nullptr;\n"
         "    }\n"
         "    private:\n    value_type m_alpha;\n    Matrix m_matrix;\n    Container m_chi;\n")
    edit(hh, "        if( m_alpha != 0)\n            blas2::symv( m_matrix, x, y);\n",
         "        if constexpr( dgb::shim::all_device_dvec<Container, ContainerType0, ContainerType1>::value)\n"
         "        {   // libdgb200 binding: operator and chi x - alpha y epilogue in one kernel\n"
         "            if( dgb::shim::cptr(x) != dgb::shim::cptr(y))\n"
         "            if( dgb_elliptic2d* plan = dgb_plan())\n"
         "            {\n"
         "                dgb::shim::check( dgb_elliptic2d_symv( plan, 1., dgb::shim::cptr(x), 0., dgb::shim::mptr(y), nullptr), \"dg::GeneralHelmholtz::symv\");\n"
         "                dgb::shim::note_library();\n"
         "                return;\n"
         "            }\n"
         "        }\n"
         "        if( m_alpha != 0)\n            blas2::symv( m_matrix, x, y);\n")
    # (e) PCG::solve
    pcg = os.path.join(out, "dg", "pcg.h")
    edit(pcg, '#include "blas.h"', '#include "blas.h"\n#include "backend/dgb_fused.h" // libdgb200 binding: fused PCG')
    edit(pcg, "    unsigned max_iter;\n    bool m_verbose = false, m_throw_on_fail = true;\n};",
         "    unsigned max_iter;\n    bool m_verbose = false, m_throw_on_fail = true;\n"
         "    dgb::shim::PcgCache m_dgb; //!< libdgb200 solver workspace (created on first fused solve)\n};")
    edit(pcg, "    // self-adjoint: apply PCG algorithm to (P 1/W) (W A) x = (P 1/W) (W b) : P' A' x = P' b'\n",
         "    if constexpr( dgb::shim::has_dgb_plan<std::remove_reference_t<Matrix>>::value &&\n"
         "                  dgb::shim::all_device_dvec<ContainerType, ContainerType0, ContainerType1, std::remove_reference_t<Preconditioner>, ContainerType2>::value)\n"
         "    {   // libdgb200 binding: the whole solve in the library (3 kernels per iteration, scalars on the device)\n"
         "        unsigned its = 0;\n"
         "        if( !m_verbose && dgb::shim::pcg_solve( m_dgb, A, x, b, P, W, eps, nrmb_correction, save_on_dots, max_iter, m_throw_on_fail, its))\n"
         "            return its;\n"
         "    }\n"
         "    // self-adjoint: apply PCG algorithm to (P 1/W) (W A) x = (P 1/W) (W b) : P' A' x = P' b'\n")

    # (f) Advection::upwind and ArakawaX::operator(): one / two kernels instead of six / eight launches
    adv = os.path.join(out, "dg", "advection.h")
    edit(adv, '#include "topology/derivativesA.h"', '#include "topology/derivativesA.h"\n#include "backend/dgb_fused.h" // libdgb200 binding: fused upwind')
    edit(adv, "    blas2::symv( m_dxb, f, m_temp0);\n    blas2::symv( m_dxf, f, m_temp1);\n",
         "    if constexpr( dgb::shim::is_device_ell<Matrix>::value &&\n"
         "                  dgb::shim::all_device_dvec<Container, ContainerType0, ContainerType1, ContainerType2, ContainerType3>::value)\n"
         "    {   // libdgb200 binding: the four derivatives and both upwind updates in one kernel\n"
         "        if( dgb::shim::advection_upwind( m_dxb, m_dxf, m_dyb, m_dyf, alpha, dgb::shim::cptr(vx), dgb::shim::cptr(vy), dgb::shim::cptr(f), beta, dgb::shim::mptr(result)))\n"
         "            return;\n"
         "    }\n"
         "    blas2::symv( m_dxb, f, m_temp0);\n    blas2::symv( m_dxf, f, m_temp1);\n")
    ara = os.path.join(out, "dg", "arakawa.h")
    edit(ara, '#include "topology/derivativesA.h"', '#include "topology/derivativesA.h"\n#include "backend/dgb_fused.h" // libdgb200 binding: fused bracket')
    edit(ara, "    Container m_chi, m_perp_vol;\n};", "    Container m_chi, m_perp_vol;\n"
         "    dgb::shim::ScratchHolder m_dgb; //!< libdgb200 scratch of the fused bracket (3 vectors, allocated on first use)\n};")
    edit(ara, "    //compute derivatives in x-space\n    blas2::symv( m_bdxf, lhs, m_dxlhs);\n",
         "    if constexpr( dgb::shim::is_device_ell<Matrix>::value &&\n"
         "                  dgb::shim::all_device_dvec<Container, ContainerType0, ContainerType1, ContainerType2>::value)\n"
         "    {   // libdgb200 binding: derivatives + ArakawaFunctor in one kernel, the two outer derivatives + chi in a second\n"
         "        if( dgb::shim::arakawa( m_dgb.work, m_bdxf, m_bdyf, alpha, dgb::shim::cptr(lhs), dgb::shim::cptr(rhs), dgb::shim::cptr(m_chi), m_chi.size(), beta, dgb::shim::mptr(result)))\n"
         "            return;\n"
         "    }\n"
         "    //compute derivatives in x-space\n    blas2::symv( m_bdxf, lhs, m_dxlhs);\n")

    # (g) MultiMatrix::symv of two block matrices: the factor-2 projection / interpolation of NestedGrids in one kernel
    fi = os.path.join(out, "dg", "topology", "fast_interpolation.h")
    edit(fi, '#include "dg/blas.h"', '#include "dg/blas.h"\n#include "dg/backend/dgb_fused.h" // libdgb200 binding: one-pass projection / interpolation')
    edit(fi, "        dg::blas2::symv( m_inter[0], x,m_temp[0]);\n",
         "        if constexpr( dgb::shim::is_device_ell<MatrixType>::value && dgb::shim::all_device_dvec<ContainerType, ContainerType0, ContainerType1>::value)\n"
         "        {   // libdgb200 binding: x- and y-matrix in one kernel, no temporary\n"
         "            if( dims == 2 && dgb::shim::multimatrix2_symv( m_dgb, m_inter[0], m_inter[1], alpha, dgb::shim::cptr(x), beta, dgb::shim::mptr(y)))\n"
         "                return;\n"
         "        }\n"
         "        dg::blas2::symv( m_inter[0], x,m_temp[0]);\n")
    edit(fi, "    mutable std::vector<ContainerType > m_temp; // OMG mutable exits !? write even in const\n",
         "    mutable std::vector<ContainerType > m_temp; // OMG mutable exits !? write even in const\n"
         "    mutable dgb::shim::MultiCache m_dgb; //!< libdgb200 binding: classification of the matrix pair\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(HERE, "_build", "inc"))
    ap.add_argument("--no-fusion", action="store_true", help="backend files only: leave elliptic.h / helmholtz.h / pcg.h untouched")
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
    # 3c. SparseMatrix::operator* of host matrices (the reference has no device spgemm): large products run on the device,
    #     bit-identical to detail::spgemm_cpu_kernel (feltor_b200/csrc/spgemm.cu); dg::geo::Fieldaligned's setup is the customer
    sm = os.path.join(a.out, "dg", "backend", "sparsematrix.h")
    edit(sm, '#include "blas2_stencil.h"', '#include "blas2_stencil.h"\n#include "dgb_shim.h" // libdgb200 binding: device spgemm for host matrices')
    edit(sm, "        Vector<Index> row_offsets, cols;\n        Vector<Value> vals;\n\n        detail::spgemm_cpu_kernel(",
         "        Vector<Index> row_offsets, cols;\n        Vector<Value> vals;\n\n"
         "        if constexpr( std::is_same_v<Index, int> && std::is_same_v<Value, double>)\n"
         "        {   // libdgb200 binding: large products on the device\n"
         "            if( dgb::shim::fusion_flag() && lhs.m_vals.size() + rhs.m_vals.size() >= dgb::shim::spgemm_threshold())\n"
         "            {\n"
         "                dgb_spgemm* plan = nullptr;\n"
         "                long long nnz = 0;\n"
         "                const int code = dgb_csr_spgemm_host_begin( &plan, (int)lhs.m_num_rows, (int)lhs.m_num_cols, (int)rhs.m_num_cols,\n"
         "                    thrust::raw_pointer_cast( lhs.m_row_offsets.data()), thrust::raw_pointer_cast( lhs.m_cols.data()), thrust::raw_pointer_cast( lhs.m_vals.data()),\n"
         "                    thrust::raw_pointer_cast( rhs.m_row_offsets.data()), thrust::raw_pointer_cast( rhs.m_cols.data()), thrust::raw_pointer_cast( rhs.m_vals.data()), &nnz);\n"
         "                if( code == 0)\n"
         "                {\n"
         "                    row_offsets.resize( lhs.m_num_rows + 1); cols.resize( nnz); vals.resize( nnz);\n"
         "                    dgb::shim::check( dgb_csr_spgemm_host_finish( plan, thrust::raw_pointer_cast( row_offsets.data()),\n"
         "                        thrust::raw_pointer_cast( cols.data()), thrust::raw_pointer_cast( vals.data())), \"dg::SparseMatrix::operator*\");\n"
         "                    dgb::shim::note_library();\n"
         "                    return SparseMatrix( lhs.m_num_rows, rhs.m_num_cols, row_offsets, cols, vals);\n"
         "                }\n"
         "                if( code != DGB_ERR_UNSUPPORTED) dgb::shim::check( code, \"dg::SparseMatrix::operator*\");\n"
         "                // a row with more distinct columns than the device kernel holds: the host kernel below\n"
         "            }\n"
         "        }\n"
         "        detail::spgemm_cpu_kernel(")
    if not a.no_fusion:
        fusion_hooks(a.out)
    print("make_tree.py: wrote", a.out)


if __name__ == "__main__":
    main()
