// TEST INFRASTRUCTURE ONLY -- not part of the product.
//
// C-ABI wrapper around the UNMODIFIED reference (feltor `dg` header library, OpenMP backend),
// compiled from the sources where they lie under /root/reference/inc by oracle/Makefile into
// oracle/_ref/libdgref.so.  It exists so that (a) the C restatement in oracle/dgoracle.c can be
// validated against the real reference, (b) golden fixtures under tests/golden/ can be generated
// (tests/golden/make_golden.py), and (c) bench.py can time the reference's own CPU path
// (`cpu_baseline.kind == "reference"`).  Nothing under feltor_b200/ may link or load this file.
//
// Build (see oracle/Makefile):
//   g++ -std=c++17 -O2 -mavx2 -mfma -fopenmp -DTHRUST_DEVICE_SYSTEM=THRUST_DEVICE_SYSTEM_OMP
//       -DWITHOUT_VCL -I/root/reference/inc -I/usr/local/cuda/include -shared -fPIC ...
//
// Every entry point takes plain host pointers; vectors are wrapped without copies in
// dg::View<dg::DVec> (reference: inc/dg/backend/view.h:44) where the reference API allows it.
#include <cstdint>
#include <cstring>
#include <vector>
#include <array>
#include <chrono>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "dg/algorithm.h"

using DVec = dg::DVec;
using HVec = dg::HVec;
using DMatrix = dg::DMatrix;
using HMatrix = dg::HMatrix;
#if THRUST_DEVICE_SYSTEM == THRUST_DEVICE_SYSTEM_CUDA
// The same wrapper compiled on a DEVICE backend (integration/Makefile builds it against the libdgb200 binding): the
// caller's host arrays are staged through device vectors; the views below are what the host-backend build gets for free.
struct CStage : dg::View<const DVec> {
    DVec buf;
    CStage(const double* p, size_t n) : buf(p, p + n) { this->construct(thrust::raw_pointer_cast(buf.data()), n); }
    CStage(const CStage&) = delete;
};
using CView = const CStage;
struct VView : dg::View<DVec> {
    DVec buf;
    double* host;
    VView(double* p, size_t n) : buf(p, p + n), host(p) { this->construct(thrust::raw_pointer_cast(buf.data()), n); }
    VView(const VView&) = delete;
    ~VView() { thrust::copy(buf.begin(), buf.end(), host); }
};
namespace dg {
template <> struct TensorTraits<CStage> : TensorTraits<View<const DVec>> {};
template <> struct TensorTraits<VView> : TensorTraits<View<DVec>> {};
}  // namespace dg
#else
using VView = dg::View<DVec>;
using CView = const dg::View<const DVec>;
#endif

extern "C" {

struct RefGrid {
    int ndim;
    double x0[3], x1[3];
    int n[3], N[3], bc[3];
};

int ref_version() { return 1; }
#ifdef _OPENMP
void ref_set_num_threads(int t) { omp_set_num_threads(t); }
int ref_get_max_threads() { return omp_get_max_threads(); }
#else
void ref_set_num_threads(int) {}
int ref_get_max_threads() { return 1; }
#endif
// device backend: switch the fused Elliptic2d / PCG hooks of the binding on or off (integration/dgb_shim/dg/backend/dgb_fused.h)
void ref_set_fusion(int on) {
#if THRUST_DEVICE_SYSTEM == THRUST_DEVICE_SYSTEM_CUDA
    dgb::shim::fusion_flag() = on ? 1 : 0;
#else
    (void)on;
#endif
}
// how many dispatches of this library went to libdgb200.so entry points / to generic kernel templates (device backend only)
void ref_dispatch_counters(long long* library, long long* generic) {
#if THRUST_DEVICE_SYSTEM == THRUST_DEVICE_SYSTEM_CUDA
    *library = dgb::shim::counters().library;
    *generic = dgb::shim::counters().generic;
#else
    *library = 0;
    *generic = 0;
#endif
}
// 0: the reference's OpenMP backend; 1: a device backend (the libdgb200 binding of integration/)
int ref_backend_is_device() { return THRUST_DEVICE_SYSTEM == THRUST_DEVICE_SYSTEM_CUDA ? 1 : 0; }

} // extern C

namespace {
dg::Grid1d g1(const RefGrid* g, int u = 0) {
    return dg::Grid1d(g->x0[u], g->x1[u], g->n[u], g->N[u], (dg::bc)g->bc[u]);
}
dg::CartesianGrid2d g2(const RefGrid* g) {
    return dg::CartesianGrid2d(g->x0[0], g->x1[0], g->x0[1], g->x1[1], g->n[0], g->N[0], g->N[1],
                               (dg::bc)g->bc[0], (dg::bc)g->bc[1]);
}
dg::CartesianGrid3d g3(const RefGrid* g) {
    return dg::CartesianGrid3d(g->x0[0], g->x1[0], g->x0[1], g->x1[1], g->x0[2], g->x1[2], g->n[0],
                               g->N[0], g->N[1], g->N[2], (dg::bc)g->bc[0], (dg::bc)g->bc[1],
                               (dg::bc)g->bc[2]);
}
template <class V>
void copy_out(const V& v, double* out) {
    thrust::copy(v.begin(), v.end(), out);
}
double now() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
// function menu for evaluate
double f_zero(double, double) { return 0.; }
double f_one(double, double) { return 1.; }
double f_sinsin(double x, double y) { return sin(x) * sin(y); }
double f_cosxsiny(double x, double y) { return cos(x) * sin(y); }
double f_cosysinx(double x, double y) { return cos(y) * sin(x); }
double f_sincos(double x, double y) { return sin(x) * cos(y); }
double f_cossin(double x, double y) { return cos(x) * sin(y); }
const double amp = 0.9;
double f_pol(double x, double y) { return 1. + amp * sin(x) * sin(y); }
double f_rhs(double x, double y) {
    return 2. * sin(x) * sin(y) * (amp * sin(x) * sin(y) + 1) - amp * sin(x) * sin(x) * cos(y) * cos(y) -
           amp * cos(x) * cos(x) * sin(y) * sin(y);
}
double f_expexp(double x, double y) { return exp(x) * exp(y); }
double f_shear(double x, double y) {  // evaluation_t.cpp:22-29
    double rho = 0.20943951023931953;
    double delta = 0.050000000000000003;
    if (y <= M_PI) return delta * cos(x) - 1. / rho / cosh((y - M_PI / 2.) / rho) / cosh((y - M_PI / 2.) / rho);
    return delta * cos(x) + 1. / rho / cosh((3. * M_PI / 2. - y) / rho) / cosh((3. * M_PI / 2. - y) / rho);
}
typedef double (*fun2)(double, double);
fun2 menu2[] = {f_zero, f_one, f_sinsin, f_cosxsiny, f_cosysinx, f_sincos, f_cossin, f_pol, f_rhs, f_expexp, f_shear};
double h_sin3(double x, double y, double z) { return sin(x) * sin(y) * sin(z); }
double h_cosx3(double x, double y, double z) { return cos(x) * sin(y) * sin(z); }
double h_cosy3(double x, double y, double z) { return cos(y) * sin(x) * sin(z); }
double h_cosz3(double x, double y, double z) { return cos(z) * sin(x) * sin(y); }
double h_exp3(double x, double y, double z) { return exp(x) * exp(y) * exp(z); }
double h_zero3(double, double, double) { return 0; }
typedef double (*fun3)(double, double, double);
fun3 menu3[] = {h_zero3, h_sin3, h_cosx3, h_cosy3, h_cosz3, h_exp3};
double e_exp(double x) { return exp(x); }
double e_sin(double x) { return sin(x); }
typedef double (*fun1)(double);
fun1 menu1[] = {e_exp, e_sin};
}  // namespace

extern "C" {

// ---------------------------------------------------------------- topology (inc/dg/topology/grid.h:128-166)
int ref_abscissas(const RefGrid* g, int u, double* out) {
    auto a = g1(g, u).abscissas(0);
    copy_out(a, out);
    return (int)a.size();
}
int ref_weights1d(const RefGrid* g, int u, double* out) {
    auto a = g1(g, u).weights(0);
    copy_out(a, out);
    return (int)a.size();
}
int ref_weights(const RefGrid* g, double* out) {  // inc/dg/topology/weights.h:60
    HVec w;
    if (g->ndim == 1) w = dg::create::weights(g1(g));
    else if (g->ndim == 2) w = dg::create::weights(g2(g));
    else w = dg::create::weights(g3(g));
    copy_out(w, out);
    return (int)w.size();
}
int ref_evaluate(const RefGrid* g, int func, double* out) {  // inc/dg/topology/evaluation.h:74
    HVec w;
    if (g->ndim == 1) w = dg::evaluate(menu1[func], g1(g));
    else if (g->ndim == 2) w = dg::evaluate(menu2[func], g2(g));
    else w = dg::evaluate(menu3[func], g3(g));
    copy_out(w, out);
    return (int)w.size();
}
// DLT tables (inc/dg/topology/dlt.h): which = 0 abscissas, 1 weights, 2 backward, 3 forward
int ref_dlt(int which, int n, double* out) {
    std::vector<double> v;
    if (which == 0) v = dg::DLT<double>::abscissas(n);
    else if (which == 1) v = dg::DLT<double>::weights(n);
    else if (which == 2) v = dg::DLT<double>::backward(n);
    else v = dg::DLT<double>::forward(n);
    for (size_t i = 0; i < v.size(); i++) out[i] = v[i];
    return (int)v.size();
}

// ---------------------------------------------------------------- EllSparseBlockMat (inc/dg/backend/sparseblockmat.h:44-188)
// kind 0: derivative(coord, g, bc, dir)            inc/dg/topology/derivatives.h:47
// kind 1: jump(coord, g, bc)                       inc/dg/topology/derivatives.h:68
// kind 2: fast_projection(coord, g, a, b)          inc/dg/topology/fast_interpolation.h:326
// kind 3: fast_interpolation(coord, g, a, b)       inc/dg/topology/fast_interpolation.h:315
void* ref_ell_create(const RefGrid* g, int kind, int coord, int bc, int dir, int a, int b) {
    HMatrix m;
    auto make = [&](const auto& grid) {
        if (kind == 0) m = dg::create::derivative(coord, grid, (dg::bc)bc, (dg::direction)dir);
        else if (kind == 1) m = dg::create::jump(coord, grid, (dg::bc)bc);
        else if (kind == 2) m = dg::create::fast_projection(coord, grid, a, b);
        else m = dg::create::fast_interpolation(coord, grid, a, b);
    };
    if (g->ndim == 1) make(g1(g));
    else if (g->ndim == 2) make(g2(g));
    else make(g3(g));
    return new DMatrix(m);
}
void ref_ell_meta(void* h, int* meta) {
    auto& m = *(DMatrix*)h;
    meta[0] = m.num_rows; meta[1] = m.num_cols; meta[2] = m.blocks_per_line; meta[3] = m.n;
    meta[4] = m.left_size; meta[5] = m.right_size; meta[6] = (int)(m.data.size() / (m.n * m.n));
    meta[7] = m.right_range[0]; meta[8] = m.right_range[1];
}
void ref_ell_arrays(void* h, double* data, int* cols, int* didx) {
    auto& m = *(DMatrix*)h;
    for (size_t i = 0; i < m.data.size(); i++) data[i] = m.data[i];
    for (size_t i = 0; i < m.cols_idx.size(); i++) cols[i] = m.cols_idx[i];
    for (size_t i = 0; i < m.data_idx.size(); i++) didx[i] = m.data_idx[i];
}
void* ref_ell_from_arrays(const int* meta, const double* data, const int* cols, const int* didx) {
    HMatrix m(meta[0], meta[1], meta[2], meta[6], meta[3]);
    for (size_t i = 0; i < m.data.size(); i++) m.data[i] = data[i];
    for (size_t i = 0; i < m.cols_idx.size(); i++) m.cols_idx[i] = cols[i];
    for (size_t i = 0; i < m.data_idx.size(); i++) m.data_idx[i] = didx[i];
    m.set_left_size(meta[4]);
    m.set_right_size(meta[5]);
    m.right_range[0] = meta[7];
    m.right_range[1] = meta[8];
    return new DMatrix(m);
}
void ref_ell_free(void* h) { delete (DMatrix*)h; }
int ref_ell_symv(void* h, double alpha, const double* x, double beta, double* y) {  // inc/dg/blas2.h:325
    auto& m = *(DMatrix*)h;
    CView vx(x, m.total_num_cols());
    VView vy(y, m.total_num_rows());
    try { dg::blas2::symv(alpha, m, vx, beta, vy); } catch (std::exception& e) { return 1; }
    return 0;
}

// ---------------------------------------------------------------- CSR (inc/dg/backend/sparsematrix.h:305, sparsematrix_omp.h:17)
int ref_csr_symv(int nrows, int ncols, int nnz, const int* pos, const int* idx, const double* val, double alpha,
                 const double* x, double beta, double* y) {
    thrust::host_vector<int> p(pos, pos + nrows + 1), c(idx, idx + nnz);
    thrust::host_vector<double> v(val, val + nnz);
    dg::IHMatrix hm(nrows, ncols, p, c, v);
    dg::IDMatrix m(hm);
    CView vx(x, ncols);
    VView vy(y, nrows);
    try { dg::blas2::symv(alpha, m, vx, beta, vy); } catch (std::exception& e) { return 1; }
    return 0;
}

// handle-based variant: the matrix is built once (what Fieldaligned does), only the symv is timed / repeated
void* ref_csr_create(int nrows, int ncols, int nnz, const int* pos, const int* idx, const double* val) {
    thrust::host_vector<int> p(pos, pos + nrows + 1), c(idx, idx + nnz);
    thrust::host_vector<double> v(val, val + nnz);
    dg::IHMatrix hm(nrows, ncols, p, c, v);
    return new dg::IDMatrix(hm);
}
void ref_csr_free(void* h) { delete (dg::IDMatrix*)h; }
int ref_csr_apply(void* h, int nrows, int ncols, double alpha, const double* x, double beta, double* y) {
    CView vx(x, ncols);
    VView vy(y, nrows);
    try { dg::blas2::symv(alpha, *(dg::IDMatrix*)h, vx, beta, vy); } catch (std::exception& e) { return 1; }
    return 0;
}

// ---------------------------------------------------------------- blas1 (inc/dg/blas1.h)
void ref_copy(int n, const double* x, double* y) { CView a(x, n); VView b(y, n); dg::blas1::copy(a, b); }
void ref_scal(int n, double* x, double a) { VView v(x, n); dg::blas1::scal(v, a); }
void ref_plus(int n, double* x, double a) { VView v(x, n); dg::blas1::plus(v, a); }
void ref_axpby(int n, double a, const double* x, double b, double* y) {
    CView vx(x, n); VView vy(y, n); dg::blas1::axpby(a, vx, b, vy);
}
void ref_axpbyz(int n, double a, const double* x, double b, const double* y, double* z) {
    CView vx(x, n), vy(y, n); VView vz(z, n); dg::blas1::axpby(a, vx, b, vy, vz);
}
void ref_axpbypgz(int n, double a, const double* x, double b, const double* y, double g, double* z) {
    CView vx(x, n), vy(y, n); VView vz(z, n); dg::blas1::axpbypgz(a, vx, b, vy, g, vz);
}
void ref_pointwiseDot(int n, double a, const double* x1, const double* x2, double b, double* y) {
    CView v1(x1, n), v2(x2, n); VView vy(y, n); dg::blas1::pointwiseDot(a, v1, v2, b, vy);
}
void ref_pointwiseDot_xy(int n, const double* x1, const double* x2, double* y) {
    CView v1(x1, n), v2(x2, n); VView vy(y, n); dg::blas1::pointwiseDot(v1, v2, vy);
}
void ref_pointwiseDot3(int n, double a, const double* x1, const double* x2, const double* x3, double b, double* y) {
    CView v1(x1, n), v2(x2, n), v3(x3, n); VView vy(y, n); dg::blas1::pointwiseDot(a, v1, v2, v3, b, vy);
}
void ref_pointwiseDot2(int n, double a, const double* x1, const double* y1, double b, const double* x2,
                       const double* y2, double g, double* z) {
    CView a1(x1, n), b1(y1, n), a2(x2, n), b2(y2, n); VView vz(z, n);
    dg::blas1::pointwiseDot(a, a1, b1, b, a2, b2, g, vz);
}
void ref_pointwiseDivide(int n, double a, const double* x1, const double* x2, double b, double* y) {
    CView v1(x1, n), v2(x2, n); VView vy(y, n); dg::blas1::pointwiseDivide(a, v1, v2, b, vy);
}
void ref_pointwiseDivide_xy(int n, const double* x1, const double* x2, double* y) {
    CView v1(x1, n), v2(x2, n); VView vy(y, n); dg::blas1::pointwiseDivide(v1, v2, vy);
}
// y[i] = exp(x[i]) through blas1::transform (blas1_t.cpp:174)
void ref_transform_exp(int n, const double* x, double* y) {
    CView v1(x, n); VView vy(y, n); dg::blas1::transform(v1, vy, dg::EXP<double>());
}
// tensor::multiply2d (inc/dg/topology/multiply.h:215) with explicit t-components
void ref_tensor_multiply2d(int n, const double* lambda, const double* t00, const double* t01, const double* t10,
                           const double* t11, const double* in0, const double* in1, double mu, double* out0,
                           double* out1) {
    CView l(lambda, n), a(t00, n), b(t01, n), c(t10, n), d(t11, n), i0(in0, n), i1(in1, n);
    VView o0(out0, n), o1(out1, n);
    dg::blas1::subroutine(dg::TensorMultiply2d(), l, a, b, c, d, i0, i1, mu, o0, o1);
}

// ---------------------------------------------------------------- vdot / reduce (inc/dg/blas1.h:90-130, 215-222)
// dg::blas1::vdot( f, x, y): extended-precision (FPE) sum of f(x_i, y_i); kind 0: dg::Product  1: a user functor
// dg::blas1::reduce( x, zero, op, unary): kind 0: sum of squares  1: maximum of |x|  2: minimum
struct RefUserBinary {
    DG_DEVICE double operator()(double a, double b) const { return a * b + 0.25 * a; }
};
struct RefSquare {
    DG_DEVICE double operator()(double a) const { return a * a; }
};
struct RefAbs {
    DG_DEVICE double operator()(double a) const { return a < 0 ? -a : a; }
};
double ref_vdot(int kind, int n, const double* x, const double* y) {
    CView a(x, n), b(y, n);
    if (kind == 0) return dg::blas1::vdot(dg::Product(), a, b);
    return dg::blas1::vdot(RefUserBinary(), a, b);
}
double ref_reduce(int kind, int n, const double* x) {
    CView a(x, n);
    if (kind == 0) return dg::blas1::reduce(a, 0., thrust::plus<double>(), RefSquare());
    if (kind == 1) return dg::blas1::reduce(a, 0., thrust::maximum<double>(), RefAbs());
    return dg::blas1::reduce(a, 1e300, thrust::minimum<double>());
}

// ---------------------------------------------------------------- exblas dot (inc/dg/blas1.h:152, blas2.h:94)
// returns status; acc = un-normalised superaccumulator as doDot_superacc returns it
int ref_dot2(int n, const double* x, const double* y, int64_t* acc) {
    int status = 0;
#if THRUST_DEVICE_SYSTEM == THRUST_DEVICE_SYSTEM_CUDA
    CView a(x, n), b(y, n);
    std::vector<int64_t> v = dg::blas1::detail::doDot_superacc(&status, a, b);  // what blas1::dot rounds (blas1.h:159)
    std::copy(v.begin(), v.end(), acc);
#else
    dg::exblas::exdot_omp((unsigned)n, x, y, acc, &status);
#endif
    return status;
}
int ref_dot3(int n, const double* x, const double* w, const double* y, int64_t* acc) {
    int status = 0;
#if THRUST_DEVICE_SYSTEM == THRUST_DEVICE_SYSTEM_CUDA
    CView a(x, n), b(w, n), c(y, n);
    std::vector<int64_t> v = dg::blas2::detail::doDot_superacc(&status, a, b, c);  // what blas2::dot rounds (blas2.h:94)
    std::copy(v.begin(), v.end(), acc);
#else
    dg::exblas::exdot_omp((unsigned)n, x, w, y, acc, &status);
#endif
    return status;
}
double ref_round(const int64_t* acc) {
    int64_t tmp[dg::exblas::BIN_COUNT];
    memcpy(tmp, acc, sizeof(tmp));
    return dg::exblas::cpu::Round(tmp);
}
// public API variants (throw on NaN -> status 1)
int ref_blas1_dot(int n, const double* x, const double* y, double* out) {
    CView a(x, n), b(y, n);
    try { *out = dg::blas1::dot(a, b); } catch (std::exception&) { return 1; }
    return 0;
}
int ref_blas2_dot(int n, const double* x, const double* w, const double* y, double* out) {
    CView a(x, n), b(w, n), c(y, n);
    try { *out = dg::blas2::dot(a, b, c); } catch (std::exception&) { return 1; }
    return 0;
}

// ---------------------------------------------------------------- Elliptic2d (inc/dg/elliptic.h:233-516)
typedef dg::Elliptic2d<dg::CartesianGrid2d, DMatrix, DVec> Ell2d;
void* ref_elliptic2d_create(const RefGrid* g, int bcx, int bcy, int dir, double jfactor, int chi_weight_jump) {
    return new Ell2d(g2(g), (dg::bc)bcx, (dg::bc)bcy, (dg::direction)dir, jfactor, (bool)chi_weight_jump);
}
void ref_elliptic2d_free(void* h) { delete (Ell2d*)h; }
void ref_elliptic2d_set_chi(void* h, const double* sigma) {
    auto& e = *(Ell2d*)h;
    CView s(sigma, e.weights().size());
    e.set_chi(s);
}
void ref_elliptic2d_symv(void* h, double alpha, const double* x, double beta, double* y) {
    auto& e = *(Ell2d*)h;
    size_t n = e.weights().size();
    CView vx(x, n);
    VView vy(y, n);
    e.symv(alpha, vx, beta, vy);
}
void ref_elliptic2d_variation(void* h, double alpha, const double* lambda, const double* phi, double beta,
                              double* sigma) {
    auto& e = *(Ell2d*)h;
    size_t n = e.weights().size();
    CView l(lambda, n), p(phi, n);
    VView s(sigma, n);
    e.variation(alpha, l, p, beta, s);
}
void ref_elliptic2d_weights(void* h, double* out) { copy_out(((Ell2d*)h)->weights(), out); }
void ref_elliptic2d_precond(void* h, double* out) { copy_out(((Ell2d*)h)->precond(), out); }

// ---------------------------------------------------------------- Elliptic1d (inc/dg/elliptic.h:65-200)
void ref_elliptic1d_symv(const RefGrid* g, int bcx, int dir, double jfactor, const double* chi, double alpha, const double* x,
                         double beta, double* y, double* weights, double* precond) {
    dg::Elliptic1d<dg::Grid1d, DMatrix, DVec> e(g1(g), (dg::bc)bcx, (dg::direction)dir, jfactor);
    size_t n = e.weights().size();
    if (chi) { CView s(chi, n); e.set_chi(s); }
    CView vx(x, n);
    VView vy(y, n);
    e.symv(alpha, vx, beta, vy);
    if (weights) copy_out(e.weights(), weights);
    if (precond) copy_out(e.precond(), precond);
}

// ---------------------------------------------------------------- Elliptic3d (inc/dg/elliptic.h:557-797)
// y = alpha Elliptic3d(x) + beta y with set_compute_in_2d(true) (the mode the feltor application uses, src/feltor/feltor.h)
// on a CartesianGrid3d (cylindrical = 0) or a CylindricalGrid3d (x = R, y = Z, z = phi; vol = R); chi = scalar field or NULL.
// weights/precond (optional outputs) as the class reports them
void ref_elliptic3d_symv(const RefGrid* g, int cylindrical, int dir, double jfactor, int chi_weight_jump, const double* chi,
                         double alpha, const double* x, double beta, double* y, double* weights, double* precond) {
    auto run = [&](auto& e) {
        size_t n = e.weights().size();
        e.set_compute_in_2d(true);
        if (chi) { CView s(chi, n); e.set_chi(s); }
        CView vx(x, n);
        VView vy(y, n);
        e.symv(alpha, vx, beta, vy);
        if (weights) copy_out(e.weights(), weights);
        if (precond) copy_out(e.precond(), precond);
    };
    if (cylindrical) {
        dg::CylindricalGrid3d grid(g->x0[0], g->x1[0], g->x0[1], g->x1[1], g->x0[2], g->x1[2], g->n[0], g->N[0], g->N[1], g->N[2],
                                   (dg::bc)g->bc[0], (dg::bc)g->bc[1], (dg::bc)g->bc[2]);
        dg::Elliptic3d<dg::CylindricalGrid3d, DMatrix, DVec> e(grid, (dg::direction)dir, jfactor, (bool)chi_weight_jump);
        run(e);
    } else {
        dg::CartesianGrid3d grid = g3(g);
        dg::Elliptic3d<dg::CartesianGrid3d, DMatrix, DVec> e(grid, (dg::direction)dir, jfactor, (bool)chi_weight_jump);
        run(e);
    }
}

// The same class in its FULL 3-d mode (compute_in_2d = 0: the z derivative, the 3-d tensor product elliptic.h:693 and the
// z jump term elliptic.h:730,743 take part) or restricted to the planes (compute_in_2d = 1).  variation != NULL additionally
// returns Elliptic3d::variation(x) (elliptic.h:758-766).
void ref_elliptic3d_symv_mode(const RefGrid* g, int cylindrical, int dir, double jfactor, int chi_weight_jump, int compute_in_2d,
                              const double* chi, double alpha, const double* x, double beta, double* y, double* variation) {
    auto run = [&](auto& e) {
        size_t n = e.weights().size();
        e.set_compute_in_2d(compute_in_2d != 0);
        if (chi) { CView s(chi, n); e.set_chi(s); }
        CView vx(x, n);
        VView vy(y, n);
        e.symv(alpha, vx, beta, vy);
        if (variation) { DVec v(n); e.variation(vx, v); copy_out(v, variation); }
    };
    if (cylindrical) {
        dg::CylindricalGrid3d grid(g->x0[0], g->x1[0], g->x0[1], g->x1[1], g->x0[2], g->x1[2], g->n[0], g->N[0], g->N[1], g->N[2],
                                   (dg::bc)g->bc[0], (dg::bc)g->bc[1], (dg::bc)g->bc[2]);
        dg::Elliptic3d<dg::CylindricalGrid3d, DMatrix, DVec> e(grid, (dg::direction)dir, jfactor, (bool)chi_weight_jump);
        run(e);
    } else {
        dg::CartesianGrid3d grid = g3(g);
        dg::Elliptic3d<dg::CartesianGrid3d, DMatrix, DVec> e(grid, (dg::direction)dir, jfactor, (bool)chi_weight_jump);
        run(e);
    }
}

// ---------------------------------------------------------------- PCG (inc/dg/pcg.h:136-195)
// returns number of iterations (max_iter if not converged; throw_on_fail is disabled).
// seconds (optional) receives the wall time of solve() only.
int ref_pcg_solve(void* h, double* x, const double* b, const double* P, const double* W, double eps,
                  double nrmb_correction, int test_frequency, int max_iter, double* seconds) {
    auto& e = *(Ell2d*)h;
    size_t n = e.weights().size();
    DVec vx(x, x + n), vb(b, b + n), vP(P, P + n), vW(W, W + n);
    dg::PCG<DVec> pcg(vx, max_iter);
    pcg.set_throw_on_fail(false);
    double t0 = now();
    unsigned it = pcg.solve(e, vx, vb, vP, vW, eps, nrmb_correction, test_frequency);
    if (seconds) *seconds = now() - t0;
    copy_out(vx, x);
    return (int)it;
}
// time `reps` applications of the Elliptic operator (y = A x), returns seconds per application
double ref_elliptic2d_time(void* h, const double* x, double* y, int reps) {
    auto& e = *(Ell2d*)h;
    size_t n = e.weights().size();
    DVec vx(x, x + n), vy(n);
    e.symv(1., vx, 0., vy);
    double t0 = now();
    for (int i = 0; i < reps; i++) e.symv(1., vx, 0., vy);
    double t = (now() - t0) / reps;
    copy_out(vy, y);
    return t;
}

// ---------------------------------------------------------------- MultigridCG2d (inc/dg/multigrid.h:500-668)
struct RefMG {
    dg::CartesianGrid2d grid;
    dg::MultigridCG2d<dg::aGeometry2d, DMatrix, DVec> mg;
    std::vector<dg::Elliptic2d<dg::aGeometry2d, DMatrix, DVec>> ops;
    unsigned stages;
    RefMG(const RefGrid* g, unsigned st) : grid(g2(g)), mg(grid, st), ops(st), stages(st) { mg.set_benchmark(false); }
};
void* ref_multigrid_create(const RefGrid* g, int stages, int dir, double jfactor) {
    auto* m = new RefMG(g, stages);
    for (unsigned u = 0; u < m->stages; u++)
        m->ops[u].construct(m->mg.grid(u), (dg::direction)dir, jfactor);
    return m;
}
void ref_multigrid_free(void* h) { delete (RefMG*)h; }
int ref_multigrid_stage_size(void* h, int u) { return (int)((RefMG*)h)->mg.grid(u).size(); }
// project src (fine) to all stages; out[u] must have stage size
void ref_multigrid_project(void* h, const double* src, double** out) {
    auto& m = *(RefMG*)h;
    DVec s(src, src + m.mg.grid(0).size());
    auto v = m.mg.project(s);
    for (unsigned u = 0; u < m.stages; u++) copy_out(v[u], out[u]);
}
void ref_multigrid_set_chi(void* h, const double* chi) {
    auto& m = *(RefMG*)h;
    DVec s(chi, chi + m.mg.grid(0).size());
    auto v = m.mg.project(s);
    for (unsigned u = 0; u < m.stages; u++) m.ops[u].set_chi(v[u]);
}
// returns 0 ok; numbers[u] = iterations on stage u
int ref_multigrid_solve(void* h, double* x, const double* b, const double* eps, int* numbers, double* seconds) {
    auto& m = *(RefMG*)h;
    size_t n = m.mg.grid(0).size();
    DVec vx(x, x + n), vb(b, b + n);
    std::vector<double> ve(eps, eps + m.stages);
    double t0 = now();
    try {
        auto num = m.mg.solve(m.ops, vx, vb, ve);
        for (unsigned u = 0; u < m.stages; u++) numbers[u] = num[u];
    } catch (std::exception& e) {
        return 1;
    }
    if (seconds) *seconds = now() - t0;
    copy_out(vx, x);
    return 0;
}

}  // extern "C"
