/* dgb200.h -- C ABI of libdgb200.so: the B200-native (sm_100a) data-parallel core of the Feltor `dg` library.
 *
 * This is the drop-in boundary.  Each entry point replaces one overload the reference resolves on
 * `dg::CudaTag`; the citation next to each declaration names the reference
 * interface it stands in for (paths relative to the feltor source tree, v8.2.2).
 *
 * Conventions
 *  - All vector/matrix pointers are DEVICE pointers unless the name ends in `_host` / the comment says host.
 *  - Every call enqueues on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream, which
 *    is what the reference uses everywhere) and returns asynchronously, except the functions documented
 *    as synchronous (they return host data, like the reference's dot).
 *  - Return value: 0 = success, otherwise a DGB_ERR_* code or a cudaError_t value (>0);
 *    dgb_last_error() returns a thread-local message.  Nothing throws across the boundary; the reference-side binding
 *    (integration/dgb_shim/dg/backend/dgb_shim.h, dgb::shim::check) converts non-zero codes into dg::Error like
 *    blas1_cuda.cuh:39-41 does, and DGB_ERR_NOCONVERGE into dg::Fail.
 *  - Like the reference (static scratch buffers in blas1_cuda.cuh:18,33,50) the library is host-thread-affine:
 *    one host thread per device context.  Workspaces are explicit handles so several may coexist.
 *  - There is NO CPU fallback: every compute entry point fails with a CUDA error when no sm_100 device is present.
 */
#ifndef DGB200_H
#define DGB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DGB_API __attribute__((visibility("default")))

#define DGB_BIN_COUNT 39 /* exblas::BIN_COUNT, inc/dg/backend/exblas/config.h:90 */

enum {
    DGB_OK = 0,
    DGB_ERR_INVALID = -1,     /* invalid argument (size mismatch, aliasing that the reference forbids, ...) */
    DGB_ERR_UNSUPPORTED = -2, /* shape outside what the kernels were written for (n > DGB_MAX_N ...) */
    DGB_ERR_NOTFINITE = -3,   /* dot product met NaN/Inf: maps to dg::Error in blas1.h:161 */
    DGB_ERR_NOCONVERGE = -4   /* solver hit max_iter: maps to dg::Fail in pcg.h:189 */
};
#define DGB_MAX_N 8 /* polynomial coefficients per cell and dimension supported by the block kernels */

/* boundary conditions and directions, same numeric values as inc/dg/enums.h:15-21,97-101 */
enum { DGB_PER = 0, DGB_DIR = 1, DGB_DIR_NEU = 2, DGB_NEU_DIR = 3, DGB_NEU = 4 };
enum { DGB_FORWARD = 0, DGB_BACKWARD = 1, DGB_CENTERED = 2 };

typedef void* dgb_stream_t;

/* ---------------------------------------------------------------------------------------------------
 * runtime (replaces thrust::device_vector allocation/copies used by dg::assign/construct,
 * inc/dg/backend/blas1_dispatch_shared.h:26-45)
 * ------------------------------------------------------------------------------------------------- */
DGB_API int dgb_version(void);
DGB_API const char* dgb_last_error(void);
DGB_API int dgb_device_count(int* count);
DGB_API int dgb_set_device(int device);
DGB_API int dgb_sm_count(int* count);
DGB_API int dgb_malloc(void** ptr, size_t bytes);
DGB_API int dgb_free(void* ptr);
DGB_API int dgb_malloc_host(void** ptr, size_t bytes); /* pinned */
DGB_API int dgb_free_host(void* ptr);
DGB_API int dgb_memcpy_h2d(void* dst, const void* src_host, size_t bytes, dgb_stream_t stream);
DGB_API int dgb_memcpy_d2h(void* dst_host, const void* src, size_t bytes, dgb_stream_t stream);
DGB_API int dgb_memcpy_d2d(void* dst, const void* src, size_t bytes, dgb_stream_t stream);
DGB_API int dgb_memset(void* dst, int value, size_t bytes, dgb_stream_t stream);
DGB_API int dgb_stream_synchronize(dgb_stream_t stream);
DGB_API int dgb_device_synchronize(void);
/* number of kernels this library has launched so far in this process (bench.py's gpu_launches claim) */
DGB_API long long dgb_launch_count(void);

/* ---------------------------------------------------------------------------------------------------
 * blas1: replaces doSubroutine_dispatch(CudaTag, size, functor, pointers...) inc/dg/backend/blas1_cuda.cuh:88
 * for the closed set of library functors in inc/dg/subroutines.h:231-384 and inc/dg/topology/multiply.h:18-52.
 * The arithmetic (order of roundings, explicit FMAs) is the functor's; aliasing between arguments is legal.
 * The call-site shortcuts of inc/dg/blas1.h (alpha == 0, &x == &y ...) are applied by the host layer above.
 * ------------------------------------------------------------------------------------------------- */
DGB_API int dgb_copy(size_t n, const double* x, double* y, dgb_stream_t s);                      /* blas1.h:243  equals */
DGB_API int dgb_fill(size_t n, double value, double* y, dgb_stream_t s);                         /* blas1.h:243  copy(scalar, y) */
DGB_API int dgb_scal(size_t n, double* x, double alpha, dgb_stream_t s);                         /* subroutines.h:233 Scal */
DGB_API int dgb_plus(size_t n, double* x, double alpha, dgb_stream_t s);                         /* subroutines.h:247 Plus */
DGB_API int dgb_axpby(size_t n, double alpha, const double* x, double beta, double* y, dgb_stream_t s); /* :260 Axpby */
DGB_API int dgb_axpbyz(size_t n, double alpha, const double* x, double beta, const double* y, double* z,
                       dgb_stream_t s);                                                          /* blas1.h:382 PairSum */
DGB_API int dgb_axpbypgz(size_t n, double alpha, const double* x, double beta, const double* y, double gamma,
                         double* z, dgb_stream_t s);                                             /* :294 Axpbypgz */
DGB_API int dgb_pointwise_dot(size_t n, double alpha, const double* x1, const double* x2, double beta, double* y,
                              dgb_stream_t s); /* :313 PointwiseDot; x1==y or x2==y -> :276 AxyPby (blas1.h:413-421) */
DGB_API int dgb_pointwise_dot_xy(size_t n, const double* x1, const double* x2, double* y, dgb_stream_t s); /* blas1.h:441 */
DGB_API int dgb_pointwise_dot3(size_t n, double alpha, const double* x1, const double* x2, const double* x3,
                               double beta, double* y, dgb_stream_t s);                          /* :325 */
DGB_API int dgb_pointwise_dot2(size_t n, double alpha, const double* x1, const double* y1, double beta,
                               const double* x2, const double* y2, double gamma, double* z, dgb_stream_t s); /* :336 */
DGB_API int dgb_pointwise_divide(size_t n, double alpha, const double* x1, const double* x2, double beta, double* y,
                                 dgb_stream_t s); /* :365 PointwiseDivide; x1==y -> 2-argument overload (blas1.h:501) */
DGB_API int dgb_pointwise_divide_xy(size_t n, const double* x1, const double* x2, double* y, dgb_stream_t s); /* blas1.h:525 */
/* TensorMultiply2d (multiply.h:18-32): out_i = lambda * T_ij in_j + mu*out_i.  lambda: vector or NULL (then the
 * scalar lambda_s); t00..t11: vectors or NULL = the SparseTensor's implicit 1 (diagonal) / 0 (off-diagonal). */
DGB_API int dgb_tensor_multiply2d(size_t n, const double* lambda, double lambda_s, const double* t00,
                                  const double* t01, const double* t10, const double* t11, const double* in0,
                                  const double* in1, double mu, double* out0, double* out1, dgb_stream_t s);
/* detail::Tolerance of dg::Adaptive (adaptive.h:123-134, used at :300): delta = delta / (rtol*|u0| + atol); the caller
 * passes rtol, atol already scaled by sqrt(size) as the functor's constructor does */
DGB_API int dgb_adaptive_tolerance(size_t n, double rtol, double atol, const double* u0, double* delta, dgb_stream_t s);
/* TensorMultiply3d (multiply.h:34-58): the same for 3 components; t = HOST array of 9 device pointers (row major,
 * NULL entries / NULL t = identity), in / out = HOST arrays of 3 device pointers (out may alias in) */
DGB_API int dgb_tensor_multiply3d(size_t n, const double* lambda, double lambda_s, const double* const t[9],
                                  const double* const in[3], double mu, double* const out[3], dgb_stream_t s);
/* EmbeddedPairSum (subroutines.h:179-204, used by ERKStep runge_kutta.h:35-62):
 * y = b0*y + sum_s b[s]*k[s], yt = bt0*yt + sum_s bt[s]*k[s];  b, bt, k are HOST arrays (k of device pointers), nk <= 16 */
DGB_API int dgb_embedded_pair_sum(size_t n, double* y, double* yt, double b0, double bt0, int nk, const double* b_host,
                                  const double* bt_host, const double* const* k_host, dgb_stream_t s);
/* blas1::transform with a unary functor of inc/dg/functors.h (y = op(x)); op codes below */
enum { DGB_OP_EXP = 0, DGB_OP_LN = 1, DGB_OP_SQRT = 2, DGB_OP_INVERT = 3, DGB_OP_ABS = 4, DGB_OP_SQUARE = 5,
       DGB_OP_INVSQRT = 6 };
/* the pointwise step of dg::ArakawaX::operator() (arakawa.h:125-145,156): dylhs, dxrhs, dyrhs are overwritten */
DGB_API int dgb_arakawa_functor(size_t n, const double* lhs, const double* rhs, const double* dxlhs, double* dylhs,
                                double* dxrhs, double* dyrhs, dgb_stream_t s);
/* y = alpha * v * (v >= 0 ? back : forw) + beta * y: blas1::evaluate(y, Axpby(alpha,beta), UpwindProduct(), v, back, forw)
 * as used by dg::Advection::upwind (advection.h:112-120, functors.h:312-337) */
DGB_API int dgb_upwind_axpby(size_t n, double alpha, const double* v, const double* back, const double* forw, double beta,
                             double* y, dgb_stream_t s);
/* y = alpha * lambda mu (v . T w) + beta y: dg::tensor::scalar_product2d (multiply.h:493-512).  NULL lambda / mu mean the
 * scalars given next to them, NULL tensor entries the identity tensor. */
DGB_API int dgb_tensor_dot2d(size_t n, double alpha, const double* lambda, double lambda_s, const double* v0, const double* v1,
                             const double* t00, const double* t01, const double* t10, const double* t11, const double* mu,
                             double mu_s, const double* w0, const double* w1, double beta, double* y, dgb_stream_t s);
/* y = alpha * PairSum(a_0, x_0, ..., a_{nk-1}, x_{nk-1}) + beta * y with the reference's nesting
 * fma(a_0, x_0, fma(a_1, x_1, ... a_last x_last)) (subroutines.h:123-143): the dense-matrix gemv behind the
 * Runge-Kutta / multistep stage sums (blas2_densematrix.h:38-74); a: host array, x: host array of device pointers, nk <= 8 */
DGB_API int dgb_pair_sum_axpby(size_t n, double alpha, int nk, const double* a_host, const double* const* x, double beta,
                               double* y, dgb_stream_t s);
DGB_API int dgb_transform(size_t n, int op, const double* x, double* y, dgb_stream_t s);          /* blas1.h:585 */

/* ---------------------------------------------------------------------------------------------------
 * exblas dot: replaces doDot_dispatch(CudaTag, status, size, x, y[, z]) inc/dg/backend/blas1_cuda.cuh:30-64
 * (exdot_gpu, inc/dg/backend/exblas/exdot_cuda.cuh:319-357).
 * Contract: the exact sum of the individually rounded products  round(x*y)  resp.  round(round(x*w)*y)
 * accumulated in a 39 x int64 superaccumulator (word i has weight 2^(56*(i-20))), returned NORMALISED
 * (accumulate.h:267-285) so exblas::cpu::Round / Normalize accept it unchanged, plus the correctly rounded double
 * (accumulate.h:297-349) computed on the device.  Bit-reproducible for any launch geometry / GPU count.
 * ------------------------------------------------------------------------------------------------- */
typedef struct dgb_dot_result {
    int64_t acc[DGB_BIN_COUNT]; /* normalised superaccumulator */
    double value;               /* Round(acc) */
    int32_t status;             /* 0 ok, 1 a product was NaN/Inf (blas1.h:161) */
    int32_t pad;
} dgb_dot_result;
/* dg::blas1::reduce(x, init, binary_op, unary_op) (blas1.h:213-223, blas1_cuda.cuh:96-102) for the (op, unary) pairs the
 * library and the applications use; synchronous like the reference (returns the value on the host).  Max / min / or are
 * order independent; SUM is a plain floating-point sum (not reproducible, as in the reference) -- the exact sums are
 * dgb_dot2/3 with a scalar operand (blas1::vdot, dot(1., x)). */
enum { DGB_REDUCE_SUM = 0, DGB_REDUCE_MAX = 1, DGB_REDUCE_MIN = 2, DGB_REDUCE_OR = 3 };
enum { DGB_UNARY_IDENTITY = 0, DGB_UNARY_ABS = 1, DGB_UNARY_SQUARE = 2, DGB_UNARY_ISNAN = 3, DGB_UNARY_ISNOTFINITE = 4 };
DGB_API int dgb_reduce(size_t n, const double* x, int op, int unary, double init, double* result_host, dgb_stream_t s);
typedef struct dgb_dot_ws dgb_dot_ws; /* device scratch: per-block partial superaccumulators + ticket */
DGB_API int dgb_dot_ws_create(dgb_dot_ws** ws);
DGB_API int dgb_dot_ws_destroy(dgb_dot_ws* ws);
/* asynchronous: result (a DEVICE pointer to dgb_dot_result) is complete when the stream reaches this point.
 * A NULL operand pointer means the scalar given next to it (dot(1., v)). */
DGB_API int dgb_exdot2(dgb_dot_ws* ws, size_t n, const double* x, double xs, const double* y, double ys,
                       dgb_dot_result* result_dev, dgb_stream_t s);
DGB_API int dgb_exdot3(dgb_dot_ws* ws, size_t n, const double* x, double xs, const double* w, double ws_,
                       const double* y, double ys, dgb_dot_result* result_dev, dgb_stream_t s);
/* synchronous convenience = the reference seam: superaccumulator on the host (acc_host[39], may be NULL),
 * rounded value (may be NULL), *status as the reference sets it.  Returns DGB_ERR_NOTFINITE if status != 0. */
DGB_API int dgb_dot2(dgb_dot_ws* ws, size_t n, const double* x, const double* y, int64_t* acc_host, double* value,
                     int* status, dgb_stream_t s);
DGB_API int dgb_dot3(dgb_dot_ws* ws, size_t n, const double* x, const double* w, const double* y, int64_t* acc_host,
                     double* value, int* status, dgb_stream_t s);
/* host-side helpers with the reference's arithmetic (accumulate.h:267-349), for callers that combine
 * accumulators of std::vector<DVec> / ranks (blas1_dispatch_vector.h:153-176, mpi_accumulate.h:94-125) */
/* returns 0; *negative (may be NULL) receives what exblas::cpu::Normalize returns: 1 if the accumulator is negative */
DGB_API int dgb_superacc_normalize_host(int64_t* acc, int* negative);
DGB_API double dgb_superacc_round_host(const int64_t* acc);
/* device-side combine for multi-GPU: acc[i] = Normalize(sum_r parts[r][i]) then Round; parts = nparts normalised
 * accumulators (<= 256, mpi_accumulate.h:75-77) contiguous in device memory; result_dev filled like dgb_exdot */
DGB_API int dgb_superacc_combine(const int64_t* parts_dev, int nparts, const int32_t* status_parts_dev,
                                 dgb_dot_result* result_dev, dgb_stream_t s);

/* ---------------------------------------------------------------------------------------------------
 * EllSparseBlockMat / CooSparseBlockMat symv: replaces launch_multiply_kernel(CudaTag, alpha, x, beta, y)
 * inc/dg/backend/sparseblockmat.h:180-186,349-357 (kernels sparseblockmat_gpu_kernels.cuh:8-354).
 * y = alpha (1_left (x) M (x) 1_right) x + beta y, fields exactly as sparseblockmat.h:168-177.
 * Per output element: temp_d = fma-chain over q for block d; y = beta==0 ? 0 : y*beta; y = fma(alpha,temp_d,y)
 * for d ascending (sparseblockmat_omp_kernels.h:36-50) -- bit-identical to the reference's OpenMP backend.
 * x must not alias y.  beta == 0 does not read y (NaN is overwritten).
 * ------------------------------------------------------------------------------------------------- */
typedef struct dgb_ell_host { /* HOST description, fields as sparseblockmat.h:168-177 */
    int num_rows, num_cols, blocks_per_line, n, left_size, right_size;
    int num_blocks;      /* data holds num_blocks * n * n doubles */
    int right_range[2];  /* sparseblockmat.h:175 */
    const double* data;  /* host */
    const int* cols_idx; /* host, num_rows*blocks_per_line, -1 = padding (sparseblockmat.h:49) */
    const int* data_idx; /* host */
} dgb_ell_host;
typedef struct dgb_ell dgb_ell; /* device-resident matrix + launch plan (row classification, constant blocks) */
DGB_API int dgb_ell_create(dgb_ell** m, const dgb_ell_host* host);
DGB_API int dgb_ell_destroy(dgb_ell* m);
/* the reference lets applications re-purpose a matrix for another dimension (sparseblockmat.h:139-166) */
DGB_API int dgb_ell_set_left_size(dgb_ell* m, int left_size);
DGB_API int dgb_ell_set_right_size(dgb_ell* m, int right_size); /* resets right_range to [0,right_size) */
DGB_API int dgb_ell_set_right_range(dgb_ell* m, int begin, int end);
DGB_API int dgb_ell_total_num_rows(const dgb_ell* m, size_t* rows);
DGB_API int dgb_ell_total_num_cols(const dgb_ell* m, size_t* cols);
DGB_API int dgb_ell_symv(const dgb_ell* m, double alpha, const double* x, double beta, double* y, dgb_stream_t s);
/* same arithmetic through the fully general one-thread-per-element kernel (any n, any pattern); A/B testing */
DGB_API int dgb_ell_symv_generic(const dgb_ell* m, double alpha, const double* x, double beta, double* y,
                                 dgb_stream_t s);
/* dg::Advection::upwind( alpha, vx, vy, f, beta, result) (inc/dg/advection.h:112-120) in ONE kernel: the four block matrices are
 * the backward / forward derivatives in x and y of one 2-d grid (dg::create::dx / dy, n = 2..4, two blocks per row); bitwise the
 * result of the reference's sequence of four symv and two evaluate( Axpby, UpwindProduct) calls.  DGB_ERR_UNSUPPORTED for
 * other matrices (compose dgb_ell_symv + dgb_upwind_axpby then).  f must not alias result. */
DGB_API int dgb_advection_upwind(const dgb_ell* dxb, const dgb_ell* dxf, const dgb_ell* dyb, const dgb_ell* dyf, double alpha,
                                 const double* vx, const double* vy, const double* f, double beta, double* result, dgb_stream_t s);
/* dg::ArakawaX::operator()( alpha, lhs, rhs, beta, result) (inc/dg/arakawa.h:147-162) in TWO kernels instead of eight launches:
 * bdx / bdy are the centered derivatives of one 2-d grid (n = 2..4, three blocks per row), chi = 1 / perp_vol (the class'
 * m_chi), work3 = 3 N doubles of scratch (the class' m_dylhs, m_dxrhs, m_dyrhs).  Bitwise the result of the reference's sequence
 * (six symv, ArakawaFunctor, pointwiseDot).  DGB_ERR_UNSUPPORTED for other matrices (compose dgb_ell_symv +
 * dgb_arakawa_functor + dgb_pointwise_dot then).  result must not alias work3; it may alias lhs or rhs (neither is read any
 * more when result is written). */
DGB_API int dgb_arakawa(const dgb_ell* bdx, const dgb_ell* bdy, double alpha, const double* lhs, const double* rhs, const double* chi,
                        double beta, double* result, double* work3, dgb_stream_t s);

/* dg::MultiMatrix::symv of two block matrices, x-matrix first (fast_interpolation.h:71-84): y = alpha my (mx x) + beta y.
 * dgb_multimatrix2_fused reports in *kind whether the pair is a factor-2 projection (1) or interpolation (2) of
 * dg::create::fast_projection / fast_interpolation (n = 2..4) -- those run as ONE kernel without the temporary (10 B per fine
 * element instead of four vector passes), bitwise equal to the two products; kind 0 runs the two products through `temp`
 * (mx.total_num_rows() doubles). */
DGB_API int dgb_multimatrix2_fused(const dgb_ell* mx, const dgb_ell* my, int* kind);
DGB_API int dgb_multimatrix2_symv(const dgb_ell* mx, const dgb_ell* my, int kind, double alpha, const double* x, double beta, double* y,
                                  double* temp, dgb_stream_t s);

typedef struct dgb_coo {
    int num_rows, num_cols, num_entries, n, left_size, right_size;
    const double* data;  /* device */
    const int* rows_idx; /* device, num_entries */
    const int* cols_idx; /* device: index into the pointer table x */
    const int* data_idx; /* device */
} dgb_coo;
/* x = DEVICE array of device pointers, chunk c laid out [q][s][j] (sparseblockmat_omp_kernels.h:370); beta must be 1 */
DGB_API int dgb_coo_symv(const dgb_coo* m, double alpha, const double* const* x_ptrs_dev, double beta, double* y,
                         dgb_stream_t s);

/* ---------------------------------------------------------------------------------------------------
 * Topology (HOST-side setup, no device work): what the applications obtain from inc/dg/topology of the reference.
 * The coefficients are bit-identical to the reference's (tests/test_topology.py).
 * ------------------------------------------------------------------------------------------------- */
typedef struct dgb_grid { /* aRealTopology<double,Nd>, inc/dg/topology/grid.h:92; x is the fastest dimension */
    int ndim;
    double x0[3], x1[3];
    int n[3], N[3], bc[3];
} dgb_grid;
typedef struct dgb_ellh dgb_ellh; /* host EllSparseBlockMat owning its arrays */
DGB_API int dgb_topo_dlt(int which, int n, double* out_host); /* dlt.h: 0 abscissas, 1 weights, 2 backward, 3 forward */
/* dg::create::window_stencil (topology/stencil.h:177-237): neighbourhood matrix for dgb_csr_stencil.  window[ndim] = points per
 * axis; outputs are caller-allocated HOST arrays: row_offsets[size + 1], cols / vals[size * prod(window)] (entries unsorted,
 * duplicates kept, value -1 for points mirrored at a Dirichlet boundary, exactly as the reference builds it) */
DGB_API int dgb_topo_window_stencil(const dgb_grid* g, const int* window, int* row_offsets, int* cols, double* vals);
/* dg::create::limiter_stencil (topology/stencil.h:89-137,199-256): the matrix of dg::CSRSlopeLimiter (DGB_STENCIL_SLOPE) for a 1-d
 * grid or along `direction` (0 x, 1 y) of a 2-d grid with boundary condition `bound`; row_offsets[size + 1], cols / vals[3 size] */
DGB_API int dgb_topo_limiter_stencil(const dgb_grid* g, int direction, int bound, int* row_offsets, int* cols, double* vals);
DGB_API int dgb_topo_size(const dgb_grid* g, size_t* size);
DGB_API int dgb_topo_abscissas(const dgb_grid* g, int axis, double* out_host); /* grid.h:128 */
DGB_API int dgb_topo_weights1d(const dgb_grid* g, int axis, double* out_host); /* grid.h:155 */
DGB_API int dgb_topo_weights(const dgb_grid* g, double* out_host);             /* weights.h:60 create::weights */
DGB_API int dgb_topo_dx(dgb_ellh** m, int n, int N, double h, int bc, int dir); /* dx.h:389 dx_normed */
DGB_API int dgb_topo_jump(dgb_ellh** m, int n, int N, double h, int bc);        /* dx.h:301 jump */
DGB_API int dgb_topo_derivative(dgb_ellh** m, const dgb_grid* g, int coord, int bc, int dir); /* derivatives.h:47 */
DGB_API int dgb_topo_jump_nd(dgb_ellh** m, const dgb_grid* g, int coord, int bc);            /* derivatives.h:68 */
DGB_API int dgb_topo_fast_projection(dgb_ellh** m, const dgb_grid* g, int coord, int dividen, int divideN);
                                                                                /* fast_interpolation.h:326 */
DGB_API int dgb_topo_fast_interpolation(dgb_ellh** m, const dgb_grid* g, int coord, int multiplyn, int multiplyN);
                                                                                /* fast_interpolation.h:315 */
DGB_API int dgb_ellh_view(const dgb_ellh* m, dgb_ell_host* view); /* pointers stay owned by m */
DGB_API int dgb_ellh_destroy(dgb_ellh* m);

/* ---------------------------------------------------------------------------------------------------
 * CSR spmv: replaces detail::spmv_gpu_kernel (cuSPARSE) inc/dg/backend/sparsematrix_gpu.cuh:190-214 with the
 * reference OpenMP order (sparsematrix_omp.h:17-52): beta==1: y = fma(alpha*v, x[j], y) sequentially over the
 * row, else t = sum fma(alpha*v, x[j], t); y = fma(beta, y, t) (beta == 0 does not read y).  Bit-reproducible.
 * dgb_csr_spmv_planes applies the same 2-D matrix to nplanes planes in ONE launch (Fieldaligned::ePlus/eMinus,
 * inc/geometries/fieldaligned.h:850-912): y[p] = alpha A x[(p + shift) mod nplanes] + beta y[p].
 * ------------------------------------------------------------------------------------------------- */
DGB_API int dgb_csr_spmv(int num_rows, int num_cols, const int* row_offsets, const int* cols, const double* vals,
                         double alpha, const double* x, double beta, double* y, dgb_stream_t s);
DGB_API int dgb_csr_spmv_planes(int num_rows, int num_cols, const int* row_offsets, const int* cols,
                                const double* vals, double alpha, const double* x, double beta, double* y,
                                int nplanes, int shift, dgb_stream_t s);

/* A = B C for CSR matrices ON THE DEVICE, bit-identical to the reference's host kernel dg::detail::spgemm_cpu_kernel
 * (inc/dg/backend/sparsematrix_cpu.h:19-95, called by SparseMatrix::operator*, sparsematrix.h:549-566): rows of A sorted by
 * column, duplicates and unsorted input allowed, every entry accumulated as w = fma(b, c, w) in the serial loop's candidate order.
 * The reference has no device version; dg::geo::Fieldaligned calls it three times per construction (fieldaligned.h:645-658).
 *   symbolic  counts the distinct columns of every row, returns the plan and the number of entries of A (synchronises);
 *   numeric   fills A_pos[B_rows + 1], A_idx[nnz], A_val[nnz] (device arrays);
 *   host_begin / host_finish  the same for HOST arrays (uploads, multiplies, downloads; finish releases the plan).
 * DGB_ERR_UNSUPPORTED if a row of A has more than 4096 distinct columns or A more than 2^31 - 1 entries. */
typedef struct dgb_spgemm dgb_spgemm;
DGB_API int dgb_csr_spgemm_symbolic(dgb_spgemm** plan, int B_rows, int B_cols, int C_cols, const int* B_pos, const int* B_idx,
                                    const int* C_pos, const int* C_idx, long long* nnz, dgb_stream_t s);
DGB_API int dgb_csr_spgemm_numeric(dgb_spgemm* plan, const int* B_pos, const int* B_idx, const double* B_val, const int* C_pos,
                                   const int* C_idx, const double* C_val, int* A_pos, int* A_idx, double* A_val, dgb_stream_t s);
DGB_API int dgb_csr_spgemm_destroy(dgb_spgemm* plan);
DGB_API int dgb_csr_spgemm_host_begin(dgb_spgemm** plan, int B_rows, int B_cols, int C_cols, const int* B_pos, const int* B_idx,
                                      const double* B_val, const int* C_pos, const int* C_idx, const double* C_val, long long* nnz);
DGB_API int dgb_csr_spgemm_host_finish(dgb_spgemm* plan, int* A_pos, int* A_idx, double* A_val);

/* The pieces of dg::MPIDistMat::symv (inc/dg/backend/mpi_matrix.h:478-523) a row-distributed CSR matrix needs besides
 * dgb_csr_spmv for its inner part and dgb_comm_gather for the exchange (feltor_b200/dist_csr.py puts them together):
 *   dgb_gather_indexed        out[i] = x[idx[i]]: packs the values other ranks asked for (MPIGather, mpi_gather.h:454-705)
 *   dgb_csr_spmv_scatter_add  y[scatter[i]] += (outer row i) . buffer -- the outer matrix product and the scatter of its rows in
 *                             one kernel (the fusion mpi_matrix.h:513-520 asks for), CSR order of sparsematrix_omp.h:39-48 */
DGB_API int dgb_gather_indexed(size_t n, const int* idx, const double* x, double* out, dgb_stream_t s);
DGB_API int dgb_csr_spmv_scatter_add(int num_rows, const int* row_offsets, const int* cols, const double* vals, const double* buffer,
                                     const int* scatter, double* y, dgb_stream_t s);
/* allreduce mode of MPIDistMat (mpi_matrix.h:438-441,487-490; dg::Average over a distributed axis): parts = nranks vectors of m
 * doubles, rank-major (what dgb_comm_gather delivers when every rank sends its partial result to every rank);
 * y[i] = (..(parts[0][i] + parts[1][i]) + ..) + parts[nranks-1][i] -- rank order, hence the same bits on every rank */
DGB_API int dgb_sum_ranks(int nranks, size_t m, const double* parts, double* y, dgb_stream_t s);

/* dg::blas2::stencil(f, M, x, y) / parallel_for (blas2.h:413-454, blas2_stencil.h:13-70) for the library's CSR stencil
 * functors (topology/filter.h:174-266): the matrix only encodes the neighbourhood of each row (create::window_stencil).
 * y[i] = lower median / switching median (alpha) / average of x over the stencil, or sum x*vals (test filter).
 * DGB_STENCIL_SLOPE = dg::CSRSlopeLimiter( alpha) (filter.h:288-336) on the matrix of create::limiter_stencil: the generalised
 * minmod slope limiter of every cell along one axis; it writes exactly the entries of y the reference writes (all of them for
 * a limiter stencil).  x must not alias y. */
enum { DGB_STENCIL_MEDIAN = 0, DGB_STENCIL_SWM = 1, DGB_STENCIL_AVERAGE = 2, DGB_STENCIL_SYMV = 3, DGB_STENCIL_SLOPE = 4 };
DGB_API int dgb_csr_stencil(int kind, int num_rows, const int* row_offsets, const int* cols, const double* vals,
                            double alpha, const double* x, double* y, dgb_stream_t s);

/* ---------------------------------------------------------------------------------------------------
 * dg::geo::Fieldaligned / dg::geo::DS apply path (inc/geometries/fieldaligned.h:806-912, ds.h:170-330,744-852).  The
 * 2-d interpolation matrices m_plus / m_minus are CSR arrays on the device (their construction by field-line tracing
 * stays on the host, out of scope).
 * ------------------------------------------------------------------------------------------------- */
/* ePlus (plus=1: out[k] = M f[k+1]) / eMinus (plus=0: out[k] = M f[k-1]) on all planes in one launch + the ghost-cell
 * fix-up of the last/first plane for bcz != DGB_PER (bnd = m_right/m_left, limiter = m_limiter, ghost = scratch) */
DGB_API int dgb_fa_shift(int plus, int num_rows, int nplanes, const int* pos, const int* idx, const double* val,
                         const double* f, double* out, int bcz, const double* bnd, const double* limiter, double* ghost,
                         double delta_phi, dgb_stream_t s);
/* ds_forward / ds_backward / ds_centered / ds_forward2 / ds_backward2 / dss_centered (ds.h:744-852): kind 0..5,
 * operands (a, b, c) in the argument order of the reference functions, see feltor_b200/csrc/ds.cu */
DGB_API int dgb_ds_apply(int kind, size_t n, double alpha, const double* a, const double* b, const double* c,
                         const double* bphi_m, const double* bphi, const double* bphi_p, double delta_phi, double beta,
                         double* g, dgb_stream_t s);
/* dssd_centered / ds_divBackward / ds_divForward / ds_divCentered / ds_average (ds.h:881-1000): kind 6..10, operands in
 * the argument order of the reference functions (6: fm,f,fp  7: fm,f  8: f,fp  9: fm,fp  10: fm,fp); sqrtG_* = fa.sqrtGm/
 * sqrtG/sqrtGp, bphi_* = fa.bphiM/bphi/bphiP; fields a kind does not use may be NULL */
DGB_API int dgb_ds_apply_vol(int kind, size_t n, double alpha, const double* a, const double* b, const double* c,
                             const double* sqrtG_m, const double* sqrtG, const double* sqrtG_p, const double* bphi_m,
                             const double* bphi, const double* bphi_p, double delta_phi, double beta, double* g,
                             dgb_stream_t s);
/* assign_bc_along_field_2nd (order = 2; fm, f, fp) / _1st (order = 1; fm, fp, f unused) (ds.h:169-296): ghost values fmg / fpg of
 * the shifted fields where the field line leaves the domain.  bc = DGB_NEU (boundary values = derivatives) or DGB_DIR (values);
 * bv_minus / bv_plus = boundary_value[0] / [1]; hbm, hbp, bbm, bbo, bbp = the fields of the Fieldaligned object (the first-order
 * Neumann form needs only bbm, bbp).  fmg may alias fm, fpg may alias fp.  dgb_swap_bc_perp: swap_bc_perp (ds.h:307-318). */
DGB_API int dgb_assign_bc_along_field(int order, int bc, size_t n, double delta_phi, const double* fm, const double* f, const double* fp,
                                      const double* hbm, const double* hbp, const double* bbm, const double* bbo, const double* bbp,
                                      double bv_minus, double bv_plus, double* fmg, double* fpg, dgb_stream_t s);
DGB_API int dgb_swap_bc_perp(size_t n, const double* fm, const double* fp, const double* bbm, const double* bbo, const double* bbp,
                             double* fmg, double* fpg, dgb_stream_t s);
/* DS::centered(alpha, f, beta, g) (ds.h:481-485) for periodic z fused into one kernel (gather f+, f-, formula) */
DGB_API int dgb_ds_centered_fused(int num_rows, int nplanes, const int* plus_pos, const int* plus_idx,
                                  const double* plus_val, const int* minus_pos, const int* minus_idx,
                                  const double* minus_val, double alpha, const double* f, const double* bphi,
                                  double delta_phi, double beta, double* g, dgb_stream_t s);

/* Dedicated gather layout for the field-line interpolation matrices (north star item 4): the CSR matrix is converted
 * once into sliced ELL (32-row slices, entries of a slice stored [k][lane]) so that a warp reads indices / values
 * coalesced; summation order = CSR order, results bitwise equal to dgb_csr_spmv_planes / dgb_ds_centered_fused. */
typedef struct dgb_gather_plan dgb_gather_plan;
DGB_API int dgb_gather_plan_create(dgb_gather_plan** plan, int num_rows, int num_cols, const int* row_offsets_dev,
                                   const int* cols_dev, const double* vals_dev, dgb_stream_t s);
DGB_API int dgb_gather_plan_destroy(dgb_gather_plan* plan);
/* y[pl] = alpha M x[(pl + shift) mod nplanes] + beta y[pl] for all planes (Fieldaligned::ePlus/eMinus core) */
DGB_API int dgb_gather_spmv_planes(const dgb_gather_plan* plan, double alpha, const double* x, double beta, double* y,
                                   int nplanes, int shift, dgb_stream_t s);
/* DS::centered(alpha, f, beta, g) (ds.h:481-485), periodic z, one launch */
DGB_API int dgb_gather_ds_centered(const dgb_gather_plan* plus, const dgb_gather_plan* minus, int nplanes, double alpha,
                                   const double* f, const double* bphi, double delta_phi, double beta, double* g,
                                   dgb_stream_t s);

/* Cell-tiled layout for the matrices dg::geo::Fieldaligned actually builds (inc/geometries/fieldaligned.h:631-657, projection *
 * interpolation): the n^2 rows of one target cell share ONE column list (the nodes of the 2..6 source cells its field lines end
 * in), i.e. the matrix is a dense n^2 x U block per target cell.  The plan keeps that block and 16-bit indices into a
 * shared-memory staging buffer; a thread owns one target cell and two planes, so every gathered operand feeds 2 n^2 FMAs
 * instead of one (feltor_b200/csrc/celltile.cu).  n, Nx, Ny: polynomial order and cells of the perpendicular grid
 * (num_rows = n^2 Nx Ny, x fastest).  A CSR matrix without that structure -> DGB_ERR_UNSUPPORTED (use dgb_gather_plan_*).
 * Results are bitwise those of dgb_csr_spmv_planes / dgb_gather_* (column order of the CSR rows). */
typedef struct dgb_celltile_plan dgb_celltile_plan;
DGB_API int dgb_celltile_plan_create(dgb_celltile_plan** plan, int n, int Nx, int Ny, const int* row_offsets_dev, const int* cols_dev,
                                     const double* vals_dev, dgb_stream_t s);
DGB_API int dgb_celltile_plan_destroy(dgb_celltile_plan* plan);
DGB_API int dgb_celltile_plan_info(const dgb_celltile_plan* plan, int* ntiles, int* max_source_cells, int* planes_per_cta, long long* nnz);
/* y[pl] = alpha M x[(pl + shift) mod nplanes] + beta y[pl]; alpha = +-1, beta != 1 (the cases Fieldaligned::ePlus/eMinus use) */
DGB_API int dgb_celltile_spmv_planes(const dgb_celltile_plan* plan, double alpha, const double* x, double beta, double* y, int nplanes,
                                     int shift, dgb_stream_t s);
/* DS::centered(alpha, f, beta, g) (ds.h:481-485), periodic z, ONE launch: both gathers and the formula */
DGB_API int dgb_celltile_ds_centered(const dgb_celltile_plan* plus, const dgb_celltile_plan* minus, int nplanes, double alpha,
                                     const double* f, const double* bphi, double delta_phi, double beta, double* g, dgb_stream_t s);

/* ---------------------------------------------------------------------------------------------------
 * Fused Elliptic2d: replaces the 8-kernel composition of Elliptic2d::symv inc/dg/elliptic.h:428-458
 *   y = alpha/vol * [ -Lx sigma (chi_xx Rx + chi_xy Ry) x - Ly sigma (chi_yx Rx + chi_yy Ry) x
 *                     + jfactor (Jx + Jy) x ] + beta y
 * by ONE kernel that replays the reference's rounding sequence in registers (bit-identical result).
 * The six matrices are the ones Elliptic2d builds (elliptic.h:285-290); they must have the near-diagonal
 * block structure dx.h produces (checked at creation; other matrices -> DGB_ERR_UNSUPPORTED, use dgb_ell_symv).
 * Vector pointers set with the setters are BORROWED (the host layer owns them, like m_sigma/m_vol/m_chi).
 * ------------------------------------------------------------------------------------------------- */
typedef struct dgb_elliptic2d dgb_elliptic2d;
/* matrices are given on the HOST (the plan uploads what it needs) */
DGB_API int dgb_elliptic2d_create(dgb_elliptic2d** plan, const dgb_ell_host* leftx, const dgb_ell_host* lefty,
                                  const dgb_ell_host* rightx, const dgb_ell_host* righty, const dgb_ell_host* jumpx,
                                  const dgb_ell_host* jumpy, double jfactor, int chi_weight_jump);
DGB_API int dgb_elliptic2d_destroy(dgb_elliptic2d* plan);
DGB_API int dgb_elliptic2d_set_sigma(dgb_elliptic2d* plan, const double* sigma);  /* m_sigma (elliptic.h:327) */
DGB_API int dgb_elliptic2d_set_vol(dgb_elliptic2d* plan, const double* vol);      /* m_vol; NULL = 1 (Cartesian) */
DGB_API int dgb_elliptic2d_set_chi(dgb_elliptic2d* plan, const double* xx, const double* xy, const double* yx,
                                   const double* yy);                             /* m_chi; NULL = identity entry */
DGB_API int dgb_elliptic2d_set_jfactor(dgb_elliptic2d* plan, double jfactor);
/* Which kernel applies the operator (and runs inside dgb_pcg_solve_* / dgb_multigrid2d_solve on this plan).  AUTO picks
 * by problem size (walker from ~400^2 cells on, tile kernel below, the unfused composition for matrices / chi tensors the
 * fused kernels do not cover); the explicit modes exist so that every kernel can be parity-tested on any grid and A/B-timed.
 * All modes give bitwise the same result.  Errors: DGB_ERR_UNSUPPORTED if the plan cannot run on that kernel. */
#define DGB_ELLIPTIC_KERNEL_AUTO 0
#define DGB_ELLIPTIC_KERNEL_TILE 1     /* elliptic2d_fused_kernel: persistent CTAs, TMA-staged cell tiles            */
#define DGB_ELLIPTIC_KERNEL_WALKER 2   /* elliptic2d_walker_kernel: warp-private TMA sliding window (n = 2, 3)       */
#define DGB_ELLIPTIC_KERNEL_UNFUSED 3  /* the reference's launch sequence elliptic.h:431-458 on dgb_ell_symv / blas1 */
DGB_API int dgb_elliptic2d_set_kernel(dgb_elliptic2d* plan, int kernel);
/* Operation order of the fused kernels.  REFERENCE (default): the rounding sequence of the reference's OpenMP backend, result
 * bitwise equal to it.  RELAXED (opt-in): interior rows of the walker kernel accumulate every output in one fused-multiply-add
 * chain (27 % fewer FP64 operations; the kernel is FP64-pipe bound); results agree with the reference to <= 1e-13 relative,
 * inside the 1e-12 the reference's own tests allow for symv, but NOT bit for bit -- PCG iteration counts may differ by a few. */
#define DGB_ORDER_REFERENCE 0
#define DGB_ORDER_RELAXED 1
DGB_API int dgb_elliptic2d_set_ordering(dgb_elliptic2d* plan, int ordering);
/* *kernel = the DGB_ELLIPTIC_KERNEL_* the next symv (with_dot = 0) / PCG iteration (with_dot = 1) on this plan will launch */
DGB_API int dgb_elliptic2d_get_kernel(const dgb_elliptic2d* plan, int with_dot, int* kernel);
/* test hook, host only: the work partition the warp-walker kernel would use (tasks_out: column, first row, end row per piece) */
DGB_API int dgb_debug_walker_partition(int Nx, int Ny, int centered, int nwarps, int fx_lo, int fx_hi, int wrapx, int tma, int dot,
                                       int* tasks_out, int max_tasks, int* ntasks, int* tbegin_out);
/* dg::GeneralHelmholtz<Elliptic2d> (helmholtz.h:27-82): with enable != 0 the two-operand symv of the plan (and with it
 * PCG / MultigridCG2d on the plan) computes  y = chi x - alpha (Elliptic x)  exactly as
 * `symv(m_matrix, x, y); pointwiseDot(1., m_chi, x, -m_alpha, y)`; chi == NULL is the default chi = 1. */
DGB_API int dgb_elliptic2d_set_helmholtz(dgb_elliptic2d* plan, int enable, double alpha, const double* chi);
/* Elliptic2d::variation (elliptic.h:497-502): sigma = alpha lambda^2 (grad phi . chi . grad phi) + beta sigma with the
 * plan's right derivatives and chi tensor; lambda == NULL means 1 */
DGB_API int dgb_elliptic2d_variation(dgb_elliptic2d* plan, double alpha, const double* lambda, const double* phi, double beta,
                                     double* sigma, dgb_stream_t s);
/* size of the operator; *fused = 1 if the one-pass kernel applies to these matrices (else the composition runs) */
DGB_API int dgb_elliptic2d_size(const dgb_elliptic2d* plan, size_t* size, int* fused);
DGB_API int dgb_elliptic2d_symv(dgb_elliptic2d* plan, double alpha, const double* x, double beta, double* y,
                                dgb_stream_t s);
/* the unfused composition on the same plan (6 Ell symv + 2 blas1), kept for A/B tests and as the general path */
/* dg::Elliptic3d with set_compute_in_2d(true) (elliptic.h:557-797,677): nplanes applications of the 2-d operator on the
 * consecutive planes of 3-d vectors x, y.  sigma3d = chi*vol on every plane (what Elliptic3d::set_chi stores in m_sigma);
 * a Helmholtz chi set on the plan is likewise read per plane; vol / the chi tensor of the plan are 2-d fields. */
DGB_API int dgb_elliptic2d_symv_planes(dgb_elliptic2d* plan, int nplanes, const double* sigma3d, double alpha,
                                       const double* x, double beta, double* y, dgb_stream_t s);
DGB_API int dgb_elliptic2d_symv_unfused(dgb_elliptic2d* plan, double alpha, const double* x, double beta, double* y,
                                        dgb_stream_t s);

/* ---------------------------------------------------------------------------------------------------
 * PCG: replaces PCG<DVec>::solve inc/dg/pcg.h:136-195 for A = Elliptic2d plan, vector preconditioner P and
 * weights W.  Same recurrences, same exact dots, alpha/beta from the rounded doubles; all scalars stay on the
 * device, the host only polls a convergence flag every `check_every` iterations.  *iterations receives what the
 * reference returns (max_iter when not converged -> also returns DGB_ERR_NOCONVERGE).
 * ------------------------------------------------------------------------------------------------- */
typedef struct dgb_pcg dgb_pcg;
DGB_API int dgb_pcg_create(dgb_pcg** pcg, size_t n);
DGB_API int dgb_pcg_destroy(dgb_pcg* pcg);
DGB_API int dgb_pcg_solve_elliptic2d(dgb_pcg* pcg, dgb_elliptic2d* A, double* x, const double* b, const double* P,
                                     const double* W, double eps, double nrmb_correction, int test_frequency,
                                     int max_iter, int* iterations, dgb_stream_t s);

/* ---------------------------------------------------------------------------------------------------
 * MultigridCG2d: replaces NestedGrids + nested_iterations + MultigridCG2d::solve inc/dg/multigrid.h:28-171,197-245,
 * 500-668.  The caller owns one Elliptic2d plan per stage (as multi_pol[u].construct(multigrid.grid(u), ...) does in
 * src/toefl/toefl.h) and passes each stage's preconditioner / weights vectors (pol.precond(), pol.weights()).
 * ------------------------------------------------------------------------------------------------- */
typedef struct dgb_multigrid2d dgb_multigrid2d;
DGB_API int dgb_multigrid2d_create(dgb_multigrid2d** mg, const dgb_grid* grid, int stages);
DGB_API int dgb_multigrid2d_destroy(dgb_multigrid2d* mg);
DGB_API int dgb_multigrid2d_stages(const dgb_multigrid2d* mg);
/* A/B switch for tests: 1 = projections / interpolations (dg::MultiMatrix::symv, fast_interpolation.h:71-84) run as the
 * reference's two block-matrix products through a temporary, 0 (default) = the one-pass factor-2 kernels (bitwise equal) */
DGB_API int dgb_multigrid2d_set_two_pass(int on);
DGB_API int dgb_multigrid2d_grid(const dgb_multigrid2d* mg, int stage, dgb_grid* grid, size_t* size); /* multigrid.h:128 */
/* project(src, out): out = HOST array of `stages` device vectors (multigrid.h:94-99) */
DGB_API int dgb_multigrid2d_project(dgb_multigrid2d* mg, const double* src, double* const* out, dgb_stream_t s);
/* xf = alpha * interpolation(coarse_stage-1) xc + beta xf (multigrid.h:232) */
DGB_API int dgb_multigrid2d_interpolate(dgb_multigrid2d* mg, int coarse_stage, double alpha, const double* xc, double beta,
                                        double* xf, dgb_stream_t s);
/* solve(ops, x, b, eps[stages]) -> numbers[stages] = PCG iterations per stage (multigrid.h:627-658); ops, precond,
 * weights are HOST arrays of length `stages` */
DGB_API int dgb_multigrid2d_solve(dgb_multigrid2d* mg, dgb_elliptic2d* const* ops, const double* const* precond,
                                  const double* const* weights, double* x, const double* b, const double* eps,
                                  int* numbers, dgb_stream_t s);

/* ---------------------------------------------------------------------------------------------------
 * Multi-GPU (one process per GPU, NCCL over NVLink/NVSwitch): replaces MPI_Vector / MPISparseBlockMat /
 * MPIKroneckerGather / reduce_mpi_cpu (inc/dg/backend/mpi_vector.h, mpi_matrix.h:183-217, mpi_gather_kron.h:143-291,
 * exblas/mpi_accumulate.h:42-125) for a decomposition of the 2-d grid into slabs of cell rows (y direction).
 * ------------------------------------------------------------------------------------------------- */
typedef struct dgb_comm dgb_comm;
DGB_API int dgb_comm_unique_id(char* id128);                 /* ncclGetUniqueId on rank 0; ship the 128 bytes to the others */
DGB_API int dgb_comm_create(dgb_comm** comm, const char* id128, int rank, int nranks);
DGB_API int dgb_comm_destroy(dgb_comm* comm);
/* *peer_memory = 1 if the data plane of this communicator is CUDA-IPC peer memory (dot records and halo rows stored straight
 * into the peers' buffers over NVLink), 0 if it is NCCL (ncclSend/Recv halo, ncclAllReduce of the int64 records) */
DGB_API int dgb_comm_info(const dgb_comm* comm, int* rank, int* size, int* peer_memory);
/* halo exchange of `ghost_rows` rows with the lower/upper neighbour; buffer layout [ghost | nrows | ghost], `interior`
 * points to the first interior row; periodic closes the ring (mpi_gather_kron.h global_gather_init/wait) */
DGB_API int dgb_comm_halo_rows(dgb_comm* comm, double* interior, size_t row_len, size_t nrows, size_t ghost_rows,
                               int periodic, dgb_stream_t s);
/* MPIGather exchange of packed buffers (mpi_gather.h:454-705): this rank sends send_counts[p] doubles to rank p and receives
 * recv_counts[p] from it, both buffers ordered by rank; the counts are HOST arrays of `size` ints.  One grouped
 * ncclSend / ncclRecv round on stream s (the message to itself is a device copy; a size-1 communicator needs no NCCL). */
DGB_API int dgb_comm_gather(dgb_comm* comm, const double* send, const int* send_counts, double* recv, const int* recv_counts,
                            dgb_stream_t s);
/* exact global dot: sums the normalised superaccumulators of `nrecords` device records over all ranks (integer
 * allreduce), renormalises and rounds on the device (blas1_dispatch_mpi.h:91-109) */
DGB_API int dgb_comm_allreduce_dot(dgb_comm* comm, dgb_dot_result* result_dev, int nrecords, dgb_stream_t s);
/* declare the plan (built from the GLOBAL matrices) to act on the slab of cell rows [yoff, yoff+rows) whose x and
 * sigma operands carry `ghost` cell rows on either side; vectors passed to symv/PCG then have rows*n*Nx*n elements */
DGB_API int dgb_elliptic2d_set_slab(dgb_elliptic2d* plan, int yoff, int rows, int ghost);
/* PCG on the decomposed problem: same arithmetic, halo exchange of the search direction and integer allreduce of the
 * three dots per iteration; the result is bit-identical to the single-GPU solve for any number of ranks */
DGB_API int dgb_pcg_solve_elliptic2d_dist(dgb_pcg* pcg, dgb_comm* comm, dgb_elliptic2d* A, double* x, const double* b,
                                          const double* P, const double* W, double eps, double nrmb_correction,
                                          int test_frequency, int max_iter, int* iterations, dgb_stream_t s);

/* measurement aid: when on, CUDA events on the launching stream bracket the three kernels of every iteration
 * (operator+dot, update+dots, direction); get returns the accumulated milliseconds and the iterations timed */
DGB_API int dgb_pcg_set_profile(dgb_pcg* pcg, int on);
DGB_API int dgb_pcg_get_profile(dgb_pcg* pcg, double* ms_apply_dot, double* ms_update, double* ms_direction,
                                long long* iterations);
/* *folded = 1 if the last solve on this workspace ran the two-kernel iteration: the direction update p = z + beta p
 * (pcg.h:182) formed inside the operator kernel's loader (walker kernel, TMA operands) instead of a third launch;
 * the arithmetic and the results are the same either way.  DGB_PCG_NO_FOLD=1 in the environment forces three kernels. */
DGB_API int dgb_pcg_last_folded(dgb_pcg* pcg, int* folded);

#ifdef __cplusplus
}
#endif
#endif /* DGB200_H */
