// dg::blas2::stencil / parallel_for (inc/dg/blas2.h:413-454, backend/blas2_stencil.h:13-70) for the library's CSR stencil
// functors CSRMedianFilter, CSRSWMFilter, CSRAverageFilter, CSRSymvFilter, CSRSlopeLimiter (inc/dg/topology/filter.h:84-336).
// One thread per row.  The (lower) median is the rank-(n+1)/2 element of the stencil values; it is found by counting
// ranks (no scratch memory, stencils are 3..25 points), which selects the same element as the reference's networks /
// bisection.  The matrix values are ignored by all but the symv filter, exactly as in the reference.
#include "common.cuh"

namespace dgb {

enum { ST_MEDIAN = 0, ST_SWM = 1, ST_AVERAGE = 2, ST_SYMV = 3, ST_SLOPE = 4 };

template <bool DEV>
__device__ __forceinline__ double stencil_value(const double* __restrict__ x, int col, double center) {
    double v = __ldg(x + col);
    return DEV ? fabs(__dsub_rn(v, center)) : v;
}
template <bool DEV>
__device__ double row_median(int b, int e, const int* __restrict__ idx, const double* __restrict__ x, double center) {
    const int rank = (e - b + 1) / 2;
    double v = 0.;
    for (int j = b; j < e; j++) {
        v = stencil_value<DEV>(x, __ldg(idx + j), center);
        int less = 0, equal = 0;
        for (int l = b; l < e; l++) {
            double u = stencil_value<DEV>(x, __ldg(idx + l), center);
            less += u < v;
            equal += u == v;
        }
        if (less < rank && rank <= less + equal) return v;
    }
    return v;
}

__device__ __forceinline__ double minmod2(double a, double b) {  // dg::MinMod (functors.h:255-285)
    if (a > 0. && b > 0.) return fmin(a, b);
    if (a < 0. && b < 0.) return fmax(a, b);
    return 0.;
}
// CSRSlopeLimiter (filter.h:288-336) on the matrix of create::limiter_stencil (stencil.h:89-137): the first row of every cell
// holds 3 n entries -- the modal transform rows towards the cell averages of the left neighbour, the cell, and the linear
// coefficient, and the back-transform column -- the other n - 1 rows are empty.  The thread of a non-empty row copies the cell's
// n values and replaces them by the limited linear polynomial when minmod picks a neighbour slope.  The sums are FMAs, as both
// of the reference's compilers contract `a += b*c` (gcc -mfma, nvcc -fmad=true); the back transform is pinned on the golden
// vectors of the reference's OpenMP build (tests/test_limiter.py, tests/test_gpu_core.py::test_slope_limiter).
__global__ void __launch_bounds__(128)
csr_slope_limiter_kernel(int num_rows, const int* __restrict__ pos, const int* __restrict__ idx, const double* __restrict__ val, double mod,
                         const double* __restrict__ x, double* __restrict__ y) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < num_rows; i += gridDim.x * blockDim.x) {
        const int k = __ldg(pos + i), n = (__ldg(pos + i + 1) - k) / 3;
        if (n == 0) continue;
        double uM = 0., u0 = 0., uP = 0., u1 = 0.;
        for (int u = 0; u < n; u++) {
            const double a = fabs(__ldg(val + k + u)), xc = __ldg(x + __ldg(idx + k + n + u));
            y[__ldg(idx + k + n + u)] = xc;
            uM = __fma_rn(__ldg(x + __ldg(idx + k + u)), a, uM);
            u0 = __fma_rn(xc, a, u0);
            u1 = __fma_rn(xc, __ldg(val + k + n + u), u1);
            uP = __fma_rn(__ldg(x + __ldg(idx + k + 2 * n + u)), a, uP);
        }
        const bool flip = __ldg(val + k + 2 * n) > 0.;
        if (__ldg(val + k) < 0.) uM = -uM;  // DIR boundary condition
        if (flip) uP = -uP;
        if (fabs(u1) <= mod) continue;
        const double m = minmod2(minmod2(u1, __dsub_rn(uP, u0)), __dsub_rn(u0, uM));
        if (m == u1) continue;
        for (int u = 0; u < n; u++) {  // the product is rounded on its own: the reference's host compiler shares it between the arms
            const double t = __dmul_rn(m, __ldg(val + k + 2 * n + u));
            y[__ldg(idx + k + n + u)] = flip ? __dsub_rn(u0, t) : __dadd_rn(u0, t);
        }
    }
}

template <int KIND>
__global__ void __launch_bounds__(128)
csr_stencil_kernel(int num_rows, const int* __restrict__ pos, const int* __restrict__ idx, const double* __restrict__ val,
                   double alpha, const double* __restrict__ x, double* __restrict__ y) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < num_rows; i += gridDim.x * blockDim.x) {
        const int b = __ldg(pos + i), e = __ldg(pos + i + 1);
        if (KIND == ST_MEDIAN) y[i] = row_median<false>(b, e, idx, x, 0.);
        else if (KIND == ST_SWM) {
            const double med = row_median<false>(b, e, idx, x, 0.);
            const double amd = row_median<true>(b, e, idx, x, med);
            const double xi = __ldg(x + i);
            y[i] = fabs(__dsub_rn(xi, med)) > __dmul_rn(alpha, amd) ? med : xi;
        } else if (KIND == ST_AVERAGE) {
            const double n = (double)(e - b);
            double t = 0.;
            for (int k = b; k < e; k++) t = __dadd_rn(t, __ddiv_rn(__ldg(x + __ldg(idx + k)), n));
            y[i] = t;
        } else {
            double t = 0.;
            for (int k = b; k < e; k++) t = __dadd_rn(t, __dmul_rn(__ldg(x + __ldg(idx + k)), __ldg(val + k)));
            y[i] = t;
        }
    }
}

}  // namespace dgb

using namespace dgb;

// x must not alias y (every row reads its neighbours' x)
extern "C" int dgb_csr_stencil(int kind, int num_rows, const int* row_offsets, const int* cols, const double* vals, double alpha,
                               const double* x, double* y, dgb_stream_t s) {
    if (num_rows == 0) return 0;
    if (!row_offsets || !cols || !x || !y || ((kind == ST_SYMV || kind == ST_SLOPE) && !vals)) { set_error("dgb_csr_stencil: missing operand"); return DGB_ERR_INVALID; }
    if (x == y) { set_error("dgb_csr_stencil: x must not alias y"); return DGB_ERR_INVALID; }
    unsigned grid = (unsigned)((num_rows + 127) / 128);
    cudaStream_t st = as_stream(s);
    switch (kind) {
        case ST_MEDIAN: csr_stencil_kernel<ST_MEDIAN><<<grid, 128, 0, st>>>(num_rows, row_offsets, cols, vals, alpha, x, y); break;
        case ST_SWM: csr_stencil_kernel<ST_SWM><<<grid, 128, 0, st>>>(num_rows, row_offsets, cols, vals, alpha, x, y); break;
        case ST_AVERAGE: csr_stencil_kernel<ST_AVERAGE><<<grid, 128, 0, st>>>(num_rows, row_offsets, cols, vals, alpha, x, y); break;
        case ST_SYMV: csr_stencil_kernel<ST_SYMV><<<grid, 128, 0, st>>>(num_rows, row_offsets, cols, vals, alpha, x, y); break;
        case ST_SLOPE: csr_slope_limiter_kernel<<<grid, 128, 0, st>>>(num_rows, row_offsets, cols, vals, alpha, x, y); break;
        default: set_error("dgb_csr_stencil: unknown kind %d", kind); return DGB_ERR_INVALID;
    }
    DGB_LAUNCHED();
    return 0;
}
