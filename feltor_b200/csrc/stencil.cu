// dg::blas2::stencil / parallel_for (inc/dg/blas2.h:413-454, backend/blas2_stencil.h:13-70) for the library's CSR stencil
// functors CSRMedianFilter, CSRSWMFilter, CSRAverageFilter, CSRSymvFilter (inc/dg/topology/filter.h:84-266).
// One thread per row.  The (lower) median is the rank-(n+1)/2 element of the stencil values; it is found by counting
// ranks (no scratch memory, stencils are 3..25 points), which selects the same element as the reference's networks /
// bisection.  The matrix values are ignored by all but the symv filter, exactly as in the reference.
#include "common.cuh"

namespace dgb {

enum { ST_MEDIAN = 0, ST_SWM = 1, ST_AVERAGE = 2, ST_SYMV = 3 };

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
    if (!row_offsets || !cols || !x || !y || (kind == ST_SYMV && !vals)) { set_error("dgb_csr_stencil: missing operand"); return DGB_ERR_INVALID; }
    if (x == y) { set_error("dgb_csr_stencil: x must not alias y"); return DGB_ERR_INVALID; }
    unsigned grid = (unsigned)((num_rows + 127) / 128);
    cudaStream_t st = as_stream(s);
    switch (kind) {
        case ST_MEDIAN: csr_stencil_kernel<ST_MEDIAN><<<grid, 128, 0, st>>>(num_rows, row_offsets, cols, vals, alpha, x, y); break;
        case ST_SWM: csr_stencil_kernel<ST_SWM><<<grid, 128, 0, st>>>(num_rows, row_offsets, cols, vals, alpha, x, y); break;
        case ST_AVERAGE: csr_stencil_kernel<ST_AVERAGE><<<grid, 128, 0, st>>>(num_rows, row_offsets, cols, vals, alpha, x, y); break;
        case ST_SYMV: csr_stencil_kernel<ST_SYMV><<<grid, 128, 0, st>>>(num_rows, row_offsets, cols, vals, alpha, x, y); break;
        default: set_error("dgb_csr_stencil: unknown kind %d", kind); return DGB_ERR_INVALID;
    }
    DGB_LAUNCHED();
    return 0;
}
