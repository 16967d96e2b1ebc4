// blas1: fused elementwise FP64 kernels (128-bit vectorised, front-batched loads for memory-level parallelism).
// Replaces doSubroutine_dispatch(CudaTag,...) of the reference (inc/dg/backend/blas1_cuda.cuh:76-94) for the
// closed set of library functors (inc/dg/subroutines.h:231-384, inc/dg/topology/multiply.h:18-32).
// The arithmetic of every functor is written with explicit __dmul_rn/__fma_rn/__dadd_rn so that the rounding
// sequence is exactly the reference functor's (the library is also compiled with -fmad=false).
#include "common.cuh"

namespace dgb {

template <int NV>
struct Pack {
    double* p[NV];
};

// One "unit" = one double2 (vector path) or one double (scalar path) per array.
// F: static constexpr int NV; unsigned RMASK (arrays read), WMASK (arrays written);
//    __device__ void operator()(double (&v)[NV]) const
template <class F, int UNROLL>
__global__ void __launch_bounds__(256) ew_vec2_kernel(F f, Pack<F::NV> pk, size_t nvec, size_t n) {
    constexpr int NV = F::NV;
    const size_t T = (size_t)gridDim.x * blockDim.x;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (size_t base = tid; base < nvec; base += (size_t)UNROLL * T) {
        double2 v[UNROLL][NV];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            size_t idx = base + (size_t)u * T;
            if (idx < nvec) {
#pragma unroll
                for (int k = 0; k < NV; k++)
                    if (((F::RMASK >> k) & 1u) && pk.p[k]) v[u][k] = ld2(pk.p[k] + 2 * idx);
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            size_t idx = base + (size_t)u * T;
            if (idx < nvec) {
                double a[NV], b[NV];
#pragma unroll
                for (int k = 0; k < NV; k++) { a[k] = v[u][k].x; b[k] = v[u][k].y; }
                f(a);
                f(b);
#pragma unroll
                for (int k = 0; k < NV; k++)
                    if ((F::WMASK >> k) & 1u) st2(pk.p[k] + 2 * idx, make_double2(a[k], b[k]));
            }
        }
    }
    // odd tail
    if ((n & 1) && tid == 0) {
        double a[NV];
#pragma unroll
        for (int k = 0; k < NV; k++)
            if (((F::RMASK >> k) & 1u) && pk.p[k]) a[k] = pk.p[k][n - 1];
        f(a);
#pragma unroll
        for (int k = 0; k < NV; k++)
            if ((F::WMASK >> k) & 1u) pk.p[k][n - 1] = a[k];
    }
}

template <class F, int UNROLL>
__global__ void __launch_bounds__(256) ew_scalar_kernel(F f, Pack<F::NV> pk, size_t n) {
    constexpr int NV = F::NV;
    const size_t T = (size_t)gridDim.x * blockDim.x;
    const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (size_t base = tid; base < n; base += (size_t)UNROLL * T) {
        double v[UNROLL][NV];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            size_t idx = base + (size_t)u * T;
            if (idx < n) {
#pragma unroll
                for (int k = 0; k < NV; k++)
                    if (((F::RMASK >> k) & 1u) && pk.p[k]) v[u][k] = pk.p[k][idx];
            }
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            size_t idx = base + (size_t)u * T;
            if (idx < n) {
                f(v[u]);
#pragma unroll
                for (int k = 0; k < NV; k++)
                    if ((F::WMASK >> k) & 1u) pk.p[k][idx] = v[u][k];
            }
        }
    }
}

template <class F>
int launch_ew(F f, Pack<F::NV> pk, size_t n, dgb_stream_t s) {
    if (n == 0) return 0;
    constexpr int UNROLL = F::NV <= 3 ? 4 : (F::NV <= 6 ? 2 : 1);
    bool vec = true;
    for (int k = 0; k < F::NV; k++)
        if (pk.p[k] && !aligned16(pk.p[k])) vec = false;
    const int sms = sm_count();
    if (vec) {
        size_t nvec = n / 2;
        size_t want = (nvec + 256ull * UNROLL - 1) / (256ull * UNROLL);
        if (want == 0) want = 1;
        size_t cap = (size_t)sms * 8;
        unsigned grid = (unsigned)(want < cap ? want : cap);
        ew_vec2_kernel<F, UNROLL><<<grid, 256, 0, as_stream(s)>>>(f, pk, nvec, n);
    } else {
        size_t want = (n + 256ull * UNROLL - 1) / (256ull * UNROLL);
        size_t cap = (size_t)sms * 8;
        unsigned grid = (unsigned)(want < cap ? want : cap);
        ew_scalar_kernel<F, UNROLL><<<grid, 256, 0, as_stream(s)>>>(f, pk, n);
    }
    DGB_LAUNCHED();
    return 0;
}

// ------------------------------------------------------------------------------------------------ functors
struct FCopy {  // v0 = x (in), v1 = y (out)
    static constexpr int NV = 2; static constexpr unsigned RMASK = 1, WMASK = 2;
    __device__ void operator()(double (&v)[2]) const { v[1] = v[0]; }
};
struct FFill {
    static constexpr int NV = 1; static constexpr unsigned RMASK = 0, WMASK = 1;
    double a;
    __device__ void operator()(double (&v)[1]) const { v[0] = a; }
};
struct FScal {  // subroutines.h:233
    static constexpr int NV = 1; static constexpr unsigned RMASK = 1, WMASK = 1;
    double a;
    __device__ void operator()(double (&v)[1]) const { v[0] = __dmul_rn(v[0], a); }
};
struct FPlus {  // subroutines.h:247
    static constexpr int NV = 1; static constexpr unsigned RMASK = 1, WMASK = 1;
    double a;
    __device__ void operator()(double (&v)[1]) const { v[0] = __dadd_rn(v[0], a); }
};
struct FAxpby {  // subroutines.h:260: y *= b; y = fma(a,x,y)
    static constexpr int NV = 2; static constexpr unsigned RMASK = 3, WMASK = 2;
    double a, b;
    __device__ void operator()(double (&v)[2]) const { v[1] = __fma_rn(a, v[0], __dmul_rn(v[1], b)); }
};
struct FAxpbyz {  // PairSum subroutines.h:124: fma(a, x, b*y)
    static constexpr int NV = 3; static constexpr unsigned RMASK = 3, WMASK = 4;
    double a, b;
    __device__ void operator()(double (&v)[3]) const { v[2] = __fma_rn(a, v[0], __dmul_rn(b, v[1])); }
};
struct FAxpbypgz {  // subroutines.h:294
    static constexpr int NV = 3; static constexpr unsigned RMASK = 7, WMASK = 4;
    double a, b, g;
    __device__ void operator()(double (&v)[3]) const {
        double z = __dmul_rn(v[2], g);
        z = __fma_rn(a, v[0], z);
        v[2] = __fma_rn(b, v[1], z);
    }
};
struct FPointwiseDot {  // subroutines.h:313: z *= b; z = fma(a*x, y, z)
    static constexpr int NV = 3; static constexpr unsigned RMASK = 7, WMASK = 4;
    double a, b;
    __device__ void operator()(double (&v)[3]) const {
        v[2] = __fma_rn(__dmul_rn(a, v[0]), v[1], __dmul_rn(v[2], b));
    }
};
struct FAxyPby {  // subroutines.h:276: tmp = y; y *= b; y = fma(a*x, tmp, y)
    static constexpr int NV = 2; static constexpr unsigned RMASK = 3, WMASK = 2;
    double a, b;
    __device__ void operator()(double (&v)[2]) const {
        double tmp = v[1];
        v[1] = __fma_rn(__dmul_rn(a, v[0]), tmp, __dmul_rn(tmp, b));
    }
};
struct FMulXY {  // blas1.h:441
    static constexpr int NV = 3; static constexpr unsigned RMASK = 3, WMASK = 4;
    __device__ void operator()(double (&v)[3]) const { v[2] = __dmul_rn(v[0], v[1]); }
};
struct FPointwiseDot3 {  // subroutines.h:325: y *= b; y = fma(a*x1, x2*x3, y)
    static constexpr int NV = 4; static constexpr unsigned RMASK = 15, WMASK = 8;
    double a, b;
    __device__ void operator()(double (&v)[4]) const {
        v[3] = __fma_rn(__dmul_rn(a, v[0]), __dmul_rn(v[1], v[2]), __dmul_rn(v[3], b));
    }
};
struct FPointwiseDot2 {  // subroutines.h:336
    static constexpr int NV = 5; static constexpr unsigned RMASK = 31, WMASK = 16;
    double a, b, g;
    __device__ void operator()(double (&v)[5]) const {
        double z = __dmul_rn(v[4], g);
        z = __fma_rn(__dmul_rn(a, v[0]), v[1], z);
        v[4] = __fma_rn(__dmul_rn(b, v[2]), v[3], z);
    }
};
struct FPointwiseDivide {  // subroutines.h:376: z *= b; z = fma(a, x/y, z)
    static constexpr int NV = 3; static constexpr unsigned RMASK = 7, WMASK = 4;
    double a, b;
    __device__ void operator()(double (&v)[3]) const {
        v[2] = __fma_rn(a, __ddiv_rn(v[0], v[1]), __dmul_rn(v[2], b));
    }
};
struct FPointwiseDivideAlias {  // subroutines.h:369: tmp = z; z *= b; z = fma(a, tmp/y, z)
    static constexpr int NV = 2; static constexpr unsigned RMASK = 3, WMASK = 2;
    double a, b;
    __device__ void operator()(double (&v)[2]) const {
        double tmp = v[1];
        v[1] = __fma_rn(a, __ddiv_rn(tmp, v[0]), __dmul_rn(tmp, b));
    }
};
struct FDivXY {
    static constexpr int NV = 3; static constexpr unsigned RMASK = 3, WMASK = 4;
    __device__ void operator()(double (&v)[3]) const { v[2] = __ddiv_rn(v[0], v[1]); }
};
struct FTensorMul2d {  // multiply.h:18-32; v = {lambda,t00,t01,t10,t11,in0,in1,out0,out1}
    static constexpr int NV = 9; static constexpr unsigned RMASK = 0x1ff, WMASK = 0x180;
    double lambda_s, mu;
    unsigned present;  // bit k set: array k present, else implicit constant
    __device__ void operator()(double (&v)[9]) const {
        double l = (present & 1u) ? v[0] : lambda_s;
        double t00 = (present & 2u) ? v[1] : 1., t01 = (present & 4u) ? v[2] : 0.;
        double t10 = (present & 8u) ? v[3] : 0., t11 = (present & 16u) ? v[4] : 1.;
        double tmp0 = __fma_rn(t00, v[5], __dmul_rn(t01, v[6]));
        double tmp1 = __fma_rn(t10, v[5], __dmul_rn(t11, v[6]));
        double temp = __dmul_rn(v[8], mu);
        v[8] = __fma_rn(l, tmp1, temp);
        temp = __dmul_rn(v[7], mu);
        v[7] = __fma_rn(l, tmp0, temp);
    }
};
struct FTolerance {  // detail::Tolerance (adaptive.h:123-134): delta = delta / (rtol*|u0| + atol); v = {u0, delta}
    static constexpr int NV = 2; static constexpr unsigned RMASK = 3, WMASK = 2;
    double rtol, atol;
    __device__ void operator()(double (&v)[2]) const { v[1] = __ddiv_rn(v[1], __fma_rn(rtol, fabs(v[0]), atol)); }
};
struct FTensorMul3d {  // multiply.h:34-58; v = {lambda, t00..t22 (row major), in0,in1,in2, out0,out1,out2}
    static constexpr int NV = 16; static constexpr unsigned RMASK = 0xffff, WMASK = 0xe000;
    double lambda_s, mu;
    unsigned present;  // bit 0: lambda array, bit 1+k: tensor component k present, else the identity's constant
    __device__ void operator()(double (&v)[16]) const {
        double l = (present & 1u) ? v[0] : lambda_s;
        double t[9];
#pragma unroll
        for (int k = 0; k < 9; k++) t[k] = (present >> (1 + k)) & 1u ? v[1 + k] : (k % 4 == 0 ? 1. : 0.);
        double tmp0 = __fma_rn(t[0], v[10], __fma_rn(t[1], v[11], __dmul_rn(t[2], v[12])));
        double tmp1 = __fma_rn(t[3], v[10], __fma_rn(t[4], v[11], __dmul_rn(t[5], v[12])));
        double tmp2 = __fma_rn(t[6], v[10], __fma_rn(t[7], v[11], __dmul_rn(t[8], v[12])));
        double temp = __dmul_rn(v[15], mu);
        v[15] = __fma_rn(l, tmp2, temp);
        temp = __dmul_rn(v[14], mu);
        v[14] = __fma_rn(l, tmp1, temp);
        temp = __dmul_rn(v[13], mu);
        v[13] = __fma_rn(l, tmp0, temp);
    }
};
struct FArakawa {  // ArakawaFunctor (arakawa.h:125-145); v = {lhs, rhs, dxlhs, dylhs, dxrhs, dyrhs}, the last three are in/out
    static constexpr int NV = 6; static constexpr unsigned RMASK = 0x3f, WMASK = 0x38;
    __device__ void operator()(double (&v)[6]) const {
        const double third = 1. / 3., mthird = -(1. / 3.);
        const double lhs = v[0], rhs = v[1], dxlhs = v[2], dylhs = v[3], dxrhs = v[4], dyrhs = v[5];
        double result = 0.;
        result = __fma_rn(__dmul_rn(third, dxlhs), dyrhs, result);
        result = __fma_rn(__dmul_rn(mthird, dylhs), dxrhs, result);
        double temp = 0.;
        temp = __fma_rn(__dmul_rn(third, lhs), dyrhs, temp);
        v[5] = result;
        temp = __fma_rn(__dmul_rn(mthird, dylhs), rhs, temp);
        v[3] = temp;
        temp = 0.;
        temp = __fma_rn(__dmul_rn(third, dxlhs), rhs, temp);
        temp = __fma_rn(__dmul_rn(mthird, lhs), dxrhs, temp);
        v[4] = temp;
    }
};
struct FUpwindAxpby {  // evaluate(y, Axpby(a,b), UpwindProduct(), v, back, forw): advection.h:112-120, functors.h:312-337
    static constexpr int NV = 4; static constexpr unsigned RMASK = 0xf, WMASK = 0x8;
    double a, b;
    __device__ void operator()(double (&v)[4]) const {
        const double up = __dmul_rn(v[0], v[0] >= 0. ? v[1] : v[2]);
        v[3] = __fma_rn(a, up, __dmul_rn(v[3], b));
    }
};
struct FTensorDot2dAxpby {  // scalar_product2d (multiply.h:493-512, TensorDot2d :135-150); v = {lambda,v0,v1,t00,t01,t10,t11,mu,w0,w1,y}
    static constexpr int NV = 11; static constexpr unsigned RMASK = 0x7ff, WMASK = 0x400;
    double a, b, lambda_s, mu_s;
    unsigned present;  // bit k set: array k present, else implicit constant (lambda_s, identity tensor, mu_s)
    __device__ void operator()(double (&v)[11]) const {
        const double l = (present & 1u) ? v[0] : lambda_s, m = (present & 128u) ? v[7] : mu_s;
        const double t00 = (present & 8u) ? v[3] : 1., t01 = (present & 16u) ? v[4] : 0.;
        const double t10 = (present & 32u) ? v[5] : 0., t11 = (present & 64u) ? v[6] : 1.;
        const double tmp0 = __fma_rn(t00, v[8], __dmul_rn(t01, v[9]));
        const double tmp1 = __fma_rn(t10, v[8], __dmul_rn(t11, v[9]));
        const double r = __dmul_rn(__dmul_rn(l, m), __fma_rn(v[1], tmp0, __dmul_rn(v[2], tmp1)));
        v[10] = __fma_rn(a, r, __dmul_rn(v[10], b));
    }
};
template <int OP>
struct FUnary {
    static constexpr int NV = 2; static constexpr unsigned RMASK = 1, WMASK = 2;
    __device__ void operator()(double (&v)[2]) const {
        double x = v[0];
        if (OP == DGB_OP_EXP) v[1] = exp(x);
        else if (OP == DGB_OP_LN) v[1] = log(x);
        else if (OP == DGB_OP_SQRT) v[1] = __dsqrt_rn(x);
        else if (OP == DGB_OP_INVERT) v[1] = __ddiv_rn(1., x);
        else if (OP == DGB_OP_ABS) v[1] = fabs(x);
        else if (OP == DGB_OP_SQUARE) v[1] = __dmul_rn(x, x);
        else v[1] = __ddiv_rn(1., __dsqrt_rn(x));
    }
};

// EmbeddedPairSum (subroutines.h:179-204) with up to 16 stage vectors
struct EpsParams {
    const double* k[16];
    double b[16], bt[16];
    double b0, bt0;
    int nk;
};
__global__ void __launch_bounds__(256) embedded_pair_sum_kernel(EpsParams P, double* __restrict__ y,
                                                                double* __restrict__ yt, size_t n) {
    const size_t T = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += T) {
        double a = __dmul_rn(P.b0, y[i]), at = __dmul_rn(P.bt0, yt[i]);
        for (int s = 0; s < P.nk; s++) {
            double ks = P.k[s][i];
            a = __fma_rn(P.b[s], ks, a);
            at = __fma_rn(P.bt[s], ks, at);
        }
        y[i] = a;
        yt[i] = at;
    }
}

// evaluate(y, Axpby(alpha, beta), PairSum(), a_0, x_0, ..., a_{nk-1}, x_{nk-1}) (subroutines.h:123-143, 260-274):
// y = fma(alpha, fma(a_0, x_0, fma(a_1, x_1, ... a_last * x_last)), y * beta) -- the dense-matrix gemv of the time steppers
// (blas2_densematrix.h:38-74)
__global__ void __launch_bounds__(256) pair_sum_axpby_kernel(EpsParams P, double alpha, double beta, double* __restrict__ y, size_t n) {
    const size_t T = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += T) {
        double sum = __dmul_rn(P.b[P.nk - 1], P.k[P.nk - 1][i]);
        for (int s = P.nk - 2; s >= 0; s--) sum = __fma_rn(P.b[s], P.k[s][i], sum);
        y[i] = __fma_rn(alpha, sum, __dmul_rn(y[i], beta));
    }
}

template <int NV>
Pack<NV> pack(std::initializer_list<const double*> l) {
    Pack<NV> pk;
    int k = 0;
    for (auto p : l) pk.p[k++] = const_cast<double*>(p);
    return pk;
}

}  // namespace dgb

using namespace dgb;

extern "C" {

int dgb_copy(size_t n, const double* x, double* y, dgb_stream_t s) {
    if (x == y) return 0;  // blas1.h:244
    return launch_ew(FCopy{}, pack<2>({x, y}), n, s);
}
int dgb_fill(size_t n, double value, double* y, dgb_stream_t s) { return launch_ew(FFill{value}, pack<1>({y}), n, s); }
int dgb_scal(size_t n, double* x, double alpha, dgb_stream_t s) { return launch_ew(FScal{alpha}, pack<1>({x}), n, s); }
int dgb_plus(size_t n, double* x, double alpha, dgb_stream_t s) { return launch_ew(FPlus{alpha}, pack<1>({x}), n, s); }
int dgb_axpby(size_t n, double alpha, const double* x, double beta, double* y, dgb_stream_t s) {
    return launch_ew(FAxpby{alpha, beta}, pack<2>({x, y}), n, s);
}
int dgb_axpbyz(size_t n, double alpha, const double* x, double beta, const double* y, double* z, dgb_stream_t s) {
    return launch_ew(FAxpbyz{alpha, beta}, pack<3>({x, y, z}), n, s);
}
int dgb_axpbypgz(size_t n, double alpha, const double* x, double beta, const double* y, double gamma, double* z,
                 dgb_stream_t s) {
    return launch_ew(FAxpbypgz{alpha, beta, gamma}, pack<3>({x, y, z}), n, s);
}
int dgb_pointwise_dot(size_t n, double alpha, const double* x1, const double* x2, double beta, double* y,
                      dgb_stream_t s) {
    if (x1 == y) return launch_ew(FAxyPby{alpha, beta}, pack<2>({x2, y}), n, s);  // blas1.h:413
    if (x2 == y) return launch_ew(FAxyPby{alpha, beta}, pack<2>({x1, y}), n, s);  // blas1.h:418
    return launch_ew(FPointwiseDot{alpha, beta}, pack<3>({x1, x2, y}), n, s);
}
int dgb_pointwise_dot_xy(size_t n, const double* x1, const double* x2, double* y, dgb_stream_t s) {
    return launch_ew(FMulXY{}, pack<3>({x1, x2, y}), n, s);
}
int dgb_pointwise_dot3(size_t n, double alpha, const double* x1, const double* x2, const double* x3, double beta,
                       double* y, dgb_stream_t s) {
    return launch_ew(FPointwiseDot3{alpha, beta}, pack<4>({x1, x2, x3, y}), n, s);
}
int dgb_pointwise_dot2(size_t n, double alpha, const double* x1, const double* y1, double beta, const double* x2,
                       const double* y2, double gamma, double* z, dgb_stream_t s) {
    return launch_ew(FPointwiseDot2{alpha, beta, gamma}, pack<5>({x1, y1, x2, y2, z}), n, s);
}
int dgb_pointwise_divide(size_t n, double alpha, const double* x1, const double* x2, double beta, double* y,
                         dgb_stream_t s) {
    if (x1 == y) return launch_ew(FPointwiseDivideAlias{alpha, beta}, pack<2>({x2, y}), n, s);  // blas1.h:501
    return launch_ew(FPointwiseDivide{alpha, beta}, pack<3>({x1, x2, y}), n, s);
}
int dgb_pointwise_divide_xy(size_t n, const double* x1, const double* x2, double* y, dgb_stream_t s) {
    return launch_ew(FDivXY{}, pack<3>({x1, x2, y}), n, s);
}
int dgb_tensor_multiply2d(size_t n, const double* lambda, double lambda_s, const double* t00, const double* t01,
                          const double* t10, const double* t11, const double* in0, const double* in1, double mu,
                          double* out0, double* out1, dgb_stream_t s) {
    unsigned present = (lambda ? 1u : 0u) | (t00 ? 2u : 0u) | (t01 ? 4u : 0u) | (t10 ? 8u : 0u) | (t11 ? 16u : 0u);
    return launch_ew(FTensorMul2d{lambda_s, mu, present}, pack<9>({lambda, t00, t01, t10, t11, in0, in1, out0, out1}),
                     n, s);
}
int dgb_tensor_multiply3d(size_t n, const double* lambda, double lambda_s, const double* const t[9], const double* const in[3],
                          double mu, double* const out[3], dgb_stream_t s) {
    if (!in || !out || !in[0] || !in[1] || !in[2] || !out[0] || !out[1] || !out[2]) {
        set_error("dgb_tensor_multiply3d: in/out vectors required");
        return DGB_ERR_INVALID;
    }
    unsigned present = lambda ? 1u : 0u;
    Pack<16> pk;
    pk.p[0] = const_cast<double*>(lambda);
    for (int k = 0; k < 9; k++) {
        pk.p[1 + k] = t ? const_cast<double*>(t[k]) : nullptr;
        if (pk.p[1 + k]) present |= 2u << k;
    }
    for (int k = 0; k < 3; k++) { pk.p[10 + k] = const_cast<double*>(in[k]); pk.p[13 + k] = out[k]; }
    return launch_ew(FTensorMul3d{lambda_s, mu, present}, pk, n, s);
}
int dgb_adaptive_tolerance(size_t n, double rtol, double atol, const double* u0, double* delta, dgb_stream_t s) {
    return launch_ew(FTolerance{rtol, atol}, pack<2>({u0, delta}), n, s);
}
int dgb_arakawa_functor(size_t n, const double* lhs, const double* rhs, const double* dxlhs, double* dylhs, double* dxrhs,
                        double* dyrhs, dgb_stream_t s) {
    return launch_ew(FArakawa{}, pack<6>({lhs, rhs, dxlhs, dylhs, dxrhs, dyrhs}), n, s);
}
int dgb_upwind_axpby(size_t n, double alpha, const double* v, const double* back, const double* forw, double beta, double* y,
                     dgb_stream_t s) {
    return launch_ew(FUpwindAxpby{alpha, beta}, pack<4>({v, back, forw, y}), n, s);
}
int dgb_tensor_dot2d(size_t n, double alpha, const double* lambda, double lambda_s, const double* v0, const double* v1,
                     const double* t00, const double* t01, const double* t10, const double* t11, const double* mu, double mu_s,
                     const double* w0, const double* w1, double beta, double* y, dgb_stream_t s) {
    unsigned present = (lambda ? 1u : 0u) | (t00 ? 8u : 0u) | (t01 ? 16u : 0u) | (t10 ? 32u : 0u) | (t11 ? 64u : 0u) | (mu ? 128u : 0u);
    return launch_ew(FTensorDot2dAxpby{alpha, beta, lambda_s, mu_s, present},
                     pack<11>({lambda, v0, v1, t00, t01, t10, t11, mu, w0, w1, y}), n, s);
}
int dgb_embedded_pair_sum(size_t n, double* y, double* yt, double b0, double bt0, int nk, const double* b,
                          const double* bt, const double* const* k, dgb_stream_t s) {
    if (nk < 0 || nk > 16) { set_error("dgb_embedded_pair_sum: nk=%d outside [0,16]", nk); return DGB_ERR_UNSUPPORTED; }
    if (n == 0) return 0;
    EpsParams P;
    for (int i = 0; i < nk; i++) { P.k[i] = k[i]; P.b[i] = b[i]; P.bt[i] = bt[i]; }
    P.b0 = b0; P.bt0 = bt0; P.nk = nk;
    size_t want = (n + 255) / 256, cap = (size_t)sm_count() * 8;
    embedded_pair_sum_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, as_stream(s)>>>(P, y, yt, n);
    DGB_LAUNCHED();
    return 0;
}
int dgb_pair_sum_axpby(size_t n, double alpha, int nk, const double* a, const double* const* x, double beta, double* y,
                       dgb_stream_t s) {
    if (nk < 1 || nk > 8) { set_error("dgb_pair_sum_axpby: nk = %d not in 1..8", nk); return DGB_ERR_INVALID; }
    if (n == 0) return 0;
    EpsParams P;
    for (int i = 0; i < nk; i++) { P.k[i] = x[i]; P.b[i] = a[i]; P.bt[i] = 0.; }
    P.b0 = 0.; P.bt0 = 0.; P.nk = nk;
    size_t want = (n + 255) / 256, cap = (size_t)sm_count() * 8;
    pair_sum_axpby_kernel<<<(unsigned)(want < cap ? want : cap), 256, 0, as_stream(s)>>>(P, alpha, beta, y, n);
    DGB_LAUNCHED();
    return 0;
}
int dgb_transform(size_t n, int op, const double* x, double* y, dgb_stream_t s) {
    auto pk = pack<2>({x, y});
    switch (op) {
        case DGB_OP_EXP: return launch_ew(FUnary<DGB_OP_EXP>{}, pk, n, s);
        case DGB_OP_LN: return launch_ew(FUnary<DGB_OP_LN>{}, pk, n, s);
        case DGB_OP_SQRT: return launch_ew(FUnary<DGB_OP_SQRT>{}, pk, n, s);
        case DGB_OP_INVERT: return launch_ew(FUnary<DGB_OP_INVERT>{}, pk, n, s);
        case DGB_OP_ABS: return launch_ew(FUnary<DGB_OP_ABS>{}, pk, n, s);
        case DGB_OP_SQUARE: return launch_ew(FUnary<DGB_OP_SQUARE>{}, pk, n, s);
        case DGB_OP_INVSQRT: return launch_ew(FUnary<DGB_OP_INVSQRT>{}, pk, n, s);
    }
    set_error("dgb_transform: unknown op %d", op);
    return DGB_ERR_INVALID;
}
}
