// minimal TMA 2-d float64 tile load test: nvcc -gencode arch=compute_100a,code=sm_100a -o tma_test tma_test.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int BR, int BC>
__global__ void k(const __grid_constant__ CUtensorMap map, double* out, int c0, int c1) {
    extern __shared__ __align__(128) double sm[];
    unsigned long long* bar = (unsigned long long*)(sm + ((BR * BC + 15) / 16) * 16);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(BR * BC * 8) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];\n" ::"r"(smem_u32(sm)),
                     "l"(&map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
    }
    unsigned ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(0) : "memory");
    } while (!ok);
    for (int i = threadIdx.x; i < BR * BC; i += blockDim.x) out[i] = sm[i];
}
typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
template <int BR, int BC>
int run(Fn fn, double* d, int rows, int ld, int c0, int c1, CUtensorMapDataType dt) {
    CUtensorMap m;
    cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
    cuuint32_t box[2] = {BC, BR};
    cuuint32_t es[2] = {1, 1};
    CUresult r = fn(&m, dt, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box %dx%d dt=%d encode=%d ", BR, BC, (int)dt, (int)r);
    if (r) { printf("\n"); return 1; }
    double* out;
    cudaMalloc(&out, BR * BC * 8);
    size_t smem = ((BR * BC + 15) / 16) * 16 * 8 + 16;
    cudaFuncSetAttribute(k<BR, BC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<BR, BC><<<1, 128, smem>>>(m, out, c0, c1);
    cudaError_t e = cudaDeviceSynchronize();
    printf("launch=%s ", cudaGetErrorString(e));
    if (e) { printf("\n"); return 2; }
    std::vector<double> h(BR * BC);
    cudaMemcpy(h.data(), out, BR * BC * 8, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r2 = 0; r2 < BR; r2++)
        for (int c = 0; c < BC; c++) {
            int gr = c1 + r2, gc = c0 + c;
            double want = (gr >= 0 && gr < rows && gc >= 0 && gc < ld) ? gr * 10000. + gc : 0.;
            if (h[r2 * BC + c] != want) bad++;
        }
    printf("mismatches=%d\n", bad);
    return 0;
}
int main() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaFree(0);
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    Fn fn = (Fn)p;
    int rows = 192, ld = 192;
    std::vector<double> h(rows * ld);
    for (int r = 0; r < rows; r++) for (int c = 0; c < ld; c++) h[r * ld + c] = r * 10000. + c;
    double* d;
    cudaMalloc(&d, rows * ld * 8);
    cudaMemcpy(d, h.data(), rows * ld * 8, cudaMemcpyHostToDevice);
    run<8, 32>(fn, d, rows, ld, 0, 0, CU_TENSOR_MAP_DATA_TYPE_FLOAT64);
    run<8, 32>(fn, d, rows, ld, -2, -3, CU_TENSOR_MAP_DATA_TYPE_FLOAT64);
    run<30, 102>(fn, d, rows, ld, 0, 0, CU_TENSOR_MAP_DATA_TYPE_FLOAT64);
    run<30, 104>(fn, d, rows, ld, -4, -3, CU_TENSOR_MAP_DATA_TYPE_FLOAT64);
    run<30, 102>(fn, d, rows, ld, 94, 165, CU_TENSOR_MAP_DATA_TYPE_FLOAT64);
    run<30, 96>(fn, d, rows, ld, 0, 0, CU_TENSOR_MAP_DATA_TYPE_FLOAT64);
    run<30, 128>(fn, d, rows, ld, 0, 0, CU_TENSOR_MAP_DATA_TYPE_FLOAT64);
    run<30, 102>(fn, d, rows, ld, -2, -3, CU_TENSOR_MAP_DATA_TYPE_FLOAT64);
    return 0;
}
