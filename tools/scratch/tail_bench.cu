// Experiment: where does the time of the exact-dot tail go?  Each CTA records %globaltimer at the phases of
// flush_warp / block_finish; the host prints, per grid size, the phase boundaries of the LAST block relative to kernel start.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -I../../feltor_b200/csrc tail_bench.cu -o tail_bench
#include <cuda_runtime.h>
__device__ unsigned long long g_tr[8];
#define DGB_TRACE(k) do { if (threadIdx.x == 0) { unsigned long long _t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(_t)); g_tr[k] = _t; } } while (0)
#include "superacc.cuh"
#include <vector>
#include <algorithm>
namespace dgb { long long g_launches = 0; void set_error(const char*, ...) {} int cuda_fail(cudaError_t, const char*, const char*, int) { return 1; } }
using namespace dgb;
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
constexpr int NPH = 8;
template <int THREADS>
__global__ void __launch_bounds__(THREADS) tail_kernel(sa::DotSlot slot, unsigned long long* trace, const double* x, size_t n) {
    __shared__ long long smem[sa::BINS];
    unsigned long long t[NPH];
    t[0] = gtime();
    sa::block_init<1>(smem);
    sa::Fpe fpe[2];
    fpe[0].clear(); fpe[1].clear();
    int bad = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double r = fpe[i & 1].add_lazy(x[i]);
        if (r != 0.) sa::accumulate(smem, r, 1);
    }
    t[1] = gtime();
    fpe[0].merge(fpe[1], smem);
    t[2] = gtime();
    {   // flush_warp, opened up for timing
        const int lane = threadIdx.x & 31;
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            double v[sa::NF];
#pragma unroll
            for (int i = 0; i < sa::NF; i++) v[i] = __shfl_down_sync(0xffffffffu, fpe[0].a[i], off);
            if (lane < off) fpe[0].absorb(v, smem);
        }
        t[5] = gtime();
        if (lane == 0) fpe[0].flush(smem);
        else fpe[0].clear();
        __syncwarp();
    }
    t[3] = gtime();
    bool last = sa::block_finish<1>(smem, bad, slot);
    t[4] = gtime();
    if (threadIdx.x == 0) {
        for (int k = 0; k < 5; k++) trace[(size_t)blockIdx.x * NPH + k] = t[k];
        trace[(size_t)blockIdx.x * NPH + 5] = last ? 1 : 0;
        trace[(size_t)blockIdx.x * NPH + 6] = t[5];
    }
}
int main() {
    sa::DotSlot slot;
    cudaMalloc(&slot.gacc, 4 * sa::GACC_WORDS * 8); cudaMemset(slot.gacc, 0, 4 * sa::GACC_WORDS * 8);
    cudaMalloc(&slot.gstatus, 16); cudaMemset(slot.gstatus, 0, 16);
    cudaMalloc(&slot.ticket, 16); cudaMemset(slot.ticket, 0, 16);
    cudaMalloc(&slot.result, 4 * sizeof(dgb_dot_result));
    const size_t nmax = 1 << 22;
    double* x; cudaMalloc(&x, nmax * 8);
    std::vector<double> hx(nmax);
    for (size_t i = 0; i < nmax; i++) hx[i] = 1.0 / (1 + (i % 977)) * ((i % 3) ? 1 : -1);
    cudaMemcpy(x, hx.data(), nmax * 8, cudaMemcpyHostToDevice);
    unsigned long long* trace; cudaMalloc(&trace, 4096 * NPH * 8);
    std::vector<unsigned long long> h(4096 * NPH);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int grid : {1, 16, 74, 148, 296, 444, 888}) {
        for (size_t n : {(size_t)grid * 256, (size_t)147456, (size_t)2359296}) {
            float best = 1e9; unsigned long long ph[6] = {0};
            for (int rep = 0; rep < 6; rep++) {
                cudaEventRecord(e0);
                tail_kernel<256><<<grid, 256>>>(slot, trace, x, n);
                cudaEventRecord(e1);
                cudaDeviceSynchronize();
                float ms; cudaEventElapsedTime(&ms, e0, e1);
                if (rep < 2) continue;
                if (ms < best) {
                    best = ms;
                    cudaMemcpy(h.data(), trace, (size_t)grid * NPH * 8, cudaMemcpyDeviceToHost);
                    unsigned long long t0 = ~0ull; int lastb = 0;
                    for (int b = 0; b < grid; b++) { t0 = std::min(t0, h[b * NPH]); if (h[b * NPH + 5]) lastb = b; }
                    unsigned long long mx[5] = {0};
                    for (int b = 0; b < grid; b++) for (int k = 0; k < 5; k++) mx[k] = std::max(mx[k], h[b * NPH + k] - t0);
                    for (int k = 0; k < 5; k++) ph[k] = mx[k];
                    unsigned long long m6 = 0; for (int b = 0; b < grid; b++) m6 = std::max(m6, h[b * NPH + 6] - t0);
                    unsigned long long tr[8]; cudaMemcpyFromSymbol(tr, g_tr, sizeof(tr));
                    printf("   shuffle tree done at %llu | last writer of trace: before atomics %llu after atomics+sync %llu after ticket %llu fetched %llu published %llu\n", m6,
                           tr[0] - t0, tr[1] - t0, tr[2] - t0, tr[3] - t0, tr[4] - t0);
                    ph[5] = h[lastb * NPH + 4] - t0;
                }
            }
            printf("grid %4d n %8zu: event %6.1f us | max over CTAs (ns since first CTA start): start %5llu loop %6llu merge %6llu flush_warp %6llu block_finish %6llu | last block done %6llu\n",
                   grid, n, best * 1e3, ph[0], ph[1], ph[2], ph[3], ph[4], ph[5]);
        }
    }
    // empty kernel floor
    for (int rep = 0; rep < 3; rep++) { cudaEventRecord(e0); tail_kernel<256><<<1, 256>>>(slot, trace, x, 0); cudaEventRecord(e1); cudaDeviceSynchronize(); float ms; cudaEventElapsedTime(&ms, e0, e1); printf("1 CTA n=0: %.1f us\n", ms * 1e3); }
    return 0;
}
