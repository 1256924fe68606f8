// B200 microbenchmarks behind the C-ABI (hq_microbench_*): the numbers the evaluator's cost model and the
// roofline denominators for compute-bound groups are calibrated from.  They stand in for the reference's
// evaluator-preprocess tool (evaluator-preprocess/process.cpp:85-165: cublasZgemm and cuTT timings) -- here
// the primitives being priced are this library's own: FP64 FMA issue rate, FP64 tensor (DMMA) rate, and the
// HBM copy bandwidth as seen by a plain grid-stride kernel.
#include <algorithm>
#include <functional>

#include "hq_internal.h"

namespace hq {

// 8 independent FMA chains per thread: 16 FP64 flops per inner iteration per thread
__global__ void dfma_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 12345.678) out[0] = s;   // keeps the chains alive without a store in the common case
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// 8 independent m8n8k4 accumulators per warp: 8 * 512 flops per inner iteration per warp
__global__ void dmma_kernel(double* out, int iters, double a, double b) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma884(c[2 * i], c[2 * i + 1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 12345.678) out[0] = s;
}

__global__ void copy_kernel(const double2* __restrict__ src, double2* __restrict__ dst, uint64_t n) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}

static int time_ms(cudaStream_t s, float* ms, const std::function<void()>& fn) {
    cudaEvent_t e0, e1;
    HQ_CUDA(cudaEventCreate(&e0));
    HQ_CUDA(cudaEventCreate(&e1));
    fn();   // warm-up
    HQ_CUDA(cudaEventRecord(e0, s));
    fn();
    HQ_CUDA(cudaEventRecord(e1, s));
    HQ_CUDA(cudaEventSynchronize(e1));
    HQ_CUDA(cudaEventElapsedTime(ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    return HQ_OK;
}

}  // namespace hq

using namespace hq;

// kind 0: FP64 FMA (CUDA cores), kind 1: FP64 tensor cores (mma.sync m8n8k4).  Reports dense TFLOP/s.
// kind 10+w: DMMA with only w warps (of 8 independent accumulators) resident per SM -- how many warps the dense kernel
// needs in its compute phase to keep the FP64 tensor pipe busy.
extern "C" int hq_microbench_fp64(int kind, double* tflops) {
    HQ_REQUIRE(rt().ready && tflops && (kind == 0 || kind == 1 || (kind > 10 && kind <= 74)), "bad arguments to hq_microbench_fp64");
    double* d = nullptr;
    HQ_CUDA(cudaMalloc(&d, 64));
    int iters = 20000, block = 256, grid = rt().sm_count * 8;
    if (kind > 10) {   // one CTA per SM with (kind - 10) warps; extra dynamic smem keeps a second CTA off the SM
        block = (kind - 10) * 32;
        if (block > 1024) { grid = rt().sm_count * 2; block /= 2; }
        else grid = rt().sm_count;
    }
    if (kind > 10) HQ_CUDA(cudaFuncSetAttribute(dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
    float ms = 0;
    int rc = time_ms(rt().compute, &ms, [&] {
        if (kind == 0) dfma_kernel<<<grid, block, 0, rt().compute>>>(d, iters, 1.0000001, 1e-9);
        else if (kind == 1) dmma_kernel<<<grid, block, 0, rt().compute>>>(d, iters, 1.0000001, 1e-9);
        else dmma_kernel<<<grid, block, 120 * 1024, rt().compute>>>(d, iters, 1.0000001, 1e-9);
    });
    cudaFree(d);
    if (rc != HQ_OK) return rc;
    HQ_CUDA(cudaGetLastError());
    const double flops = kind == 0 ? (double)grid * block * iters * 16.0 : (double)grid * (block / 32) * iters * 8.0 * 512.0;
    *tflops = flops / (ms * 1e-3) / 1e12;
    return HQ_OK;
}

// Plain device-to-device copy of 2^L amplitudes inside `state` (first half -> second half); read+write GB/s.
extern "C" int hq_microbench_copy(void* state, int L, double* gbs) {
    HQ_REQUIRE(rt().ready && state && gbs && L >= 2, "bad arguments to hq_microbench_copy");
    const uint64_t n = 1ull << (L - 1);
    double2* s = static_cast<double2*>(state);
    float ms = 0;
    int rc = time_ms(rt().compute, &ms, [&] { copy_kernel<<<rt().sm_count * 16, 512, 0, rt().compute>>>(s, s + n, n); });
    if (rc != HQ_OK) return rc;
    HQ_CUDA(cudaGetLastError());
    *gbs = 32.0 * (double)n / (ms * 1e-3) / 1e9;
    return HQ_OK;
}
