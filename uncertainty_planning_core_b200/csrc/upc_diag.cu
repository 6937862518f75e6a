/*
 * upc_diag.cu - roofline denominators that MEASURED_PEAKS.json does not carry:
 * random 32-byte-sector gather bandwidth out of L2 (the ceiling of the SDF-collision kernel's gathers) and the
 * FP64 FMA rate (the ceiling of the fused distance/adjacency kernel).  Diagnostics only; nothing on the hot path.
 */
#include "upc_internal.h"

namespace upc
{
__global__ void __launch_bounds__(256) l2_gather_kernel(const float* __restrict__ buf, uint32_t sector_mask, int iters, float* sink)
{
    uint32_t x = (blockIdx.x * 256u + threadIdx.x) * 2654435761u + 12345u;
    float acc = 0.0f;
    for (int it = 0; it < iters; it++)
    {
        uint32_t idx[8];
#pragma unroll
        for (int k = 0; k < 8; k++)
        {
            x ^= x << 13; x ^= x >> 17; x ^= x << 5;
            idx[k] = (x & sector_mask) * 8u; /* first float of a random 32-byte sector */
        }
#pragma unroll
        for (int k = 0; k < 8; k++) acc += __ldcg(buf + idx[k]);
    }
    if (acc == 123.456f) *sink = acc;
}

__global__ void __launch_bounds__(256) fp64_fma_kernel(int iters, double* sink)
{
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1.0, a2 = a0 + 2.0, a3 = a0 + 3.0, a4 = a0 + 4.0, a5 = a0 + 5.0, a6 = a0 + 6.0, a7 = a0 + 7.0;
    const double m = 1.0000001, c = 1e-7;
    for (int it = 0; it < iters; it++)
    {
        a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
        a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
    }
    const double s = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
    if (s == 123.456) *sink = s;
}
} /* namespace upc */

using namespace upc;

extern "C" {

int32_t upc_measure_l2_gather_gbs(upc_handle* h, int64_t buffer_bytes, double* out_gbs)
{
    if (!h || !out_gbs || buffer_bytes < 4096) { set_error("bad argument"); return UPC_ERR_INVALID_ARGUMENT; }
    UPC_CUDA_TRY(cudaSetDevice(h->device));
    int64_t sectors = 1;
    while (sectors * 2 * 32 <= buffer_bytes) sectors *= 2;
    float* buf = nullptr;
    float* sink = nullptr;
    UPC_CUDA_TRY(cudaMalloc(&buf, (size_t)sectors * 32));
    UPC_CUDA_TRY(cudaMalloc(&sink, 4));
    UPC_CUDA_TRY(cudaMemsetAsync(buf, 0, (size_t)sectors * 32, h->stream));
    const int blocks = h->sm_count * 8, iters = 256;
    cudaEvent_t e0, e1;
    UPC_CUDA_TRY(cudaEventCreate(&e0));
    UPC_CUDA_TRY(cudaEventCreate(&e1));
    float best_ms = 1e30f;
    for (int rep = 0; rep < 5; rep++)
    {
        UPC_CUDA_TRY(cudaEventRecord(e0, h->stream));
        l2_gather_kernel<<<blocks, 256, 0, h->stream>>>(buf, (uint32_t)(sectors - 1), iters, sink);
        UPC_CUDA_TRY(cudaEventRecord(e1, h->stream));
        UPC_CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        UPC_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best_ms) best_ms = ms;
    }
    const double bytes = (double)blocks * 256.0 * iters * 8.0 * 32.0;
    *out_gbs = bytes / (best_ms * 1e-3) / 1e9;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(buf); cudaFree(sink);
    return UPC_OK;
}

int32_t upc_measure_fp64_tflops(upc_handle* h, double* out_tflops)
{
    if (!h || !out_tflops) { set_error("bad argument"); return UPC_ERR_INVALID_ARGUMENT; }
    UPC_CUDA_TRY(cudaSetDevice(h->device));
    double* sink = nullptr;
    UPC_CUDA_TRY(cudaMalloc(&sink, 8));
    const int blocks = h->sm_count * 8, iters = 8192;
    cudaEvent_t e0, e1;
    UPC_CUDA_TRY(cudaEventCreate(&e0));
    UPC_CUDA_TRY(cudaEventCreate(&e1));
    float best_ms = 1e30f;
    for (int rep = 0; rep < 5; rep++)
    {
        UPC_CUDA_TRY(cudaEventRecord(e0, h->stream));
        fp64_fma_kernel<<<blocks, 256, 0, h->stream>>>(iters, sink);
        UPC_CUDA_TRY(cudaEventRecord(e1, h->stream));
        UPC_CUDA_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        UPC_CUDA_TRY(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best_ms) best_ms = ms;
    }
    const double flops = (double)blocks * 256.0 * iters * 8.0 * 2.0;
    *out_tflops = flops / (best_ms * 1e-3) / 1e12;
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(sink);
    return UPC_OK;
}
}
