// Dependent-chain latency of FP64 operations on one warp (cycles per op), and throughput with 1..16 warps per SMSP.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP> __device__ __forceinline__ double op(double x, double y)
{
    if (OP == 0) return __fma_rn(x, y, 1e-9);
    if (OP == 1) return x + y;
    if (OP == 2) return x * y;
    if (OP == 3) return x / y + 1.0;
    if (OP == 4) return sqrt(x) + 1.5;
    if (OP == 5) return rint(x) + 0.3;
    if (OP == 6) return floor(x) + 0.3;
    return x;
}
template <int OP> __global__ void chain(double* out, long long* cyc, int iters, double y)
{
    double x = 1.0 + threadIdx.x * 1e-6;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++) { x = op<OP>(x, y); x = op<OP>(x, y); x = op<OP>(x, y); x = op<OP>(x, y); }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int OP> void run(const char* name)
{
    double* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    for (int warps : {1, 4, 8, 16, 32})
    {
        chain<OP><<<148, warps * 32>>>(out, cyc, iters, 1.0000001);
        chain<OP><<<148, warps * 32>>>(out, cyc, iters, 1.0000001);
        long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("%-6s warps/SM %2d: %.2f cycles per dependent op (per warp)  -> %.3f ops/clk/SM\n", name, warps, (double)h / (4.0 * iters), warps * 4.0 * iters / (double)h);
    }
}
int main()
{
    run<0>("DFMA"); run<1>("DADD"); run<2>("DMUL"); run<3>("DDIV"); run<4>("DSQRT"); run<5>("RINT"); run<6>("FLOOR");
    return 0;
}
