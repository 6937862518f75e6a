// Single-warp latency (cycles per call) of the state-math building blocks, as the simulation kernel calls them.
#include <cstdio>
#include <cuda_runtime.h>
#include "upc_device.cuh"
using namespace upc;
template <int OP> __global__ void chain(double* out, long long* cyc, int iters)
{
    double x = 0.3 + threadIdx.x * 1e-9;
    double T[12] = {1, 0, 0, 0.1, 0, 1, 0, 0.2, 0, 0, 1, 0.3};
    double tw[6] = {0.01, 0.02, 0.03, 0.1 + x * 1e-3, 0.05, 0.02};
    DevAxis ax = {5.0, 0.0, 0.0, 0.0, 1.0, 0.1};
    double pi = 0.0, pe = 0.0;
    long long t0 = clock64();
    for (int i = 0; i < iters; i++)
    {
        if (OP == 0) x = x / 1.0000001 + 1e-9;
        if (OP == 1) x = sqrt(x) + 0.1;
        if (OP == 2) { double s, c; upc_sincos(x, &s, &c); x = s * 0.5 + c * 0.3; }
        if (OP == 3) x = upc_atan2(x, 0.7) + 0.2;
        if (OP == 4) { se3_exp(tw, T); tw[3] = 0.1 + T[1] * 1e-3; }
        if (OP == 5) { se3_log(T, tw); T[3] = 0.1 + tw[0] * 1e-3; T[1] = -0.01 + tw[5] * 1e-6; T[4] = 0.01 - tw[5] * 1e-6; }
        if (OP == 6) x = actuator_clamp(ax, pid_step(ax, pi, pe, x, 0.01)) * 0.01 + 0.3;
        if (OP == 7) x = upc_acos(0.5 + x * 1e-3) * 0.1 + 0.3;
        if (OP == 8) x = rint(x * 0.6366) * 1e-9 + x;
        if (OP == 9) x = floor(x + 0.5) * 1e-9 + 0.3;
    }
    long long t1 = clock64();
    out[threadIdx.x] = x + T[3] + tw[0];
    if (threadIdx.x == 0) *cyc = t1 - t0;
}
template <int OP> void run(const char* name)
{
    double* out; long long* cyc; cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
    const int iters = 2048;
    chain<OP><<<1, 32>>>(out, cyc, iters); chain<OP><<<1, 32>>>(out, cyc, iters);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-10s %.1f cycles per call (err %s)\n", name, (double)h / iters, cudaGetErrorString(cudaGetLastError()));
}
int main()
{
    run<0>("ddiv"); run<1>("dsqrt"); run<2>("sincos"); run<3>("atan2"); run<4>("se3_exp"); run<5>("se3_log"); run<6>("pid+clamp"); run<7>("acos");
    run<8>("rint"); run<9>("floor");
    return 0;
}
