// fp64_rate.cu -- lab: DADD / DMUL / DFMA issue rate of one SM sub-partition as a function of warps per scheduler and
// independent chains per thread.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a --fmad=false fp64_rate.cu -o fp64_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int CH, int KIND>
__global__ void k(double *out, int iters, double a, double b)
{
    double x[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = a + i + threadIdx.x;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                if (KIND == 0) x[i] = __dadd_rn(x[i], b);
                else if (KIND == 1) x[i] = __dmul_rn(x[i], b);
                else x[i] = __fma_rn(x[i], b, a);
            }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += x[i];
    if (s == 12345.678) out[0] = s;
}

template <int CH, int KIND>
void run(const char *name, int warps_per_sm)
{
    double *d;
    cudaMalloc(&d, 8);
    const int iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<CH, KIND><<<148, warps_per_sm * 32>>>(d, 16, 1.0, 1.0000001);
    cudaEventRecord(e0);
    k<CH, KIND><<<148, warps_per_sm * 32>>>(d, iters, 1.0, 1.0000001);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double winstr = (double)iters * 8 * CH * warps_per_sm;          // warp instructions per SM
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double cycles = ms * 1e-3 * clk * 1e3;
    printf("%s chains %d warps/SM %2d: %.3f ms  %.3f warp-instr/cycle/SM (at %d MHz nominal)\n", name, CH, warps_per_sm, ms,
           winstr / cycles, clk / 1000);
    cudaFree(d);
}

int main()
{
    for (int w : {4, 8, 16, 32}) {
        run<1, 0>("DADD", w);
        run<4, 0>("DADD", w);
        run<8, 0>("DADD", w);
        run<8, 1>("DMUL", w);
        run<8, 2>("DFMA", w);
    }
    return 0;
}
