// fp64_mix.cu -- lab: does a DADD occupy the scheduler's issue port for one cycle or for two?  8 independent DADD chains
// per thread plus NI independent FFMA (fp32) or IMAD chains; 2 and 4 warps per scheduler.
#include <cstdio>
#include <cuda_runtime.h>

template <int ND, int NI, int KIND>
__global__ void k(double *out, int iters, double a, double b, float fa, float fb, int ia)
{
    double x[ND > 0 ? ND : 1];
    float f[NI > 0 ? NI : 1];
    int n[NI > 0 ? NI : 1];
#pragma unroll
    for (int i = 0; i < ND; ++i) x[i] = a + i + threadIdx.x;
#pragma unroll
    for (int i = 0; i < NI; ++i) { f[i] = fa + i + threadIdx.x; n[i] = ia + i + threadIdx.x; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int i = 0; i < (ND > NI ? ND : NI); ++i) {
                if (i < ND) x[i] = __dadd_rn(x[i], b);
                if (i < NI) {
                    if (KIND == 0) f[i] = __fmaf_rn(f[i], fb, fa);
                    else if (KIND == 1) n[i] = n[i] * ia + 7;
                    else n[i] = (n[i] ^ ia) + i;          // ALU pipe: one LOP3 / IADD3 pair
                }
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ND; ++i) s += x[i];
#pragma unroll
    for (int i = 0; i < NI; ++i) s += f[i] + n[i];
    if (s == 12345.678) out[0] = s;
}

template <int ND, int NI, int KIND>
void run(int warps_per_sm)
{
    double *d;
    cudaMalloc(&d, 8);
    const int iters = 4096;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<ND, NI, KIND><<<148, warps_per_sm * 32>>>(d, 16, 1.0, 1.0000001, 1.f, 1.0001f, 3);
    cudaEventRecord(e0);
    k<ND, NI, KIND><<<148, warps_per_sm * 32>>>(d, iters, 1.0, 1.0000001, 1.f, 1.0001f, 3);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double cycles = ms * 1e-3 * clk * 1e3;
    const double per_smsp_iter = cycles / iters / 4 / (warps_per_sm / 4);   // cycles per (warp, r-iteration) on one scheduler
    printf("fp64 %d + %s %d per iteration, %d warps/scheduler: %.2f cycles per warp-iteration (fp64 alone would be %d, other alone %d)\n",
           ND, KIND == 0 ? "FFMA" : (KIND == 1 ? "IMAD" : "LOP3+IADD3 pairs"), NI, warps_per_sm / 4, per_smsp_iter, 2 * ND, NI);
    cudaFree(d);
}

int main()
{
    for (int w : {8, 16}) {
        run<8, 0, 0>(w);
        run<8, 4, 0>(w);
        run<8, 8, 0>(w);
        run<8, 12, 0>(w);
        run<8, 16, 0>(w);
        run<0, 16, 0>(w);
        run<8, 8, 1>(w);
        run<8, 16, 1>(w);
        run<0, 16, 1>(w);
        run<8, 4, 2>(w);
        run<8, 8, 2>(w);
        run<0, 8, 2>(w);
        run<8, 4, 1>(w);
    }
    return 0;
}
