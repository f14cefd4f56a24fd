// FP64 issue-rate probe: DFMA vs DADD vs DMUL vs (DMUL + DADD) per SM per clock.  nvcc -arch=sm_100a -O3 fp64_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(double* out, int iters, double x, double y) {
    double a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = threadIdx.x * 1e-3 + j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (MODE == 0) a[j] = __fma_rn(a[j], x, y);
            if (MODE == 1) a[j] = __dadd_rn(a[j], y);
            if (MODE == 2) a[j] = __dmul_rn(a[j], x);
            if (MODE == 3) a[j] = __dadd_rn(a[j], __dmul_rn(a[(j + 1) & 7], x));
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += a[j];
    if (s == 12345.678) out[0] = s;
}
template <int MODE> void run(const char* name, int warps_per_sm, double ops_per_iter) {
    int sms = 148, iters = 20000;
    double* d; cudaMalloc(&d, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<sms, warps_per_sm * 32>>>(d, 100, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    k<MODE><<<sms, warps_per_sm * 32>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double instr = (double)sms * warps_per_sm * 32 * iters * ops_per_iter;      // thread-level DP instructions
    printf("%-12s warps/SM %2d: %.1f DP lane-instr / clk / SM (at 1.965 GHz)\n", name, warps_per_sm, instr / (ms * 1e-3) / sms / 1.965e9);
}
int main() {
    for (int w : {4, 8, 16, 32}) {
        if (w == 4) { run<0>("DFMA", 4, 8); run<1>("DADD", 4, 8); run<2>("DMUL", 4, 8); run<3>("DMUL+DADD", 4, 16); }
        if (w == 8) { run<0>("DFMA", 8, 8); run<1>("DADD", 8, 8); run<2>("DMUL", 8, 8); run<3>("DMUL+DADD", 8, 16); }
        if (w == 16) { run<0>("DFMA", 16, 8); run<1>("DADD", 16, 8); run<2>("DMUL", 16, 8); run<3>("DMUL+DADD", 16, 16); }
        if (w == 32) { run<0>("DFMA", 32, 8); run<1>("DADD", 32, 8); run<2>("DMUL", 32, 8); run<3>("DMUL+DADD", 32, 16); }
    }
    return 0;
}
