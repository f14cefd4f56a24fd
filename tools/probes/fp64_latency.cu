// Dependent-chain latencies on one SM: DFMA, DADD, 1.0/x, rsqrt(x), FFMA, shared-memory round trip, __syncthreads with 16 warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency tools/probes/fp64_latency.cu && ./fp64_latency
#include <cstdio>
#include <cuda_runtime.h>
__global__ void probe(double* out, long long* cyc, double seed, int warps_busy) {
    __shared__ double sh[64];
    const int tid = threadIdx.x;
    double x = seed + tid * 1e-9, y = 1.0000001;
    long long t0, t1;
    // DFMA chain
    __syncthreads();
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) x = fma(x, y, 1e-9);
    }
    t1 = clock64();
    if (tid == 0) cyc[0] = (t1 - t0);
    // division chain
    __syncthreads();
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) x = 1.0 / (x + 1.5);
    }
    t1 = clock64();
    if (tid == 0) cyc[1] = (t1 - t0);
    // rsqrt chain
    __syncthreads();
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j) x = rsqrt(x + 1.5);
    }
    t1 = clock64();
    if (tid == 0) cyc[2] = (t1 - t0);
    // FFMA chain
    float f = (float)x, g = 1.0001f;
    __syncthreads();
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 64; ++i) {
#pragma unroll
        for (int j = 0; j < 16; ++j) f = fmaf(f, g, 1e-6f);
    }
    t1 = clock64();
    if (tid == 0) cyc[3] = (t1 - t0);
    // shared-memory round trip (store -> load, same thread)
    __syncthreads();
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) {
        sh[tid & 63] = x;
        x = *(volatile double*)&sh[(tid + (int)x) & 63] + 1.0;
    }
    t1 = clock64();
    if (tid == 0) cyc[4] = (t1 - t0);
    // barrier
    __syncthreads();
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) __syncthreads();
    t1 = clock64();
    if (tid == 0) cyc[5] = (t1 - t0);
    // barrier + store/load through shared memory by different warps (publish / consume)
    __syncthreads();
    t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < 256; ++i) {
        if (tid == ((i * 37) & (blockDim.x - 1))) sh[0] = x + 1.0;
        __syncthreads();
        x = sh[0];
    }
    t1 = clock64();
    if (tid == 0) cyc[6] = (t1 - t0);
    out[tid] = x + f;
}
int main() {
    double* out; long long* cyc;
    cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 64);
    for (int threads : {32, 512}) {
        probe<<<1, threads>>>(out, cyc, 0.5, 0);
        cudaDeviceSynchronize();
        probe<<<1, threads>>>(out, cyc, 0.5, 0);
        cudaDeviceSynchronize();
        long long h[8];
        cudaMemcpy(h, cyc, 56, cudaMemcpyDeviceToHost);
        printf("%d threads: DFMA %.1f  1.0/x %.1f  rsqrt %.1f  FFMA %.1f  smem st->ld %.1f  __syncthreads %.1f  publish+barrier+load %.1f  (cycles per dependent op)\n",
               threads, h[0] / 1024.0, h[1] / 256.0, h[2] / 256.0, h[3] / 1024.0, h[4] / 256.0, h[5] / 256.0, h[6] / 256.0);
    }
    return 0;
}
