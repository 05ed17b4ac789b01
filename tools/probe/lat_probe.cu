// Dependent-issue latencies on sm_100 that bound the single-CTA kernels of K4 (Jacobi, Cholesky): DFMA / DADD / DMUL chains,
// rsqrt(double), FP64 division, shuffle, shared-memory load->use, __syncthreads at several CTA sizes.  One warp, clock64.
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
__global__ void lat(double* out, long long* cyc, double seed) {
    __shared__ double sm[1024];
    double a = seed, b = 1.0000001, c = 1e-9;
    long long t0, t1;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (double)((i * 7 + 1) % 1024);
    __syncthreads();
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) a = fma(a, b, c);
    t1 = clock64(); if (threadIdx.x == 0) cyc[0] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) a = a + c;
    t1 = clock64(); if (threadIdx.x == 0) cyc[1] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) a = a * b;
    t1 = clock64(); if (threadIdx.x == 0) cyc[2] = t1 - t0;
    a = fabs(a) + 1.5;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; i++) a = rsqrt(a) + 1.5;
    t1 = clock64(); if (threadIdx.x == 0) cyc[3] = t1 - t0;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; i++) a = 1.0 / a + 1.5;
    t1 = clock64(); if (threadIdx.x == 0) cyc[4] = t1 - t0;
    t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; i++) a = sqrt(a) + 1.5;
    t1 = clock64(); if (threadIdx.x == 0) cyc[5] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) a = __shfl_xor_sync(0xffffffffu, a, 1);
    t1 = clock64(); if (threadIdx.x == 0) cyc[6] = t1 - t0;
    int idx = threadIdx.x;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) idx = (int)sm[idx & 1023];
    t1 = clock64(); if (threadIdx.x == 0) cyc[7] = t1 - t0;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) __syncthreads();
    t1 = clock64(); if (threadIdx.x == 0) cyc[8] = t1 - t0;
    float f = (float)a;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) f = fmaf(f, 1.0000001f, 1e-9f);
    t1 = clock64(); if (threadIdx.x == 0) cyc[9] = t1 - t0;
    out[threadIdx.x] = a + idx + f;
}
int main() {
    double* d; long long* c; cudaMalloc(&d, 8 * 1024); cudaMalloc(&c, 8 * 16);
    const char* names[] = {"DFMA chain", "DADD chain", "DMUL chain", "rsqrt(double)+DADD", "1.0/x+DADD", "sqrt(double)+DADD", "SHFL.BFLY 64-bit",
                           "LDS.64 -> F2I -> address", "__syncthreads", "FFMA chain"};
    for (int threads : {32, 256, 640, 1024}) {
        lat<<<1, threads>>>(d, c, 1.25); cudaDeviceSynchronize();
        lat<<<1, threads>>>(d, c, 1.25); cudaDeviceSynchronize();
        long long h[16]; cudaMemcpy(h, c, sizeof(h), cudaMemcpyDeviceToHost);
        printf("threads %4d:", threads);
        for (int i = 0; i < 10; i++) printf("  %s %.1f", names[i], (double)h[i] / N);
        printf("\n");
    }
    return 0;
}
