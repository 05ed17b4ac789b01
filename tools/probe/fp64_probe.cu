// FP64 pipe probe for sm_100a: DMMA (mma.sync f64) shapes vs DFMA throughput.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o fp64_probe fp64_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double (&c)[4], const double (&a)[4], const double (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
}
__device__ __forceinline__ void dmma16816(double (&c)[4], const double (&a)[8], const double (&b)[4]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int CH>
__global__ void k_dmma884(double* out, int iters, double a0, double b0) {
    double c[CH][2];
#pragma unroll
    for (int i = 0; i < CH; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) dmma884(c[i][0], c[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
__global__ void k_dmma1688(double* out, int iters, double a0, double b0) {
    double c[CH][4];
#pragma unroll
    for (int i = 0; i < CH; i++) for (int j = 0; j < 4; j++) c[i][j] = threadIdx.x + j;
    double a[4] = {a0, a0 + 1e-9, a0 + 2e-9, a0 + threadIdx.x * 1e-9};
    double b[2] = {b0, b0 + 1e-9};
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) dmma1688(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) for (int j = 0; j < 4; j++) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
__global__ void k_dmma16816(double* out, int iters, double a0, double b0) {
    double c[CH][4];
#pragma unroll
    for (int i = 0; i < CH; i++) for (int j = 0; j < 4; j++) c[i][j] = threadIdx.x + j;
    double a[8], b[4];
    for (int j = 0; j < 8; j++) a[j] = a0 + j * 1e-9 + threadIdx.x * 1e-9;
    for (int j = 0; j < 4; j++) b[j] = b0 + j * 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) dmma16816(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) for (int j = 0; j < 4; j++) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
__global__ void k_dfma(double* out, int iters, double a0, double b0) {
    double c[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) c[i] = threadIdx.x + i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) c[i] = fma(a, c[i], b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static double timeit(F launch, int reps) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best * 1e-3;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int nsm = p.multiProcessorCount;
    printf("device %s sm_%d%d SMs=%d clock(kHz)=%d\n", p.name, p.major, p.minor, nsm, p.clockRate);
    double* out; CK(cudaMalloc(&out, sizeof(double) * nsm * 32 * 1024));
    const int iters = 20000;
    for (int wps : {4, 8, 16, 32}) {   // warps per SM (one CTA per SM... use ctas of 128 thr)
        int threads = 128; int ctas = nsm * (wps / 4);
        double warps = (double)ctas * 4;
        double t;
        t = timeit([&] { k_dmma884<8><<<ctas, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
        printf("warps/SM=%2d dmma.m8n8k4   x8 chains: %.2f TFLOP/s\n", wps, warps * iters * 8 * 512.0 / t * 1e-12);
        t = timeit([&] { k_dmma884<16><<<ctas, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
        printf("warps/SM=%2d dmma.m8n8k4  x16 chains: %.2f TFLOP/s\n", wps, warps * iters * 16 * 512.0 / t * 1e-12);
        t = timeit([&] { k_dmma1688<8><<<ctas, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
        printf("warps/SM=%2d dmma.m16n8k8  x8 chains: %.2f TFLOP/s\n", wps, warps * iters * 8 * 2048.0 / t * 1e-12);
        t = timeit([&] { k_dmma16816<8><<<ctas, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
        printf("warps/SM=%2d dmma.m16n8k16 x8 chains: %.2f TFLOP/s\n", wps, warps * iters * 8 * 4096.0 / t * 1e-12);
        t = timeit([&] { k_dfma<16><<<ctas, threads>>>(out, iters, 1.0000001, 1e-9); }, 5);
        printf("warps/SM=%2d dfma         x16 chains: %.2f TFLOP/s\n", wps, warps * 32 * iters * 16 * 2.0 / t * 1e-12);
    }
    CK(cudaGetLastError());
    return 0;
}
