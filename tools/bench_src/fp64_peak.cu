// Micro-benchmark: FP64 peak of the vector pipe (DFMA) and of the legacy FP64 MMA (mma.sync m8n8k4)
// on this GPU.  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_dfma(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
        x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

__global__ void k_dmma(double* out, int iters, double a, double b) {
    double c0[4] = {0, 0, 0, 0}, c1[4] = {0, 0, 0, 0};
    double av = a + threadIdx.x, bv = b;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[j]), "+d"(c1[j]) : "d"(av), "d"(bv));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0[0] + c1[0] + c0[1] + c1[1] + c0[2] + c1[2] + c0[3] + c1[3];
}

__global__ void k_ffma(float* out, int iters, float a, float b) {
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
        x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s SMs=%d\n", p.name, p.multiProcessorCount);
    const int blocks = p.multiProcessorCount * 8, threads = 256, iters = 20000;
    double* out; cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0); k_dfma<<<blocks, threads>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("DFMA  : %.3f ms  %.2f TFLOP/s\n", ms, 2.0 * 8 * iters * (double)blocks * threads / ms / 1e9);
        cudaEventRecord(e0); k_dmma<<<blocks, threads>>>(out, iters / 4, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("DMMA  : %.3f ms  %.2f TFLOP/s\n", ms, 512.0 * 4 * (iters / 4) * (double)blocks * threads / 32 / ms / 1e9);
        cudaEventRecord(e0); k_ffma<<<blocks, threads>>>((float*)out, iters, 1.0000001f, 1e-9f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        printf("FFMA  : %.3f ms  %.2f TFLOP/s\n", ms, 2.0 * 8 * iters * (double)blocks * threads / ms / 1e9);
    }
    printf("cuda error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
