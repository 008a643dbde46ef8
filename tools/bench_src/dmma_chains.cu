// Micro-benchmark: FP64 MMA (mma.sync m8n8k4) throughput of ONE CTA per SM as a function of warps per CTA and independent
// accumulator chains per warp (how much parallelism a kernel with few resident warps needs to fill the FP64 tensor pipe).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o dmma_chains dmma_chains.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int C>
__global__ void k(double* out, int iters, double a, double b) {
    double c0[C], c1[C];
#pragma unroll
    for (int j = 0; j < C; ++j) c0[j] = c1[j] = 0.0;
    double av = a + threadIdx.x, bv = b;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < C; ++j)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[j]), "+d"(c1[j]) : "d"(av), "d"(bv));
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < C; ++j) s += c0[j] + c1[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int C>
void run(int sms, int warps, double* out) {
    const int iters = 40000 / C;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); k<C><<<sms, warps * 32>>>(out, iters, 1.0000001, 1e-9); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
    }
    const double dm = (double)iters * C * warps * sms;
    printf("warps/SM %2d chains/warp %d : %7.2f TFLOP/s  (%.1f cycles per DMMA per warp-chain at 1.965 GHz)\n", warps, C,
           512.0 * dm / ms / 1e9, ms * 1e-3 * 1.965e9 / ((double)iters));
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    double* out; cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 1024);
    for (int w : {1, 4, 8, 16, 32}) {
        run<1>(p.multiProcessorCount, w, out);
        run<2>(p.multiProcessorCount, w, out);
        run<4>(p.multiProcessorCount, w, out);
        run<8>(p.multiProcessorCount, w, out);
    }
    printf("cuda error: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
