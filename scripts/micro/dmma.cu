// Microbenchmark: FP64 throughput of DFMA vs mma.sync.m8n8k4.f64 (DMMA) on this GPU.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dfma(double *out, int iters) {
    double a = threadIdx.x * 1e-3, b = 1.000001, c0 = 0, c1 = 1, c2 = 2, c3 = 3, c4 = 4, c5 = 5, c6 = 6, c7 = 7;
    for (int i = 0; i < iters; ++i) {
        c0 = fma(a, b, c0); c1 = fma(a, b, c1); c2 = fma(a, b, c2); c3 = fma(a, b, c3);
        c4 = fma(a, b, c4); c5 = fma(a, b, c5); c6 = fma(a, b, c6); c7 = fma(a, b, c7);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7;
}
__global__ void k_dmma(double *out, int iters) {
    double a = threadIdx.x * 1e-3, b = 1.000001;
    double c[8][2];
    for (int j = 0; j < 8; ++j) c[j][0] = c[j][1] = j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
    }
    double s = 0;
    for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
    double *out; cudaMalloc(&out, 8 * 148 * 8 * 1024);
    const int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int warps : {4, 8, 16, 32}) {
        float ms;
        k_dfma<<<148 * 2, warps * 16>>>(out, 100);
        cudaEventRecord(e0); k_dfma<<<148 * 2, warps * 16>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        double fl = 2.0 * 8 * iters * 148 * 2 * warps * 16;
        printf("DFMA  warps/SM=%2d: %.2f TFLOP/s\n", warps, fl / ms / 1e9);
        k_dmma<<<148 * 2, warps * 16>>>(out, 100);
        cudaEventRecord(e0); k_dmma<<<148 * 2, warps * 16>>>(out, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1);
        fl = 2.0 * 256 * 8 * iters * 148 * 2 * (warps * 16 / 32);
        printf("DMMA  warps/SM=%2d: %.2f TFLOP/s  (%s)\n", warps, fl / ms / 1e9, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
