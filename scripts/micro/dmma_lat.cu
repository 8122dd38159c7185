// Microbenchmark: latency of a dependent chain of DMMA.8x8x4 / DFMA (one warp), and the issue interval of
// independent DMMAs from one warp.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(long long *out, double *sink, int chains) {
    double a = threadIdx.x * 1e-3, b = 1.000001;
    double c[8][2];
    for (int j = 0; j < 8; ++j) c[j][0] = c[j][1] = j;
    const int iters = 2000;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (j < chains)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
    }
    long long t1 = clock64();
    double s = 0;
    for (int j = 0; j < 8; ++j) s += c[j][0] + c[j][1];
    sink[threadIdx.x] = s;
    if (threadIdx.x == 0) out[0] = (t1 - t0) / iters;
}
__global__ void kf(long long *out, double *sink) {
    double a = threadIdx.x * 1e-3, b = 1.000001, c = 0;
    const int iters = 2000;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) { c = fma(a, b, c); c = fma(a, b, c); c = fma(a, b, c); c = fma(a, b, c); }
    long long t1 = clock64();
    sink[threadIdx.x] = c;
    if (threadIdx.x == 0) out[0] = (t1 - t0) / iters / 4;
}
int main() {
    long long *out, h; double *sink; cudaMalloc(&out, 8); cudaMalloc(&sink, 8 * 1024);
    for (int chains : {1, 2, 4, 8}) {
        k<<<1, 32>>>(out, sink, chains); cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
        printf("DMMA: %d independent chains in one warp: %lld cycles per round (%.1f per DMMA)\n", chains, h, (double)h / chains);
    }
    kf<<<1, 32>>>(out, sink); cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
    printf("DFMA dependent latency: %lld cycles\n", h);
    return 0;
}
