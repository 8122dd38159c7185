// libcb_micro.so — measurement helper of bench.py (not part of the product ABI): the FP64 FMA rate of this GPU, the
// denominator of the FP64-pipe rooflines (MEASURED_PEAKS.json holds HBM and bf16 peaks only).
#include <cuda_runtime.h>

__global__ void __launch_bounds__(512) dfma_chain_kernel(double *out, int iters) {
    double a = threadIdx.x * 1e-3, b = 1.000001, c0 = 0, c1 = 1, c2 = 2, c3 = 3, c4 = 4, c5 = 5, c6 = 6, c7 = 7;
    for (int i = 0; i < iters; ++i) {
        c0 = fma(a, b, c0); c1 = fma(a, b, c1); c2 = fma(a, b, c2); c3 = fma(a, b, c3);
        c4 = fma(a, b, c4); c5 = fma(a, b, c5); c6 = fma(a, b, c6); c7 = fma(a, b, c7);
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = c0 + c1 + c2 + c3 + c4 + c5 + c6 + c7;
}

extern "C" int cbm_fp64_peak(double *tflops) {
    int dev = 0, sms = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 2;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int threads = 512, blocks = sms * 4, iters = 40000;
    double *out = nullptr;
    if (cudaMalloc(&out, sizeof(double) * (size_t)threads * blocks) != cudaSuccess) return 2;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double best = 0;
    dfma_chain_kernel<<<blocks, threads>>>(out, 1000);
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        dfma_chain_kernel<<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(out); return 2; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double tf = 2.0 * 8 * (double)iters * threads * blocks / (ms * 1e-3) / 1e12;
        if (tf > best) best = tf;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = best;
    return 0;
}
