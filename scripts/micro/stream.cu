// Microbenchmark: how fast can a tile-shaped (ROWS x 32 vectors, leading dimension ld) read stream go
// for different load instructions and occupancies?  Development aid (not part of the product).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef long long i64;
__device__ __forceinline__ void cp8(double *d, const double *s) {
    unsigned a = (unsigned)__cvta_generic_to_shared(d);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(a), "l"(s) : "memory");
}
__device__ __forceinline__ void cp16(double *d, const double *s) {
    unsigned a = (unsigned)__cvta_generic_to_shared(d);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a), "l"(s) : "memory");
}
// MODE 0: LDG.64 -> registers (sum); 1: LDGSTS.64 -> smem; 2: LDGSTS.128 (ld even only); 3: LDG.64 -> STS.64
// 4: LDG.64 -> registers, store back STG.64 to Y (copy through registers, no smem)
// 5: LDGSTS.64 -> smem -> LDS.64 -> STG.64 (staged copy, the mover pair of the apply kernels)
// 6: LDG.128 -> STG.128 (ld even only)
template <int MODE>
__global__ void k(const double *X, double *Y, i64 ld, i64 n, i64 nvec, int rows, double *sink) {
    extern __shared__ double sm[];
    const int pitch = rows | 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const i64 nrt = n / rows, nvt = nvec / 32;
    double acc = 0;
    for (i64 tile = blockIdx.x; tile < nrt * nvt; tile += gridDim.x) {
        const i64 vt = tile / nrt, rt = tile - vt * nrt;
        const i64 g0 = rt * rows, v0 = vt * 32;
        if (MODE != 0 && MODE != 4 && MODE != 6) __syncthreads();
        for (int v = warp; v < 32; v += nw) {
            const double *xc = X + (v0 + v) * ld + g0;
            if (MODE == 6) {
                for (int r = 2 * lane; r < rows; r += 64) *reinterpret_cast<double2 *>(Y + (v0 + v) * ld + g0 + r) = *reinterpret_cast<const double2 *>(xc + r);
            } else if (MODE == 2) {
                for (int r = 2 * lane; r < rows; r += 64) cp16(sm + v * (pitch + 1) + r, xc + r);
            } else {
                for (int r = lane; r < rows; r += 32) {
                    if (MODE == 0) acc += xc[r];
                    else if (MODE == 1 || MODE == 5) cp8(sm + v * pitch + r, xc + r);
                    else if (MODE == 3) sm[v * pitch + r] = xc[r];
                    else if (MODE == 4) Y[(v0 + v) * ld + g0 + r] = xc[r];
                }
            }
        }
        if (MODE == 1 || MODE == 2 || MODE == 5) { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }
        if (MODE == 5) {
            __syncthreads();
            for (int v = warp; v < 32; v += nw) {
                double *yc = Y + (v0 + v) * ld + g0;
                for (int r = lane; r < rows; r += 32) yc[r] = sm[v * pitch + r];
            }
        } else if (MODE != 0 && MODE != 4 && MODE != 6) { __syncthreads(); acc += sm[threadIdx.x]; }
    }
    if (acc == 1.234e-300) *sink = acc;
}
template <int MODE>
void run(const char *name, const double *X, double *Y, i64 ld, i64 n, i64 nvec, int rows, int threads, double *sink) {
    size_t smem = (MODE == 0 || MODE == 4 || MODE == 6) ? 0 : sizeof(double) * 32 * ((rows | 1) + 1);
    cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k<MODE>, threads, smem);
    int grid = 148 * per_sm;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 2; ++i) k<MODE><<<grid, threads, smem>>>(X, Y, ld, n, nvec, rows, sink);
    cudaEventRecord(e0);
    const int reps = 5;
    for (int i = 0; i < reps; ++i) k<MODE><<<grid, threads, smem>>>(X, Y, ld, n, nvec, rows, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    double bytes = 8.0 * (n / rows * rows) * nvec * ((MODE == 4 || MODE == 5 || MODE == 6) ? 2 : 1);
    printf("%-28s rows=%4d thr=%4d cta/sm=%2d  %7.1f us  %7.1f GB/s  err=%s\n", name, rows, threads, per_sm, ms * 1e3, bytes / ms / 1e6,
           cudaGetErrorString(cudaGetLastError()));
}
int main() {
    const i64 nvec = 1024;
    for (i64 ld : {90001LL, 90112LL}) {
        const i64 n = 90000;
        double *X, *Y, *sink;
        cudaMalloc(&X, 8 * ld * nvec + 1024); cudaMalloc(&Y, 8 * ld * nvec + 1024); cudaMalloc(&sink, 8);
        cudaMemset(X, 0, 8 * ld * nvec); cudaMemset(Y, 0, 8 * ld * nvec);
        printf("ld = %lld\n", ld);
        for (int rows : {72, 144, 288}) {
            for (int threads : {128, 256, 512}) {
                run<0>("LDG.64 -> reg", X, Y, ld, n, nvec, rows, threads, sink);
                run<1>("LDGSTS.64 -> smem", X, Y, ld, n, nvec, rows, threads, sink);
                if (ld % 2 == 0) run<2>("LDGSTS.128 -> smem", X, Y, ld, n, nvec, rows, threads, sink);
                run<3>("LDG.64 -> STS.64", X, Y, ld, n, nvec, rows, threads, sink);
                run<4>("LDG.64 -> STG.64 (copy)", X, Y, ld, n, nvec, rows, threads, sink);
                run<5>("LDGSTS -> smem -> STG (copy)", X, Y, ld, n, nvec, rows, threads, sink);
                if (ld % 2 == 0) run<6>("LDG.128 -> STG.128 (copy)", X, Y, ld, n, nvec, rows, threads, sink);
            }
        }
        cudaFree(X); cudaFree(Y); cudaFree(sink);
    }
    return 0;
}
