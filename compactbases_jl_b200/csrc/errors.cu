// errors.cu — status plumbing shared by every entry point.
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void cb_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// Scratch arrays of the library come from the device's stream-ordered pool (cudaMallocAsync).  With the default
// release threshold of 0 the pool hands its memory back to the driver at every synchronisation, which makes every
// call that needs scratch allocation-bound (B[x,:] as CSC: 2.9 ms for 1e6 points, two orders of magnitude above
// the evaluation itself; 0.36 ms with the pool kept).  Called once per device by the entry points that allocate.
int cb_keep_scratch_pool() {
    static bool kept[64] = {false};
    int dev = 0;
    CB_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !kept[dev]) {
        cudaMemPool_t pool;
        CB_CUDA(cudaDeviceGetDefaultMemPool(&pool, dev));
        unsigned long long keep = ~0ull;
        CB_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        kept[dev] = true;
    }
    return CB_OK;
}

extern "C" {
const char *cb_version(void) { return "compactbases_cuda 0.1.0 (sm_100a)"; }
const char *cb_last_error(void) { return g_err; }
int cb_device_count(int *count) {
    CB_REQUIRE(count != nullptr, "cb_device_count: NULL");
    CB_CUDA(cudaGetDeviceCount(count));
    return CB_OK;
}
}
