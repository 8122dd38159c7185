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

extern "C" {
const char *cb_version(void) { return "compactbases_cuda 0.1.0 (sm_100a)"; }
const char *cb_last_error(void) { return g_err; }
int cb_device_count(int *count) {
    CB_REQUIRE(count != nullptr, "cb_device_count: NULL");
    CB_CUDA(cudaGetDeviceCount(count));
    return CB_OK;
}
}
