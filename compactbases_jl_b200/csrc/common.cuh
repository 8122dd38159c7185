// common.cuh — shared helpers for libcompactbases_cuda (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/compactbases_cuda.h"

typedef int64_t i64;
#define CB_HD __host__ __device__ __forceinline__

#define CB_NUM_SMS 148  // B200; grid sizes are multiples of this

// ---- error plumbing -------------------------------------------------------
void cb_set_error(const char *fmt, ...);
int cb_keep_scratch_pool();                    // errors.cu: keep the stream-ordered pool's memory between calls

#define CB_CUDA(call)                                                            \
    do {                                                                         \
        cudaError_t e__ = (call);                                                \
        if (e__ != cudaSuccess) {                                                \
            cb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,           \
                         cudaGetErrorString(e__));                               \
            (void)cudaGetLastError(); /* reported here: a later launch check must not see it again */ \
            return CB_ERR_CUDA;                                                  \
        }                                                                        \
    } while (0)

#define CB_REQUIRE(cond, ...)                                                    \
    do {                                                                         \
        if (!(cond)) {                                                           \
            cb_set_error(__VA_ARGS__);                                           \
            return CB_ERR_ARG;                                                   \
        }                                                                        \
    } while (0)

#define CB_LAUNCH_CHECK() CB_CUDA(cudaGetLastError())

// ---- B-spline basis handle --------------------------------------------------
struct cb_bspline_s {
    int k, ml, mr, q;
    i64 n_tt, nt, nf, ni;        // ni = number of non-empty intervals
    int device;
    double *d_tt;                // t.t on the device
    double *d_xi, *d_wr;         // reference Gauss-Legendre rule
    i64 *d_ivlist;               // [ni] expanded (1-based) index of each non-empty interval
    int32_t *d_ord;              // [nt] ordinal of interval I (entry I-1) or -1 if empty
    double *d_x, *d_w;           // lgwt nodes / weights, [ni*q]
    std::vector<double> h_tt, h_xi, h_wr;
    std::vector<i64> h_ivlist;
};

// Virtual knot vector with implicit end multiplicities (src/knot_sets.jl:138-151).
// i is the 0-based index into the expanded vector; indices outside are clamped
// to the end knots, which is what the all-functions-at-once recursion needs.
struct KnotView {
    const double *tt;
    i64 n_tt;
    int ml, mr;
    i64 nt;
    CB_HD double at(i64 i) const {
        i64 j = i - (ml - 1);
        j = j < 0 ? 0 : (j > n_tt - 1 ? n_tt - 1 : j);
        return tt[j];
    }
};

static inline KnotView make_knot_view(const cb_bspline_s *B) {
    KnotView kv;
    kv.tt = B->d_tt; kv.n_tt = B->n_tt; kv.ml = B->ml; kv.mr = B->mr; kv.nt = B->nt;
    return kv;
}

// ---- device math ------------------------------------------------------------
// 1/d to within ~1 ulp: MUFU.RCP64H seed y0 (relative error e0 <= 2^-19.9: it reads the upper 32 bits of d) and one
// cubically convergent step  y = y0 (1 + e + e^2),  e = 1 - d y0  (error e0^3 < 2^-59), three DFMAs instead of the four
// of two Newton steps.  No denormal/inf special-casing (knot differences are normal numbers); d == 0 must be guarded
// by the caller.
CB_HD double fast_rcp(double d) {
#ifdef __CUDA_ARCH__
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
    const double e = fma(-d, y, 1.0);
    const double t = fma(e, e, e);
    return fma(y, t, y);
#else
    return 1.0 / d;   // host instantiation exists only for tests/hostcheck
#endif
}

CB_HD double guarded_rcp(double d) { return d != 0.0 ? fast_rcp(d) : 0.0; }

// y = alpha*v + beta*y with the BLAS convention that beta == 0 does not read y.
CB_HD double axpby(double v, double alpha, double beta, const double *y) {
    return beta == 0.0 ? alpha * v : fma(alpha, v, beta * (*y));
}

static inline int cb_grid_for(i64 work_items, int per_block, int max_waves = 8) {
    i64 g = (work_items + per_block - 1) / per_block;
    i64 cap = (i64)CB_NUM_SMS * max_waves;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}
