// tests/hostcheck/hostcheck.cu — TEST INFRASTRUCTURE.
// Compiles the product's __host__ __device__ math (bspline_core.cuh, ...) for the
// host so that the algorithms can be checked against the oracle on a machine
// without a GPU.  Never loaded by the product package.
#include "../../compactbases_jl_b200/csrc/bspline_core.cuh"

void cb_set_error(const char *, ...) {}

template <int K>
static void eval_k(const double *tt, i64 n_tt, int ml, int mr, i64 nt, const double *x, i64 nx, int m,
                   int32_t *iv, double *vals) {
    i64 nf = nt - K;
    for (i64 p = 0; p < nx; ++p) {
        int I = cb_find_interval(tt, n_tt, ml, mr, nt, x[p]);
        iv[p] = I;
        double N[K];
        for (int r = 0; r < K; ++r) N[r] = 0;
        if (I > 0) {
            double tw[2 * K];
            cb_knot_window<K>(tt, n_tt, ml, I, tw);
            cb_bspline_all<K>(tw, x[p], m, N);
        }
        for (int r = 0; r < K; ++r) {
            i64 j = (i64)I - K + 1 + r;
            vals[p * K + r] = (I > 0 && j >= 1 && j <= nf) ? N[r] : 0.0;
        }
    }
}

extern "C" int hc_bspline_eval(const double *tt, i64 n_tt, int k, int ml, int mr, const double *x, i64 nx, int m,
                               int32_t *iv, double *vals) {
    i64 nt = n_tt + ml + mr - 2;
    switch (k) {
#define C(KK) case KK: eval_k<KK>(tt, n_tt, ml, mr, nt, x, nx, m, iv, vals); return 0;
        C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9) C(10) C(11) C(12)
#undef C
    }
    return 3;
}
