// bspline_core.cuh — device-side B-spline primitives.
//
// Replaces, for all K functions that are non-zero on one knot interval at once,
//   deBoor(t, e_j, x, i, m)            src/bsplines.jl:159-202
//   find_interval / within_support     src/knot_sets.jl:209-238
// The reference runs one triangular recursion per (point, function) — O(k^3) per
// point; here one Cox–de Boor sweep gives the K values in O(k^2) and m derivative
// steps cost O(m k) more.  Denominators are support widths that contain the
// (non-empty) interval, so they never vanish; knots outside 1..nt are clamped
// (KnotView::at) and only feed functions that do not exist (masked by callers),
// which is what the reference's r==1 / r==nf+j / jjj>nt branches (:173-192) do.
#pragma once
#include "common.cuh"

// 1-based expanded interval index containing x, 0 if outside the support.
// Intervals are [t_i, t_{i+1}); the last non-empty one is closed on the right
// (x == last(t) -> length(t) - mr, src/knot_sets.jl:211,233).
CB_HD int cb_find_interval(const double *__restrict__ tt, i64 n_tt,
                                                int ml, int mr, i64 nt, double x) {
    double first = tt[0], last = tt[n_tt - 1];
    if (!(x >= first) || !(x <= last)) return 0;   // also rejects NaN
    if (x == last) return (int)(nt - mr);
    i64 lo = 0, hi = n_tt;                         // first index with tt[idx] > x
    while (lo < hi) {
        i64 mid = (lo + hi) >> 1;
        if (tt[mid] > x) hi = mid; else lo = mid + 1;
    }
    return (int)(lo + (ml - 1));
}

// Knot window of interval I (1-based): tw[s] = t_{I-K+1+s}, s = 0..2K-1,
// so t_I = tw[K-1], t_{I+1} = tw[K].
template <int K>
CB_HD void cb_knot_window(const double *__restrict__ tt, i64 n_tt, int ml,
                                               int I, double (&tw)[2 * K]) {
#pragma unroll
    for (int s = 0; s < 2 * K; ++s) {
        i64 j = (i64)(I - K + s) - (ml - 1);       // 0-based expanded index I-K+1+s-1
        j = j < 0 ? 0 : (j > n_tt - 1 ? n_tt - 1 : j);
        tw[s] = tt[j];
    }
}

// N[r], r = 0..K-1: m-th derivative at x of function I-K+1+r (order K).
// MT >= 0 fixes the derivative order at compile time (m is then ignored): the level loops lose
// their run-time selects, which matters in the FP64-bound evaluation kernel.
// The sweep works on the 2(K-1) distances  dh[r] = t_{I+1+r} - x  and  dl[i] = x - t_{I-K+1+i}  (both >= 0 for x in
// the interval): the support widths are their sums, dh[r] + dl[K-j+r] = t_{I+1+r} - t_{I+1+r-j}, so a level entry
// costs one DADD instead of three.
template <int K, int MT = -1>
CB_HD void cb_bspline_all(const double (&tw)[2 * K], double x, int m_rt,
                                               double (&N)[K]) {
    const int m = MT >= 0 ? MT : m_rt;
#pragma unroll
    for (int r = 0; r < K; ++r) N[r] = 0.0;
    if (m >= K) return;                            // k == 1 && m > 0 -> 0 (src/bsplines.jl:164)
    N[0] = 1.0;
    double dh[K], dl[K];
#pragma unroll
    for (int r = 0; r < K; ++r) { dh[r] = tw[K + r] - x; dl[r] = x - tw[r]; }
    const int jmax = K - 1 - m;                    // value recursion up to order K-m
#pragma unroll
    for (int j = 1; j <= K - 1; ++j) {
        if (j <= jmax) {
            double saved = 0.0;
#pragma unroll
            for (int r = 0; r < j; ++r) {
                const double term = N[r] * fast_rcp(dh[r] + dl[K - j + r]);
                N[r] = fma(dh[r], term, saved);
                saved = dl[K - j + r] * term;
            }
            N[j] = saved;
        }
    }
    // derivative steps: order o -> o+1,  new[r] = o*(q[r-1] - q[r]),
    // q[r] = old[r] / (t_{f+o} - t_f) with f the function of old[r].
#pragma unroll
    for (int o = 1; o <= K - 1; ++o) {
        if (o >= K - m) {
            double qprev = 0.0;                    // q[r-1]
#pragma unroll
            for (int r = 0; r <= o; ++r) {
                double qr = 0.0;
                if (r < o) qr = N[r] * fast_rcp(dh[r] + dl[K - o + r]);
                N[r] = (double)o * (qprev - qr);
                qprev = qr;
            }
        }
    }
}

// Same recursion with the reciprocal knot differences supplied by the caller:
// rc[j*(j-1)/2 + r] = 1 / (t_{I+1+r} - t_{I+1+r-j}), j = 1..K-1, r = 0..j-1 (tw[K+r] - tw[K-j+r]).
// They depend on the interval only, so quadrature assembly computes them once per interval and
// shares them among its q nodes.
template <int K>
__device__ __forceinline__ void cb_knot_rcp(const double (&tw)[2 * K], int idx, double &out) {
    // idx -> (j, r)
    int j = 1, base = 0;
    while (base + j <= idx) { base += j; ++j; }
    const int r = idx - base;
    out = fast_rcp(tw[K + r] - tw[K - j + r]);
}

template <int K>
__device__ __forceinline__ void cb_bspline_all_rc(const double (&tw)[2 * K], const double *__restrict__ rc, double x, int m,
                                                  double (&N)[K]) {
#pragma unroll
    for (int r = 0; r < K; ++r) N[r] = 0.0;
    if (m >= K) return;
    N[0] = 1.0;
    const int jmax = K - 1 - m;
    // level j is a value step (order j -> j+1) up to jmax and a derivative step after it: one
    // warp-uniform branch per level instead of per-entry selects
#pragma unroll
    for (int j = 1; j <= K - 1; ++j) {
        if (j <= jmax) {
            double saved = 0.0;
#pragma unroll
            for (int r = 0; r < j; ++r) {
                const double term = N[r] * rc[j * (j - 1) / 2 + r];
                N[r] = fma(tw[K + r] - x, term, saved);
                saved = (x - tw[K - j + r]) * term;
            }
            N[j] = saved;
        } else {
            double qprev = 0.0;
#pragma unroll
            for (int r = 0; r <= j; ++r) {
                double qr = 0.0;
                if (r < j) qr = N[r] * rc[j * (j - 1) / 2 + r];
                N[r] = (double)j * (qprev - qr);
                qprev = qr;
            }
        }
    }
}
