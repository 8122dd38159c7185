/*
 * oracle/cbo.c — CPU restatement ("oracle") of the CompactBases.jl hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under compactbases_jl_b200/ may import,
 * link or call this file; only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, and there only as
 * the checker / the CPU arm.
 *
 * Every function cites the reference file:line (relative to the reference
 * tree, CompactBases.jl v0.3.1) whose algorithm it restates.  The code keeps
 * the reference's arithmetic order (Julia does not contract a*b+c into an
 * fma; the one explicit fma is lerp), so build with -ffp-contract=off.
 *
 * Parity pinned: checked in tests/test_oracle_golden.py against the
 * reference's own known-answer matrices (test/bsplines/splines.jl:3-16,
 * test/bsplines/derivatives.jl:19-150, README.md:160-182, ...).
 *
 * Index conventions: all public entry points take/return 1-BASED interval and
 * function indices exactly like the Julia code; C arrays are 0-based.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int64_t i64;

/* ------------------------------------------------------------------ */
/* a1. Knot sets with implicit end multiplicities                      */
/* ------------------------------------------------------------------ */

/* src/knot_sets.jl:116  length(t) = length(t.t) + ml + mr - 2 */
i64 cbo_knot_length(i64 n_tt, int ml, int mr) { return n_tt + ml + mr - 2; }

/* src/knot_sets.jl:138-151  getindex(t, i), i 1-based.
 * ni = numintervals = n_tt-1;  i<ml -> first; i<ml+ni -> t.t[i-ml+1];
 * i<ml+ni+mr -> last. */
static inline double knot_get(const double *tt, i64 n_tt, int ml, int mr, i64 i)
{
    i64 ni = n_tt - 1;
    if (i < ml) return tt[0];
    if (i < ml + ni) return tt[i - ml];          /* t.t[i-ml+1], 1-based */
    (void)mr;
    return tt[n_tt - 1];
}

/* Materialise the virtual knot vector (what collect(t) gives). */
void cbo_expand_knots(const double *tt, i64 n_tt, int ml, int mr, double *t)
{
    i64 nt = cbo_knot_length(n_tt, ml, mr);
    for (i64 i = 1; i <= nt; ++i) t[i - 1] = knot_get(tt, n_tt, ml, mr, i);
}

/* src/knot_sets.jl:196-198  nonempty_intervals: j in 1:length(t)-1 with
 * t[j] != t[j+1].  Returns the count; out (may be NULL) gets the 1-based js. */
i64 cbo_nonempty_intervals(const double *t, i64 nt, i64 *out)
{
    i64 c = 0;
    for (i64 j = 1; j <= nt - 1; ++j)
        if (t[j - 1] != t[j]) { if (out) out[c] = j; ++c; }
    return c;
}

/* src/knot_sets.jl:209-216  find_interval(t, x, i=ml): literal linear scan.
 * Returns 0 for `nothing`. */
i64 cbo_find_interval(const double *t, i64 nt, int mr, double x, i64 i)
{
    double first = t[0], last = t[nt - 1];
    if (x < first || x > last || i > nt || x < t[i - 1]) return 0;
    if (x == last) return nt - mr;
    for (i64 r = i; r <= nt; ++r)
        if (t[r - 1] > x) return r - 1;
    return -1; /* @assert false */
}

/* Linear-cost twin of the above for big inputs: upper_bound by bisection.
 * Same result as cbo_find_interval(t, nt, mr, x, ml) for any non-decreasing t
 * (checked in tests). */
i64 cbo_find_interval_bisect(const double *t, i64 nt, int mr, double x)
{
    double first = t[0], last = t[nt - 1];
    if (!(x >= first) || !(x <= last)) return 0;
    if (x == last) return nt - mr;
    i64 lo = 0, hi = nt;            /* first index (0-based) with t[idx] > x */
    while (lo < hi) {
        i64 mid = (lo + hi) >> 1;
        if (t[mid] > x) hi = mid; else lo = mid + 1;
    }
    return lo;                      /* 1-based (r-1) where r = lo+1 */
}

/* ------------------------------------------------------------------ */
/* a3. de Boor's algorithm                                             */
/* ------------------------------------------------------------------ */

#define CBO_KMAX 64

/* src/bsplines.jl:159-202  deBoor(t, c, x, i, m) with c = UnitVector(nc, j)
 * (src/unit_vectors.jl:7-13).  i is the 1-based interval (0 = nothing).
 * Loop bounds, the r==1 / r==nf+j / jjj>nt edge branches and the order of
 * the floating-point operations are the reference's. */
double cbo_deboor_unit(const double *t, i64 nt, int k, i64 nc, i64 junit,
                       double x, i64 i, int m)
{
    if (i == 0) return 0.0;                                   /* :161 */
    if (k == 1) return (nc < i || m > 0) ? 0.0 : (i == junit ? 1.0 : 0.0); /* :164 */
    double alpha[CBO_KMAX];
    for (int q = 0; q < k; ++q) {                             /* :166-167 */
        i64 r = i - k + 1 + q;
        alpha[q] = (r > 0 && r <= nc && r == junit) ? 1.0 : 0.0;
    }
    i64 nf = nt - k;                                          /* :169 */
    for (int j = 1; j <= k - 1; ++j) {                        /* :170 */
        i64 rlo = i - k + j; if (rlo < 1) rlo = 1;
        for (i64 r = i; r >= rlo; --r) {                      /* :171 */
            i64 jjj = r + k - j;
            if (jjj > nt || jjj < 1) continue;                /* :173 */
            i64 rp = r - i + k;                               /* r', 1-based */
            if (r != 1 && rp == 1) continue;                  /* :175 */
            double a = t[r + k - j - 1];
            double b = t[r - 1];
            double v;
            if (j <= m) {                                     /* :180-187 */
                double d;
                if (r == 1) d = -alpha[rp - 1];
                else if (r == nf + j) d = alpha[rp - 2];
                else d = alpha[rp - 2] - alpha[rp - 1];
                v = (double)(k - j) * d;
            } else {                                          /* :188-196 */
                if (r == 1) v = (b - x) * alpha[rp - 1];
                else if (r == nf + j) v = (x - a) * alpha[rp - 2];
                else v = ((x - a) * alpha[rp - 2] + (b - x) * alpha[rp - 1]);
            }
            alpha[rp - 1] = v / (b - a);                      /* :197 */
        }
    }
    return alpha[k - 1];                                      /* :201 */
}

/* Same routine for a general coefficient vector c[0..nc) (spline evaluation
 * (B*c)[x], test/bsplines/splines.jl:81). */
double cbo_deboor(const double *t, i64 nt, int k, const double *c, i64 nc,
                  double x, i64 i, int m)
{
    if (i == 0) return 0.0;
    if (k == 1) return (nc < i || m > 0) ? 0.0 : c[i - 1];
    double alpha[CBO_KMAX];
    for (int q = 0; q < k; ++q) {
        i64 r = i - k + 1 + q;
        alpha[q] = (r > 0 && r <= nc) ? c[r - 1] : 0.0;
    }
    i64 nf = nt - k;
    for (int j = 1; j <= k - 1; ++j) {
        i64 rlo = i - k + j; if (rlo < 1) rlo = 1;
        for (i64 r = i; r >= rlo; --r) {
            i64 jjj = r + k - j;
            if (jjj > nt || jjj < 1) continue;
            i64 rp = r - i + k;
            if (r != 1 && rp == 1) continue;
            double a = t[r + k - j - 1];
            double b = t[r - 1];
            double v;
            if (j <= m) {
                double d;
                if (r == 1) d = -alpha[rp - 1];
                else if (r == nf + j) d = alpha[rp - 2];
                else d = alpha[rp - 2] - alpha[rp - 1];
                v = (double)(k - j) * d;
            } else {
                if (r == 1) v = (b - x) * alpha[rp - 1];
                else if (r == nf + j) v = (x - a) * alpha[rp - 2];
                else v = ((x - a) * alpha[rp - 2] + (b - x) * alpha[rp - 1]);
            }
            alpha[rp - 1] = v / (b - a);
        }
    }
    return alpha[k - 1];
}

/* ------------------------------------------------------------------ */
/* a2/a4. B[x, sel]                                                    */
/* ------------------------------------------------------------------ */

/* Row ("ELL") form of src/bsplines.jl:204-224: for every point the interval
 * (0 = outside the support; intervals are right-open except the last,
 * src/knot_sets.jl:228-238, which makes within_support and find_interval
 * agree) and the k values deBoor(t, e_j, x, i, m), j = i-k+1 .. i
 * (0 where j is not a function index 1..nf).  vals is nx*k row-major. */
void cbo_bspline_eval_ell(const double *t, i64 nt, int k, int mr,
                          const double *x, i64 nx, int m,
                          int32_t *interval, double *vals)
{
    i64 nf = nt - k;
    for (i64 p = 0; p < nx; ++p) {
        i64 i = cbo_find_interval_bisect(t, nt, mr, x[p]);
        interval[p] = (int32_t)i;
        for (int a = 0; a < k; ++a) {
            i64 j = i - k + 1 + a;
            vals[p * k + a] = (i > 0 && j >= 1 && j <= nf)
                ? cbo_deboor_unit(t, nt, k, nf, j, x[p], i, m) : 0.0;
        }
    }
}

/* Literal cost model of src/bsplines.jl:208-224 + src/knot_sets.jl:228-238:
 * for every selected function and each of its k intervals scan ALL points.
 * Produces the same numbers as the ELL form; used only to time the
 * reference's real O(nf*k*nx) behaviour on a reduced size.  Dense output
 * chi[nx * ncols] column-major (zero where nothing is stored). */
void cbo_bspline_eval_literal(const double *tt, i64 n_tt, int k, int ml, int mr,
                              const double *x, i64 nx, i64 col0, i64 ncols,
                              double *chi)
{
    i64 nt = cbo_knot_length(n_tt, ml, mr);
    double *t = (double *)malloc(sizeof(double) * nt);
    cbo_expand_knots(tt, n_tt, ml, mr, t);
    i64 nf = nt - k, ni = n_tt - 1;
    memset(chi, 0, sizeof(double) * nx * ncols);
    for (i64 jj = 0; jj < ncols; ++jj) {
        i64 j = col0 + 1 + jj;
        for (i64 i = j; i <= j + k - 1; ++i) {
            double lo = t[i - 1], hi = t[i];
            int closed = (i - ml + 1 == ni);
            for (i64 p = 0; p < nx; ++p) {
                int in = closed ? (x[p] >= lo && x[p] <= hi)
                                : (x[p] >= lo && x[p] < hi);
                if (in) chi[jj * nx + p] = cbo_deboor_unit(t, nt, k, nf, j, x[p], i, 0);
            }
        }
    }
    free(t);
}

/* SparseMatrixCSC{Float64,Int64} form of B[x, col0+1 : col0+ncols]
 * (src/bsplines.jl:217-224): 1-based colptr/rowval, rows ascending within a
 * column, exact zeros not stored (sparse setindex! of a zero into an unstored
 * slot is a no-op).  Two passes (count, fill); returns nnz.  rowval/nzval
 * need room for min(k,ncols)*nx entries. */
i64 cbo_bspline_eval_csc(const double *t, i64 nt, int k, int mr,
                         const double *x, i64 nx, i64 col0, i64 ncols,
                         i64 *colptr, i64 *rowval, double *nzval)
{
    i64 nf = nt - k;
    i64 *cnt = (i64 *)calloc((size_t)ncols + 1, sizeof(i64));
    int32_t *iv = (int32_t *)malloc(sizeof(int32_t) * (size_t)(nx ? nx : 1));
    double *v = (double *)malloc(sizeof(double) * (size_t)(nx ? nx : 1) * k);
    cbo_bspline_eval_ell(t, nt, k, mr, x, nx, 0, iv, v);
    for (i64 p = 0; p < nx; ++p) {
        if (iv[p] == 0) continue;
        for (int a = 0; a < k; ++a) {
            i64 j = (i64)iv[p] - k + 1 + a;
            if (j < 1 || j > nf) continue;
            i64 c = j - 1 - col0;
            if (c < 0 || c >= ncols) continue;
            if (v[p * k + a] != 0.0) cnt[c + 1]++;
        }
    }
    colptr[0] = 1;
    for (i64 c = 0; c < ncols; ++c) colptr[c + 1] = colptr[c] + cnt[c + 1];
    i64 nnz = colptr[ncols] - 1;
    for (i64 c = 0; c < ncols; ++c) cnt[c] = colptr[c] - 1;   /* write cursors */
    for (i64 p = 0; p < nx; ++p) {
        if (iv[p] == 0) continue;
        for (int a = 0; a < k; ++a) {
            i64 j = (i64)iv[p] - k + 1 + a;
            if (j < 1 || j > nf) continue;
            i64 c = j - 1 - col0;
            if (c < 0 || c >= ncols) continue;
            double val = v[p * k + a];
            if (val != 0.0) { rowval[cnt[c]] = p + 1; nzval[cnt[c]] = val; cnt[c]++; }
        }
    }
    free(cnt); free(iv); free(v);
    return nnz;
}

/* ------------------------------------------------------------------ */
/* a5. Gauss–Legendre nodes per non-empty interval                     */
/* ------------------------------------------------------------------ */

/* src/quadrature.jl:4  lerp(a,b,t) = fma(t, b, fma(-t, a, a)) */
static inline double lerp_(double a, double b, double t) { return fma(t, b, fma(-t, a, a)); }

/* src/quadrature.jl:38-41 */
int cbo_num_quadrature_points(int k, int kp) { int N2 = 2 * (k - 1) + kp; return (N2 >> 1) + (N2 & 1); }

/* src/quadrature.jl:60-77 (lgwt) + :18-23 (change_interval!):
 * xs = lerp(a, b, (xi+1)/2), ws = (b-a)*w/2 for every non-empty interval.
 * Returns the number of non-empty intervals; x,w have ni*q entries. */
i64 cbo_lgwt(const double *t, i64 nt, const double *xi, const double *wr, int q,
             double *x, double *w)
{
    i64 c = 0;
    for (i64 j = 1; j <= nt - 1; ++j) {
        double a = t[j - 1], b = t[j];
        if (a == b) continue;
        for (int l = 0; l < q; ++l) {
            x[c * q + l] = lerp_(a, b, (xi[l] + 1) / 2);
            w[c * q + l] = (b - a) * wr[l] / 2;
        }
        ++c;
    }
    return c;
}

/* ------------------------------------------------------------------ */
/* a6. basis_functions(t, x, m) as dense per-interval tables            */
/* ------------------------------------------------------------------ */

/* src/bsplines.jl:28-52.  For the c-th non-empty interval i and node l the
 * functions j = max(1,i-k+1)..min(i,nf) are evaluated with deBoor(t,e_j,x,i,m)
 * (:45).  table[(c*q + l)*k + a], a = j-(i-k+1); 0 where j is no function. */
void cbo_basis_table(const double *t, i64 nt, int k, const double *x, int q,
                     int m, double *table)
{
    extern int cbo_get_threads(void);
    i64 nf = nt - k;
    i64 ni = cbo_nonempty_intervals(t, nt, NULL);
    i64 *iv = (i64 *)malloc(sizeof(i64) * (size_t)(ni ? ni : 1));
    cbo_nonempty_intervals(t, nt, iv);
    int nthreads = cbo_get_threads();
    (void)nthreads;
#pragma omp parallel for schedule(static) num_threads(nthreads)
    for (i64 c = 0; c < ni; ++c) {
        i64 i = iv[c];
        for (int l = 0; l < q; ++l)
            for (int a = 0; a < k; ++a) {
                i64 j = i - k + 1 + a;
                table[(c * q + l) * k + a] = (j >= 1 && j <= nf)
                    ? cbo_deboor_unit(t, nt, k, nf, j, x[c * q + l], i, m) : 0.0;
            }
    }
    free(iv);
}

/* ------------------------------------------------------------------ */
/* a7-a10. overlap_matrix! / diffop! into BandedMatrix storage          */
/* ------------------------------------------------------------------ */

/* src/bsplines.jl:58-73 with chi = (-1)^a basis_functions(L,a),
 * xi = op * basis_functions(R,b) (src/bspline_derivatives.jl:14-17):
 *   S[i,j] = sum_l (chi[l,i]*w[l]) * (op[l]*xi[l,j]),  l increasing,
 * for |i-j| <= p = max(l,u) and inside the (l,u) band; rows are functions
 * rowL0+1.., columns colR0+1.. (restrictions, src/bsplines.jl:258-270).
 * tL/tR are the expanded knot vectors of the two bases (same t.t, same nodes;
 * src/bsplines.jl:75-82).  band is (l+u+1) x n column-major,
 * band[(u+i-j) + (l+u+1)*j] (0-based i,j); slots outside the matrix are 0. */
/* Number of threads the *_mt entry points and the table / assembly loops may use (1 = the single-threaded
 * reference; bench.py raises it for the "OpenMP over all host cores" baseline).  Results do not depend on it:
 * every output entry is computed by one thread with the reference's summation order. */
static int g_threads = 1;
void cbo_set_threads(int n) { g_threads = n < 1 ? 1 : n; }
int cbo_get_threads(void) { return g_threads; }

void cbo_bspline_assemble(const double *tL, i64 ntL, int kL,
                          const double *tR, i64 ntR, int kR,
                          const double *x, const double *w, int q,
                          int da, int db, const double *op,
                          i64 rowL0, i64 m, i64 colR0, i64 n,
                          int l, int u, double *band)
{
    i64 nfL = ntL - kL, nfR = ntR - kR;
    int ld = l + u + 1;
    memset(band, 0, sizeof(double) * (size_t)ld * (size_t)n);
    /* non-empty interval numbering: c-th non-empty interval <-> knot interval
     * index in each expanded vector.  The same t.t => same sequence of
     * non-empty intervals; only the leading offset (ml-1) differs. */
    i64 niL = cbo_nonempty_intervals(tL, ntL, NULL);
    i64 niR = cbo_nonempty_intervals(tR, ntR, NULL);
    i64 ni = niL < niR ? niL : niR;
    i64 *ivL = (i64 *)malloc(sizeof(i64) * (size_t)(niL ? niL : 1));
    i64 *ivR = (i64 *)malloc(sizeof(i64) * (size_t)(niR ? niR : 1));
    cbo_nonempty_intervals(tL, ntL, ivL);
    cbo_nonempty_intervals(tR, ntR, ivR);
    /* first ordinal c with ivL[c] >= I, for every expanded interval index I of L: the reference's sparse dot
     * product (src/bsplines.jl:67-71) only meets the stored rows of column i, i.e. the intervals of function fi;
     * visiting exactly those, in increasing order, gives the same sum at linear instead of quadratic cost */
    i64 *firstL = (i64 *)malloc(sizeof(i64) * (size_t)(ntL + 2));
    { i64 c = 0; for (i64 I = 0; I <= ntL + 1; ++I) { while (c < niL && ivL[c] < I) ++c; firstL[I] = c; } }
    double *TL = (double *)malloc(sizeof(double) * (size_t)(niL ? niL : 1) * q * kL);
    double *TR = (double *)malloc(sizeof(double) * (size_t)(niR ? niR : 1) * q * kR);
    cbo_basis_table(tL, ntL, kL, x, q, da, TL);
    cbo_basis_table(tR, ntR, kR, x, q, db, TR);
    double sgn = (da % 2 == 0) ? 1.0 : -1.0;       /* src/bspline_derivatives.jl:14 */
    int p = l > u ? l : u;
    i64 mm = m < nfL - rowL0 ? m : nfL - rowL0;    /* min(size(S,1), size(chi,2)) */
    i64 nn = n < nfR - colR0 ? n : nfR - colR0;
#pragma omp parallel for schedule(static) num_threads(g_threads)
    for (i64 i = 0; i < mm; ++i) {
        i64 jlo = i - p < 0 ? 0 : i - p;
        i64 jhi = i + p > nn - 1 ? nn - 1 : i + p;
        i64 fi = rowL0 + 1 + i;                    /* 1-based function of L: lives on expanded intervals fi..fi+kL-1 */
        i64 Ihi = fi + kL > ntL + 1 ? ntL + 1 : fi + kL;
        i64 c0 = firstL[fi], c1 = firstL[Ihi];
        if (c1 > ni) c1 = ni;
        for (i64 j = jlo; j <= jhi; ++j) {
            if (j - i > u || i - j > l) continue;  /* outside the band: zero */
            i64 fj = colR0 + 1 + j;
            double s = 0.0;
            for (i64 c = c0; c < c1; ++c) {
                i64 aL = fi - (ivL[c] - kL + 1), aR = fj - (ivR[c] - kR + 1);
                if (aL < 0 || aL >= kL || aR < 0 || aR >= kR) continue;
                for (int ll = 0; ll < q; ++ll) {
                    double ch = sgn * TL[(c * q + ll) * kL + aL];
                    double xv = TR[(c * q + ll) * kR + aR];
                    if (op) xv = op[c * q + ll] * xv;
                    s += (ch * w[c * q + ll]) * xv;
                }
            }
            band[(u + i - j) + (i64)ld * j] = s;
        }
    }
    free(ivL); free(ivR); free(firstL); free(TL); free(TR);
}

/* The literal loop of overlap_matrix! (src/bsplines.jl:58-73) on sparse tables costs, per band entry, a sparse
 * column copy and a sparse dot over ALL stored rows; restated here as the scan over every non-empty interval
 * (quadratic in the number of intervals).  Kept for the reduced-size "literal algorithm" baseline of bench.py
 * and to pin the linear-cost form above (tests/test_oracle_golden.py). */
void cbo_bspline_assemble_literal(const double *tL, i64 ntL, int kL,
                                  const double *tR, i64 ntR, int kR,
                                  const double *x, const double *w, int q,
                                  int da, int db, const double *op,
                                  i64 rowL0, i64 m, i64 colR0, i64 n,
                                  int l, int u, double *band)
{
    i64 nfL = ntL - kL, nfR = ntR - kR;
    int ld = l + u + 1;
    memset(band, 0, sizeof(double) * (size_t)ld * (size_t)n);
    i64 niL = cbo_nonempty_intervals(tL, ntL, NULL);
    i64 niR = cbo_nonempty_intervals(tR, ntR, NULL);
    i64 ni = niL < niR ? niL : niR;
    i64 *ivL = (i64 *)malloc(sizeof(i64) * (size_t)(niL ? niL : 1));
    i64 *ivR = (i64 *)malloc(sizeof(i64) * (size_t)(niR ? niR : 1));
    cbo_nonempty_intervals(tL, ntL, ivL);
    cbo_nonempty_intervals(tR, ntR, ivR);
    double *TL = (double *)malloc(sizeof(double) * (size_t)(niL ? niL : 1) * q * kL);
    double *TR = (double *)malloc(sizeof(double) * (size_t)(niR ? niR : 1) * q * kR);
    cbo_basis_table(tL, ntL, kL, x, q, da, TL);
    cbo_basis_table(tR, ntR, kR, x, q, db, TR);
    double sgn = (da % 2 == 0) ? 1.0 : -1.0;
    int p = l > u ? l : u;
    i64 mm = m < nfL - rowL0 ? m : nfL - rowL0;
    i64 nn = n < nfR - colR0 ? n : nfR - colR0;
    for (i64 i = 0; i < mm; ++i) {
        i64 jlo = i - p < 0 ? 0 : i - p;
        i64 jhi = i + p > nn - 1 ? nn - 1 : i + p;
        i64 fi = rowL0 + 1 + i;
        for (i64 j = jlo; j <= jhi; ++j) {
            if (j - i > u || i - j > l) continue;
            i64 fj = colR0 + 1 + j;
            double s = 0.0;
            for (i64 c = 0; c < ni; ++c) {
                i64 aL = fi - (ivL[c] - kL + 1), aR = fj - (ivR[c] - kR + 1);
                if (aL < 0 || aL >= kL || aR < 0 || aR >= kR) continue;
                for (int ll = 0; ll < q; ++ll) {
                    double ch = sgn * TL[(c * q + ll) * kL + aL];
                    double xv = TR[(c * q + ll) * kR + aR];
                    if (op) xv = op[c * q + ll] * xv;
                    s += (ch * w[c * q + ll]) * xv;
                }
            }
            band[(u + i - j) + (i64)ld * j] = s;
        }
    }
    free(ivL); free(ivR); free(TL); free(TR);
}

/* ------------------------------------------------------------------ */
/* a11-a14. FE-DVR element matrices                                     */
/* ------------------------------------------------------------------ */

/* src/fedvr_derivatives.jl:15-31  lagrangeder!(x, w, L').  Lp is o x o
 * column-major: Lp[m + o*mp] = L'[m,m'] = d/dx L_m at node m'. */
void cbo_lagrangeder(const double *xe, const double *we, int o, double *Lp)
{
    for (int m = 0; m < o; ++m) {
        int dmn = (m == o - 1), dm1 = (m == 0);
        Lp[m + o * m] = (double)(dmn - dm1) / (2 * we[m]);
        for (int mp = 0; mp < o; ++mp) {
            if (mp == m) continue;
            double f = 1;
            for (int kk = 0; kk < o; ++kk) {
                if (kk == m || kk == mp) continue;
                f *= (xe[mp] - xe[kk]) / (xe[m] - xe[kk]);
            }
            Lp[m + o * mp] = f / (xe[m] - xe[mp]);
        }
    }
}

/* src/fedvr_derivatives.jl:33-57  diff(B, n, i): D[m,m'] = n_m n_m' *
 * dot(Lt[m,:], w .* L'[m',:]), Lt = I (nder=1) or -L' (nder=2).
 * D is o x o column-major. */
void cbo_fedvr_element(const double *xe, const double *we, const double *ne,
                       int o, int nder, double *D)
{
    double *Lp = (double *)malloc(sizeof(double) * o * o);
    cbo_lagrangeder(xe, we, o, Lp);
    for (int m = 0; m < o; ++m)
        for (int mp = 0; mp < o; ++mp) {
            double s = 0.0;
            for (int qd = 0; qd < o; ++qd) {
                double lt = (nder == 1) ? (m == qd ? 1.0 : 0.0) : -Lp[m + o * qd];
                s += lt * (we[qd] * Lp[mp + o * qd]);
            }
            D[m + o * mp] = ne[m] * ne[mp] * s;
        }
    free(Lp);
}

/* All element blocks of derop!(A, B, nder) (src/fedvr_derivatives.jl:62-63):
 * blocks[boff[e] ...] = diff(B, nder, e), element e uses nodes
 * estart[e] .. estart[e]+order[e]-1 (0-based, neighbours share one node),
 * w is the concatenation of the per-element weight vectors B.w^i. */
void cbo_fedvr_blocks(const double *x, const double *wcat, const double *nrm,
                      const i64 *estart, const int32_t *order, i64 nel,
                      int nder, const i64 *boff, double *blocks)
{
    i64 woff = 0;
    for (i64 e = 0; e < nel; ++e) {
        cbo_fedvr_element(x + estart[e], wcat + woff, nrm + estart[e], order[e],
                          nder, blocks + boff[e]);
        woff += order[e];
    }
}

/* Dense n x n matrix equal to what set_elements! (src/fedvr_operators.jl:
 * 40-154) builds block by block: every element block is placed at its node
 * range, neighbouring corners are summed on the shared bridge node; then the
 * restriction a..b (1-based, inclusive) is cut out.  A is (b-a+1)^2
 * column-major. */
void cbo_fedvr_dense(const double *blocks, const i64 *boff, const i64 *estart,
                     const int32_t *order, i64 nel, i64 a, i64 b, double *A)
{
    i64 n = b - a + 1;
    memset(A, 0, sizeof(double) * (size_t)n * (size_t)n);
    for (i64 e = 0; e < nel; ++e) {
        int o = order[e];
        for (int mp = 0; mp < o; ++mp)
            for (int m = 0; m < o; ++m) {
                i64 gi = estart[e] + m + 1, gj = estart[e] + mp + 1; /* 1-based */
                if (gi < a || gi > b || gj < a || gj > b) continue;
                A[(gi - a) + n * (gj - a)] += blocks[boff[e] + m + o * mp];
            }
    }
}

/* ------------------------------------------------------------------ */
/* a17. Operator application  Y = alpha*A*X + beta*Y                     */
/* ------------------------------------------------------------------ */
/* The arithmetic of mul! lives in BandedMatrices / BlockBandedMatrices /
 * LinearAlgebra (not vendored).  y = A x is unambiguous up to summation
 * order; these loops follow the column-by-column, row-ascending pattern of
 * the only in-tree analogue, src/skewtridiag.jl:106-139. */

static inline void axpby_store(double *y, double v, double alpha, double beta)
{
    *y = (beta == 0.0) ? alpha * v : alpha * v + beta * (*y);
}

/* BandedMatrix (l,u), data (l+u+1) x n column-major, m rows. */
void cbo_banded_apply(const double *band, i64 m, i64 n, int l, int u,
                      const double *X, i64 ldx, i64 nvec,
                      double alpha, double beta, double *Y, i64 ldy)
{
    int ld = l + u + 1;
    for (i64 v = 0; v < nvec; ++v) {
        const double *xv = X + v * ldx; double *yv = Y + v * ldy;
        for (i64 i = 0; i < m; ++i) {
            i64 jlo = i - l < 0 ? 0 : i - l, jhi = i + u > n - 1 ? n - 1 : i + u;
            double s = 0.0;
            for (i64 j = jlo; j <= jhi; ++j) s += band[(u + i - j) + (i64)ld * j] * xv[j];
            axpby_store(&yv[i], s, alpha, beta);
        }
    }
}

/* Tridiagonal(dl, d, du) (n-1, n, n-1); SymTridiagonal(dv, ev) = dl=du=ev. */
void cbo_tridiag_apply(const double *dl, const double *d, const double *du, i64 n,
                       const double *X, i64 ldx, i64 nvec,
                       double alpha, double beta, double *Y, i64 ldy)
{
    for (i64 v = 0; v < nvec; ++v) {
        const double *xv = X + v * ldx; double *yv = Y + v * ldy;
        for (i64 i = 0; i < n; ++i) {
            double s = d[i] * xv[i];
            if (i > 0) s = dl[i - 1] * xv[i - 1] + s;
            if (i < n - 1) s += du[i] * xv[i + 1];
            axpby_store(&yv[i], s, alpha, beta);
        }
    }
}

/* FE-DVR operator given as element blocks (bridge corners NOT pre-summed):
 * y[rows of e] += block_e * x[rows of e]. */
void cbo_fedvr_apply(const double *blocks, const i64 *boff, const i64 *estart,
                     const int32_t *order, i64 nel, i64 n,
                     const double *X, i64 ldx, i64 nvec,
                     double alpha, double beta, double *Y, i64 ldy)
{
    double *tmp = (double *)malloc(sizeof(double) * (size_t)n);
    for (i64 v = 0; v < nvec; ++v) {
        const double *xv = X + v * ldx; double *yv = Y + v * ldy;
        memset(tmp, 0, sizeof(double) * (size_t)n);
        for (i64 e = 0; e < nel; ++e) {
            int o = order[e];
            const double *b = blocks + boff[e];
            const double *xe = xv + estart[e];
            double *te = tmp + estart[e];
            for (int m = 0; m < o; ++m) {
                double s = 0.0;
                for (int mp = 0; mp < o; ++mp) s += b[m + o * mp] * xe[mp];
                te[m] += s;
            }
        }
        for (i64 i = 0; i < n; ++i) axpby_store(&yv[i], tmp[i], alpha, beta);
    }
    free(tmp);
}

/* ------------------------------------------------------------------ */
/* a18 at scale: pinv(V) * b by a banded QR                             */
/* ------------------------------------------------------------------ */

/* The reference forms C = pinv(Matrix(RV)) * w as a DENSE nf x nquad matrix (src/densities.jl:96) and applies
 * it with a dense product (:163).  V = B[locs, sel] has full column rank, so pinv(V) b is the least-squares
 * solution of V c = b, which this routine computes at linear cost by a row-sequential Givens QR of V (each node
 * row holds the k live functions of its interval, so R stays upper banded with k diagonals) — the algorithm
 * class of Julia's sparse `\` (SPQR), not the normal equations the CUDA path uses: an independent check.
 *   table[(c*q + l)*k + a]: basis_functions table (cbo_basis_table, m = 0)
 *   iv[c]: 1-based expanded index of the c-th non-empty interval
 *   columns col0+1 .. col0+ncols of the basis (restriction)
 *   B: nquad x P column-major right-hand sides (ldb), overwritten;  out: ncols x P (ldo). */
void cbo_banded_lstsq(const double *table, const i64 *iv, i64 ni, int q, int k,
                      i64 col0, i64 ncols, const double *Bm, i64 ldb, i64 P, double *out, i64 ldo)
{
    double *R = (double *)calloc((size_t)ncols * k, sizeof(double));      /* R[j*k + d] = R(j, j+d) */
    double *y = (double *)calloc((size_t)ncols * P, sizeof(double));      /* (Q'b)[j], row-major [j][p] */
    double *row = (double *)malloc(sizeof(double) * (size_t)(k + 1));
    double *br = (double *)malloc(sizeof(double) * (size_t)(P ? P : 1));
    for (i64 c = 0; c < ni; ++c) {
        for (int l = 0; l < q; ++l) {
            const i64 e = c * q + l;
            /* the row's k entries at absolute (restricted, 0-based) columns j0 .. j0+k-1 */
            const i64 j0 = iv[c] - k + 1 - 1 - col0;
            for (int a = 0; a < k; ++a) {
                const i64 j = j0 + a;
                row[a] = (j >= 0 && j < ncols) ? table[e * k + a] : 0.0;
            }
            for (i64 p = 0; p < P; ++p) br[p] = Bm[e + ldb * p];
            for (int a = 0; a < k; ++a) {
                const i64 j = j0 + a;
                if (j < 0 || j >= ncols || row[a] == 0.0) continue;
                double *Rj = R + j * k;
                const double r0 = Rj[0], v = row[a];
                const double h = hypot(r0, v);
                const double cs = r0 / h, sn = v / h;
                /* rotate (R row j, working row) over columns j .. j+k-1; the working row's entries live at a.. */
                Rj[0] = h;
                for (int d = 1; d < k; ++d) {
                    const int aa = a + d;
                    const double rv = Rj[d], wv = aa < k ? row[aa] : 0.0;
                    Rj[d] = cs * rv + sn * wv;
                    if (aa < k) row[aa] = -sn * rv + cs * wv;
                }
                double *yj = y + j * P;
                for (i64 p = 0; p < P; ++p) {
                    const double yv = yj[p], bv = br[p];
                    yj[p] = cs * yv + sn * bv;
                    br[p] = -sn * yv + cs * bv;
                }
                row[a] = 0.0;
            }
        }
    }
    /* back substitution R c = Q'b */
    for (i64 j = ncols - 1; j >= 0; --j) {
        const double *Rj = R + j * k;
        for (i64 p = 0; p < P; ++p) {
            double sacc = y[j * P + p];
            for (int d = 1; d < k && j + d < ncols; ++d) sacc -= Rj[d] * out[(j + d) + ldo * p];
            out[j + ldo * p] = sacc / Rj[0];
        }
    }
    free(R); free(y); free(row); free(br);
}

/* The same over chunks of right-hand sides on several threads (every thread repeats the QR of V for its chunk):
 * the "OpenMP over all host cores" form of the CPU baseline. */
void cbo_banded_lstsq_mt(const double *table, const i64 *iv, i64 ni, int q, int k,
                         i64 col0, i64 ncols, const double *Bm, i64 ldb, i64 P, double *out, i64 ldo)
{
    int nt = g_threads;
    if (nt > P) nt = (int)(P > 0 ? P : 1);
#pragma omp parallel for schedule(static) num_threads(nt)
    for (int c = 0; c < nt; ++c) {
        i64 p0 = P * c / nt, p1 = P * (c + 1) / nt;
        if (p1 > p0) cbo_banded_lstsq(table, iv, ni, q, k, col0, ncols, Bm + ldb * p0, ldb, p1 - p0, out + ldo * p0, ldo);
    }
}

/* b = w .* conj(V f) .* (V g) at the nodes (src/densities.jl:158-162), V restricted to columns col0+1..col0+ncols;
 * f: ncols x Pf, g: ncols x Pg with Pf == Pg or one of them 1 (:57-72).  Bm: nquad x max(Pf,Pg). */
void cbo_density_nodal(const double *table, const i64 *iv, i64 ni, int q, int k, i64 col0, i64 ncols,
                       const double *wvals, const double *f, i64 ldf, i64 Pf, const double *g, i64 ldg, i64 Pg,
                       double *Bm, i64 ldb)
{
    const i64 P = Pf > Pg ? Pf : Pg;
#pragma omp parallel for schedule(static) num_threads(g_threads)
    for (i64 p = 0; p < P; ++p) {
        const double *fc = f + ldf * (Pf == 1 ? 0 : p), *gc = g + ldg * (Pg == 1 ? 0 : p);
        for (i64 c = 0; c < ni; ++c)
            for (int l = 0; l < q; ++l) {
                const i64 e = c * q + l, j0 = iv[c] - k + 1 - 1 - col0;
                double uf = 0.0, ug = 0.0;
                for (int a = 0; a < k; ++a) {
                    const i64 j = j0 + a;
                    if (j < 0 || j >= ncols) continue;
                    uf += table[e * k + a] * fc[j];
                    ug += table[e * k + a] * gc[j];
                }
                double t = uf * ug;
                if (wvals) t = wvals[e] * t;
                Bm[e + ldb * p] = t;
            }
    }
}

/* Multi-threaded variants for the CPU baseline (split the vector batch). */
#ifdef _OPENMP
#include <omp.h>
int cbo_max_threads(void) { return omp_get_max_threads(); }
#else
int cbo_max_threads(void) { return 1; }
#endif

void cbo_fedvr_apply_mt(const double *blocks, const i64 *boff, const i64 *estart,
                        const int32_t *order, i64 nel, i64 n,
                        const double *X, i64 ldx, i64 nvec,
                        double alpha, double beta, double *Y, i64 ldy)
{
#pragma omp parallel for schedule(static)
    for (i64 v = 0; v < nvec; ++v)
        cbo_fedvr_apply(blocks, boff, estart, order, nel, n, X + v * ldx, ldx, 1,
                        alpha, beta, Y + v * ldy, ldy);
}

void cbo_tridiag_apply_mt(const double *dl, const double *d, const double *du, i64 n,
                          const double *X, i64 ldx, i64 nvec,
                          double alpha, double beta, double *Y, i64 ldy)
{
#pragma omp parallel for schedule(static)
    for (i64 v = 0; v < nvec; ++v)
        cbo_tridiag_apply(dl, d, du, n, X + v * ldx, ldx, 1, alpha, beta, Y + v * ldy, ldy);
}

void cbo_banded_apply_mt(const double *band, i64 m, i64 n, int l, int u,
                         const double *X, i64 ldx, i64 nvec,
                         double alpha, double beta, double *Y, i64 ldy)
{
#pragma omp parallel for schedule(static)
    for (i64 v = 0; v < nvec; ++v)
        cbo_banded_apply(band, m, n, l, u, X + v * ldx, ldx, 1, alpha, beta, Y + v * ldy, ldy);
}

void cbo_bspline_eval_ell_mt(const double *t, i64 nt, int k, int mr,
                             const double *x, i64 nx, int m,
                             int32_t *interval, double *vals)
{
    i64 chunk = 4096;
    i64 nchunk = (nx + chunk - 1) / chunk;
#pragma omp parallel for schedule(dynamic)
    for (i64 c = 0; c < nchunk; ++c) {
        i64 p0 = c * chunk, p1 = p0 + chunk > nx ? nx : p0 + chunk;
        cbo_bspline_eval_ell(t, nt, k, mr, x + p0, p1 - p0, m, interval + p0, vals + p0 * k);
    }
}
