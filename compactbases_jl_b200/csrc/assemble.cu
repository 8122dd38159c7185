// assemble.cu — operator assembly:
//   K2+K3  B-spline banded operators  <(-1)^a d^a L_i | op | d^b R_j>   (fused:
//          the per-interval k x q tables live only in shared memory)
//   K4     FE-DVR element blocks  diff(B, n, i)
//   K5     finite-difference stencils (Cartesian / staggered uniform / staggered
//          non-uniform, Coulomb correction of the first diagonal entry)
#include "bspline_core.cuh"

// ---------------------------------------------------------------------------
// Runtime-order twin of cb_bspline_all (used only when the two bases have
// different orders, e.g. test/bsplines/derivatives.jl:36-79).  tw has 2k
// entries, N has k entries; arrays are sized for CB_KMAX.
// ---------------------------------------------------------------------------
__device__ void bspline_all_rt(int k, const double *tw, double x, int m, double *N) {
    for (int r = 0; r < k; ++r) N[r] = 0.0;
    if (m >= k) return;
    N[0] = 1.0;
    const int jmax = k - 1 - m;
    for (int j = 1; j <= jmax; ++j) {
        double saved = 0.0;
        for (int r = 0; r < j; ++r) {
            double d = tw[k + r] - tw[k - j + r];
            double term = N[r] * fast_rcp(d);
            N[r] = fma(tw[k + r] - x, term, saved);
            saved = (x - tw[k - j + r]) * term;
        }
        N[j] = saved;
    }
    for (int o = k - m; o <= k - 1; ++o) {
        if (o < 1) continue;
        double qprev = 0.0;
        for (int r = 0; r <= o; ++r) {
            double qr = 0.0;
            if (r < o) {
                double d = tw[k + r] - tw[k - o + r];
                qr = N[r] * fast_rcp(d);
            }
            N[r] = (double)o * (qprev - qr);
            qprev = qr;
        }
    }
}

// ---------------------------------------------------------------------------
// K2+K3: fused B-spline band assembly.
//
// A CTA owns TJ consecutive columns (functions of R).  Phase 1 evaluates, for
// every non-empty knot interval s that one of those columns lives on and every
// quadrature node l of it, the kL live functions of L (a-th derivative, times
// (-1)^a w_l) and the kR live functions of R (b-th derivative, times op(x_l))
// into shared memory — basis_functions(t, x, m), src/bsplines.jl:28-52, which
// the reference stores as a sparse matrix in memory, never leaves the SM.
// Phase 2 is the gather form of overlap_matrix! (src/bsplines.jl:58-73): one
// thread per band entry sums over the common intervals and their nodes in
// increasing node order, the order of the reference's sparse dot product, so no
// atomics and a deterministic result.
//
// s is the 0-based index of an interval of t.t (shared by L and R, src/
// bsplines.jl:75-82); its 1-based index in an expanded knot vector is s + ml.
// ---------------------------------------------------------------------------
struct AsmArgs {
    KnotView kvL, kvR;
    int kL, kR;
    i64 nfL, nfR;
    int q, da, db;
    const double *x, *w;          // nodes / weights, [ni*q] (of R's parent)
    const int32_t *ordR;          // R's d_ord: ordinal of expanded interval I (entry I-1) or -1
    const double *opvals;         // NULL or op at the nodes
    int coulomb; double Z;        // 1: op(x) = -Z/x evaluated on chip; 2: unit quadrature weights (Gram matrix V'V)
    i64 rowL0, mm, colR0, nn;     // restriction offsets and clipped sizes
    int l, u;
    i64 col_lo, col_hi;           // columns written (0-based, within the restriction)
    i64 s_lo, s_hi;               // t.t intervals summed
    int TJ;
    double *band;                 // column col_lo is at band[0]
};

constexpr int ASM_THREADS = 256;
constexpr int ASM_SMAX = 32 + CB_KMAX;             // most intervals a column tile can touch

// idx = j*(j-1)/2 + r  ->  (j, r) for the reciprocal table (j = 1..CB_KMAX-1, r < j)
struct RcIndex { unsigned char j[CB_KMAX * (CB_KMAX - 1) / 2], r[CB_KMAX * (CB_KMAX - 1) / 2]; };
static RcIndex make_rc_index() {
    RcIndex t; int idx = 0;
    for (int j = 1; j < CB_KMAX; ++j) for (int r = 0; r < j; ++r) { t.j[idx] = (unsigned char)j; t.r[idx] = (unsigned char)r; ++idx; }
    return t;
}
__constant__ RcIndex c_rc_index;

// padded node count: even (16-byte pairs) and = 2 mod 4 so that rows of different functions fall
// into different banks for the 16-byte reads of phase 2
__host__ __device__ inline int asm_qpad(int q) { int p = q + (q & 1); return (p & 3) == 0 ? p + 2 : p; }

template <int KT>
__global__ void __launch_bounds__(ASM_THREADS)
bspline_assemble_kernel(AsmArgs A) {
    extern __shared__ __align__(16) double sm[];
    const int kL = KT > 0 ? KT : A.kL, kR = KT > 0 ? KT : A.kR;
    const int q = A.q, W = A.l + A.u + 1, TJ = A.TJ;
    const int QP = asm_qpad(q);
    const int mlL = A.kvL.ml, mlR = A.kvR.ml;
    const i64 n_tt = A.kvR.n_tt;
    const int SMAX = TJ + kR - 1;                  // intervals a column tile can touch
    constexpr int NRC = KT > 0 ? KT * (KT - 1) / 2 : 0;            // reciprocal knot differences per interval
    double *TL = sm;                               // [SMAX][kL][QP]  (node index fastest, zero padded)
    double *TR = TL + (size_t)SMAX * kL * QP;      // [SMAX][kR][QP]
    double *RC = TR + (size_t)SMAX * kR * QP;      // [SMAX][NRC]    (same-order fast path only)
    const double sgn = (A.da & 1) ? -1.0 : 1.0;    // src/bspline_derivatives.jl:14
    const i64 ntiles = (A.col_hi - A.col_lo + TJ - 1) / TJ;
    const bool same_tab = A.da == A.db && mlL == mlR;
    __shared__ int sOrd[ASM_SMAX];                  // ordinal of each interval of the tile among the non-empty ones, -1 if empty

    for (i64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const i64 j0 = A.col_lo + tile * TJ;
        const i64 j1 = j0 + TJ < A.col_hi ? j0 + TJ : A.col_hi;      // exclusive
        // intervals of the tile's columns: function fj (1-based) lives on s = fj-mlR .. fj-mlR+kR-1
        i64 sA = (A.colR0 + 1 + j0) - mlR;
        i64 sB = (A.colR0 + j1) - mlR + kR - 1;                       // inclusive
        if (sA < 0) sA = 0;
        if (sB > n_tt - 2) sB = n_tt - 2;
        if (sA < A.s_lo) sA = A.s_lo;
        if (sB > A.s_hi - 1) sB = A.s_hi - 1;
        const int ns = sB >= sA ? (int)(sB - sA + 1) : 0;
        __syncthreads();                            // previous tile consumed
        for (int e = threadIdx.x; e < ns; e += ASM_THREADS) sOrd[e] = A.ordR[sA + e + mlR - 1];
        __syncthreads();
        // ---- phase 0: reciprocal knot differences, once per interval -----------
        if (KT > 0) {
            constexpr int K = KT > 0 ? KT : 1;
            for (int e = threadIdx.x; e < ns * NRC; e += ASM_THREADS) {
                const int sl = e / (NRC > 0 ? NRC : 1), idx = e - sl * NRC;
                const i64 s = sA + sl;
                if (sOrd[sl] < 0) continue;
                const int j = c_rc_index.j[idx], r = c_rc_index.r[idx];
                // tw[K + r] - tw[K - j + r] with tw[t] = knot (I - K + t), I = s + mlR (0-based expanded index I-K+t)
                const i64 I = s + mlR;
                const double hi = A.kvR.at(I - K + (K + r)), lo = A.kvR.at(I - K + (K - j + r));
                RC[e] = fast_rcp(hi - lo);
            }
            // zero the node padding of both tables (never written in phase 1)
            for (int e = threadIdx.x; e < ns * (kL + kR) * (QP - q); e += ASM_THREADS) {
                const int row = e / (QP - q), c = e - row * (QP - q);
                (row < ns * kL ? TL + (size_t)row * QP : TR + (size_t)(row - ns * kL) * QP)[q + c] = 0.0;
            }
            __syncthreads();
        } else {
            for (int e = threadIdx.x; e < ns * (kL + kR) * (QP - q); e += ASM_THREADS) {
                const int row = e / (QP - q), c = e - row * (QP - q);
                (row < ns * kL ? TL + (size_t)row * QP : TR + (size_t)(row - ns * kL) * QP)[q + c] = 0.0;
            }
        }
        // ---- phase 1: tables ----------------------------------------------
        for (int e = threadIdx.x; e < ns * q; e += ASM_THREADS) {
            const int sl = e / q, ll = e - sl * q;
            const i64 s = sA + sl;
            const int c = sOrd[sl];
            double *tl = TL + (size_t)sl * kL * QP + ll, *tr = TR + (size_t)sl * kR * QP + ll;
            if (c < 0) continue;                    // empty interval: never read in phase 2
            const double xv = A.x[(i64)c * q + ll];
            const double wv = A.coulomb == 2 ? 1.0 : A.w[(i64)c * q + ll];
            double ov = 1.0;
            if (A.opvals) ov = A.opvals[(i64)c * q + ll];
            else if (A.coulomb == 1) ov = -A.Z / xv;
            const int IL = (int)(s + mlL), IR = (int)(s + mlR);
            if (KT > 0) {
                constexpr int K = KT > 0 ? KT : 1;
                const double *rc = RC + (size_t)sl * NRC;
                double tw[2 * K], N[K];
                cb_knot_window<K>(A.kvR.tt, n_tt, mlR, IR, tw);
                cb_bspline_all_rc<K>(tw, rc, xv, A.db, N);
#pragma unroll
                for (int a = 0; a < K; ++a) {
                    i64 f = (i64)IR - K + 1 + a;
                    tr[a * QP] = (f >= 1 && f <= A.nfR) ? ov * N[a] : 0.0;
                }
                if (!same_tab) {
                    // the knot window of L is the same set of knots when mlL == mlR; otherwise the
                    // windows differ only by clamping at the ends, where the reciprocals are recomputed
                    if (mlL == mlR) {
                        cb_bspline_all_rc<K>(tw, rc, xv, A.da, N);
                    } else {
                        cb_knot_window<K>(A.kvL.tt, n_tt, mlL, IL, tw);
                        cb_bspline_all<K>(tw, xv, A.da, N);
                    }
                }
#pragma unroll
                for (int a = 0; a < K; ++a) {
                    i64 f = (i64)IL - K + 1 + a;
                    tl[a * QP] = (f >= 1 && f <= A.nfL) ? (sgn * N[a]) * wv : 0.0;
                }
            } else {
                double tw[2 * CB_KMAX], N[CB_KMAX];
                for (int t = 0; t < 2 * kR; ++t) tw[t] = A.kvR.at((i64)IR - kR + t);
                bspline_all_rt(kR, tw, xv, A.db, N);
                for (int a = 0; a < kR; ++a) {
                    i64 f = (i64)IR - kR + 1 + a;
                    tr[a * QP] = (f >= 1 && f <= A.nfR) ? ov * N[a] : 0.0;
                }
                for (int t = 0; t < 2 * kL; ++t) tw[t] = A.kvL.at((i64)IL - kL + t);
                bspline_all_rt(kL, tw, xv, A.da, N);
                for (int a = 0; a < kL; ++a) {
                    i64 f = (i64)IL - kL + 1 + a;
                    tl[a * QP] = (f >= 1 && f <= A.nfL) ? (sgn * N[a]) * wv : 0.0;
                }
            }
        }
        __syncthreads();
        // ---- phase 2: gather into the band ----------------------------------
        // one thread per band entry; the sum runs over the common intervals and, inside an interval,
        // over the nodes in increasing order (two nodes per 16-byte shared-memory read)
        const int nent = (int)(j1 - j0) * W;
        for (int e = threadIdx.x; e < nent; e += ASM_THREADS) {
            const int jj = e / W, d = e - jj * W;
            const i64 j = j0 + jj, i = j - A.u + d;
            double acc = 0.0;
            if (i >= 0 && i < A.mm && j < A.nn) {
                const i64 fi = A.rowL0 + 1 + i, fj = A.colR0 + 1 + j;
                i64 lo = fi - mlL > fj - mlR ? fi - mlL : fj - mlR;
                i64 hi = fi - mlL + kL - 1 < fj - mlR + kR - 1 ? fi - mlL + kL - 1 : fj - mlR + kR - 1;
                if (lo < sA) lo = sA;
                if (hi > sB) hi = sB;
                const int sl0 = (int)(lo - sA), sl1 = (int)(hi - sA);
                // table slot of the functions in interval slot sl: a = (f - (sA + sl) - ml) + k - 1
                const int aL0 = (int)(fi - sA - mlL) + kL - 1, aR0 = (int)(fj - sA - mlR) + kR - 1;
                for (int sl = sl0; sl <= sl1; ++sl) {
                    if (sOrd[sl] < 0) continue;
                    const double2 *tl = reinterpret_cast<const double2 *>(TL + ((size_t)sl * kL + (aL0 - sl)) * QP);
                    const double2 *tr = reinterpret_cast<const double2 *>(TR + ((size_t)sl * kR + (aR0 - sl)) * QP);
#pragma unroll 4
                    for (int l2 = 0; l2 < QP / 2; ++l2) {
                        const double2 a2 = tl[l2], b2 = tr[l2];
                        acc = fma(a2.x, b2.x, acc);
                        acc = fma(a2.y, b2.y, acc);
                    }
                }
            }
            A.band[(i64)d + (i64)W * (j - A.col_lo)] = acc;
        }
    }
}

// BandedMatrix (l = u = 1) -> Tridiagonal(dl, d, du): the k == 2 && m == n case of
// Matrix(undef, A, B) (src/bsplines.jl:258-262).
__global__ void band_to_tridiag_kernel(const double *__restrict__ band, i64 n, double *__restrict__ dl,
                                       double *__restrict__ d, double *__restrict__ du) {
    for (i64 j = blockIdx.x * (i64)blockDim.x + threadIdx.x; j < n; j += (i64)gridDim.x * blockDim.x) {
        d[j] = band[1 + 3 * j];
        if (j > 0) du[j - 1] = band[0 + 3 * j];        // A(j-1, j)
        if (j < n - 1) dl[j] = band[2 + 3 * j];        // A(j+1, j)
    }
}

// ---------------------------------------------------------------------------
// K4: FE-DVR element blocks.  One warp per element:
//   L'[m,m'] = lagrangeder!  (src/fedvr_derivatives.jl:15-31)
//   D[m,m']  = n_m n_m' sum_q Lt[m,q] w_q L'[m',q], Lt = I | -L'   (:33-57)
// IEEE divisions (this kernel is latency-bound, 10 MB of output at C2).
// ---------------------------------------------------------------------------
constexpr int FEA_WARPS = 4;

__global__ void __launch_bounds__(FEA_WARPS * 32)
fedvr_assemble_kernel(const double *__restrict__ x, const double *__restrict__ wcat, const double *__restrict__ nrm,
                      const i64 *__restrict__ estart, const int32_t *__restrict__ order,
                      const i64 *__restrict__ boff, const i64 *__restrict__ woff, int uniform_order,
                      i64 nel, int nder, int omax, double *__restrict__ blocks) {
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *Lp = sm + (size_t)warp * (omax * omax + 3 * omax);   // [o][o] column-major L'[m + o*m']
    double *xe = Lp + omax * omax, *we = xe + omax, *ne = we + omax;
    for (i64 e = (i64)blockIdx.x * FEA_WARPS + warp; e < nel; e += (i64)gridDim.x * FEA_WARPS) {
        const int o = uniform_order > 0 ? uniform_order : order[e];
        const i64 es = uniform_order > 0 ? e * (o - 1) : estart[e];
        const i64 ws = uniform_order > 0 ? e * o : woff[e];
        const i64 bo = uniform_order > 0 ? e * (i64)o * o : boff[e];
        __syncwarp();
        for (int t = lane; t < o; t += 32) { xe[t] = x[es + t]; we[t] = wcat[ws + t]; ne[t] = nrm[es + t]; }
        __syncwarp();
        for (int t = lane; t < o * o; t += 32) {
            const int m = t % o, mp = t / o;
            double v;
            if (m == mp) {
                v = (double)((m == o - 1) - (m == 0)) / (2 * we[m]);
            } else {
                double f = 1.0;
                for (int kk = 0; kk < o; ++kk) {
                    if (kk == m || kk == mp) continue;
                    f *= (xe[mp] - xe[kk]) / (xe[m] - xe[kk]);
                }
                v = f / (xe[m] - xe[mp]);
            }
            Lp[t] = v;
        }
        __syncwarp();
        for (int t = lane; t < o * o; t += 32) {
            const int m = t % o, mp = t / o;
            double s;
            if (nder == 1) {
                s = we[m] * Lp[mp + o * m];
            } else {
                s = 0.0;
                for (int qd = 0; qd < o; ++qd) s += (-Lp[m + o * qd]) * (we[qd] * Lp[mp + o * qd]);
            }
            blocks[bo + t] = ne[m] * ne[mp] * s;
        }
    }
}

// ---------------------------------------------------------------------------
// K5: finite-difference stencils, one thread per node.  Round-to-nearest
// intrinsics keep the reference's operation order (no contraction), so the
// diagonals are bit-identical to the CPU restatement.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double sq_(double a) { return __dmul_rn(a, a); }

// r~_j for j = 0..N+1 (1-based nodes; r~_0 = 2r_1 - r_2, r~_{N+1} = 2r_N - r_{N-1})
__device__ __forceinline__ double rext(const double *__restrict__ r, i64 N, i64 j) {
    if (j < 1) return __dsub_rn(__dmul_rn(2.0, r[0]), r[1]);
    if (j > N) return __dsub_rn(__dmul_rn(2.0, r[N - 1]), r[N - 2]);
    return r[j - 1];
}

__global__ void fd_stencil_kernel(int scheme, int nder, const double *__restrict__ r, const double *__restrict__ fhalf,
                                  i64 N, double step, i64 i0, i64 n, double Z, int ell,
                                  double *__restrict__ dl, double *__restrict__ d, double *__restrict__ du) {
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const i64 j = i0 + i + 1;                   // 1-based node of the parent grid
        double al = 1.0, be = 1.0, ga = 0.0, de = 1.0;   // Cartesian, src/fd_derivatives.jl:12-15
        if (scheme == 1) {                          // :20-25
            const double j2 = (double)(j * j);
            al = __ddiv_rn(j2, __dsub_rn(j2, 0.25));
            de = al;
            const double t = (double)(j * j - j);
            be = __ddiv_rn(__dadd_rn(t, 0.5), __dadd_rn(t, 0.25));
        } else if (scheme == 2) {                   // :27-92
            const double a = rext(r, N, j - 1), b = rext(r, N, j), c = rext(r, N, j + 1);
            const double dd = rext(r, N, j + 2 > N + 1 ? N + 1 : j + 2);
            const double rp = __ddiv_rn(__dadd_rn(b, c), 2.0);
            if (j < N) {
                const double root = __dsqrt_rn(__dmul_rn(__dsub_rn(dd, b), __dsub_rn(c, a)));
                // delta: 2/sqrt(..)*((b+c)/2)^2/(b*c)
                de = __ddiv_rn(__dmul_rn(__ddiv_rn(2.0, root), sq_(rp)), __dmul_rn(b, c));
                if (fhalf) al = __ddiv_rn(__dmul_rn(__ddiv_rn(__dmul_rn(2.0, fhalf[j]), __dmul_rn(__dsub_rn(c, b), root)), sq_(rp)), __dmul_rn(b, c));
                else al = __ddiv_rn(__dmul_rn(__ddiv_rn(2.0, __dmul_rn(__dsub_rn(c, b), root)), sq_(rp)), __dmul_rn(b, c));
            }
            if (fhalf) {
                const double bm2 = __ddiv_rn(1.0, sq_(b));
                const double rm = __ddiv_rn(__dadd_rn(a, b), 2.0);
                const double fp = __dmul_rn(sq_(rp), bm2), fm = __dmul_rn(sq_(rm), bm2);
                const double t1 = __dmul_rn(__ddiv_rn(fhalf[j], __dsub_rn(c, b)), fp);
                const double t2 = __dmul_rn(__ddiv_rn(fhalf[j - 1], __dsub_rn(b, a)), fm);
                be = __dmul_rn(__ddiv_rn(1.0, __dsub_rn(c, a)), __dadd_rn(t1, t2));
            } else {
                const double tb = __dmul_rn(2.0, b);
                const double fp = sq_(__ddiv_rn(__dadd_rn(c, b), tb));
                const double fm = sq_(__ddiv_rn(__dadd_rn(b, a), tb));
                const double t1 = __dmul_rn(__ddiv_rn(1.0, __dsub_rn(c, b)), fp);
                const double t2 = __dmul_rn(__ddiv_rn(1.0, __dsub_rn(b, a)), fm);
                be = __dmul_rn(__ddiv_rn(1.0, __dsub_rn(c, a)), __dadd_rn(t1, t2));
            }
        }
        if (nder == 1) {                            // :233-237
            const double mu = __dmul_rn(2.0, step);
            d[i] = __ddiv_rn(ga, mu);
            if (i < n - 1) {
                if (du) du[i] = __ddiv_rn(de, mu);
                if (dl) dl[i] = __ddiv_rn(-de, mu);
            }
        } else {                                    // :265-277
            const double h2 = sq_(step);
            double dv = __ddiv_rn(__dmul_rn(-2.0, be), h2);
            if (j == 1 && scheme != 0 && ell == 0 && Z != 0.0) {   // laplacian_correction!, :542-551
                const double rho = scheme == 1 ? step : __dsub_rn(r[1], r[0]);
                // 1/rho^2*Z*rho/8*(1 + Z*rho), left to right
                double db = __ddiv_rn(1.0, sq_(rho));
                db = __dmul_rn(db, Z);
                db = __dmul_rn(db, rho);
                db = __ddiv_rn(db, 8.0);
                db = __dmul_rn(db, __dadd_rn(1.0, __dmul_rn(Z, rho)));
                dv = __dsub_rn(dv, __dmul_rn(2.0, db));
            }
            d[i] = dv;
            if (i < n - 1) {
                const double ev = __ddiv_rn(al, h2);
                if (du) du[i] = ev;
                if (dl) dl[i] = ev;
            }
        }
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static int ensure_rc_index() {
    static bool done = false;                       // per process; every device of the process gets its copy below
    static int devs_done[64] = {0};
    int dev = 0;
    CB_CUDA(cudaGetDevice(&dev));
    (void)done;
    if (dev >= 0 && dev < 64 && !devs_done[dev]) {
        RcIndex t = make_rc_index();
        CB_CUDA(cudaMemcpyToSymbol(c_rc_index, &t, sizeof(t)));
        devs_done[dev] = 1;
    }
    return CB_OK;
}

static int assemble_common(const cb_bspline *L, const cb_bspline *R, int a, int b, const double *opvals,
                           int coulomb, double Z, i64 rowL0, i64 m, i64 colR0, i64 n, int l, int u,
                           i64 col_lo, i64 col_hi, i64 s_lo, i64 s_hi, double *band, void *stream) {
    CB_REQUIRE(L != nullptr && R != nullptr && band != nullptr, "cb_bspline_assemble: NULL argument");
    // assert_compatible_bases, src/bsplines.jl:75-82
    bool same = L->n_tt == R->n_tt && L->q == R->q && L->h_tt == R->h_tt && L->h_xi == R->h_xi && L->h_wr == R->h_wr;
    CB_REQUIRE(same, "Can only multiply B-spline bases with identical knot sets, resolved on the same Gauss-Legendre points");
    CB_REQUIRE(L->device == R->device, "cb_bspline_assemble: bases live on different devices");
    CB_REQUIRE(a >= 0 && b >= 0, "cb_bspline_assemble: negative derivative order");
    CB_REQUIRE(rowL0 >= 0 && colR0 >= 0 && m >= 0 && n >= 0 && rowL0 + m <= L->nf && colR0 + n <= R->nf,
               "cb_bspline_assemble: restriction outside 1..nf");
    const int k = L->k > R->k ? L->k : R->k;
    // Matrix(undef, A, B) (src/bsplines.jl:264-269) and the check of diffop! (src/bspline_derivatives.jl:9-10)
    const i64 ij = colR0 - rowL0;
    CB_REQUIRE((i64)l == k - 1 + ij && (i64)u == k - 1 - ij,
               "DimensionMismatch: bandwidths (%d,%d) do not match Matrix(undef, A, B) = (%lld,%lld)", l, u,
               (long long)(k - 1 + ij), (long long)(k - 1 - ij));
    CB_REQUIRE(l + u + 1 >= 1, "cb_bspline_assemble: empty band");
    CB_REQUIRE(col_lo >= 0 && col_hi <= n && col_lo <= col_hi, "cb_bspline_assemble: column range outside 0..n");
    CB_REQUIRE(s_lo >= 0 && s_hi <= R->n_tt - 1 && s_lo <= s_hi, "cb_bspline_assemble: interval range outside 0..numintervals");
    if (col_hi == col_lo) return CB_OK;
    { int rc_ = ensure_rc_index(); if (rc_) return rc_; }
    cudaStream_t st = (cudaStream_t)stream;
    AsmArgs A;
    A.kvL = make_knot_view(L); A.kvR = make_knot_view(R);
    A.kL = L->k; A.kR = R->k; A.nfL = L->nf; A.nfR = R->nf;
    A.q = R->q; A.da = a; A.db = b;
    A.x = R->d_x; A.w = R->d_w; A.ordR = R->d_ord;
    A.opvals = opvals; A.coulomb = coulomb; A.Z = Z;
    A.rowL0 = rowL0; A.mm = m; A.colR0 = colR0; A.nn = n;
    A.l = l; A.u = u; A.col_lo = col_lo; A.col_hi = col_hi; A.s_lo = s_lo; A.s_hi = s_hi;
    A.band = band;
    // column tile: as many columns as fit ~96 KB (two CTAs per SM), at most 32
    const int NRC = (L->k == R->k) ? L->k * (L->k - 1) / 2 : 0;
    const size_t per_iv = sizeof(double) * ((size_t)asm_qpad(A.q) * (A.kL + A.kR) + NRC);
    int TJ = 32;
    while (TJ > 1 && (size_t)(TJ + A.kR - 1) * per_iv > 112 * 1024) TJ /= 2;   // two CTAs per SM
    size_t smem = (size_t)(TJ + A.kR - 1) * per_iv;
    if (smem > 220 * 1024) { cb_set_error("cb_bspline_assemble: k=%d, q=%d tables exceed shared memory", k, A.q); return CB_ERR_UNSUPPORTED; }
    A.TJ = TJ;
    i64 ntiles = (col_hi - col_lo + TJ - 1) / TJ;
    int grid = (int)(ntiles < (i64)CB_NUM_SMS * 32 ? ntiles : (i64)CB_NUM_SMS * 32);
#define CB_ASM_LAUNCH(KT)                                                                         \
    do {                                                                                          \
        CB_CUDA(cudaFuncSetAttribute(bspline_assemble_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        bspline_assemble_kernel<KT><<<grid, ASM_THREADS, smem, st>>>(A);                          \
    } while (0)
    if (L->k != R->k) {
        CB_ASM_LAUNCH(0);
    } else {
        switch (L->k) {
            case 1: CB_ASM_LAUNCH(1); break;   case 2: CB_ASM_LAUNCH(2); break;
            case 3: CB_ASM_LAUNCH(3); break;   case 4: CB_ASM_LAUNCH(4); break;
            case 5: CB_ASM_LAUNCH(5); break;   case 6: CB_ASM_LAUNCH(6); break;
            case 7: CB_ASM_LAUNCH(7); break;   case 8: CB_ASM_LAUNCH(8); break;
            case 9: CB_ASM_LAUNCH(9); break;   case 10: CB_ASM_LAUNCH(10); break;
            case 11: CB_ASM_LAUNCH(11); break; case 12: CB_ASM_LAUNCH(12); break;
            case 13: CB_ASM_LAUNCH(13); break; case 14: CB_ASM_LAUNCH(14); break;
            case 15: CB_ASM_LAUNCH(15); break; case 16: CB_ASM_LAUNCH(16); break;
            default: CB_ASM_LAUNCH(0); break;
        }
    }
#undef CB_ASM_LAUNCH
    CB_LAUNCH_CHECK();
    return CB_OK;
}

// Gram matrix V'V of the restricted Vandermonde matrix at the nodes (density.cu).
int cb_bspline_gram_internal(const cb_bspline *B, i64 col0, i64 ncols, double *band, void *stream) {
    return assemble_common(B, B, 0, 0, nullptr, 2, 0.0, col0, ncols, col0, ncols, B->k - 1, B->k - 1, 0, ncols, 0,
                           B->n_tt - 1, band, stream);
}

extern "C" {

int cb_bspline_assemble(const cb_bspline *L, const cb_bspline *R, int a, int b, const double *opvals,
                        i64 rowL0, i64 m, i64 colR0, i64 n, int l, int u, double *band, void *stream) {
    if (!R) { cb_set_error("cb_bspline_assemble: NULL basis"); return CB_ERR_ARG; }
    return assemble_common(L, R, a, b, opvals, 0, 0.0, rowL0, m, colR0, n, l, u, 0, n, 0, R->n_tt - 1, band, stream);
}

int cb_bspline_assemble_coulomb(const cb_bspline *L, const cb_bspline *R, double Z,
                                i64 rowL0, i64 m, i64 colR0, i64 n, int l, int u, double *band, void *stream) {
    if (!R) { cb_set_error("cb_bspline_assemble_coulomb: NULL basis"); return CB_ERR_ARG; }
    return assemble_common(L, R, 0, 0, nullptr, 1, Z, rowL0, m, colR0, n, l, u, 0, n, 0, R->n_tt - 1, band, stream);
}

int cb_bspline_assemble_part(const cb_bspline *L, const cb_bspline *R, int a, int b, const double *opvals,
                             int coulomb, double Z, i64 rowL0, i64 m, i64 colR0, i64 n, int l, int u,
                             i64 col_lo, i64 col_hi, i64 s_lo, i64 s_hi, double *band_cols, void *stream) {
    return assemble_common(L, R, a, b, opvals, coulomb, Z, rowL0, m, colR0, n, l, u, col_lo, col_hi, s_lo, s_hi,
                           band_cols, stream);
}

int cb_band_to_tridiag(const double *band, i64 n, double *dl, double *d, double *du, void *stream) {
    CB_REQUIRE(n >= 0, "cb_band_to_tridiag: negative size");
    if (n == 0) return CB_OK;
    CB_REQUIRE(band && d && (n == 1 || (dl && du)), "cb_band_to_tridiag: NULL array");
    band_to_tridiag_kernel<<<cb_grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(band, n, dl, d, du);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int cb_fedvr_assemble(const double *x, const double *wcat, const double *nrm,
                      const i64 *estart, const int32_t *order, const i64 *boff, const i64 *woff,
                      i64 nel, int uniform_order, int max_order, int nder, double *blocks, void *stream) {
    CB_REQUIRE(x && wcat && nrm && blocks, "cb_fedvr_assemble: NULL array");
    CB_REQUIRE(nel >= 1, "cb_fedvr_assemble: no elements");
    CB_REQUIRE(nder == 1 || nder == 2, "cb_fedvr_assemble: derivative order must be 1 or 2");
    int omax = uniform_order > 0 ? uniform_order : max_order;
    CB_REQUIRE(omax >= 2, "cb_fedvr_assemble: element order must be >= 2");
    if (omax > CB_FEDVR_OMAX) { cb_set_error("cb_fedvr_assemble: element order %d > %d", omax, CB_FEDVR_OMAX); return CB_ERR_UNSUPPORTED; }
    if (uniform_order <= 0) CB_REQUIRE(estart && order && boff && woff, "cb_fedvr_assemble: per-element arrays required");
    size_t smem = sizeof(double) * FEA_WARPS * ((size_t)omax * omax + 3 * omax);
    int grid = cb_grid_for(nel, FEA_WARPS, 16);
    fedvr_assemble_kernel<<<grid, FEA_WARPS * 32, smem, (cudaStream_t)stream>>>(x, wcat, nrm, estart, order, boff, woff,
                                                                                uniform_order, nel, nder, omax, blocks);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int cb_fd_derivative(int scheme, int nder, const double *r, const double *fhalf,
                     i64 N, double step, i64 i0, i64 n, double Z, int ell,
                     double *dl, double *d, double *du, void *stream) {
    CB_REQUIRE(scheme >= 0 && scheme <= 2, "cb_fd_derivative: unknown scheme %d", scheme);
    CB_REQUIRE(nder == 1 || nder == 2, "cb_fd_derivative: derivative order must be 1 or 2");
    CB_REQUIRE(N >= 1 && i0 >= 0 && n >= 0 && i0 + n <= N, "cb_fd_derivative: restriction outside 1..N");
    CB_REQUIRE(scheme != 2 || (r != nullptr && N >= 3), "cb_fd_derivative: non-uniform scheme needs the node vector (N >= 3)");
    CB_REQUIRE(fhalf == nullptr || scheme == 2, "cb_fd_derivative: weighted stencils exist for the non-uniform staggered scheme only");
    CB_REQUIRE(step > 0.0, "cb_fd_derivative: step must be positive");
    if (n == 0) return CB_OK;
    CB_REQUIRE(d != nullptr, "cb_fd_derivative: NULL diagonal");
    fd_stencil_kernel<<<cb_grid_for(n, 256, 16), 256, 0, (cudaStream_t)stream>>>(scheme, nder, r, fhalf, N, step, i0, n, Z, ell, dl, d, du);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

}  // extern "C"
