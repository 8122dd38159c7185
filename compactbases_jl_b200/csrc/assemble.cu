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
// s is the 0-based index of an interval of t.t (shared by L and R, src/
// bsplines.jl:75-82); its 1-based index in an expanded knot vector is s + ml.
// A CTA owns TJ consecutive columns (functions of R) and the TJ + k - 1 knot
// intervals those columns live on; basis_functions(t, x, m) (src/bsplines.jl:
// 28-52), which the reference stores as a sparse matrix in memory, never
// leaves the SM.
// ---------------------------------------------------------------------------
#include "asm_halo.cuh"

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
    int two_tab;                  // tiled kernel: L and R need different tables (da != db or different left multiplicity)
    double *band;                 // column col_lo is at band[0]
    AsmHalo H;                    // interval-sharded assembly: fused halo exchange over peer memory (all zero: none)
};

constexpr int ASM_THREADS = 256;
constexpr int ASM_SMAX = 32 + CB_KMAX;             // most intervals a column tile of the generic kernel can touch

// idx = j*(j-1)/2 + r  ->  (j, r) for the reciprocal table (j = 1..CB_KMAX-1, r < j)
struct RcIndex { unsigned char j[CB_KMAX * (CB_KMAX - 1) / 2], r[CB_KMAX * (CB_KMAX - 1) / 2]; };
static RcIndex make_rc_index() {
    RcIndex t; int idx = 0;
    for (int j = 1; j < CB_KMAX; ++j) for (int r = 0; r < j; ++r) { t.j[idx] = (unsigned char)j; t.r[idx] = (unsigned char)r; ++idx; }
    return t;
}
__constant__ RcIndex c_rc_index;

// ---------------------------------------------------------------------------
// Tiled kernel (both bases of order K, q <= 2 QH nodes per interval).  Everything below is
// bound by shared-memory wavefronts, not by FP64 issue (ncu, profiles/r1l): the phases are laid
// out so that every shared-memory access is a 16-byte one and every loaded value is used twice.
//
//   phase 0  per interval: its 2K-knot window and the K(K-1)/2 reciprocal knot differences of
//            the Cox-de Boor recursion (shared by the interval's q nodes), one 16-byte aligned
//            record IT[sl]
//   phase 1  one thread per (interval, node pair): the K live functions (db-th derivative; da-th
//            as well when the two tables differ) at both nodes by one interleaved sweep; rows
//            of 2 QH node values, stored as 16-byte pairs; the node weight (-1)^a w_l op(x_l)
//            is kept separately.  The pad node of an odd q gets weight and values 0.
//   phase 2a one thread per (interval, row pair a, a + ceil(K/2)): 2K entries of the interval's
//            local matrix  M_s[a][b] = sum_l (L_a w)_l R_b,l ; the weighted L rows stay in
//            registers, the R rows are shared-memory broadcasts within the interval
//   phase 2b one thread per band entry: sum of the <= K local matrices that hold it, intervals
//            in increasing order — overlap_matrix! (src/bsplines.jl:58-73) in gather form, so
//            no atomics, and every entry is the same sum whatever the tiling or restriction
//
// QP/2 is odd so that the 16-byte reads of consecutive (interval, row) threads fall into
// different banks; local matrices have the odd row pitch K|1.
// ---------------------------------------------------------------------------
template <int QH> struct AsmPitch { static constexpr int QP = (QH & 1) ? 2 * QH : 2 * QH + 2; };

// Knot and reciprocal tables of a tile, each in two copies so that a run starting at ANY index
// can be read as 16-byte pairs: copy 0 holds v[0], v[1], ..., copy 1 holds v[1], v[2], ...
// (run starting at odd p: pairs of copy 1 from index p-1).
//   KN[c][e]      knot with expanded 0-based index sA + ml - K + e (clamped), e < NS + 2K
//   RD[c][j][p]   1 / (KN[p + j] - KN[p]), j = 1..K-1: the reciprocal of level j, entry r of the
//                 Cox-de Boor sweep on interval slot sl is RD[j][sl + K - j + r]
struct AsmTab {
    double *sm;                                     // the CTA's dynamic shared memory (keeps the accesses LDS/STS)
    int kn, rd, pitch;                              // offsets of KN / RD in doubles; doubles per level row, even
    int kn_copy, rd_copy;                           // doubles between the two copies: an odd number of 16-byte units, so that
                                                    // threads of neighbouring intervals (different copies) hit different banks
};

// Two interleaved Cox-de Boor sweeps (cb_bspline_all_rc, bspline_core.cuh) for two nodes x0, x1
// of interval slot sl.
template <int K>
__device__ __forceinline__ void asm_bspline_pair(const AsmTab &T, int sl, double x0, double x1, int m,
                                                 double (&N0)[K], double (&N1)[K]) {
    double tw[2 * K];                               // knot window of the interval
    {
        const double2 *k2 = reinterpret_cast<const double2 *>(T.sm + T.kn + (sl & 1) * T.kn_copy + (sl & ~1));
#pragma unroll
        for (int t = 0; t < K; ++t) { const double2 v = k2[t]; tw[2 * t] = v.x; tw[2 * t + 1] = v.y; }
    }
#pragma unroll
    for (int r = 0; r < K; ++r) { N0[r] = 0.0; N1[r] = 0.0; }
    if (m >= K) return;
    N0[0] = 1.0; N1[0] = 1.0;
    const int jmax = K - 1 - m;
#pragma unroll
    for (int j = 1; j <= K - 1; ++j) {
        const int p0 = sl + K - j;                  // first entry of the level's reciprocal run
        const double2 *r2 = reinterpret_cast<const double2 *>(T.sm + T.rd + (p0 & 1) * T.rd_copy + j * T.pitch + (p0 & ~1));
        double rc[K];
#pragma unroll
        for (int h = 0; h < (j + 1) / 2; ++h) { const double2 v = r2[h]; rc[2 * h] = v.x; if (2 * h + 1 < K) rc[2 * h + 1] = v.y; }
        if (j <= jmax) {
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int r = 0; r < j; ++r) {
                const double t0 = N0[r] * rc[r], t1 = N1[r] * rc[r];
                N0[r] = fma(tw[K + r] - x0, t0, s0); N1[r] = fma(tw[K + r] - x1, t1, s1);
                s0 = (x0 - tw[K - j + r]) * t0; s1 = (x1 - tw[K - j + r]) * t1;
            }
            N0[j] = s0; N1[j] = s1;
        } else {
            double p0_ = 0.0, p1_ = 0.0;
#pragma unroll
            for (int r = 0; r <= j; ++r) {
                double q0 = 0.0, q1 = 0.0;
                if (r < j) { q0 = N0[r] * rc[r]; q1 = N1[r] * rc[r]; }
                N0[r] = (double)j * (p0_ - q0); N1[r] = (double)j * (p1_ - q1);
                p0_ = q0; p1_ = q1;
            }
        }
    }
}

// Local matrices: K columns of pitch KP2 (even) per interval; the interval pitch is an odd multiple (= KH mod 8) of
// 16 bytes so that the 16-byte stores of consecutive (interval, row pair) threads are bank-conflict free.
template <int K> struct AsmMs {
    static constexpr int KP2 = (K + 1) / 2 * 2, KH = (K + 1) / 2;
    static constexpr int U0 = K * KP2 / 2;                     // 16-byte units needed
    static constexpr int U = U0 + ((KH - U0) % 8 + 8) % 8;     // smallest U >= U0 with U = KH (mod 8)
    static constexpr int MSP = 2 * U;
};

// Geometry of one column tile: columns [j0, j1), intervals [sA, sB] of t.t.
struct AsmTile { i64 j0, j1, sA, sB; int ns; };

// Halo hooks of a column tile [j0, j1) (CTA-uniform; contain CTA barriers).
__device__ __forceinline__ bool halo_tile_pushes(const AsmArgs &A, i64 j0) { return A.H.n_push > 0 && j0 < A.col_lo + A.H.n_push; }
__device__ __forceinline__ bool halo_tile_receives(const AsmArgs &A, i64 j1) { return A.H.n_recv > 0 && j1 > A.col_hi - A.H.n_recv; }

__device__ __forceinline__ void halo_before_stores(const AsmArgs &A, i64 j0, i64 j1) {
    const bool rcv = halo_tile_receives(A, j1), psh = halo_tile_pushes(A, j0) && A.H.epoch > 2;
    if (!rcv && !psh) return;
    if (threadIdx.x == 0) {
        if (rcv) halo_spin(A.H.recv_flag, A.H.epoch, A.H.status);
        if (psh) halo_spin(A.H.ack_in, A.H.epoch - 2, A.H.status);     // the buffer of this parity has been consumed
    }
    __syncthreads();
}

// band entry (d, j): pushed columns go to the neighbour's mailbox, receiving columns add what the neighbour sent
__device__ __forceinline__ void halo_store(const AsmArgs &A, int W, i64 j, int d, double v, bool in_run) {
    if (A.H.n_push > 0 && j < A.col_lo + A.H.n_push) {
        A.H.push_cols[d + (i64)W * (j - A.col_lo)] = v;
        return;
    }
    if (in_run && A.H.n_recv > 0 && j >= A.col_hi - A.H.n_recv)
        v += __ldcv(A.H.recv_cols + d + (i64)W * (j - (A.col_hi - A.H.n_recv)));
    A.band[d + (i64)W * (j - A.col_lo)] = v;
}

__device__ __forceinline__ void halo_after_stores(const AsmArgs &A, i64 j0, i64 j1) {
    const bool rcv = halo_tile_receives(A, j1), psh = halo_tile_pushes(A, j0);
    if (!rcv && !psh) return;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (psh && atomicAdd(&A.H.counters[0], 1u) == (unsigned)(A.H.n_push_tiles - 1)) {
            A.H.counters[0] = 0;
            __threadfence_system();
            st_release_sys(A.H.push_flag, A.H.epoch);
        }
        if (rcv && atomicAdd(&A.H.counters[1], 1u) == (unsigned)(A.H.n_recv_tiles - 1)) {
            A.H.counters[1] = 0;
            __threadfence_system();
            st_release_sys(A.H.ack_out, A.H.epoch);
        }
    }
}

template <int K>
__device__ __forceinline__ AsmTile asm_tile(const AsmArgs &A, i64 tile) {
    AsmTile t;
    const int mlR = A.kvR.ml;
    t.j0 = A.col_lo + tile * A.TJ;
    t.j1 = t.j0 + A.TJ < A.col_hi ? t.j0 + A.TJ : A.col_hi;           // exclusive
    // function fj (1-based) lives on s = fj-mlR .. fj-mlR+K-1
    t.sA = (A.colR0 + 1 + t.j0) - mlR;
    t.sB = (A.colR0 + t.j1) - mlR + K - 1;                             // inclusive
    if (t.sA < 0) t.sA = 0;
    if (t.sB > A.kvR.n_tt - 2) t.sB = A.kvR.n_tt - 2;
    if (t.sA < A.s_lo) t.sA = A.s_lo;
    if (t.sB > A.s_hi - 1) t.sB = A.s_hi - 1;
    t.ns = t.sB >= t.sA ? (int)(t.sB - t.sA + 1) : 0;
    return t;
}

template <int K, int QH>
__global__ void __launch_bounds__(ASM_THREADS, 2)
bspline_assemble_tiled_kernel(AsmArgs A) {
    extern __shared__ __align__(16) double sm[];
    constexpr int QP = AsmPitch<QH>::QP, KP2 = (K + 1) / 2 * 2, KH = (K + 1) / 2, W2 = 2 * K - 1;
    constexpr int TS = KH * QP;                    // doubles per interval in each half table: KH * QP/2 16-byte units, odd multiple per row
    constexpr int MSP = AsmMs<K>::MSP;             // doubles per interval of the local matrices
    const int q = A.q, W = A.l + A.u + 1, TJ = A.TJ;
    const int NS = TJ + K - 1;
    const int NKP = (NS + 2 * K + 2) / 2 * 2;      // pitch of the knot / reciprocal rows
    const int mlL = A.kvL.ml, mlR = A.kvR.ml;
    const i64 n_tt = A.kvR.n_tt;
    // tables: row a of interval sl is at  T?[(a & 1) * NS * TS + sl * TS + (a >> 1) * QP]  — even rows and odd rows in
    // separate halves, so that the row pairs (2 ap, 2 ap + 1) of consecutive (interval, ap) threads are QP/2 (odd) 16-byte
    // units apart in each half
    double *TR = sm;                               // [2][NS][KH][QP]  raw db-th derivatives
    double *TL = A.two_tab ? TR + (size_t)2 * NS * TS : TR;   //               raw da-th derivatives
    double *WL = TL + (size_t)2 * NS * TS;         // [NS][QP]         (-1)^a w op
    double *MS = WL + (size_t)NS * QP;             // [NS][MSP]        local matrices, TRANSPOSED: MS[sl][b * KP2 + a]
    double *KN = MS + (size_t)NS * MSP;            // [2][NKP + 2]
    double *RD = KN + 2 * (NKP + 2);               // [2][K * NKP + 2]
    int *sOrd = reinterpret_cast<int *>(RD + (size_t)2 * (K * NKP + 2));   // [NS] ordinal among the non-empty intervals, -1 if empty
    const int HALF = NS * TS;
    AsmTab T;
    T.sm = sm; T.kn = (int)(KN - sm); T.rd = (int)(RD - sm); T.pitch = NKP;
    T.kn_copy = ((NKP / 2) | 1) * 2; T.rd_copy = ((K * NKP / 2) | 1) * 2;
    const double sgn = (A.da & 1) ? -1.0 : 1.0;    // src/bspline_derivatives.jl:14
    const i64 ntiles = (A.col_hi - A.col_lo + TJ - 1) / TJ;

    // Three barriers per tile:  [gather of the previous tile | tables of this tile] -> phase 1 -> phase 2a
    AsmTile prev; prev.ns = -1; prev.j0 = prev.j1 = prev.sA = prev.sB = 0;
    for (i64 tile = blockIdx.x; ; tile += gridDim.x) {
        const bool live = tile < ntiles;
        AsmTile t = prev;
        if (live) t = asm_tile<K>(A, tile);
        const int ns = live ? t.ns : 0;
        const i64 sA = t.sA;
        // this thread's phase-1 item: its node data is fetched now and used after the barrier
        const int e1 = threadIdx.x, sl1 = e1 / QH, lp1 = e1 - sl1 * QH;
        int c1 = -1;
        double x0 = 1.0, x1 = 1.0, w0 = 1.0, w1 = 1.0, o0 = 1.0, o1 = 1.0;
        const bool has1 = 2 * lp1 + 1 < q;         // the second node of the last pair is padding when q is odd
        if (e1 < ns * QH) {
            c1 = A.ordR[sA + sl1 + mlR - 1];
            if (c1 >= 0) {
                const i64 g0 = (i64)c1 * q + 2 * lp1, g1 = has1 ? g0 + 1 : g0;
                x0 = A.x[g0]; x1 = A.x[g1];
                if (A.coulomb != 2) { w0 = A.w[g0]; w1 = A.w[g1]; }
                if (A.opvals) { o0 = A.opvals[g0]; o1 = A.opvals[g1]; }
            }
        }
        // ---- phase 0 (this tile) -------------------------------------------------
        for (int e = threadIdx.x; e < ns; e += ASM_THREADS) sOrd[e] = A.ordR[sA + e + mlR - 1];
        for (int j = threadIdx.x >> 5; j < K; j += ASM_THREADS / 32) {           // one warp per level
            for (int p = threadIdx.x & 31; p < ns + 2 * K; p += 32) {
                const double lo = A.kvR.at(sA + mlR - K + p);
                if (j == 0) {
                    KN[p] = lo;
                    if (p > 0) KN[T.kn_copy + p - 1] = lo;
                } else if (p + j < ns + 2 * K) {
                    const double dk = A.kvR.at(sA + mlR - K + p + j) - lo;
                    const double v = dk != 0.0 ? fast_rcp(dk) : 0.0;     // zero widths only feed empty intervals / absent functions
                    RD[j * NKP + p] = v;
                    if (p > 0) RD[T.rd_copy + j * NKP + p - 1] = v;
                }
            }
        }
        // ---- phase 2b (previous tile): gather into the band -------------------------
        // one thread per band entry of the (2K-1)-row run of each column; run entry tt of a column is
        // the sum over its intervals r of MS[slot + r][K-1-r][tt - r]
        if (prev.ns >= 0) {
            halo_before_stores(A, prev.j0, prev.j1);
            const int nent = (int)(prev.j1 - prev.j0) * W2;
            for (int e = threadIdx.x; e < nent; e += ASM_THREADS) {
                const int jj = e / W2, tt = e - jj * W2;
                const i64 j = prev.j0 + jj, fj = A.colR0 + 1 + j;
                const int slot0 = (int)(fj - mlR - prev.sA);          // slot of the first interval of the column (may be < 0)
                double s = 0.0;
#pragma unroll
                for (int r = 0; r < K; ++r) {
                    const int sl = slot0 + r, a = tt - r;
                    if (a >= 0 && a < K && sl >= 0 && sl < prev.ns)
                        s += MS[sl * MSP + (K - 1 - r) * KP2 + a];
                }
                // run entry tt is row fi = fj - mlR + mlL - K + 1 + tt; band slot d holds row i = j - u + d
                const i64 i = (fj - mlR + mlL - K + 1 + tt) - (A.rowL0 + 1);
                const i64 d = i - j + A.u;
                if (d >= 0 && d < W && i >= 0 && i < A.mm && j < A.nn) halo_store(A, W, j, (int)d, s, true);
            }
            // band slots outside the run or outside the matrix are zero
            const int nz = (int)(prev.j1 - prev.j0) * W;
            for (int e = threadIdx.x; e < nz; e += ASM_THREADS) {
                const int jj = e / W, d = e - jj * W;
                const i64 j = prev.j0 + jj, i = j - A.u + d;
                const i64 tt = (i + A.rowL0 + 1) - (A.colR0 + 1 + j - mlR + mlL - K + 1);
                if (tt < 0 || tt >= W2 || i < 0 || i >= A.mm || j >= A.nn) halo_store(A, W, j, d, 0.0, false);
            }
            halo_after_stores(A, prev.j0, prev.j1);
        }
        if (!live) break;
        prev = t;
        __syncthreads();
        // ---- phase 1 ----------------------------------------------------------
        if (c1 >= 0) {                              // empty intervals are skipped (zero local matrix in phase 2a)
            if (!A.opvals && A.coulomb == 1) { o0 = -A.Z * fast_rcp(x0); o1 = -A.Z * fast_rcp(x1); }
            const int sl = sl1, l0 = 2 * lp1;
            double2 *tr2 = reinterpret_cast<double2 *>(TR + sl * TS + l0);
            double N0[K], N1[K];
            asm_bspline_pair<K>(T, sl, x0, x1, A.db, N0, N1);
#pragma unroll
            for (int a = 0; a < K; ++a) tr2[((a & 1) * HALF + (a >> 1) * QP) / 2] = make_double2(N0[a], has1 ? N1[a] : 0.0);
            *reinterpret_cast<double2 *>(WL + (size_t)sl * QP + l0) = make_double2((sgn * w0) * o0, has1 ? (sgn * w1) * o1 : 0.0);
            if (A.two_tab) {
                double2 *tl2 = reinterpret_cast<double2 *>(TL + sl * TS + l0);
                if (mlL == mlR) {                   // same knot window, other derivative order
                    asm_bspline_pair<K>(T, sl, x0, x1, A.da, N0, N1);
                } else {                            // windows differ by the clamping at the left end
                    double tw[2 * K];
                    cb_knot_window<K>(A.kvL.tt, n_tt, mlL, (int)(sA + sl + mlL), tw);
                    cb_bspline_all<K>(tw, x0, A.da, N0);
                    cb_bspline_all<K>(tw, x1, A.da, N1);
                }
#pragma unroll
                for (int a = 0; a < K; ++a) tl2[((a & 1) * HALF + (a >> 1) * QP) / 2] = make_double2(N0[a], has1 ? N1[a] : 0.0);
            }
        }
        __syncthreads();
        // ---- phase 2a: local matrices ---------------------------------------------
        for (int e = threadIdx.x; e < ns * KH; e += ASM_THREADS) {
            const int sl = e / KH, ap = e - sl * KH;
            const int a0 = 2 * ap, a1 = a0 + 1;     // rows a0, a0 + 1: one 16-byte store per column of the local matrix
            const bool two = a1 < K;
            double2 *m2 = reinterpret_cast<double2 *>(MS + sl * MSP + a0);
            if (sOrd[sl] < 0) {
#pragma unroll
                for (int b = 0; b < K; ++b) m2[b * (KP2 / 2)] = make_double2(0.0, 0.0);
                continue;
            }
            const double2 *wl2 = reinterpret_cast<const double2 *>(WL + (size_t)sl * QP);
            const double2 *tl0 = reinterpret_cast<const double2 *>(TL + sl * TS + ap * QP);          // row 2 ap
            const double2 *tl1 = reinterpret_cast<const double2 *>(TL + (two ? HALF : 0) + sl * TS + ap * QP);   // row 2 ap + 1
            double2 lw0[QH], lw1[QH];
#pragma unroll
            for (int h = 0; h < QH; ++h) {
                const double2 w2 = wl2[h], t0 = tl0[h], t1 = tl1[h];
                lw0[h] = make_double2(t0.x * w2.x, t0.y * w2.y);
                lw1[h] = make_double2(t1.x * w2.x, t1.y * w2.y);
            }
            const double2 *tr2 = reinterpret_cast<const double2 *>(TR + sl * TS);
#pragma unroll
            for (int b = 0; b < K; ++b) {
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int h = 0; h < QH; ++h) {
                    const double2 r2 = tr2[((b & 1) * HALF + (b >> 1) * QP) / 2 + h];
                    s0 = fma(lw0[h].x, r2.x, s0); s1 = fma(lw1[h].x, r2.x, s1);
                    s0 = fma(lw0[h].y, r2.y, s0); s1 = fma(lw1[h].y, r2.y, s1);
                }
                m2[b * (KP2 / 2)] = make_double2(s0, two ? s1 : 0.0);
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------
// Generic kernel: run-time orders (kL != kR, test/bsplines/derivatives.jl:36-79), unusual node
// counts, k == 1.  Phase 1 evaluates both tables (already weighted) into shared memory, phase 2
// is one thread per band entry summing over the common intervals and their nodes in increasing
// node order, the order of the reference's sparse dot product.
// ---------------------------------------------------------------------------
// padded node count: even (16-byte pairs) and = 2 mod 4 so that rows of different functions fall
// into different banks for the 16-byte reads of phase 2
__host__ __device__ inline int asm_qpad(int q) { int p = q + (q & 1); return (p & 3) == 0 ? p + 2 : p; }

__global__ void __launch_bounds__(ASM_THREADS)
bspline_assemble_kernel(AsmArgs A) {
    extern __shared__ __align__(16) double sm[];
    const int kL = A.kL, kR = A.kR;
    const int q = A.q, W = A.l + A.u + 1, TJ = A.TJ;
    const int QP = asm_qpad(q);
    const int mlL = A.kvL.ml, mlR = A.kvR.ml;
    const i64 n_tt = A.kvR.n_tt;
    const int SMAX = TJ + kR - 1;                  // intervals a column tile can touch
    double *TL = sm;                               // [SMAX][kL][QP]  (node index fastest, zero padded)
    double *TR = TL + (size_t)SMAX * kL * QP;      // [SMAX][kR][QP]
    const double sgn = (A.da & 1) ? -1.0 : 1.0;    // src/bspline_derivatives.jl:14
    const i64 ntiles = (A.col_hi - A.col_lo + TJ - 1) / TJ;
    __shared__ int sOrd[ASM_SMAX];                  // ordinal of each interval of the tile among the non-empty ones, -1 if empty

    for (i64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const i64 j0 = A.col_lo + tile * TJ;
        const i64 j1 = j0 + TJ < A.col_hi ? j0 + TJ : A.col_hi;      // exclusive
        i64 sA = (A.colR0 + 1 + j0) - mlR;
        i64 sB = (A.colR0 + j1) - mlR + kR - 1;                       // inclusive
        if (sA < 0) sA = 0;
        if (sB > n_tt - 2) sB = n_tt - 2;
        if (sA < A.s_lo) sA = A.s_lo;
        if (sB > A.s_hi - 1) sB = A.s_hi - 1;
        const int ns = sB >= sA ? (int)(sB - sA + 1) : 0;
        __syncthreads();                            // previous tile consumed
        for (int e = threadIdx.x; e < ns; e += ASM_THREADS) sOrd[e] = A.ordR[sA + e + mlR - 1];
        // zero the node padding of both tables (never written in phase 1)
        for (int e = threadIdx.x; e < ns * (kL + kR) * (QP - q); e += ASM_THREADS) {
            const int row = e / (QP - q), c = e - row * (QP - q);
            (row < ns * kL ? TL + (size_t)row * QP : TR + (size_t)(row - ns * kL) * QP)[q + c] = 0.0;
        }
        __syncthreads();
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
        __syncthreads();
        // ---- phase 2: gather into the band ----------------------------------
        const int nent = (int)(j1 - j0) * W;
        halo_before_stores(A, j0, j1);
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
            halo_store(A, W, j, d, acc, true);
        }
        halo_after_stores(A, j0, j1);
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

// OT > 0: the (uniform) element order at compile time — index arithmetic without integer divisions, unrolled sums.
template <int OT>
__global__ void __launch_bounds__(FEA_WARPS * 32)
fedvr_assemble_kernel(const double *__restrict__ x, const double *__restrict__ wcat, const double *__restrict__ nrm,
                      const i64 *__restrict__ estart, const int32_t *__restrict__ order,
                      const i64 *__restrict__ boff, const i64 *__restrict__ woff, int uniform_order,
                      i64 nel, int nder, int omax, double *__restrict__ blocks) {
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *Lp = sm + (size_t)warp * (2 * omax * omax + 3 * omax);   // [o][o] column-major L'[m + o*m']
    double *rd = Lp + omax * omax;                  // [o][o] 1 / (x_m - x_k), m != k
    double *xe = rd + omax * omax, *we = xe + omax, *ne = we + omax;
    for (i64 e = (i64)blockIdx.x * FEA_WARPS + warp; e < nel; e += (i64)gridDim.x * FEA_WARPS) {
        const int o = OT > 0 ? OT : (uniform_order > 0 ? uniform_order : order[e]);
        const i64 es = uniform_order > 0 ? e * (o - 1) : estart[e];
        const i64 ws = uniform_order > 0 ? e * o : woff[e];
        const i64 bo = uniform_order > 0 ? e * (i64)o * o : boff[e];
        __syncwarp();
        for (int t = lane; t < o; t += 32) { xe[t] = x[es + t]; we[t] = wcat[ws + t]; ne[t] = nrm[es + t]; }
        __syncwarp();
        // the o(o-1) reciprocal node differences once per element: the product form of lagrangeder!
        // (src/fedvr_derivatives.jl:15-31) divides by them o-2 times per entry, which made this kernel issue-bound on
        // FP64 divisions (27 M warp instructions at C2)
        for (int t = lane; t < o * o; t += 32) {
            const int m = t % o, kk = t / o;
            rd[m * o + kk] = m == kk ? 0.0 : 1.0 / (xe[m] - xe[kk]);
        }
        __syncwarp();
        for (int t = lane; t < o * o; t += 32) {
            const int m = t % o, mp = t / o;
            double v;
            if (m == mp) {
                v = (double)((m == o - 1) - (m == 0)) / (2 * we[m]);
            } else {
                double f = 1.0;
                const double xmp = xe[mp];
                const double *rm = rd + m * o;
#pragma unroll
                for (int kk = 0; kk < (OT > 0 ? OT : o); ++kk) {
                    if (kk == m || kk == mp) continue;
                    f *= (xmp - xe[kk]) * rm[kk];
                }
                v = f * rm[mp];
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
#pragma unroll
                for (int qd = 0; qd < (OT > 0 ? OT : o); ++qd) s += (-Lp[m + o * qd]) * (we[qd] * Lp[mp + o * qd]);
            }
            blocks[bo + t] = ne[m] * ne[mp] * s;
        }
    }
}

// Uniform element order O <= 16: floor(32 / O) elements per warp, one lane per ROW m of an element.  The lane keeps the
// element's nodes, its reciprocal differences 1 / (x_m - x_k) and its row of L' in registers with compile-time
// indices (no shared-memory tables, no index divisions); only the rows of (w .* L') that every other row of the element
// needs pass through shared memory.  Same operations in the same order as fedvr_assemble_kernel — bit-identical —
// at a sixth of the warp instructions (the one-warp-per-element kernel is issue-bound: 1382 warp instructions per
// element at order 10, 19 of 32 lanes active).
#ifndef FEG_WARPS_PER_BLOCK
#define FEG_WARPS_PER_BLOCK 1
#endif
#ifndef FEG_MAXREG
#define FEG_MAXREG 255
#endif
constexpr int FEG_WARPS = FEG_WARPS_PER_BLOCK;

template <int O>
__global__ void __launch_bounds__(FEG_WARPS * 32)
#if FEG_MAXREG < 255
__maxnreg__(FEG_MAXREG)
#endif
fedvr_assemble_rows_kernel(const double *__restrict__ x, const double *__restrict__ wcat, const double *__restrict__ nrm,
                           i64 nel, int nder, double *__restrict__ blocks) {
    constexpr int G = 32 / O;                       // elements per warp
    constexpr int OP = O + (O & 1);                 // row stride (even: rows are read as double2)
    __shared__ __align__(16) double sD[FEG_WARPS][G * O * OP];      // rows of x_m - x_k
    __shared__ __align__(16) double sL[FEG_WARPS][G * O * OP];      // rows of L'
    __shared__ __align__(16) double sW[FEG_WARPS][G * O * OP];      // rows of w .* L' (second derivative)
    __shared__ __align__(16) double sV[FEG_WARPS][2][G * OP];       // w and n of the elements
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool lane_used = lane < G * O;
    const int g = lane_used ? lane / O : 0, m = lane_used ? lane - g * O : 0;
    double *dif = sD[warp] + g * (O * OP), *raw = sL[warp] + g * (O * OP), *rows = sW[warp] + g * (O * OP);
    double *wv = sV[warp][0] + g * OP, *nv = sV[warp][1] + g * OP;
    const i64 ngroups = (nel + G - 1) / G;
    // a dependent launch (the apply pipeline of cb_fedvr_assemble_apply) may take the SMs this grid leaves
    asm volatile("griddepcontrol.launch_dependents;");
    for (i64 grp = (i64)blockIdx.x * FEG_WARPS + warp; grp < ngroups; grp += (i64)gridDim.x * FEG_WARPS) {
        const i64 e0 = grp * G + g;
        const bool live = lane_used && e0 < nel;
        const i64 e = e0 < nel ? e0 : nel - 1;
        const double *xe = x + e * (O - 1);
        const double xm = xe[m], wm = wcat[e * O + m], nm = nrm[e * (O - 1) + m];
        double r[O], one_at_m[O];
        // r[k] = 1 / (x_m - x_k); the k == m slot holds 0 and its factor below becomes fma(d, 0, 1) = 1 exactly
#pragma unroll
        for (int k = 0; k < O; ++k) {
            const double dk = xm - xe[k];
            if (lane_used) dif[m * OP + k] = dk;
            // own slot: +0.0 becomes 1.0 through the exponent bits (a select on the denominator is folded away by the
            // compiler, and 1 / 0 sends the whole warp through the slow path of the division)
            const double q = 1.0 / __hiloint2double(__double2hiint(dk) | (k == m ? 0x3ff00000 : 0), __double2loint(dk));
            r[k] = k == m ? 0.0 : q;
            one_at_m[k] = k == m ? 1.0 : 0.0;
        }
        if (lane_used) { wv[m] = wm; nv[m] = nm; }
        // (delta_{m,O} - delta_{m,1}) / (2 w_m): the zero numerators would take the slow path of the division
        const double hw = 1.0 / (2 * wm);
        const double diag = m == O - 1 ? hw : (m == 0 ? -hw : 0.0);
        __syncwarp();
        // L'[m, mp]  (src/fedvr_derivatives.jl:15-31)
#pragma unroll
        for (int mp = 0; mp < O; ++mp) {
            double dmp[OP];
#pragma unroll
            for (int k2 = 0; k2 < OP / 2; ++k2) {
                const double2 v = reinterpret_cast<const double2 *>(dif + mp * OP)[k2];
                dmp[2 * k2] = v.x; dmp[2 * k2 + 1] = v.y;
            }
            double f = 1.0;
#pragma unroll
            for (int k = 0; k < O; ++k) {
                if (k == mp) continue;
                f = __dmul_rn(f, __fma_rn(dmp[k], r[k], one_at_m[k]));
            }
            const double v = mp == m ? diag : __dmul_rn(f, r[mp]);
            if (lane_used) {
                raw[m * OP + mp] = v;
                if (nder != 1) rows[m * OP + mp] = __dmul_rn(wv[mp], v);
            }
        }
        if ((O & 1) && lane_used) { raw[m * OP + O] = 0.0; rows[m * OP + O] = 0.0; }
        __syncwarp();
        if (nder == 1) {
            // D[m, mp] = n_m n_mp w_m L'[mp, m]
#pragma unroll
            for (int mp = 0; mp < O; ++mp) {
                const double s = __dmul_rn(wm, raw[mp * OP + m]);
                if (live) blocks[e * (i64)(O * O) + mp * O + m] = __dmul_rn(__dmul_rn(nm, nv[mp]), s);
            }
        } else {
            // D[m, mp] = n_m n_mp sum_q (-L'[m, q]) (w_q L'[mp, q])   (:33-57)
            double nl[OP];
#pragma unroll
            for (int q2 = 0; q2 < OP / 2; ++q2) {
                const double2 v = reinterpret_cast<const double2 *>(raw + m * OP)[q2];
                nl[2 * q2] = -v.x; nl[2 * q2 + 1] = -v.y;
            }
#pragma unroll
            for (int mp = 0; mp < O; ++mp) {
                const double2 *row = reinterpret_cast<const double2 *>(rows + mp * OP);
                double s = 0.0;
#pragma unroll
                for (int q2 = 0; q2 < OP / 2; ++q2) {
                    const double2 v = row[q2];
                    s = __fma_rn(nl[2 * q2], v.x, s);
                    if (2 * q2 + 1 < O) s = __fma_rn(nl[2 * q2 + 1], v.y, s);
                }
                if (live) blocks[e * (i64)(O * O) + mp * O + m] = __dmul_rn(__dmul_rn(nm, nv[mp]), s);
            }
        }
        __syncwarp();
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

// alpha_j, beta_j, gamma_j, delta_j of node j (1-based) of the parent grid.
__device__ __forceinline__ void fd_coeffs(int scheme, const double *__restrict__ r, const double *__restrict__ fhalf, i64 N,
                                          i64 j, double &al, double &be, double &ga, double &de) {
    al = 1.0; be = 1.0; ga = 0.0; de = 1.0;        // Cartesian, src/fd_derivatives.jl:12-15
    if (scheme == 1) {                              // :20-25
        const double j2 = (double)(j * j);
        al = __ddiv_rn(j2, __dsub_rn(j2, 0.25));
        de = al;
        const double t = (double)(j * j - j);
        be = __ddiv_rn(__dadd_rn(t, 0.5), __dadd_rn(t, 0.25));
    } else if (scheme == 2) {                       // :27-92
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
}

// The three entries of row/column j (1-based parent node) of the parent operator:
//   dg = A[j, j], up = A[j, j+1] (du[j]), lo = A[j+1, j] (dl[j])        (src/fd_derivatives.jl:233-237, :265-277)
__device__ __forceinline__ void fd_entries(int scheme, int nder, const double *__restrict__ r, const double *__restrict__ fhalf,
                                           i64 N, double step, i64 j, double Z, int ell, double &dg, double &up, double &lo) {
    double al, be, ga, de;
    fd_coeffs(scheme, r, fhalf, N, j, al, be, ga, de);
    if (nder == 1) {                                // :233-237
        const double mu = __dmul_rn(2.0, step);
        dg = __ddiv_rn(ga, mu);
        up = __ddiv_rn(de, mu);
        lo = __ddiv_rn(-de, mu);
    } else {                                        // :265-277
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
        dg = dv;
        up = lo = __ddiv_rn(al, h2);
    }
}

__global__ void fd_stencil_kernel(int scheme, int nder, const double *__restrict__ r, const double *__restrict__ fhalf,
                                  i64 N, double step, i64 i0, i64 n, double Z, int ell,
                                  double *__restrict__ dl, double *__restrict__ d, double *__restrict__ du) {
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const i64 j = i0 + i + 1;                   // 1-based node of the parent grid
        double dg, up, lo;
        fd_entries(scheme, nder, r, fhalf, N, step, j, Z, ell, dg, up, lo);
        d[i] = dg;
        if (i < n - 1) {
            if (du) du[i] = up;
            if (dl) dl[i] = lo;
        }
    }
}

// Operators between DIFFERENT restrictions A = R[:, ra0+1 : ra0+m], B = R[:, cb0+1 : cb0+n] (derivative_matrix,
// src/fd_derivatives.jl:214-219): the m x n sub-block of the parent tridiagonal operator as a BandedMatrix with
// bandwidths (1 - offset, 1 + offset), offset = ra0 - cb0, filled band by band (:244-249, :278-289).  In
// BandedMatrices storage band[(u + i - j) + 3 j] the three rows are du, d, dl of the parent at the column's node.
// One thread per column.
__global__ void fd_band_kernel(int scheme, int nder, const double *__restrict__ r, const double *__restrict__ fhalf,
                               i64 N, double step, i64 ra0, i64 m, i64 cb0, i64 n, double Z, int ell,
                               double *__restrict__ band) {
    for (i64 jc = blockIdx.x * (i64)blockDim.x + threadIdx.x; jc < n; jc += (i64)gridDim.x * blockDim.x) {
        const i64 pj = cb0 + jc + 1;                // parent node of this column (1-based)
        double dg, up, lo, t0, t1;
        fd_entries(scheme, nder, r, fhalf, N, step, pj, Z, ell, dg, up, lo);
        double above = 0.0;                         // A[pj-1, pj] = up of node pj-1
        if (pj >= 2) fd_entries(scheme, nder, r, fhalf, N, step, pj - 1, Z, ell, t0, above, t1);
        auto row_in = [&](i64 p) { return p >= ra0 + 1 && p <= ra0 + m; };
        band[0 + 3 * jc] = (pj >= 2 && row_in(pj - 1)) ? above : 0.0;
        band[1 + 3 * jc] = row_in(pj) ? dg : 0.0;
        band[2 + 3 * jc] = (pj + 1 <= N && row_in(pj + 1)) ? lo : 0.0;
    }
}

// Non-uniform compact scheme (Patchkovskii & Muller 2016, Eq. 33): implicit_stencil_common (src/fd_derivatives.jl:105-125)
// feeds fun(a, b, c) the consecutive local steps of the grid continued to the origin on the left and by one mirrored
// point on the right; with R_0 = 0, R_i = r_i, R_{N+1} = 2 r_N - r_{N-1} the call with shift s gives, at index j,
// a = R_{j+s-1} - R_{j+s-2}, b = R_{j+s} - R_{j+s-1}.  One thread per node writes the three right-hand-side stencils
// (:127-143) and the left-hand side of the second derivative (implicit_lhs, :410-424) with the reference's order of
// operations.
__device__ __forceinline__ double rext0(const double *__restrict__ r, i64 N, i64 i) {
    if (i <= 0) return 0.0;
    if (i > N) return __dsub_rn(__dmul_rn(2.0, r[N - 1]), r[N - 2]);
    return r[i - 1];
}

__global__ void fd_implicit_stencil_kernel(const double *__restrict__ r, i64 N, i64 i0, i64 n,
                                           double *__restrict__ alpha_t, double *__restrict__ alpha, double *__restrict__ beta,
                                           double *__restrict__ m_dl, double *__restrict__ m_d, double *__restrict__ m_du) {
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const i64 j = i0 + i + 1;
        {   // s = 1: beta_j, m_d_j (j = 1..N), alpha_j, m_du_j (j = 1..N-1)
            const double R0 = rext0(r, N, j - 1), R1 = rext0(r, N, j), R2 = rext0(r, N, j + 1);
            const double a = __dsub_rn(R1, R0), b = __dsub_rn(R2, R1);
            const double ab = __dmul_rn(a, b), apb = __dadd_rn(a, b);
            beta[i] = __ddiv_rn(1.0, ab);
            // (a^2 + 3a*b + b^2)/(6*a*b)
            m_d[i] = __ddiv_rn(__dadd_rn(__dadd_rn(sq_(a), __dmul_rn(__dmul_rn(3.0, a), b)), sq_(b)), __dmul_rn(__dmul_rn(6.0, a), b));
            if (i < n - 1) {
                alpha[i] = __ddiv_rn(2.0, __dmul_rn(b, apb));
                // (-a^2 + a*b + b^2)/(6*b*(a+b))
                m_du[i] = __ddiv_rn(__dadd_rn(__dadd_rn(-sq_(a), ab), sq_(b)), __dmul_rn(__dmul_rn(6.0, b), apb));
            }
        }
        if (i < n - 1) {   // s = 2: alpha~_j, m_dl_j (j = 1..N-1)
            const double R0 = rext0(r, N, j), R1 = rext0(r, N, j + 1), R2 = rext0(r, N, j + 2);
            const double a = __dsub_rn(R1, R0), b = __dsub_rn(R2, R1);
            const double apb = __dadd_rn(a, b);
            alpha_t[i] = __ddiv_rn(2.0, __dmul_rn(a, apb));
            // (a^2 + a*b - b^2)/(6*a*(a+b))
            m_dl[i] = __ddiv_rn(__dsub_rn(__dadd_rn(sq_(a), __dmul_rn(a, b)), sq_(b)), __dmul_rn(__dmul_rn(6.0, a), apb));
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

int cb_asm_stream_launch(const cb_bspline *B, int m, const double *opvals, int coulomb, double Z, i64 rowL0, i64 mm, i64 colR0,
                         i64 nn, int u, i64 col_lo, i64 col_hi, i64 s_lo, i64 s_hi, double *band, cudaStream_t st,
                         const AsmHalo *halo, double *band2 = nullptr);   // assemble_stream.cu

static int assemble_common(const cb_bspline *L, const cb_bspline *R, int a, int b, const double *opvals,
                           int coulomb, double Z, i64 rowL0, i64 m, i64 colR0, i64 n, int l, int u,
                           i64 col_lo, i64 col_hi, i64 s_lo, i64 s_hi, double *band, void *stream,
                           const AsmHalo *halo = nullptr) {
    CB_REQUIRE(L != nullptr && R != nullptr && band != nullptr, "cb_bspline_assemble: NULL argument");
    // assert_compatible_bases, src/bsplines.jl:75-82
    // (the same handle on both sides needs no look at the knots: comparing 1.6e6 of them cost more than the kernel)
    bool same = L == R || (L->n_tt == R->n_tt && L->q == R->q && L->h_tt == R->h_tt && L->h_xi == R->h_xi && L->h_wr == R->h_wr);
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
    // ---- streaming kernel (assemble_stream.cu): the same basis on both sides, a == b <= 1 ----
    if (a == b && L->k == R->k && L->ml == R->ml && L->nf == R->nf) {
        const int src = cb_asm_stream_launch(R, a, opvals, coulomb, Z, rowL0, m, colR0, n, u, col_lo, col_hi, s_lo, s_hi, band, st, halo);
        if (src >= 0) return src;
    }
    AsmArgs A;
    A.kvL = make_knot_view(L); A.kvR = make_knot_view(R);
    A.kL = L->k; A.kR = R->k; A.nfL = L->nf; A.nfR = R->nf;
    A.q = R->q; A.da = a; A.db = b;
    A.x = R->d_x; A.w = R->d_w; A.ordR = R->d_ord;
    A.opvals = opvals; A.coulomb = coulomb; A.Z = Z;
    A.rowL0 = rowL0; A.mm = m; A.colR0 = colR0; A.nn = n;
    A.l = l; A.u = u; A.col_lo = col_lo; A.col_hi = col_hi; A.s_lo = s_lo; A.s_hi = s_hi;
    A.band = band;
    memset(&A.H, 0, sizeof(A.H));
    if (halo) A.H = *halo;
    const i64 ncol = col_hi - col_lo;
    // tiles that hold pushed / receiving columns (the kernels count them down to raise the flags)
    auto halo_tiles = [&](int TJ) {
        if (!halo) return;
        const i64 ntl = (ncol + TJ - 1) / TJ;
        A.H.n_push_tiles = A.H.n_push > 0 ? (int)((A.H.n_push + TJ - 1) / TJ) : 0;
        const i64 first_recv_tile = A.H.n_recv > 0 ? (ncol - A.H.n_recv) / TJ : ntl;
        A.H.n_recv_tiles = (int)(ntl - first_recv_tile);
    };
    // ---- tiled kernel: same order 2..CB_KMAX on both sides and the default node count family ----
    const int K = L->k, QH = (K + 2) / 2;           // q = K + 1 (num_quadrature_points(k, 3)) -> ceil(q / 2)
    if (L->k == R->k && K >= 2 && K <= CB_KMAX && (A.q + 1) / 2 == QH) {
        const int QP = (QH & 1) ? 2 * QH : 2 * QH + 2, KP2 = (K + 1) / 2 * 2;
        A.two_tab = (a != b || L->ml != R->ml) ? 1 : 0;
        const int KH = (K + 1) / 2, U0 = K * KP2 / 2, MSP = 2 * (U0 + ((KH - U0) % 8 + 8) % 8);   // AsmMs<K>::MSP
        const size_t per_iv = sizeof(double) * ((size_t)2 * KH * QP * (A.two_tab ? 2 : 1) + QP + MSP + 2 * (K + 1)) + sizeof(int);
        const size_t fixed = sizeof(double) * (2 * (K + 1) * (2 * K + 2) + 8) + 16;
        // one round of 256 (interval, node pair) work items per tile; two CTAs per SM when that fits
        int ns = ASM_THREADS / QH;
        while (ns > 2 * K && (size_t)ns * per_iv + fixed > 112 * 1024) --ns;
        while (ns >= K && (size_t)ns * per_iv + fixed > 220 * 1024) --ns;
        if (ns >= K) {
            i64 TJ = ns - K + 1;
            if (TJ > ncol) TJ = ncol;
            A.TJ = (int)TJ;
            const size_t smem = ((size_t)(A.TJ + K - 1) * per_iv + fixed) / 16 * 16;
            const i64 ntiles = (ncol + A.TJ - 1) / A.TJ;
            halo_tiles(A.TJ);
            const int grid = (int)(ntiles < (i64)CB_NUM_SMS * 8 ? ntiles : (i64)CB_NUM_SMS * 8);
#define CB_ASM_TILED(KT)                                                                                     \
    case KT: {                                                                                               \
        auto kern = bspline_assemble_tiled_kernel<KT, (KT + 2) / 2>;                                         \
        CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
        kern<<<grid, ASM_THREADS, smem, st>>>(A);                                                            \
    } break;
            switch (K) {
                CB_ASM_TILED(2) CB_ASM_TILED(3) CB_ASM_TILED(4) CB_ASM_TILED(5) CB_ASM_TILED(6)
                CB_ASM_TILED(7) CB_ASM_TILED(8) CB_ASM_TILED(9) CB_ASM_TILED(10) CB_ASM_TILED(11)
                CB_ASM_TILED(12) CB_ASM_TILED(13) CB_ASM_TILED(14) CB_ASM_TILED(15) CB_ASM_TILED(16)
                default: break;
            }
#undef CB_ASM_TILED
            CB_LAUNCH_CHECK();
            return CB_OK;
        }
    }
    // ---- generic kernel ----
    // column tile: as many columns as fit ~96 KB (two CTAs per SM), at most 32
    A.two_tab = 1;
    const size_t per_iv = sizeof(double) * ((size_t)asm_qpad(A.q) * (A.kL + A.kR));
    int TJ = 32;
    while (TJ > 1 && (size_t)(TJ + A.kR - 1) * per_iv > 112 * 1024) TJ /= 2;   // two CTAs per SM
    size_t smem = (size_t)(TJ + A.kR - 1) * per_iv;
    if (smem > 220 * 1024) { cb_set_error("cb_bspline_assemble: k=%d, q=%d tables exceed shared memory", k, A.q); return CB_ERR_UNSUPPORTED; }
    A.TJ = TJ;
    halo_tiles(TJ);
    i64 ntiles = (ncol + TJ - 1) / TJ;
    int grid = (int)(ntiles < (i64)CB_NUM_SMS * 32 ? ntiles : (i64)CB_NUM_SMS * 32);
    CB_CUDA(cudaFuncSetAttribute(bspline_assemble_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    bspline_assemble_kernel<<<grid, ASM_THREADS, smem, st>>>(A);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

// Interval-sharded assembly with the fused halo (comm.cu).
int cb_bspline_assemble_halo_internal(const cb_bspline *L, const cb_bspline *R, int a, int b, const double *opvals,
                                      int coulomb, double Z, i64 rowL0, i64 m, i64 colR0, i64 n, int l, int u,
                                      i64 col_lo, i64 col_hi, i64 s_lo, i64 s_hi, double *band, void *stream,
                                      const void *halo) {
    return assemble_common(L, R, a, b, opvals, coulomb, Z, rowL0, m, colR0, n, l, u, col_lo, col_hi, s_lo, s_hi, band, stream,
                           static_cast<const AsmHalo *>(halo));
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

// The Hamiltonian pair of a (restricted) basis R = B[:, col0+1 : col0+n]: R'*QuasiDiagonal(op)*R and R'*D'*D*R.  One
// kernel evaluates values and first derivatives of the basis at every node with a shared Cox-de Boor sweep and builds
// both bands (assemble_stream.cu, mode 2); outside that kernel's range the two matrices are assembled one after the other.
int cb_bspline_assemble_pair(const cb_bspline *B, const double *opvals, int coulomb, double Z,
                             i64 col0, i64 n, double *bandV, double *bandT, void *stream) {
    CB_REQUIRE(B != nullptr && bandV != nullptr && bandT != nullptr, "cb_bspline_assemble_pair: NULL argument");
    CB_REQUIRE(col0 >= 0 && n >= 0 && col0 + n <= B->nf, "cb_bspline_assemble_pair: restriction outside 1..nf");
    CB_REQUIRE(coulomb == 0 || coulomb == 1, "cb_bspline_assemble_pair: coulomb must be 0 or 1");
    if (n == 0) return CB_OK;
    const int k = B->k;
    if (k >= 2) {
        const int src = cb_asm_stream_launch(B, 2, opvals, coulomb, Z, col0, n, col0, n, k - 1, 0, n, 0, B->n_tt - 1, bandV,
                                             (cudaStream_t)stream, nullptr, bandT);
        if (src >= 0) return src;
    }
    int rc = assemble_common(B, B, 0, 0, opvals, coulomb, Z, col0, n, col0, n, k - 1, k - 1, 0, n, 0, B->n_tt - 1, bandV, stream);
    if (rc) return rc;
    return assemble_common(B, B, 1, 1, nullptr, 0, 0.0, col0, n, col0, n, k - 1, k - 1, 0, n, 0, B->n_tt - 1, bandT, stream);
}

// Host-pointer convenience form: the band is assembled on the device and copied into a host
// BandedMatrix' data; opvals_host (numnodes doubles) is uploaded if given.  Synchronous.
int cb_bspline_assemble_host(const cb_bspline *L, const cb_bspline *R, int a, int b, const double *opvals_host,
                             int coulomb, double Z, i64 rowL0, i64 m, i64 colR0, i64 n, int l, int u,
                             double *band_host, void *stream) {
    CB_REQUIRE(R != nullptr && band_host != nullptr, "cb_bspline_assemble_host: NULL argument");
    CB_REQUIRE(n >= 0 && l + u + 1 >= 1, "cb_bspline_assemble_host: bad shape");
    if (n == 0) return CB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    { int prc = cb_keep_scratch_pool(); if (prc) return prc; }
    const size_t bytes = sizeof(double) * (size_t)(l + u + 1) * (size_t)n, nb = sizeof(double) * (size_t)(R->ni * R->q);
    double *d_band = nullptr, *d_op = nullptr;
    CB_CUDA(cudaMallocAsync(&d_band, bytes, st));
    cudaError_t e = cudaSuccess;
    if (opvals_host) {
        e = cudaMallocAsync(&d_op, nb ? nb : 8, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_op, opvals_host, nb, cudaMemcpyHostToDevice, st);
    }
    int rc = CB_OK;
    if (e == cudaSuccess) rc = assemble_common(L, R, a, b, d_op, coulomb, Z, rowL0, m, colR0, n, l, u, 0, n, 0, R->n_tt - 1, d_band, stream);
    if (e == cudaSuccess && rc == CB_OK) e = cudaMemcpyAsync(band_host, d_band, bytes, cudaMemcpyDeviceToHost, st);
    cudaFreeAsync(d_band, st);
    if (d_op) cudaFreeAsync(d_op, st);
    cudaError_t e2 = cudaStreamSynchronize(st);
    if (e == cudaSuccess) e = e2;
    if (e != cudaSuccess) { cb_set_error("cb_bspline_assemble_host: %s", cudaGetErrorString(e)); return CB_ERR_CUDA; }
    return rc;
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
    size_t smem = sizeof(double) * FEA_WARPS * ((size_t)2 * omax * omax + 3 * omax);
    int grid = cb_grid_for(nel, FEA_WARPS, 16);
    auto launch = [&](auto kern) -> int {
        CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, FEA_WARPS * 32, smem, (cudaStream_t)stream>>>(x, wcat, nrm, estart, order, boff, woff,
                                                                   uniform_order, nel, nder, omax, blocks);
        return CB_OK;
    };
    int lrc = CB_OK;
    // uniform orders up to 16: several elements per warp, one lane per row (CB_FEDVR_ASM_WARP=1 keeps the
    // one-warp-per-element kernel for comparisons; the two are bit-identical)
    static const bool per_warp = getenv("CB_FEDVR_ASM_WARP") != nullptr;
    if (uniform_order >= 2 && uniform_order <= 16 && !per_warp) {
        auto launch_rows = [&](auto kern, int per_task) {
            const int g = cb_grid_for((nel + per_task - 1) / per_task, FEG_WARPS, 64);
            kern<<<g, FEG_WARPS * 32, 0, (cudaStream_t)stream>>>(x, wcat, nrm, nel, nder, blocks);
        };
        switch (uniform_order) {
#define CB_FEG(OO) case OO: launch_rows(fedvr_assemble_rows_kernel<OO>, 32 / OO); break;
            CB_FEG(2) CB_FEG(3) CB_FEG(4) CB_FEG(5) CB_FEG(6) CB_FEG(7) CB_FEG(8) CB_FEG(9) CB_FEG(10)
            CB_FEG(11) CB_FEG(12) CB_FEG(13) CB_FEG(14) CB_FEG(15) CB_FEG(16)
#undef CB_FEG
        }
        CB_LAUNCH_CHECK();
        return CB_OK;
    }
    switch (uniform_order) {
#define CB_FEA(OO) case OO: lrc = launch(fedvr_assemble_kernel<OO>); break;
        CB_FEA(2) CB_FEA(3) CB_FEA(4) CB_FEA(5) CB_FEA(6) CB_FEA(7) CB_FEA(8) CB_FEA(9) CB_FEA(10)
        CB_FEA(11) CB_FEA(12) CB_FEA(13) CB_FEA(14) CB_FEA(15) CB_FEA(16)
#undef CB_FEA
        default: lrc = launch(fedvr_assemble_kernel<0>); break;
    }
    if (lrc) return lrc;
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

// derivative_matrix / materialisers between different restrictions (src/fd_derivatives.jl:214-219,244-249,278-289).
int cb_fd_derivative_banded(int scheme, int nder, const double *r, const double *fhalf, i64 N, double step,
                            i64 rowi0, i64 m, i64 colj0, i64 n, double Z, int ell, double *band, void *stream) {
    CB_REQUIRE(scheme >= 0 && scheme <= 2, "cb_fd_derivative_banded: unknown scheme %d", scheme);
    CB_REQUIRE(nder == 1 || nder == 2, "cb_fd_derivative_banded: derivative order must be 1 or 2");
    CB_REQUIRE(N >= 1 && rowi0 >= 0 && m >= 0 && rowi0 + m <= N && colj0 >= 0 && n >= 0 && colj0 + n <= N,
               "cb_fd_derivative_banded: restriction outside 1..N");
    CB_REQUIRE(scheme != 2 || (r != nullptr && N >= 3), "cb_fd_derivative_banded: non-uniform scheme needs the node vector (N >= 3)");
    CB_REQUIRE(fhalf == nullptr, "cb_fd_derivative_banded: the reference has no weighted form between different restrictions (src/fd_derivatives.jl:320-331)");
    CB_REQUIRE(step > 0.0, "cb_fd_derivative_banded: step must be positive");
    if (n == 0) return CB_OK;
    CB_REQUIRE(band != nullptr, "cb_fd_derivative_banded: NULL band");
    // the Coulomb correction of the [1,1] entry is applied only when both restrictions start at the first node (:620-622)
    const double Zc = (rowi0 == 0 && colj0 == 0) ? Z : 0.0;
    fd_band_kernel<<<cb_grid_for(n, 256, 16), 256, 0, (cudaStream_t)stream>>>(scheme, nder, r, fhalf, N, step, rowi0, m, colj0, n, Zc, ell, band);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

// Stencils of the non-uniform compact scheme for nodes i0+1 .. i0+n (src/fd_derivatives.jl:105-143, :410-424).
int cb_fd_implicit_stencils(const double *r, i64 N, i64 i0, i64 n, double *alpha_t, double *alpha, double *beta,
                            double *m_dl, double *m_d, double *m_du, void *stream) {
    CB_REQUIRE(r != nullptr && N >= 2 && i0 >= 0 && n >= 0 && i0 + n <= N, "cb_fd_implicit_stencils: restriction outside 1..N");
    if (n == 0) return CB_OK;
    CB_REQUIRE(beta && m_d && (n == 1 || (alpha_t && alpha && m_dl && m_du)), "cb_fd_implicit_stencils: NULL output");
    fd_implicit_stencil_kernel<<<cb_grid_for(n, 256, 16), 256, 0, (cudaStream_t)stream>>>(r, N, i0, n, alpha_t, alpha, beta, m_dl, m_d, m_du);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

}  // extern "C"
