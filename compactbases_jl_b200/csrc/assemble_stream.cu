// assemble_stream.cu — K3s: streaming band assembly for the symmetric case  <(-1)^m d^m B_i | op | d^m B_j>
// (both sides the same B-spline basis, same derivative order m = 0 or 1: mass, potential, Coulomb, Gram and
// Laplacian matrices — overlap_matrix! src/bsplines.jl:58-93 and diffop! src/bspline_derivatives.jl:3-20 with a == b).
//
// What bounded the tiled kernel of assemble.cu was neither FP64 nor HBM but shared-memory wavefronts and plain
// instruction count (profiles/r1p): every column tile recomputed the K-1 intervals it shares with its neighbour
// (27 % at K = 10), the Cox-de Boor sweep spent 5 FP64 operations per entry, the K x K local matrices were computed
// in full although they are symmetric, and the gather was one thread per band entry with ten predicated steps.
//
// Here a CTA owns a contiguous range of columns and walks through the knot intervals they live on in steps of NS
// intervals, keeping the local matrices of the last 2 NS + K - 1 intervals in a shared-memory ring.  A step has two
// segments (two CTA barriers):
//
//   segment A  every thread is one (interval, node pair): one Cox-de Boor sweep for both nodes on the 2(K-1) knot
//              distances (3 FP64 operations per level entry, reciprocal support widths from the step's table, shared by
//              the interval's q nodes), m derivative levels; the K values per node go to a node-major table (even
//              nodes first: conflict-free 16-byte stores).  The columns gathered one segment earlier leave the SM as
//              one contiguous, fully coalesced copy.
//   segment B  three roles side by side, warp-specialised:
//              * local matrices: two threads per interval, each a row range of the UPPER triangle of
//                M_s[a][b] = sum_l w_l op(x_l) N_a(x_l) N_b(x_l)  in registers; lanes are consecutive intervals, so the
//                16-byte table reads of a warp fall into different banks;
//              * gather: one thread per column completed by the previous steps: its 2K-1 band entries are sums over
//                the column's K intervals of compile-time-indexed ring entries, intervals in increasing order — the
//                same sum whatever the partition into CTAs, ranks or steps;
//              * tables of the NEXT step: knot windows and reciprocal knot differences from knots that were fetched
//                into registers during segment A.
// Nothing on the critical path waits for global memory: knots and interval ordinals of step t+1 are loaded at the
// start of segment A of step t, the nodes / weights of its phase-1 items arrive by cp.async during segment B.  Halo
// work is the K - 1 intervals at the end of a CTA's range (1.3 % at the C3 size).
#include "bspline_core.cuh"
#include "asm_halo.cuh"

struct StreamArgs {
    KnotView kv;
    int q;
    const double *x, *w;          // nodes / weights [ni*q]
    const int32_t *ord;           // ordinal of expanded interval I (entry I-1) among the non-empty ones, or -1
    const double *opvals;         // NULL or op at the nodes
    int coulomb; double Z;        // 1: op(x) = -Z/x on chip; 2: unit quadrature weights (Gram matrix V'V)
    i64 rowL0, mm, colR0, nn;
    int u;
    i64 col_lo, col_hi, s_lo, s_hi;
    i64 cpc;                      // columns per CTA
    int NS;                       // intervals per step
    double *band;                 // column col_lo at band[0]
    double *band2;                // pair mode: the derivative matrix (band gets the potential matrix)
    AsmHalo H;                    // interval-sharded assembly: fused halo exchange over peer memory (all zero: none)
};

constexpr int STREAM_MAX_THREADS = 384;

template <int K> struct StreamGeom {
    static constexpr int QH = (K + 2) / 2;                 // node pairs per interval (q = K + 1 nodes, or K)
    static constexpr int KP2 = (K + 1) / 2 * 2, KH = (K + 1) / 2;
    static constexpr int W = 2 * K - 1;
    static constexpr int NE = K * (K + 1) / 2, MSP = NE | 1;
    static constexpr int PIV0 = 2 * QH * KP2;
    static constexpr int PIV = PIV0 + ((2 - PIV0 % 4) + 4) % 4;   // = 2 (mod 4): lanes of consecutive intervals hit different banks
    static constexpr int QPW = (2 * QH) | 1;
    // table row of node 2 lp / 2 lp + 1 of an interval.  Any one-to-one map onto the rows 0 .. q-1 is valid (the local
    // matrices sum over the rows in storage order); for K = 10 the rows are permuted so that the 16-byte table stores of
    // a quarter-warp (6 node pairs of one interval + 2 of the next) fall into different banks
    // (scripts: exhaustive search, 128 instead of 192 wavefronts per 384 threads and store).
    __host__ __device__ static constexpr int row_even(int lp) {
        if (K == 10) return (lp & 1) ? 6 - (lp >> 1) : 2 - (lp >> 1);          // 2 6 1 5 0 4
        return lp;
    }
    __host__ __device__ static constexpr int row_odd(int lp) {
        if (K == 10) return lp == 0 ? 3 : (lp >= 4 ? 3 + lp : 7 + lp);          // 3 8 9 10 7 (8: unused pad pair)
        return QH + lp;
    }
    // upper-triangle index of (a, b), a <= b
    __host__ __device__ static constexpr int sidx(int a, int b) { return a <= b ? a * K - a * (a - 1) / 2 + (b - a) : b * K - b * (b - 1) / 2 + (a - b); }
};

struct StreamTab {                 // knot / reciprocal tables of a step, two copies each (see AsmTab in assemble.cu)
    const double *sm;
    int kn, rd, pitch, kn_copy, rd_copy;
};

// One Cox-de Boor sweep for the nodes x0, x1 of interval slot sl: N[r] = M-th derivative of the r-th live function.
// Works on the distances dh[r] = t_{I+1+r} - x, dl[i] = x - t_{I-K+1+i}; a value level entry costs one DMUL and two
// DFMA/DMUL (the reciprocal support widths come from the step table, shared by the interval's q nodes).
template <int K, int M>
__device__ __forceinline__ void stream_sweep_pair(const StreamTab &T, int sl, double x0, double x1,
                                                  double (&N0)[K], double (&N1)[K]) {
    double dh0[K], dh1[K], dl0[K], dl1[K];
    {
        const double2 *k2 = reinterpret_cast<const double2 *>(T.sm + T.kn + (sl & 1) * T.kn_copy + (sl & ~1));
#pragma unroll
        for (int t = 0; t < K; ++t) {
            const double2 v = k2[t];
            if (2 * t < K) { dl0[2 * t] = x0 - v.x; dl1[2 * t] = x1 - v.x; }
            else { dh0[2 * t - K] = v.x - x0; dh1[2 * t - K] = v.x - x1; }
            if (2 * t + 1 < K) { dl0[2 * t + 1] = x0 - v.y; dl1[2 * t + 1] = x1 - v.y; }
            else { dh0[2 * t + 1 - K] = v.y - x0; dh1[2 * t + 1 - K] = v.y - x1; }
        }
    }
#pragma unroll
    for (int r = 0; r < K; ++r) { N0[r] = 0.0; N1[r] = 0.0; }
    if (M >= K) return;
    N0[0] = 1.0; N1[0] = 1.0;
    constexpr int jmax = K - 1 - M;
#pragma unroll
    for (int j = 1; j <= K - 1; ++j) {
        const int p0 = sl + K - j;
        const double2 *r2 = reinterpret_cast<const double2 *>(T.sm + T.rd + (p0 & 1) * T.rd_copy + j * T.pitch + (p0 & ~1));
        double rc[K];
#pragma unroll
        for (int h = 0; h < (j + 1) / 2; ++h) { const double2 v = r2[h]; rc[2 * h] = v.x; if (2 * h + 1 < K) rc[2 * h + 1] = v.y; }
        if (j <= jmax) {
            double s0 = 0.0, s1 = 0.0;
#pragma unroll
            for (int r = 0; r < j; ++r) {
                const double t0 = __dmul_rn(N0[r], rc[r]), t1 = __dmul_rn(N1[r], rc[r]);
                N0[r] = fma(dh0[r], t0, s0); N1[r] = fma(dh1[r], t1, s1);
                s0 = __dmul_rn(dl0[K - j + r], t0); s1 = __dmul_rn(dl1[K - j + r], t1);
            }
            N0[j] = s0; N1[j] = s1;
        } else {
            double p0_ = 0.0, p1_ = 0.0;
#pragma unroll
            for (int r = 0; r <= j; ++r) {
                double q0 = 0.0, q1 = 0.0;
                // (rounded products: the same numbers as the value + derivative sweep below, whatever the compiler contracts)
                if (r < j) { q0 = __dmul_rn(N0[r], rc[r]); q1 = __dmul_rn(N1[r], rc[r]); }
                N0[r] = (double)j * __dsub_rn(p0_, q0); N1[r] = (double)j * __dsub_rn(p1_, q1);
                p0_ = q0; p1_ = q1;
            }
        }
    }
}

// Values AND first derivatives of the K live functions at the two nodes from one sweep (the Hamiltonian pair
// B'VB, B'D'DB): the levels 1 .. K-2 are shared, the last level gives the values (Cox-de Boor step) and the
// derivatives ((K-1) * difference of the scaled order-(K-1) functions) from the same products t_r = N_r / width_r.
// Same operations as stream_sweep_pair<K, 0> / <K, 1>, hence bit-identical tables.  Rows are stored as they are formed.
template <int K>
__device__ __forceinline__ void stream_sweep_pair_vd(const StreamTab &T, int sl, double x0, double x1, bool has1,
                                                     double2 *ve, double2 *vo, double2 *de, double2 *dd) {
    double dh0[K], dh1[K], dl0[K], dl1[K];
    {
        const double2 *k2 = reinterpret_cast<const double2 *>(T.sm + T.kn + (sl & 1) * T.kn_copy + (sl & ~1));
#pragma unroll
        for (int t = 0; t < K; ++t) {
            const double2 v = k2[t];
            if (2 * t < K) { dl0[2 * t] = x0 - v.x; dl1[2 * t] = x1 - v.x; }
            else { dh0[2 * t - K] = v.x - x0; dh1[2 * t - K] = v.x - x1; }
            if (2 * t + 1 < K) { dl0[2 * t + 1] = x0 - v.y; dl1[2 * t + 1] = x1 - v.y; }
            else { dh0[2 * t + 1 - K] = v.y - x0; dh1[2 * t + 1 - K] = v.y - x1; }
        }
    }
    double N0[K], N1[K];
#pragma unroll
    for (int r = 0; r < K; ++r) { N0[r] = 0.0; N1[r] = 0.0; }
    N0[0] = 1.0; N1[0] = 1.0;
#pragma unroll
    for (int j = 1; j <= K - 2; ++j) {
        const int p0 = sl + K - j;
        const double2 *r2 = reinterpret_cast<const double2 *>(T.sm + T.rd + (p0 & 1) * T.rd_copy + j * T.pitch + (p0 & ~1));
        double rc[K];
#pragma unroll
        for (int h = 0; h < (j + 1) / 2; ++h) { const double2 v = r2[h]; rc[2 * h] = v.x; if (2 * h + 1 < K) rc[2 * h + 1] = v.y; }
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int r = 0; r < j; ++r) {
            const double t0 = N0[r] * rc[r], t1 = N1[r] * rc[r];
            N0[r] = fma(dh0[r], t0, s0); N1[r] = fma(dh1[r], t1, s1);
            s0 = dl0[K - j + r] * t0; s1 = dl1[K - j + r] * t1;
        }
        N0[j] = s0; N1[j] = s1;
    }
    // last level, j = K - 1: t[r] = N[r] / width, r < K - 1
    double t0[K + 1], t1[K + 1];                    // t[r + 1] holds t_r; t[0] = t_{-1} = 0, t[K] = t_{K-1} = 0
    {
        constexpr int j = K - 1;
        const int p0 = sl + K - j;
        const double2 *r2 = reinterpret_cast<const double2 *>(T.sm + T.rd + (p0 & 1) * T.rd_copy + j * T.pitch + (p0 & ~1));
        double rc[K];
#pragma unroll
        for (int h = 0; h < (j + 1) / 2; ++h) { const double2 v = r2[h]; rc[2 * h] = v.x; if (2 * h + 1 < K) rc[2 * h + 1] = v.y; }
        t0[0] = 0.0; t1[0] = 0.0; t0[K] = 0.0; t1[K] = 0.0;
#pragma unroll
        for (int r = 0; r < K - 1; ++r) { t0[r + 1] = __dmul_rn(N0[r], rc[r]); t1[r + 1] = __dmul_rn(N1[r], rc[r]); }
    }
    auto val = [&](const double (&dh)[K], const double (&dl)[K], const double (&t)[K + 1], int r) -> double {
        // N_r = dh[r] t_r + dl[r] t_{r-1}   (the sweep's  fma(dh[r], t_r, s_r),  s_r = dl[r] * t_{r-1};  N_{K-1} = s_{K-1})
        if (r >= K) return 0.0;
        const double s = r >= 1 ? __dmul_rn(dl[r], t[r]) : 0.0;
        return r < K - 1 ? fma(dh[r], t[r + 1], s) : s;
    };
    auto der = [&](const double (&t)[K + 1], int r) -> double {
        return r < K ? (double)(K - 1) * __dsub_rn(t[r], t[r + 1]) : 0.0;
    };
#pragma unroll
    for (int a = 0; a < (K + 1) / 2; ++a) {
        ve[a] = make_double2(val(dh0, dl0, t0, 2 * a), val(dh0, dl0, t0, 2 * a + 1));
        de[a] = make_double2(der(t0, 2 * a), der(t0, 2 * a + 1));
        if (has1) {
            vo[a] = make_double2(val(dh1, dl1, t1, 2 * a), val(dh1, dl1, t1, 2 * a + 1));
            dd[a] = make_double2(der(t1, 2 * a), der(t1, 2 * a + 1));
        }
    }
}

// Row split of the upper triangle between the two threads of an interval in phase 2a: rows [0, AS) hold about half
// of the K(K+1)/2 entries.
template <int K> struct StreamSplit {
    static constexpr int as() {
        int best = 1, bestd = 1 << 30;
        for (int s = 1; s < K; ++s) {
            int e = 0;
            for (int a = 0; a < s; ++a) e += K - a;
            int d = 2 * e - K * (K + 1) / 2;
            d = d < 0 ? -d : d;
            if (d < bestd) { bestd = d; best = s; }
        }
        return best;
    }
    static constexpr int AS = as();
};

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// phase 2a for the rows [A0, A1) of the upper triangle of one interval's local matrix
template <int K, int A0, int A1>
__device__ __forceinline__ void stream_local_rows(const double *tv, const double *wl, int q, double *m) {
    using G = StreamGeom<K>;
    constexpr int KP2 = G::KP2, KH = G::KH;
    constexpr int E0 = G::sidx(A0, A0), E1 = A1 < K ? G::sidx(A1, A1) : G::NE;
    double acc[E1 - E0];
#pragma unroll
    for (int i = 0; i < E1 - E0; ++i) acc[i] = 0.0;
#pragma unroll 2
    for (int p = 0; p < q; ++p) {
        double n[KP2];
        const double2 *t2 = reinterpret_cast<const double2 *>(tv + p * KP2);
#pragma unroll
        for (int a = A0 / 2; a < KH; ++a) { const double2 v = t2[a]; n[2 * a] = v.x; n[2 * a + 1] = v.y; }
        const double wv = wl[p];
#pragma unroll
        for (int a = A0; a < A1; ++a) {
            const double lw = wv * n[a];
#pragma unroll
            for (int b = a; b < K; ++b) acc[G::sidx(a, b) - E0] = fma(lw, n[b], acc[G::sidx(a, b) - E0]);
        }
    }
#pragma unroll
    for (int i = 0; i < E1 - E0; ++i) m[E0 + i] = acc[i];
}

// The completed columns [pj, pj + pn) leave the SM.  Plain case: one contiguous, fully coalesced copy.  With a fused halo
// (asm_halo.cuh) the first n_push columns of the call go to the left neighbour's mailbox instead, and the last n_recv
// columns first add what the right neighbour pushed; the flags are raised / awaited by this CTA alone (the host makes
// steps long enough for each group to sit in one emission of one CTA).  All conditions are CTA-uniform.
template <int W>
__device__ __forceinline__ void stream_copy_out(const StreamArgs &A, const double *OUT, i64 pj, int pn, int tid, int NT) {
    const AsmHalo &H = A.H;
    const bool psh = H.n_push > 0 && pj < A.col_lo + H.n_push;
    const bool rcv = H.n_recv > 0 && pj + pn > A.col_hi - H.n_recv;
    double *dst = A.band + (size_t)(pj - A.col_lo) * W;
    if (!psh && !rcv) {
        for (int e = tid; e < pn * W; e += NT) dst[e] = OUT[e];
        return;
    }
    if (tid == 0) {
        if (rcv) halo_spin(H.recv_flag, H.epoch, H.status);
        if (psh && H.epoch > 2) halo_spin(H.ack_in, H.epoch - 2, H.status);      // the buffer of this parity has been consumed
    }
    __syncthreads();
    const i64 push_end = A.col_lo + H.n_push, recv_beg = A.col_hi - H.n_recv;
    for (int e = tid; e < pn * W; e += NT) {
        const int jj = e / W;
        const i64 j = pj + jj;
        double v = OUT[e];
        if (H.n_push > 0 && j < push_end) {
            H.push_cols[(size_t)(j - A.col_lo) * W + (e - jj * W)] = v;
        } else {
            if (H.n_recv > 0 && j >= recv_beg) v += __ldcv(H.recv_cols + (size_t)(j - recv_beg) * W + (e - jj * W));
            dst[e] = v;
        }
    }
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
        if (psh && pj + pn >= push_end) st_release_sys(H.push_flag, H.epoch);     // every pushed column has left
        if (rcv && pj + pn >= A.col_hi) st_release_sys(H.ack_out, H.epoch);       // mailbox consumed
    }
}

#ifdef STREAM_PROF
__device__ unsigned long long stream_prof[8];      // cycles of thread 0 of every CTA: [0] segment A, [1] segment B, [2] steps
#define SPROF(i) do { if (threadIdx.x == 0) { const long long t__ = clock64(); atomicAdd(&stream_prof[i], (unsigned long long)(t__ - tprof)); tprof = t__; } } while (0)
extern "C" int cb_debug_stream_prof(unsigned long long *out, int reset) {
    cudaMemcpyFromSymbol(out, stream_prof, sizeof(unsigned long long) * 8);
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(stream_prof, z, sizeof(z)); }
    return 0;
}
#else
#define SPROF(i) do { } while (0)
#endif

template <int K, int M>
__global__ void __launch_bounds__(STREAM_MAX_THREADS, 1)
bspline_assemble_stream_kernel(StreamArgs A) {
    extern __shared__ __align__(16) double sm[];
#ifdef STREAM_PROF
    long long tprof = clock64();
#endif
    using G = StreamGeom<K>;
    constexpr int QH = G::QH, KP2 = G::KP2, KH = G::KH, W = G::W, NE = G::NE, MSP = G::MSP, PIV = G::PIV, QPW = G::QPW;
    constexpr int AS = StreamSplit<K>::AS;
    constexpr int NM = M == 2 ? 2 : 1;                     // M == 2: the pair (values | first derivatives) from one sweep
    const int NT = blockDim.x, tid = threadIdx.x, warp = tid >> 5;
    const int q = A.q, NS = A.NS, NSR = (NS + 31) & ~31, RING = 2 * NS + K - 1, CAP = NS + K - 1;
    const int NKP = (NS + 2 * K + 2) / 2 * 2;
    const int ml = A.kv.ml;
    const i64 n_tt = A.kv.n_tt;
    // roles of the second segment of a step: warps [0, W2B) local matrices, [W2B, WP0) gather, [WP0, ..) tables
    const int W2A = 2 * NSR / 32, W2B = NM * W2A, GW = (CAP + 31) / 32, WP0 = W2B + NM * GW;
    const int P0T = NT - 32 * WP0;
    // shared memory carve-up (every region starts on a 16-byte boundary)
    const size_t TVS = (size_t)NS * PIV, WLS = ((size_t)NS * QPW + 1) & ~(size_t)1, MSS = ((size_t)RING * MSP + 1) & ~(size_t)1,
                 OUTS = ((size_t)CAP * W + 1) & ~(size_t)1;
    double *TV = sm;                                         // [NM][NS][PIV]   node-major tables, even nodes first
    double *WL = TV + NM * TVS;                              // [NM][NS][QPW]   (-1)^m w op at the node positions
    double *MS = WL + NM * WLS;                              // [NM][RING][MSP] upper triangles of the local matrices
    double *OUT = MS + NM * MSS;                             // [NM][CAP][W]    staging tiles of completed columns
    double *KN = OUT + NM * OUTS;                            // [2][NKP + 2]
    double *RD = KN + 2 * (NKP + 2);                         // [2][K * NKP + 2]
    double *RAW = RD + (size_t)2 * (K * NKP + 2);            // [NKP] knots of the next step
    double *NODE = RAW + NKP;                                // [NT][6] x0 x1 w0 w1 o0 o1 of the next step's phase-1 items
    int *sOrd = reinterpret_cast<int *>(NODE + (size_t)NT * 6);   // [2][NS]
    StreamTab T;
    T.sm = sm; T.kn = (int)(KN - sm); T.rd = (int)(RD - sm); T.pitch = NKP;
    T.kn_copy = ((NKP / 2) | 1) * 2; T.rd_copy = ((K * NKP / 2) | 1) * 2;
    const double sgn = (M & 1) ? -1.0 : 1.0;               // src/bspline_derivatives.jl:14

    const i64 j0 = A.col_lo + (i64)blockIdx.x * A.cpc;
    if (j0 >= A.col_hi) return;
    const i64 j1 = j0 + A.cpc < A.col_hi ? j0 + A.cpc : A.col_hi;
    // column j is function fj = colR0 + 1 + j, which lives on the t.t intervals f0 .. f0 + K - 1, f0 = fj - ml
    const i64 foff = A.colR0 + 1 - ml;
    i64 sA = j0 + foff, sB = (j1 - 1) + foff + K - 1;
    if (sA < 0) sA = 0;
    if (sA < A.s_lo) sA = A.s_lo;
    if (sB > n_tt - 2) sB = n_tt - 2;
    if (sB > A.s_hi - 1) sB = A.s_hi - 1;
    const i64 ntot = sB >= sA ? sB - sA + 1 : 0;
    const int nsteps = (int)((ntot + NS - 1) / NS);
    auto step_ns = [&](int t) -> int {
        if (t >= nsteps) return 0;
        const i64 left = sB - (sA + (i64)t * NS) + 1;
        return (int)(left < NS ? left : NS);
    };
    // knots / interval ordinals of step t: global -> shared (prologue), or global -> registers -> shared (steady state)
    auto raw_index = [&](int t, int i) -> i64 { return sA + (i64)t * NS + ml - K + i; };
    // tables of step t from RAW (threads first, first + stride, ...)
    auto build_tables = [&](int ns, int first, int stride) {
        const int NP = ns + 2 * K;
        const int per = stride / K;                         // threads per level (the host gives the role >= K threads)
        const int j = first / per;
        if (j >= K) return;
        for (int p = first - j * per; p < NP; p += per) {
            const double lo = RAW[p];
            if (j == 0) {
                KN[p] = lo;
                if (p > 0) KN[T.kn_copy + p - 1] = lo;
            } else if (p + j < NP) {
                const double dk = RAW[p + j] - lo;
                const double v = dk != 0.0 ? fast_rcp(dk) : 0.0;   // zero widths only feed empty intervals / absent functions
                RD[j * NKP + p] = v;
                if (p > 0) RD[T.rd_copy + j * NKP + p - 1] = v;
            }
        }
    };
    // this thread's phase-1 item
    const int sl1 = tid / QH, lp1 = tid - sl1 * QH;
    const bool has1 = 2 * lp1 + 1 < q;
    double *node = NODE + (size_t)tid * 6;
    auto prefetch_nodes = [&](int ns, const int *ord) {       // node data of the coming step: global -> NODE (async)
        if (tid < ns * QH) {
            const int c = ord[sl1];
            if (c >= 0) {
                const i64 g0 = (i64)c * q + 2 * lp1, g1 = has1 ? g0 + 1 : g0;
                cp_async8(node + 0, A.x + g0); cp_async8(node + 1, A.x + g1);
                if (A.coulomb != 2) { cp_async8(node + 2, A.w + g0); cp_async8(node + 3, A.w + g1); }
                if (A.opvals) { cp_async8(node + 4, A.opvals + g0); cp_async8(node + 5, A.opvals + g1); }
            }
        }
        cp_async_commit();
    };

    // ---- prologue: tables and node data of step 0 ----
    {
        const int ns = step_ns(0);
        if (ns > 0) {
            for (int i = tid; i < ns + 2 * K; i += NT) RAW[i] = A.kv.at(raw_index(0, i));
            for (int e = tid; e < ns; e += NT) sOrd[e] = A.ord[sA + e + ml - 1];
        }
        __syncthreads();
        if (ns > 0) build_tables(ns, tid, NT);
        prefetch_nodes(ns, sOrd);
        __syncthreads();
    }
    i64 jnext = j0, pend_j = j0;
    int pend_n = 0;

    for (int t = 0;; ++t) {
        const int ns = step_ns(t), ns_next = step_ns(t + 1);
        const i64 s0 = sA + (i64)t * NS;
        const int *ord_cur = sOrd + (t & 1) * NS;
        int *ord_next = sOrd + ((t + 1) & 1) * NS;
        // ======== segment A: node tables of step t; the columns gathered in the previous segment leave the SM ========
        double raw_next = 0.0; int ord_n = -1;
        if (tid < ns_next + 2 * K) raw_next = A.kv.at(raw_index(t + 1, tid));          // lands while phase 1 runs
        if (tid < ns_next) ord_n = A.ord[s0 + NS + tid + ml - 1];
        if (pend_n > 0) {
            stream_copy_out<W>(A, OUT, pend_j, pend_n, tid, NT);
            if (NM == 2) {
                double *dst = A.band2 + (size_t)(pend_j - A.col_lo) * W;
                for (int e = tid; e < pend_n * W; e += NT) dst[e] = OUT[OUTS + e];
            }
            pend_n = 0;
        }
        cp_async_wait_all();
        if (tid < ns * QH && ord_cur[sl1] >= 0) {
            const double x0 = node[0], x1 = node[1];
            double w0 = 1.0, w1 = 1.0, o0 = 1.0, o1 = 1.0;
            if (A.coulomb != 2) { w0 = node[2]; w1 = node[3]; }
            if (A.opvals) { o0 = node[4]; o1 = node[5]; }
            else if (A.coulomb == 1) { o0 = -A.Z * fast_rcp(x0); o1 = -A.Z * fast_rcp(x1); }
            const int re = G::row_even(lp1), ro = G::row_odd(lp1);
            double2 *te = reinterpret_cast<double2 *>(TV + (size_t)sl1 * PIV + re * KP2);
            double2 *to = reinterpret_cast<double2 *>(TV + (size_t)sl1 * PIV + ro * KP2);
            if (NM == 2) {
                stream_sweep_pair_vd<K>(T, sl1, x0, x1, has1, te, to, te + TVS / 2, to + TVS / 2);
                WL[sl1 * QPW + re] = w0 * o0;                       // potential term
                WL[WLS + sl1 * QPW + re] = -w0;                     // (-1)^1 <B'|B'>, src/bspline_derivatives.jl:14
                if (has1) { WL[sl1 * QPW + ro] = w1 * o1; WL[WLS + sl1 * QPW + ro] = -w1; }
            } else {
                double N0[K], N1[K];
                stream_sweep_pair<K, (M == 2 ? 0 : M)>(T, sl1, x0, x1, N0, N1);
#pragma unroll
                for (int a = 0; a < KH; ++a) {
                    te[a] = make_double2(N0[2 * a], 2 * a + 1 < K ? N0[2 * a + 1] : 0.0);
                    if (has1) to[a] = make_double2(N1[2 * a], 2 * a + 1 < K ? N1[2 * a + 1] : 0.0);
                }
                WL[sl1 * QPW + re] = (sgn * w0) * o0;
                if (has1) WL[sl1 * QPW + ro] = (sgn * w1) * o1;
            }
        }
        if (tid < ns_next + 2 * K) RAW[tid] = raw_next;
        if (tid < ns_next) ord_next[tid] = ord_n;
        SPROF(3);                                       // thread 0's own work in segment A
        __syncthreads();
        SPROF(0);                                       // + waiting for the slowest warp
        // ======== segment B: three roles side by side ========
        prefetch_nodes(ns_next, ord_next);
        if (warp < W2B) {
            // ---- local matrices of step t: thread (matrix, row half, interval) ----
            const int mi = NM == 2 ? warp / W2A : 0;
            const int tm = tid - mi * 2 * NSR;
            const int h = tm >= NSR ? 1 : 0, sl = tm - h * NSR;
            if (sl < ns) {
                int slot = (int)((s0 + sl - sA) % RING);
                double *m = MS + mi * MSS + (size_t)slot * MSP;
                if (ord_cur[sl] < 0) {
                    if (h == 0) { for (int i = 0; i < G::sidx(AS, AS); ++i) m[i] = 0.0; }
                    else { for (int i = G::sidx(AS, AS); i < NE; ++i) m[i] = 0.0; }
                } else {
                    const double *tv = TV + mi * TVS + (size_t)sl * PIV;
                    const double *wl = WL + mi * WLS + sl * QPW;
                    // node positions: even nodes 0..QH-1, odd nodes QH..q-1 (contiguous)
                    if (h == 0) stream_local_rows<K, 0, AS>(tv, wl, q, m);
                    else stream_local_rows<K, AS, K>(tv, wl, q, m);
                }
            }
        } else if (warp < WP0) {
            // ---- gather: columns completed by the steps before t (everything that is left once all steps are done) ----
            const i64 s_done = t == 0 ? sA - 1 : (sA + (i64)t * NS - 1 < sB ? sA + (i64)t * NS - 1 : sB);
            i64 jl = t >= nsteps ? j1 : s_done - (K - 1) - foff + 1;      // exclusive: columns with f0 + K - 1 <= s_done
            if (jl > j1) jl = j1;
            if (jl < jnext) jl = jnext;
            const int nem = (int)(jl - jnext < CAP ? jl - jnext : CAP);
            const int gi = NM == 2 ? (warp - W2B) / GW : 0;        // which matrix this gather warp serves
            const int e2 = tid - 32 * (W2B + gi * GW);
            const double *MSg = MS + gi * MSS;
            double *OUTg = OUT + gi * OUTS;
            if (e2 < nem) {
                const i64 j = jnext + e2, f0 = j + foff;
                double acc[W];
#pragma unroll
                for (int d = 0; d < W; ++d) acc[d] = 0.0;
                int slot = (int)((f0 - sA) % RING);
                if (slot < 0) slot += RING;
#pragma unroll
                for (int r = 0; r < K; ++r) {
                    const i64 s = f0 + r;
                    if (s >= sA && s <= s_done) {
                        const double *m = MSg + (size_t)slot * MSP;
#pragma unroll
                        for (int a = 0; a < K; ++a) acc[a + r] += m[G::sidx(a, K - 1 - r)];
                    }
                    slot = slot + 1 == RING ? 0 : slot + 1;
                }
                // band slot d of column j holds row i = j - u + d; with (l, u) = (K-1+ij, K-1-ij) slot d is run entry d
                const i64 i0 = j - A.u;
                if (i0 >= 0 && i0 + W <= A.mm) {
#pragma unroll
                    for (int d = 0; d < W; ++d) OUTg[e2 * W + d] = acc[d];
                } else {
#pragma unroll
                    for (int d = 0; d < W; ++d) OUTg[e2 * W + d] = (i0 + d >= 0 && i0 + d < A.mm) ? acc[d] : 0.0;
                }
            }
        } else if (ns_next > 0) {
            build_tables(ns_next, tid - 32 * WP0, P0T);
        }
        {   // CTA-uniform bookkeeping of the gather (every thread computes the same numbers)
            const i64 s_done = t == 0 ? sA - 1 : (sA + (i64)t * NS - 1 < sB ? sA + (i64)t * NS - 1 : sB);
            i64 jl = t >= nsteps ? j1 : s_done - (K - 1) - foff + 1;
            if (jl > j1) jl = j1;
            if (jl < jnext) jl = jnext;
            const int nem = (int)(jl - jnext < CAP ? jl - jnext : CAP);
            pend_j = jnext; pend_n = nem; jnext += nem;
        }
        SPROF(4);
        __syncthreads();
        SPROF(1);
#ifdef STREAM_PROF
        if (threadIdx.x == 0) atomicAdd(&stream_prof[2], 1ull);
#endif
        if (t >= nsteps && jnext >= j1) break;
    }
    if (pend_n > 0) {
        stream_copy_out<W>(A, OUT, pend_j, pend_n, tid, NT);
        if (NM == 2) {
            double *dst = A.band2 + (size_t)(pend_j - A.col_lo) * W;
            for (int e = tid; e < pend_n * W; e += NT) dst[e] = OUT[OUTS + e];
        }
    }
}

static int stream_threads(int K, int NS, int NM = 1) {
    const int QH = (K + 2) / 2, NSR = (NS + 31) & ~31, CAP = NS + K - 1;
    int nt = NS * QH;                                       // phase 1: one thread per (interval, node pair)
    const int roles = NM * (2 * NSR + (CAP + 31) / 32 * 32) + 32;  // segment B: local matrices, gather, at least one table warp (>= K threads)
    if (roles > nt) nt = roles;
    if (NS + 2 * K > nt) nt = NS + 2 * K;                   // knot prefetch: one knot per thread
    return (nt + 31) / 32 * 32;
}

static size_t stream_smem_bytes(int K, int NS, int NM = 1) {
    const int QH = (K + 2) / 2, KP2 = (K + 1) / 2 * 2, W = 2 * K - 1, NE = K * (K + 1) / 2, MSP = NE | 1;
    const int PIV0 = 2 * QH * KP2, PIV = PIV0 + ((2 - PIV0 % 4) + 4) % 4, QPW = (2 * QH) | 1;
    const int RING = 2 * NS + K - 1, CAP = NS + K - 1, NKP = (NS + 2 * K + 2) / 2 * 2;
    size_t d = NM * ((size_t)NS * PIV + (((size_t)NS * QPW + 1) & ~(size_t)1) + (((size_t)RING * MSP + 1) & ~(size_t)1) +
                     (((size_t)CAP * W + 1) & ~(size_t)1)) + 2 * (size_t)(NKP + 2) + 2 * (size_t)(K * NKP + 2) + (size_t)NKP +
               (size_t)stream_threads(K, NS, NM) * 6;
    return d * sizeof(double) + sizeof(int) * 2 * (size_t)NS + 16;
}

static int stream_ns_override() {
    static int v = -2;
    if (v == -2) {
        const char *e = getenv("CB_ASM_STREAM_NS");          // tuning knob: intervals per step (0 disables the kernel)
        v = e ? atoi(e) : -1;
    }
    return v;
}

// Returns CB_OK after launching, a CB_ERR_* code on failure, or -1 when the request is outside this kernel's range
// (the caller then takes the tiled / generic kernels of assemble.cu).
// m = 0 / 1: one matrix of derivative order m; m = 2: the pair (band <- potential matrix with op, band2 <- derivative
// matrix (-1) <B'|B'>) from one sweep.
int cb_asm_stream_launch(const cb_bspline *B, int m, const double *opvals, int coulomb, double Z,
                         i64 rowL0, i64 mm, i64 colR0, i64 nn, int u,
                         i64 col_lo, i64 col_hi, i64 s_lo, i64 s_hi, double *band, cudaStream_t st, const AsmHalo *halo,
                         double *band2 = nullptr) {
    const int K = B->k;
    const int NM = m == 2 ? 2 : 1;
    if (K < 2 || K > 11 || m < 0 || m > 2 || (m == 1 && m >= K)) return -1;
    if ((B->q + 1) / 2 != (K + 2) / 2) return -1;
    if (m == 2 && (halo || !band2)) return -1;
    // The pair kernel holds two rings and two tables, which halves its step (NS = 32 at K = 10): at the C3 size it is
    // slower than two launches with full steps (219 vs 182 us), on short column ranges (launch-bound: C1, 17 vs 26 us) it wins.
    if (m == 2 && (col_hi - col_lo) > (i64)CB_NUM_SMS * 512) return -1;
    int NS = 64;
    const int ov = stream_ns_override();
    if (ov == 0) return -1;
    if (ov > 0) NS = ov;
    if (NS < K) NS = K;
    auto fits = [&](int ns) { return stream_threads(K, ns, NM) <= STREAM_MAX_THREADS && stream_smem_bytes(K, ns, NM) <= 220 * 1024; };
    while (NS > K && !fits(NS)) --NS;
    if (!fits(NS)) return -1;
    if (col_hi <= col_lo) return CB_OK;
    const int nt = stream_threads(K, NS, NM);
    // columns without support in the interval range [s_lo, s_hi) are zero; the kernel gets the others
    // (an interval-sharded call passes exactly the columns that touch its range: nothing to clip)
    if (!halo) {
        const int W = 2 * K - 1;
        const i64 foff = colR0 + 1 - B->ml;
        i64 jf = s_lo - (K - 1) - foff, jl = s_hi - foff;
        if (jf < col_lo) jf = col_lo;
        if (jf > col_hi) jf = col_hi;
        if (jl > col_hi) jl = col_hi;
        if (jl < jf) jl = jf;
        for (int b = 0; b < NM; ++b) {
            double *bb = b == 0 ? band : band2;
            if (jf > col_lo) CB_CUDA(cudaMemsetAsync(bb, 0, sizeof(double) * (size_t)W * (size_t)(jf - col_lo), st));
            if (jl < col_hi) CB_CUDA(cudaMemsetAsync(bb + (size_t)W * (size_t)(jl - col_lo), 0, sizeof(double) * (size_t)W * (size_t)(col_hi - jl), st));
        }
        band += (size_t)W * (size_t)(jf - col_lo);
        if (band2) band2 += (size_t)W * (size_t)(jf - col_lo);
        col_lo = jf; col_hi = jl;
        if (col_hi <= col_lo) return CB_OK;
    }
    const i64 ncol = col_hi - col_lo;
    const size_t smem = stream_smem_bytes(K, NS, NM);
    i64 ncta = (ncol + NS - 1) / NS;
    StreamArgs A;
    A.kv = make_knot_view(B);
    A.q = B->q; A.x = B->d_x; A.w = B->d_w; A.ord = B->d_ord;
    A.opvals = opvals; A.coulomb = coulomb; A.Z = Z;
    A.rowL0 = rowL0; A.mm = mm; A.colR0 = colR0; A.nn = nn; A.u = u;
    A.col_lo = col_lo; A.col_hi = col_hi; A.s_lo = s_lo; A.s_hi = s_hi;
    A.NS = NS;
    A.band = band;
    A.band2 = band2;
    memset(&A.H, 0, sizeof(A.H));
    if (halo) {
        A.H = *halo;
        // the pushed columns must leave with the first emission of the first CTA, the receiving ones arrive with the
        // last emission of the last CTA: both hold when a step covers at least 2K - 1 columns and CTAs are longer than that
        if (NS < 2 * K || ncol < 4 * (i64)K) return -1;
    }
    auto launch = [&](auto kern) -> int {
        CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 1;
        CB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, nt, smem));
        if (per_sm < 1) per_sm = 1;
        if (ncta > (i64)CB_NUM_SMS * per_sm) ncta = (i64)CB_NUM_SMS * per_sm;
        A.cpc = (ncol + ncta - 1) / ncta;
        kern<<<(int)ncta, nt, smem, st>>>(A);
        return CB_OK;
    };
    int lrc = -1;
#define CB_STREAM_CASE(KT)                                                                                       \
    case KT:                                                                                                     \
        lrc = m == 0 ? launch(bspline_assemble_stream_kernel<KT, 0>)                                             \
            : m == 1 ? launch(bspline_assemble_stream_kernel<KT, 1>) : launch(bspline_assemble_stream_kernel<KT, 2>); \
        break;
    switch (K) {
        CB_STREAM_CASE(2) CB_STREAM_CASE(3) CB_STREAM_CASE(4) CB_STREAM_CASE(5) CB_STREAM_CASE(6)
        CB_STREAM_CASE(7) CB_STREAM_CASE(8) CB_STREAM_CASE(9) CB_STREAM_CASE(10) CB_STREAM_CASE(11)
        default: return -1;
    }
#undef CB_STREAM_CASE
    if (lrc) return lrc;
    CB_LAUNCH_CHECK();
    return CB_OK;
}
