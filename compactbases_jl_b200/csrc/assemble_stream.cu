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
// intervals, keeping the local matrices of the last NS + K - 1 intervals in a shared-memory ring:
//
//   phase 0  knot window and reciprocal knot differences of the step (one table per step, shared by all nodes)
//   phase 1  one thread per (interval, node pair): one Cox-de Boor sweep for both nodes on the 2(K-1) knot
//            distances (3 FP64 operations per entry), m derivative levels; the K values per node go to a
//            node-major table (even nodes first: conflict-free 16-byte stores)
//   phase 2a two threads per interval (one per half of the nodes): the K(K+1)/2 entries of the UPPER triangle of the
//            interval's local matrix  M_s[a][b] = sum_l w_l op(x_l) N_a(x_l) N_b(x_l)  in registers — each table value
//            is read once, 16 bytes at a time, consecutive lanes in different banks; the two halves are added with
//            one shuffle and stored to the ring
//   phase 2b one thread per completed column: its 2K-1 band entries are sums over the column's K intervals of
//            compile-time-indexed ring entries (no index arithmetic in the loop), intervals in increasing order —
//            the same sum whatever the partition into CTAs, ranks or steps; the columns go to a staging tile and
//            leave the SM as one contiguous, fully coalesced copy
//
// Three CTA barriers per step; the gather of step t-1 shares a segment with phase 0 of step t.  Halo work is the
// K - 1 intervals at the end of a CTA's range (1.3 % at the C3 size).
#include "bspline_core.cuh"

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
                const double t0 = N0[r] * rc[r], t1 = N1[r] * rc[r];
                N0[r] = fma(dh0[r], t0, s0); N1[r] = fma(dh1[r], t1, s1);
                s0 = dl0[K - j + r] * t0; s1 = dl1[K - j + r] * t1;
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

template <int K, int M>
__global__ void __launch_bounds__(STREAM_MAX_THREADS, 1)
bspline_assemble_stream_kernel(StreamArgs A) {
    extern __shared__ __align__(16) double sm[];
    using G = StreamGeom<K>;
    constexpr int QH = G::QH, KP2 = G::KP2, KH = G::KH, W = G::W, NE = G::NE, MSP = G::MSP, PIV = G::PIV, QPW = G::QPW;
    constexpr int NH = (NE + 1) / 2;
    const int NT = blockDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = NT >> 5;
    const int q = A.q, NS = A.NS, RING = NS + K - 1, CAP = RING;
    const int NKP = (NS + 2 * K + 2) / 2 * 2;
    const int ml = A.kv.ml;
    const i64 n_tt = A.kv.n_tt;
    // shared memory carve-up (every region starts on a 16-byte boundary)
    double *TV = sm;                                         // [NS][PIV]   node-major tables, even nodes first
    double *WL = TV + (size_t)NS * PIV;                      // [NS][QPW]   (-1)^m w op at the node positions
    double *MS = WL + (((size_t)NS * QPW + 1) & ~(size_t)1); // [RING][MSP] upper triangles of the local matrices
    double *OUT = MS + (((size_t)RING * MSP + 1) & ~(size_t)1);   // [CAP][W] staging tile of completed columns
    double *KN = OUT + (((size_t)CAP * W + 1) & ~(size_t)1); // [2][NKP + 2]
    double *RD = KN + 2 * (NKP + 2);                         // [2][K * NKP + 2]
    int *sOrd = reinterpret_cast<int *>(RD + (size_t)2 * (K * NKP + 2));   // [NS]
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
    i64 jnext = j0;

    for (int t = 0; t <= nsteps; ++t) {
        const bool live = t < nsteps;
        const i64 s0 = sA + (i64)t * NS;
        const int ns = live ? (int)(sB - s0 + 1 < NS ? sB - s0 + 1 : NS) : 0;
        // ---- this thread's phase-1 item: node data fetched now, used after the barrier ----
        const int sl1 = tid / QH, lp1 = tid - sl1 * QH;
        const bool has1 = 2 * lp1 + 1 < q;
        int c1 = -1;
        double x0 = 1.0, x1 = 1.0, w0 = 1.0, w1 = 1.0, o0 = 1.0, o1 = 1.0;
        if (tid < ns * QH) {
            c1 = A.ord[s0 + sl1 + ml - 1];
            if (c1 >= 0) {
                const i64 g0 = (i64)c1 * q + 2 * lp1, g1 = has1 ? g0 + 1 : g0;
                x0 = A.x[g0]; x1 = A.x[g1];
                if (A.coulomb != 2) { w0 = A.w[g0]; w1 = A.w[g1]; }
                if (A.opvals) { o0 = A.opvals[g0]; o1 = A.opvals[g1]; }
            }
        }
        // ---- phase 0: tables of step t ----
        for (int e = tid; e < ns; e += NT) sOrd[e] = A.ord[s0 + e + ml - 1];
        if (live) {
            for (int j = warp; j < K; j += nwarps) {
                for (int p = lane; p < ns + 2 * K; p += 32) {
                    const double lo = A.kv.at(s0 + ml - K + p);
                    if (j == 0) {
                        KN[p] = lo;
                        if (p > 0) KN[T.kn_copy + p - 1] = lo;
                    } else if (p + j < ns + 2 * K) {
                        const double dk = A.kv.at(s0 + ml - K + p + j) - lo;
                        const double v = dk != 0.0 ? fast_rcp(dk) : 0.0;   // zero widths only feed empty intervals / absent functions
                        RD[j * NKP + p] = v;
                        if (p > 0) RD[T.rd_copy + j * NKP + p - 1] = v;
                    }
                }
            }
        }
        // ---- phase 2b: columns completed by step t-1 (everything that is left after the last step) ----
        const i64 s_end = s0 - 1 < sB ? s0 - 1 : sB;          // last interval whose local matrix is in the ring
        i64 jl = live ? s_end - (K - 1) - foff + 1 : j1;      // exclusive: columns with f0 + K - 1 <= s_end
        if (jl > j1) jl = j1;
        if (jl < jnext) jl = jnext;
        bool first = true;
        while (true) {
            const int nem = (int)(jl - jnext < CAP ? jl - jnext : CAP);
            if (tid < nem) {
                const i64 j = jnext + tid, f0 = j + foff;
                double acc[W];
#pragma unroll
                for (int d = 0; d < W; ++d) acc[d] = 0.0;
                int slot = (int)((f0 - sA) % RING);
                if (slot < 0) slot += RING;
#pragma unroll
                for (int r = 0; r < K; ++r) {
                    const i64 s = f0 + r;
                    if (s >= sA && s <= s_end) {
                        const double *m = MS + (size_t)slot * MSP;
#pragma unroll
                        for (int a = 0; a < K; ++a) acc[a + r] += m[G::sidx(a, K - 1 - r)];
                    }
                    slot = slot + 1 == RING ? 0 : slot + 1;
                }
                // band slot d of column j holds row i = j - u + d; with (l, u) = (K-1+ij, K-1-ij) slot d is run entry d
#pragma unroll
                for (int d = 0; d < W; ++d) {
                    const i64 i = j - A.u + d;
                    OUT[tid * W + d] = (i >= 0 && i < A.mm) ? acc[d] : 0.0;
                }
            }
            __syncthreads();
            {
                double *dst = A.band + (size_t)(jnext - A.col_lo) * W;
                for (int e = tid; e < nem * W; e += NT) dst[e] = OUT[e];
            }
            jnext += nem;
            first = false;
            if (jnext >= jl) break;
            __syncthreads();
        }
        (void)first;
        if (!live) break;
        // ---- phase 1: node tables ----
        if (c1 >= 0) {
            if (!A.opvals && A.coulomb == 1) { o0 = -A.Z * fast_rcp(x0); o1 = -A.Z * fast_rcp(x1); }
            double N0[K], N1[K];
            stream_sweep_pair<K, M>(T, sl1, x0, x1, N0, N1);
            double2 *te = reinterpret_cast<double2 *>(TV + (size_t)sl1 * PIV + lp1 * KP2);
            double2 *to = reinterpret_cast<double2 *>(TV + (size_t)sl1 * PIV + (QH + lp1) * KP2);
#pragma unroll
            for (int a = 0; a < KH; ++a) {
                te[a] = make_double2(N0[2 * a], 2 * a + 1 < K ? N0[2 * a + 1] : 0.0);
                if (has1) to[a] = make_double2(N1[2 * a], 2 * a + 1 < K ? N1[2 * a + 1] : 0.0);
            }
            WL[sl1 * QPW + lp1] = (sgn * w0) * o0;
            if (has1) WL[sl1 * QPW + QH + lp1] = (sgn * w1) * o1;
        }
        __syncthreads();
        // ---- phase 2a: upper triangle of the local matrix, two threads (node halves) per interval ----
        if (tid < ((2 * ns + 31) & ~31)) {
            const int sl = tid >> 1, h = tid & 1;
            const bool act = sl < ns && sOrd[sl < ns ? sl : 0] >= 0;
            double acc[NE];
#pragma unroll
            for (int i = 0; i < NE; ++i) acc[i] = 0.0;
            if (act) {
                const int pb = h ? QH : 0, pe = h ? q : QH;     // even nodes | odd nodes
                const double *tv = TV + (size_t)sl * PIV;
                const double *wl = WL + sl * QPW;
                for (int p = pb; p < pe; ++p) {
                    double n[KP2];
                    const double2 *t2 = reinterpret_cast<const double2 *>(tv + p * KP2);
#pragma unroll
                    for (int a = 0; a < KH; ++a) { const double2 v = t2[a]; n[2 * a] = v.x; n[2 * a + 1] = v.y; }
                    const double wv = wl[p];
#pragma unroll
                    for (int a = 0; a < K; ++a) {
                        const double lw = wv * n[a];
#pragma unroll
                        for (int b = a; b < K; ++b) acc[G::sidx(a, b)] = fma(lw, n[b], acc[G::sidx(a, b)]);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < NE; ++i) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], 1);
            if (sl < ns) {
                int slot = (int)((s0 + sl - sA) % RING);
                double *m = MS + (size_t)slot * MSP;
                if (h == 0) {
#pragma unroll
                    for (int i = 0; i < NH; ++i) m[i] = acc[i];
                } else {
#pragma unroll
                    for (int i = NH; i < NE; ++i) m[i] = acc[i];
                }
            }
        }
        __syncthreads();
    }
}

static size_t stream_smem_bytes(int K, int NS) {
    const int QH = (K + 2) / 2, KP2 = (K + 1) / 2 * 2, W = 2 * K - 1, NE = K * (K + 1) / 2, MSP = NE | 1;
    const int PIV0 = 2 * QH * KP2, PIV = PIV0 + ((2 - PIV0 % 4) + 4) % 4, QPW = (2 * QH) | 1;
    const int RING = NS + K - 1, NKP = (NS + 2 * K + 2) / 2 * 2;
    size_t d = (size_t)NS * PIV + (((size_t)NS * QPW + 1) & ~(size_t)1) + (((size_t)RING * MSP + 1) & ~(size_t)1) +
               (((size_t)RING * W + 1) & ~(size_t)1) + 2 * (size_t)(NKP + 2) + 2 * (size_t)(K * NKP + 2);
    return d * sizeof(double) + sizeof(int) * (size_t)NS + 16;
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
int cb_asm_stream_launch(const cb_bspline *B, int m, const double *opvals, int coulomb, double Z,
                         i64 rowL0, i64 mm, i64 colR0, i64 nn, int u,
                         i64 col_lo, i64 col_hi, i64 s_lo, i64 s_hi, double *band, cudaStream_t st) {
    const int K = B->k;
    if (K < 2 || K > 11 || m < 0 || m > 1 || m >= K) return -1;
    if ((B->q + 1) / 2 != (K + 2) / 2) return -1;
    int NS = 64;
    const int QH = (K + 2) / 2;
    if (NS * QH > STREAM_MAX_THREADS) NS = STREAM_MAX_THREADS / QH;
    const int ov = stream_ns_override();
    if (ov == 0) return -1;
    if (ov > 0) NS = ov * QH > STREAM_MAX_THREADS ? STREAM_MAX_THREADS / QH : ov;
    if (NS < K) NS = K;
    while (NS > K && stream_smem_bytes(K, NS) > 220 * 1024) --NS;
    const i64 ncol = col_hi - col_lo;
    if (ncol <= 0) return CB_OK;
    int nt = NS * QH;
    if (2 * NS > nt) nt = 2 * NS;
    if (NS + K - 1 > nt) nt = NS + K - 1;
    nt = (nt + 31) / 32 * 32;
    if (nt > STREAM_MAX_THREADS) return -1;
    i64 ncta = (ncol + NS - 1) / NS;
    if (ncta > CB_NUM_SMS) ncta = CB_NUM_SMS;
    StreamArgs A;
    A.kv = make_knot_view(B);
    A.q = B->q; A.x = B->d_x; A.w = B->d_w; A.ord = B->d_ord;
    A.opvals = opvals; A.coulomb = coulomb; A.Z = Z;
    A.rowL0 = rowL0; A.mm = mm; A.colR0 = colR0; A.nn = nn; A.u = u;
    A.col_lo = col_lo; A.col_hi = col_hi; A.s_lo = s_lo; A.s_hi = s_hi;
    A.cpc = (ncol + ncta - 1) / ncta;
    A.NS = NS;
    A.band = band;
    const size_t smem = stream_smem_bytes(K, NS);
#define CB_STREAM_CASE(KT)                                                                                       \
    case KT: {                                                                                                   \
        if (m == 0) {                                                                                            \
            auto kern = bspline_assemble_stream_kernel<KT, 0>;                                                   \
            CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
            kern<<<(int)ncta, nt, smem, st>>>(A);                                                                \
        } else {                                                                                                 \
            auto kern = bspline_assemble_stream_kernel<KT, 1>;                                                   \
            CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
            kern<<<(int)ncta, nt, smem, st>>>(A);                                                                \
        }                                                                                                        \
    } break;
    switch (K) {
        CB_STREAM_CASE(2) CB_STREAM_CASE(3) CB_STREAM_CASE(4) CB_STREAM_CASE(5) CB_STREAM_CASE(6)
        CB_STREAM_CASE(7) CB_STREAM_CASE(8) CB_STREAM_CASE(9) CB_STREAM_CASE(10) CB_STREAM_CASE(11)
        default: return -1;
    }
#undef CB_STREAM_CASE
    CB_LAUNCH_CHECK();
    return CB_OK;
}
