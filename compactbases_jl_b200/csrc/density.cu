// density.cu — mutual-density projection (K10).
//
// Reference (src/densities.jl:48-101,156-164): rho = C * (conj(V f) .* (V g)),
// V = B[locs(B), sel] the Vandermonde matrix at the quadrature nodes and
// C = pinv(Matrix(V)) * Diagonal(w.(x)) a DENSE nf x nquad matrix (:96).  V has
// full column rank, so pinv(V) = (V'V)^-1 V' and
//     rho = G^-1 V' (w .* conj(V f) .* (V g)),      G = V'V
// with G symmetric positive definite of half-bandwidth k-1.  The plan holds the
// node table of V (K2), the banded Cholesky factor of G and nothing dense.
//
// Apply = three kernels over the pair batch (lanes <-> pairs so that every
// table / factor entry is a warp-uniform load):
//   density_rhs      r = V'(w .* Vf .* Vg)  sliding window over the knot intervals
//   density_solve    forward + backward banded substitution, in place
// Optional refinement (corrected semi-normal equations) repeats both on the
// node residual s - V rho.
#include "bspline_core.cuh"

int cb_bspline_gram_internal(const cb_bspline *B, i64 col0, i64 ncols, double *band, void *stream);  // assemble.cu

struct cb_density_s {
    int k, q, ml, device;
    i64 n_tt, nf, col0, ncols, ni;
    double *d_table;     // [ni][q][k] values of the live functions at the nodes
    double *d_wv;        // NULL or w(x) at the nodes [ni*q]
    int32_t *d_ord;      // [nt] ordinal of expanded interval I (entry I-1) or -1 (copy of the basis')
    i64 *d_ivlist;       // [ni] expanded 1-based index of each non-empty interval (copy)
    double *d_fw;        // [ncols][k]: fw[j][0] = 1/L_jj, fw[j][p] = L_{j,j-p}/L_jj
    double *d_bw;        // [ncols][k]: bw[j][p] = L_{j+p,j}/L_jj (p >= 1)
    // partitioned substitution (chunks of DS_CH rows), NULL when ncols < 2*DS_CH
    i64 nchunk;
    double *d_cfb;       // [ncols][2k-1]: bw[j][0..k-1], then the forward spikes SF[j][1..k-1]
    double *d_sb;        // [ncols][k-1]:  backward spikes SB[j][1..k-1]
    double *d_tf, *d_tb; // [nchunk][k-1][k-1]: incoming -> outgoing state of a chunk (forward / backward)
    int refine;
    int chol_status;     // 0 ok, else 1-based column where the factorisation broke down
};

// ---------------------------------------------------------------------------
// One-time banded Cholesky of G (lower band g[j][p] = G_{j,j-p}, p = 0..k-1).
// Sequential in j; a single warp keeps the last K rows of L in shared memory.
// ---------------------------------------------------------------------------
// indef != 0: symmetric indefinite band, G = L Sigma L' with Sigma = diag(+-1) (the L D L' factorisation without pivoting,
// written so that the tables keep their meaning: forward solve with L, the sign folded into bw[j][0], backward solve with
// L').  A pivot that has lost its digits against the terms it was formed from is reported as a breakdown (status): the
// caller then falls back to the pivoted band LU.
template <int K>
__global__ void density_cholesky_kernel(const double *__restrict__ gband, i64 n,
                                        double *__restrict__ fw, double *__restrict__ bw, int *__restrict__ status,
                                        int indef) {
    constexpr int CH = 64;                          // rows of G staged per step (lanes load, lane 0 factorises)
    constexpr int k = K, W = 2 * K - 1;
    __shared__ double ring[K][K];                   // ring[j % k][p] = L_{j,j-p}
    __shared__ double rsign[K];                     // rsign[j % k] = Sigma_j
    __shared__ double sg[CH][K];                    // sg[r][p] = G_{j0+r, j0+r-p}
    const int lane = threadIdx.x;
    for (i64 j0 = 0; j0 < n; j0 += CH) {
        const int rows = n - j0 < CH ? (int)(n - j0) : CH;
        __syncwarp();
        for (int e = lane; e < rows * k; e += 32) {
            const int r = e / k, p = e - r * k;
            const i64 j = j0 + r, c = j - p;
            // BandedMatrix storage (u + i - j') + W*j' with u = k-1, i = j, j' = c
            sg[r][p] = c >= 0 ? gband[(k - 1 + p) + (i64)W * c] : 0.0;
        }
        __syncwarp();
        if (lane == 0) {
            for (int rr = 0; rr < rows; ++rr) {
                const i64 j = j0 + rr;
                double row[K];
                double sgn = 1.0;
                // L_{j,c}, c = j-k+1 .. j  <->  p = j-c = k-1 .. 0
#pragma unroll
                for (int p = k - 1; p >= 0; --p) {
                    const i64 c = j - p;
                    if (c < 0) { row[p] = 0.0; continue; }
                    double s = sg[rr][p], mag = fabs(s);
                    // sum over c' < c, c' >= j-k+1: L_{j,c'} Sigma_c' L_{c,c'}
#pragma unroll
                    for (int pp = k - 1; pp > p; --pp) {
                        const i64 cc = j - pp;
                        if (cc < 0) continue;
                        const int pc = (int)(c - cc);   // L_{c,cc} = ring[c % k][c-cc]; row j itself is still in `row`
                        const double term = row[pp] * (p == 0 ? row[pp] : ring[c % k][pc]);
                        s -= indef ? rsign[cc % k] * term : term;
                        mag += fabs(term);
                    }
                    if (p == 0) {
                        if (indef) {
                            sgn = s < 0.0 ? -1.0 : 1.0;
                            if (!(fabs(s) > 1e-12 * mag)) { if (*status == 0) *status = (int)(j + 1); s = 1.0; }
                            row[0] = sqrt(fabs(s));
                        } else {
                            if (!(s > 0.0)) { if (*status == 0) *status = (int)(j + 1); s = 1.0; }
                            row[0] = sqrt(s);
                        }
                    } else {
                        row[p] = s / (indef ? rsign[c % k] * ring[c % k][0] : ring[c % k][0]);
                    }
                }
                for (int p = 0; p < k; ++p) ring[j % k][p] = row[p];
                rsign[j % k] = sgn;
                const double inv = 1.0 / row[0];
                fw[j * k] = inv;
                for (int p = 1; p < k; ++p) fw[j * k + p] = row[p] * inv;
                // bw[c][p] = L_{c+p,c}/L_cc : row j contributes bw[j-p][p]; the sign of the pivot rides on bw[j][0]
                bw[j * k] = sgn * inv;
                for (int p = 1; p < k; ++p)
                    if (j - p >= 0) bw[(j - p) * k + p] = row[p] / ring[(j - p) % k][0];
            }
        }
    }
    __syncwarp();
    // entries of bw that refer to rows beyond n
    if (lane == 0)
        for (i64 j = n - k + 1 > 0 ? n - k + 1 : 0; j < n; ++j)
            for (int p = 1; p < k; ++p)
                if (j + p >= n) bw[j * k + p] = 0.0;
}

__device__ __forceinline__ void ds_cp16(double *smem_dst, const double *gsrc) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}

__device__ __forceinline__ void ds_cp8(double *smem_dst, const double *gsrc) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}

// ---------------------------------------------------------------------------
// MODE 0:  r = V'(w .* (V f) .* (V g))
// MODE 1:  r = V'(w .* (V f) .* (V g) - V rho)      (refinement residual)
//
// A warp owns 32 pairs and a chunk of DR_CH consecutive functions.  Each lane
// walks the knot intervals of the chunk keeping K coefficients of f and g and K
// partial sums in registers (a function enters at window slot K-1 and leaves,
// complete, at slot 0).  Coefficient rows are staged through shared memory in
// 32x32 tiles so that global traffic is coalesced along the contiguous
// (function) dimension of f, g and rho; table and weight entries are
// warp-uniform loads.
// ---------------------------------------------------------------------------
#ifndef DR_NODES
#define DR_NODES 3
#endif
#ifndef DS_FIX_UNROLL
#define DS_FIX_UNROLL 8
#endif
constexpr int DR_WARPS = 4;
constexpr int DR_CH = 256;
constexpr int DR_TILE = 32 * 33;                   // doubles per warp tile (odd pitch)

struct DensArgs {
    const int32_t *ord; const double *table; const double *wv;
    i64 n_tt, nf, col0, ncols; int ml, q;
    const double *f; i64 ldf; i64 Pf;
    const double *g; i64 ldg; i64 Pg;
    const double *rho_in; i64 ldri;     // MODE 1
    double *out; i64 ldo;               // r [ncols x P], or the nodal product [nnodes x P]
    i64 P;
    // FUSE: the local forward sweep of the partitioned solve (density_fwd_local_kernel, F1) runs on the rows as they are
    // produced — `out` receives yloc instead of r and E the chunks' final states; chunks are the solver's (ds_chunk_*)
    const double *fwt; double *E; i64 nchunk_ds;
    // FUSE: yloc goes to the solver's tiled intermediate T[(pair group * ncols + row) * 32 + pair in group] (dst_tiled),
    // a row of 32 pairs being one coalesced 256-byte store, instead of through a transposing tile into `out`
    double *out_tiled;
};

// QT > 0: the node count is the compile-time constant QT (the default q = K + 1): the node loops unroll and the
// table reads get immediate offsets (25 % of the instructions of the run-time version were index arithmetic).
__host__ __device__ inline i64 ds_chunk_begin(i64 c);
__host__ __device__ inline i64 ds_chunk_end(i64 c, i64 nchunk, i64 n);

// The fused form (no output tile: 73 KB of shared memory per CTA) runs three CTAs per SM where the window fits 168
// registers without spilling (k <= 10 with the default node count, k <= 8 otherwise): 12 instead of 8 warps per SM,
// C5 16.8 -> 16.4 ms; the compiler used 216-236 registers only because it could.
template <int K, int MODE, int QT, bool FUSE = false>
__global__ void __launch_bounds__(DR_WARPS * 32, FUSE ? ((K <= 8 || (QT > 0 && K <= 10)) ? 3 : 2) : 1)
density_rhs_kernel(DensArgs A) {
    extern __shared__ __align__(16) double sm[];
    constexpr int NT = FUSE ? 2 : (MODE == 1 ? 4 : 3);   // FUSE writes the tiled intermediate directly: no output tile
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = QT > 0 ? QT : A.q;
    const int TS = (q * K + q + 1) & ~1;            // one interval's table + node weights (even: rows are read as double2)
    constexpr int SORD = (FUSE ? 2 : 1) * DR_CH + K + 2;         // intervals of a chunk (the solver's last chunk takes the remainder)
    constexpr int SW = (SORD + 31) / 32;           // words of the chunk's non-empty-interval mask
    double *sf = sm + (size_t)warp * (NT * DR_TILE + 2 * TS + (SW + 3) / 4 * 2), *sg = sf + DR_TILE, *so = sg + DR_TILE;
    double *sr = MODE == 1 ? so + DR_TILE : so;
    double *stab = sf + NT * DR_TILE;               // [2][TS]: the interval being used and the next one (cp.async)
    // which of the chunk's intervals are non-empty (bit e: interval s0 + e); their ordinals are consecutive, so the
    // mask and the ordinal of the first one replace a table of SORD ordinals (2 KB per warp: the third CTA per SM)
    unsigned *smask = reinterpret_cast<unsigned *>(stab + 2 * TS);
    const i64 nchunk = FUSE ? A.nchunk_ds : (A.ncols + DR_CH - 1) / DR_CH;
    const i64 npg = (A.P + 31) / 32;
    const int ml = A.ml;
    for (i64 item = (i64)blockIdx.x * DR_WARPS + warp; item < nchunk * npg; item += (i64)gridDim.x * DR_WARPS) {
        const i64 pg = item / nchunk, ch = item - pg * nchunk;
        const i64 p0 = pg * 32;
        const i64 fa = A.col0 + 1 + (FUSE ? ds_chunk_begin(ch) : ch * DR_CH);   // first function of the chunk (1-based, absolute)
        i64 fb = FUSE ? A.col0 + 1 + ds_chunk_end(ch, nchunk, A.ncols) : fa + DR_CH;
        if (fb > A.col0 + A.ncols + 1) fb = A.col0 + A.ncols + 1;   // exclusive
        double yw[K];                                               // FUSE: the last K-1 forward-swept rows
#pragma unroll
        for (int p = 0; p < K; ++p) yw[p] = 0.0;
        const i64 s0 = fa - ml, s1 = (fb - 1) - ml + K - 1;         // interval range (virtual outside 0..n_tt-2)
        double fw[K], gw[K], rw[K], acc[K];
        i64 tile_base = fa - K + 1;                                 // function of tile row 0
        // Coefficient rows stream through a 32-row ring per array (slot = row % 32, 33-double pitch per pair): the half
        // that has just been consumed is refilled with cp.async 16 intervals before it is needed again.
        // rows [tile_base + 16h .. +16) -> ring half h & 1; lanes 0-15 <-> rows, two pairs per step
        auto issue_half = [&](i64 h) {
            const int r16 = lane & 15;
            const i64 row = tile_base + 16 * h + r16 - 1 - A.col0;  // row of the restricted coefficient arrays
            const int slot = (int)((16 * h + r16) & 31);
            const bool rok = row >= 0 && row < A.ncols;
#pragma unroll 4
            for (int pp = lane >> 4; pp < 32; pp += 2) {
                const i64 p = p0 + pp;
                if (rok && p < A.P) {
                    ds_cp8(&sf[pp * 33 + slot], A.f + row + A.ldf * (A.Pf == 1 ? 0 : p));
                    ds_cp8(&sg[pp * 33 + slot], A.g + row + A.ldg * (A.Pg == 1 ? 0 : p));
                    if (MODE == 1) ds_cp8(&sr[pp * 33 + slot], A.rho_in + row + A.ldri * p);
                } else {
                    sf[pp * 33 + slot] = 0.0; sg[pp * 33 + slot] = 0.0;
                    if (MODE == 1) sr[pp * 33 + slot] = 0.0;
                }
            }
        };
        asm volatile("cp.async.wait_all;" ::: "memory");           // no copy of the previous item is still landing
        __syncwarp();                                               // the previous item is done with the rings
        issue_half(0); issue_half(1);
        int nord = 0;                                               // ordinal of the next non-empty interval
        {
            bool found = false;
#pragma unroll
            for (int w_ = 0; w_ < SW; ++w_) {
                const int e = 32 * w_ + lane;
                const i64 s_ = fa - ml + e;
                const int o = (e < SORD && s_ >= 0 && s_ <= A.n_tt - 2) ? A.ord[s_ + ml - 1] : -1;
                const unsigned m = __ballot_sync(0xffffffffu, o >= 0);
                if (lane == 0) smask[w_] = m;
                if (!found && m != 0u) { nord = __shfl_sync(0xffffffffu, o, __ffs(m) - 1); found = true; }
            }
        }
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        i64 next_half = 2;
        // window before interval s0: functions fa-K+1 .. fa-1 wait in slots 1..K-1 (shifted down on entry)
        fw[0] = gw[0] = rw[0] = acc[0] = 0.0;
#pragma unroll
        for (int a = 0; a < K - 1; ++a) {
            fw[a + 1] = sf[lane * 33 + a]; gw[a + 1] = sg[lane * 33 + a];
            rw[a + 1] = MODE == 1 ? sr[lane * 33 + a] : 0.0;
            acc[a + 1] = 0.0;
        }
        int tpos = K - 1;                                           // tile row of the next entering function
        int opos = 0;                                               // rows pending in the output tile
        i64 obase = fa;                                             // function of output tile row 0
        // The k x q table of an interval is the same for every pair: it is copied to shared memory one
        // interval ahead (its L2 latency would otherwise stall the two resident warps per scheduler).
        // ordinal of interval s0 + e or -1; to be called with e = 0, 1, 2, ... in this order
        auto take = [&](int e) -> int { return ((smask[e >> 5] >> (e & 31)) & 1u) ? nord++ : -1; };
        auto issue_tab = [&](int c_, int buf_) {
            if (c_ >= 0) {
                const double *src = A.table + (i64)c_ * q * K;
                double *dst = stab + buf_ * TS;
                if (QT > 0 && ((QT * K) & 1) == 0) {
#pragma unroll
                    for (int e = 0; e < (QT * K / 2 + 31) / 32; ++e)
                        if (32 * e + lane < QT * K / 2) ds_cp16(dst + 2 * (32 * e + lane), src + 2 * (32 * e + lane));
                } else {
                    for (int e = lane; e < q * K; e += 32) ds_cp8(dst + e, src + e);
                }
                if (A.wv) for (int e = lane; e < q; e += 32) ds_cp8(dst + q * K + e, A.wv + (i64)c_ * q + e);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        asm volatile("cp.async.wait_group 0;" ::: "memory");        // nothing of the previous item is still landing
        __syncwarp();
        int c = take(0), c_nxt = take(1), buf = 0;
        issue_tab(c, 0);
        // (Unrolling this loop K times so that the window rotates through the registers instead of being shifted — 56
        // moves per interval — was measured: 23.5 ms instead of 19.7 ms at C5; the 8-fold body thrashes the instruction
        // cache with two out-of-phase warps per scheduler.)
        // 32-bit interval counter si = s - s0: function fe = fa + si - (K-1) is completed by interval si >= K-1 (the index
        // arithmetic of the loop was a sixth of its instructions with 64-bit counters)
        const int nint = (int)(s1 - s0) + 1;
        const double *fwrow = FUSE ? A.fwt + (fa - (A.col0 + 1)) * K : nullptr;    // factor row of function fa
        for (int si = 0; si < nint; ++si) {
            const int c_nn = take(si + 2);
            double cfr[K];                                          // FUSE: factor row of the function completed by this interval
            if (FUSE) {
                const double *fr = fwrow + (si >= K - 1 ? si - (K - 1) : 0) * K;
                if ((K & 1) == 0) {
#pragma unroll
                    for (int p = 0; p < K; p += 2) {
                        const double2 c2 = __ldg(reinterpret_cast<const double2 *>(fr + p));
                        cfr[p] = c2.x; cfr[p + 1] = c2.y;
                    }
                } else {
#pragma unroll
                    for (int p = 0; p < K; ++p) cfr[p] = __ldg(fr + p);
                }
            }
            issue_tab(c_nxt, buf ^ 1);
            asm volatile("cp.async.wait_group 1;" ::: "memory");
            __syncwarp();
#pragma unroll
            for (int a = 0; a < K - 1; ++a) { fw[a] = fw[a + 1]; gw[a] = gw[a + 1]; rw[a] = rw[a + 1]; acc[a] = acc[a + 1]; }
            if (tpos == 32) tpos = 0;
            // slot tpos is read now; when the read position enters a half, the other half (fully consumed) is refilled
            // (the copies join the cp.async group committed with the next table and have landed long before slot
            // tpos + 16 is read)
            if ((tpos & 15) == 0 && si > 0) { issue_half(next_half); ++next_half; }
            fw[K - 1] = sf[lane * 33 + tpos]; gw[K - 1] = sg[lane * 33 + tpos];
            rw[K - 1] = MODE == 1 ? sr[lane * 33 + tpos] : 0.0;
            acc[K - 1] = 0.0;
            ++tpos;
            if (c >= 0) {
                const double *tab = stab + buf * TS, *wq = tab + q * K;
                // DR_NODES nodes per iteration: independent FMA chains keep the FP64 pipe busy (table rows as double2)
                int l = 0;
                auto ldrow = [&](double (&t_)[K], int l_) {
                    if ((K & 1) == 0) {
#pragma unroll
                        for (int a = 0; a < K; a += 2) {
                            const double2 v2 = *reinterpret_cast<const double2 *>(tab + l_ * K + a);
                            t_[a] = v2.x; t_[a + 1] = v2.y;
                        }
                    } else {
#pragma unroll
                        for (int a = 0; a < K; ++a) t_[a] = tab[l_ * K + a];
                    }
                };
#if DR_NODES == 3
#pragma unroll
                for (; l + 2 < q; l += 3) {
                    double u0 = 0.0, v0 = 0.0, z0 = 0.0, u1 = 0.0, v1 = 0.0, z1 = 0.0, u2 = 0.0, v2 = 0.0, z2 = 0.0, ta[K], tb[K], tc[K];
                    ldrow(ta, l); ldrow(tb, l + 1); ldrow(tc, l + 2);
#pragma unroll
                    for (int a = 0; a < K; ++a) {
                        u0 = fma(ta[a], fw[a], u0); v0 = fma(ta[a], gw[a], v0);
                        u1 = fma(tb[a], fw[a], u1); v1 = fma(tb[a], gw[a], v1);
                        u2 = fma(tc[a], fw[a], u2); v2 = fma(tc[a], gw[a], v2);
                        if (MODE == 1) { z0 = fma(ta[a], rw[a], z0); z1 = fma(tb[a], rw[a], z1); z2 = fma(tc[a], rw[a], z2); }
                    }
                    double t0 = u0 * v0, t1 = u1 * v1, t2 = u2 * v2;
                    if (A.wv) { t0 *= wq[l]; t1 *= wq[l + 1]; t2 *= wq[l + 2]; }
                    if (MODE == 1) { t0 -= z0; t1 -= z1; t2 -= z2; }
#pragma unroll
                    for (int a = 0; a < K; ++a) acc[a] = fma(tc[a], t2, fma(tb[a], t1, fma(ta[a], t0, acc[a])));
                }
#endif
#pragma unroll
                for (; l + 1 < q; l += 2) {
                    double u0 = 0.0, v0 = 0.0, z0 = 0.0, u1 = 0.0, v1 = 0.0, z1 = 0.0, ta[K], tb[K];
                    ldrow(ta, l); ldrow(tb, l + 1);
#pragma unroll
                    for (int a = 0; a < K; ++a) {
                        u0 = fma(ta[a], fw[a], u0); v0 = fma(ta[a], gw[a], v0);
                        u1 = fma(tb[a], fw[a], u1); v1 = fma(tb[a], gw[a], v1);
                        if (MODE == 1) { z0 = fma(ta[a], rw[a], z0); z1 = fma(tb[a], rw[a], z1); }
                    }
                    double t0 = u0 * v0, t1 = u1 * v1;
                    if (A.wv) { t0 *= wq[l]; t1 *= wq[l + 1]; }
                    if (MODE == 1) { t0 -= z0; t1 -= z1; }
#pragma unroll
                    for (int a = 0; a < K; ++a) acc[a] = fma(tb[a], t1, fma(ta[a], t0, acc[a]));
                }
                for (; l < q; ++l) {
                    double u = 0.0, v = 0.0, z = 0.0, tv[K];
#pragma unroll
                    for (int a = 0; a < K; ++a) {
                        tv[a] = tab[l * K + a];
                        u = fma(tv[a], fw[a], u);
                        v = fma(tv[a], gw[a], v);
                        if (MODE == 1) z = fma(tv[a], rw[a], z);
                    }
                    double t = u * v;
                    if (A.wv) t *= wq[l];
                    if (MODE == 1) t -= z;
#pragma unroll
                    for (int a = 0; a < K; ++a) acc[a] = fma(tv[a], t, acc[a]);
                }
            }
            __syncwarp();                                           // table buffer free for the copy after next
            c = c_nxt; c_nxt = c_nn; buf ^= 1;
            if (si >= K - 1) {                                      // slot 0 has seen its last interval: function fa + si - (K-1)
                double rv = acc[0];
                if (FUSE) {                                         // density_fwd_local_kernel's row step, same order of operations
                    double a0 = rv * cfr[0], a1 = 0.0;
#pragma unroll
                    for (int p = K - 1; p >= 2; --p) {
                        if (p & 1) a1 = fma(-cfr[p], yw[p], a1); else a0 = fma(-cfr[p], yw[p], a0);
                    }
                    a0 += a1;
                    if (K > 1) a0 = fma(-cfr[1], yw[1], a0);
#pragma unroll
                    for (int p = K - 1; p >= 2; --p) yw[p] = yw[p - 1];
                    if (K > 1) yw[1] = a0;
                    rv = a0;
                }
                if (FUSE) {
                    if (p0 + lane < A.P) A.out_tiled[((pg * A.ncols) + (fa - (A.col0 + 1)) + (si - (K - 1))) * 32 + lane] = rv;
                } else {
                so[lane * 33 + opos] = rv;
                ++opos;
                if (opos == 32 || si == nint - 1) {
                    __syncwarp();
                    for (int pp = 0; pp < 32; ++pp) {
                        const i64 p = p0 + pp;
                        if (p < A.P && lane < opos) A.out[(obase + lane - 1 - A.col0) + A.ldo * p] = so[pp * 33 + lane];
                    }
                    __syncwarp();
                    obase += opos; opos = 0;
                }
                }
            }
        }
        if (FUSE && p0 + lane < A.P) {
#pragma unroll
            for (int p = 1; p < K; ++p) A.E[(ch * (K - 1) + (p - 1)) * A.P + p0 + lane] = yw[p];
        }
    }
}

// Nodal product tmp = (V f) .* (V g) (ftmp .* gtmp of src/densities.jl:158-161;
// feeds copyto!(M, rho), :206-208).  Thread per (node, pair).
template <int K>
__global__ void density_nodal_kernel(DensArgs A, const i64 *__restrict__ ivlist, i64 ni) {
    const i64 nn = ni * A.q;
    for (i64 p = blockIdx.y; p < A.P; p += gridDim.y) {
        const double *fc = A.f + A.ldf * (A.Pf == 1 ? 0 : p), *gc = A.g + A.ldg * (A.Pg == 1 ? 0 : p);
        for (i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x; e < nn; e += (i64)gridDim.x * blockDim.x) {
            const i64 c = e / A.q;
            const i64 I = ivlist[c];
            double u = 0.0, v = 0.0;
#pragma unroll
            for (int a = 0; a < K; ++a) {
                const i64 row = I - K + 1 + a - 1 - A.col0;
                if (row >= 0 && row < A.ncols) {
                    const double tv = A.table[e * K + a];
                    u = fma(tv, fc[row], u);
                    v = fma(tv, gc[row], v);
                }
            }
            A.out[e + A.ldo * p] = u * v;
        }
    }
}

// ---------------------------------------------------------------------------
// In-place banded solve  L L' x = r  for 32 pairs per warp (one warp per CTA).
// The substitution is sequential along the functions and independent over the pairs, so lanes
// run over pairs; 32-row tiles of X and of the scaled factor are double-buffered in shared memory
// with cp.async (coalesced along the rows), and the dependent chain per row is one FMA: the term
// with the newest unknown is added last.
// ---------------------------------------------------------------------------

template <int K>
__global__ void __launch_bounds__(32)
density_solve_kernel(const double *__restrict__ fwt, const double *__restrict__ bwt, i64 n,
                     double *__restrict__ X, i64 ldx, i64 P, const double *__restrict__ addto, i64 lda) {
    constexpr int NB = K <= 12 ? 4 : 3;             // ring of tiles: NB-1 in flight hide the DRAM latency of one warp
    __shared__ double sx[NB][32][33];
    __shared__ double sc[NB][32 * K];
    const int lane = threadIdx.x;
    const i64 ntile = (n + 31) / 32;
    const i64 p0 = (i64)blockIdx.x * 32;
    // issues the copies of tile t (rows 32t..32t+31 of the 32 pairs + the factor rows) into buffer b
    auto issue = [&](i64 t, int b, const double *ct) {
        const i64 j = t * 32 + lane;
        if (j < n) {
#pragma unroll 8
            for (int pp = 0; pp < 32; ++pp)
                if (p0 + pp < P) ds_cp8(&sx[b][pp][lane], X + j + ldx * (p0 + pp));
        }
        const i64 c0 = t * 32 * K, cend = n * K;
#pragma unroll
        for (int i = 0; i < K; ++i) {
            const i64 e = c0 + lane + 32 * i;
            if (e < cend) ds_cp8(&sc[b][lane + 32 * i], ct + e);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto store = [&](i64 t, int b, bool last_sweep) {
        const i64 j = t * 32 + lane;
        if (j < n) {
#pragma unroll 8
            for (int pp = 0; pp < 32; ++pp) {
                const i64 p = p0 + pp;
                if (p < P) {
                    double v = sx[b][pp][lane];
                    if (last_sweep && addto) v += addto[j + lda * p];
                    X[j + ldx * p] = v;
                }
            }
        }
    };
    double yw[K];                                   // yw[p] = y_{j-p} (p >= 1) at row j
    // ---- forward: y_j = r_j/L_jj - sum_p (L_{j,j-p}/L_jj) y_{j-p} ----
#pragma unroll
    for (int p = 0; p < K; ++p) yw[p] = 0.0;
    auto issue_or_skip = [&](i64 t, int b, const double *ct) {     // always commits a group
        if (t >= 0 && t < ntile) issue(t, b, ct);
        else asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int i = 0; i < NB - 1; ++i) issue_or_skip(i, i, fwt);
    for (i64 t = 0; t < ntile; ++t) {
        const int b = (int)(t % NB);
        issue_or_skip(t + NB - 1, (int)((t + NB - 1) % NB), fwt);
        asm volatile("cp.async.wait_group %0;" ::"n"(NB - 1) : "memory");
        __syncwarp();
        const int rows = n - t * 32 < 32 ? (int)(n - t * 32) : 32;
        {   // the tile's rows go through registers so that no shared-memory store sits between the loads
            double xr[32];
#pragma unroll
            for (int r = 0; r < 32; ++r) xr[r] = sx[b][lane][r];
#pragma unroll
            for (int r = 0; r < 32; ++r) {
                if (r < rows) {
                    const double *cf = &sc[b][r * K];
                    // older unknowns in two independent partial sums, the newest one (yw[1]) last
                    double acc = xr[r] * cf[0], acc2 = 0.0;
#pragma unroll
                    for (int p = K - 1; p >= 2; --p) {
                        if (p & 1) acc2 = fma(-cf[p], yw[p], acc2); else acc = fma(-cf[p], yw[p], acc);
                    }
                    acc += acc2;
                    if (K > 1) acc = fma(-cf[1], yw[1], acc);
#pragma unroll
                    for (int p = K - 1; p >= 2; --p) yw[p] = yw[p - 1];
                    if (K > 1) yw[1] = acc;
                    xr[r] = acc;
                }
            }
#pragma unroll
            for (int r = 0; r < 32; ++r) sx[b][lane][r] = xr[r];
        }
        __syncwarp();
        store(t, b, false);
        __syncwarp();
    }
    // ---- backward: x_j = y_j/L_jj - sum_p (L_{j+p,j}/L_jj) x_{j+p} ----
#pragma unroll
    for (int p = 0; p < K; ++p) yw[p] = 0.0;
    __threadfence_block();
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
#pragma unroll
    for (int i = 0; i < NB - 1; ++i) issue_or_skip(ntile - 1 - i, i, bwt);
    for (i64 t = ntile - 1, it = 0; t >= 0; --t, ++it) {
        const int b = (int)(it % NB);
        issue_or_skip(t - (NB - 1), (int)((it + NB - 1) % NB), bwt);
        asm volatile("cp.async.wait_group %0;" ::"n"(NB - 1) : "memory");
        __syncwarp();
        const int rows = n - t * 32 < 32 ? (int)(n - t * 32) : 32;
        {   // the tile's rows go through registers so that no shared-memory store sits between the loads
            double xr[32];
#pragma unroll
            for (int r = 0; r < 32; ++r) xr[r] = sx[b][lane][r];
#pragma unroll
            for (int r = 31; r >= 0; --r) {
                if (r < rows) {
                    const double *cb = &sc[b][r * K];
                    // older unknowns in two independent partial sums, the newest one (yw[1]) last
                    double acc = xr[r] * cb[0], acc2 = 0.0;
#pragma unroll
                    for (int p = K - 1; p >= 2; --p) {
                        if (p & 1) acc2 = fma(-cb[p], yw[p], acc2); else acc = fma(-cb[p], yw[p], acc);
                    }
                    acc += acc2;
                    if (K > 1) acc = fma(-cb[1], yw[1], acc);
#pragma unroll
                    for (int p = K - 1; p >= 2; --p) yw[p] = yw[p - 1];
                    if (K > 1) yw[1] = acc;
                    xr[r] = acc;
                }
            }
#pragma unroll
            for (int r = 0; r < 32; ++r) sx[b][lane][r] = xr[r];
        }
        __syncwarp();
        store(t, b, true);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// Partitioned banded substitution.
//
// The forward sweep  y_j = r_j/L_jj - sum_p (L_{j,j-p}/L_jj) y_{j-p}  is a linear recurrence of order k-1 along
// the rows: one warp per 32 pairs walks all n rows one after the other (density_solve_kernel), which is ~50 cycles of
// dependent FP64 latency per row and leaves the GPU empty for small pair batches (C5 tile: 50 007 rows x 1024 pairs
// = 32 warps, 10.6 ms).  The rows are therefore cut into chunks of DS_CH; with s_c the k-1 unknowns just before chunk
// c ("incoming state"),
//     y_j = yloc_j + sum_i SF[j][i] s_c[i],     j in chunk c,
// where yloc is the chunk's own sweep started from a zero state and the spikes SF (the response of the chunk to a unit
// incoming state — a property of the factor only) are tabulated once in the plan.  So:
//     F1  every (chunk, 32 pairs) in parallel: yloc in place, and the chunk's final state E_c
//     F2  one thread per pair, chunk after chunk:  s_{c+1} = E_c + TF_c s_c      (TF_c = last k-1 rows of SF)
//     FB  every (chunk, 32 pairs): rows in descending order, y_j = yloc_j + SF[j].s_c on the fly, fed straight into
//         the chunk's local BACKWARD sweep; zloc in place, final state Eb_c
//     B2  as F2, downwards
//     B3  x_j = zloc_j + SB[j].sb_c  (no recurrence: one thread per entry)
// The result is the same triangular solves (exact, no truncation of the decaying spikes); the work doubles to
// 4(k-1) FMAs per entry and three passes over the batch, but every pass has n/DS_CH x P/32 independent warps.
// ---------------------------------------------------------------------------
constexpr int DS_CH = 256;                          // rows per chunk (multiple of 32); the last chunk takes the remainder

__host__ __device__ inline i64 ds_chunk_begin(i64 c) { return c * DS_CH; }
__host__ __device__ inline i64 ds_chunk_end(i64 c, i64 nchunk, i64 n) { return c == nchunk - 1 ? n : (c + 1) * DS_CH; }

// Spikes and transfer matrices: one warp per chunk, lane i-1 <-> incoming state component i = 1..K-1.
template <int K>
__global__ void __launch_bounds__(32)
density_spike_kernel(const double *__restrict__ fw, const double *__restrict__ bw, i64 n, i64 nchunk,
                     double *__restrict__ cfb, double *__restrict__ sb, double *__restrict__ tf, double *__restrict__ tb) {
    constexpr int CW = 2 * K - 1, S = K - 1;
    const i64 c = blockIdx.x;
    const int i = threadIdx.x + 1;
    const i64 a = ds_chunk_begin(c), b = ds_chunk_end(c, nchunk, n);
    // copy of the backward factor rows into the combined table
    for (i64 e = a * K + threadIdx.x; e < b * K; e += 32) { const i64 j = e / K; cfb[j * CW + (e - j * K)] = bw[e]; }
    if (i > S) return;
    double w[K];                                    // w[p] = unknown p rows back (forward) / ahead (backward)
#pragma unroll
    for (int p = 1; p < K; ++p) w[p] = p == i ? 1.0 : 0.0;
    for (i64 j = a; j < b; ++j) {
        double y = 0.0;
#pragma unroll
        for (int p = 1; p < K; ++p) y = fma(-fw[j * K + p], w[p], y);
#pragma unroll
        for (int p = K - 1; p >= 2; --p) w[p] = w[p - 1];
        if (K > 1) w[1] = y;
        cfb[j * CW + K + (i - 1)] = y;
    }
#pragma unroll
    for (int p = 1; p < K; ++p) tf[(c * S + (p - 1)) * S + (i - 1)] = w[p];       // outgoing component p <- incoming i
#pragma unroll
    for (int p = 1; p < K; ++p) w[p] = p == i ? 1.0 : 0.0;
    for (i64 j = b - 1; j >= a; --j) {
        double z = 0.0;
#pragma unroll
        for (int p = 1; p < K; ++p) z = fma(-bw[j * K + p], w[p], z);
#pragma unroll
        for (int p = K - 1; p >= 2; --p) w[p] = w[p - 1];
        if (K > 1) w[1] = z;
        sb[j * S + (i - 1)] = z;
    }
#pragma unroll
    for (int p = 1; p < K; ++p) tb[(c * S + (p - 1)) * S + (i - 1)] = w[p];
}

// Shared-memory ring of one warp: tiles of 32 rows x 32 pairs (transposed through shared memory so that global
// traffic runs along the contiguous row dimension) and the CW coefficients of each row.
#ifndef DS_TILE_LDG
#define DS_TILE_LDG 1
#endif
template <int CW, int NB>
struct DsRing {
    double sx[NB][32][33];
    double sc[NB][32 * CW];
};

template <int CW, int NB>
__device__ __forceinline__ void ds_issue(DsRing<CW, NB> &R, int b, i64 row0, i64 row_end, const double *__restrict__ X, i64 ldx,
                                         i64 p0, i64 P, const double *__restrict__ ct) {
    const int lane = threadIdx.x;
    const i64 j = row0 + lane;
    if (j < row_end) {
        // running pointers: the 64-bit index arithmetic and the bound test per copy were a third of the sweeps' instructions
        const int nv = (int)(P - p0 < 32 ? P - p0 : 32);
        const double *src = X + j + ldx * p0;
        double *dst = &R.sx[b][0][lane];
#if DS_TILE_LDG
        // plain loads, eight in flight per lane, then the transposing stores: 8-byte cp.async copies queue up in the LSU
        // (the sweeps moved 2.7 TB/s with them, whatever the occupancy, the tile shape or the instruction count)
        if (nv == 32) {
#pragma unroll
            for (int p8 = 0; p8 < 32; p8 += 8) {
                double v[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) v[u] = __ldg(src + (i64)u * ldx);
                src += 8 * ldx;
#pragma unroll
                for (int u = 0; u < 8; ++u) dst[(p8 + u) * 33] = v[u];
            }
        } else {
            for (int pp = 0; pp < nv; ++pp, src += ldx) dst[pp * 33] = __ldg(src);
        }
#else
        if (nv == 32) {
#pragma unroll 8
            for (int pp = 0; pp < 32; ++pp, src += ldx) ds_cp8(dst + pp * 33, src);
        } else {
            for (int pp = 0; pp < nv; ++pp, src += ldx) ds_cp8(dst + pp * 33, src);
        }
#endif
    }
    const i64 c0 = row0 * CW, cend = row_end * CW;
#pragma unroll
    for (int i = 0; i < CW; ++i) {
        const i64 e = c0 + lane + 32 * i;
        if (e < cend) ds_cp8(&R.sc[b][lane + 32 * i], ct + e);
    }
}

// the coefficient rows of a tile alone
template <int CW, int NB>
__device__ __forceinline__ void ds_issue_coef(DsRing<CW, NB> &R, int b, i64 row0, i64 row_end, const double *__restrict__ ct) {
    const int lane = threadIdx.x;
    const i64 c0 = row0 * CW, cend = row_end * CW;
#pragma unroll
    for (int i = 0; i < CW; ++i) {
        const i64 e = c0 + lane + 32 * i;
        if (e < cend) ds_cp8(&R.sc[b][lane + 32 * i], ct + e);
    }
}

template <int CW, int NB>
__device__ __forceinline__ void ds_store(const DsRing<CW, NB> &R, int b, i64 row0, i64 row_end, double *__restrict__ X, i64 ldx,
                                         i64 p0, i64 P) {
    const int lane = threadIdx.x;
    const i64 j = row0 + lane;
    if (j < row_end) {
        const int nv = (int)(P - p0 < 32 ? P - p0 : 32);
        double *dstg = X + j + ldx * p0;
        const double *srcs = &R.sx[b][0][lane];
        if (nv == 32) {
#pragma unroll 8
            for (int pp = 0; pp < 32; ++pp, dstg += ldx) *dstg = srcs[pp * 33];
        } else {
            for (int pp = 0; pp < nv; ++pp, dstg += ldx) *dstg = srcs[pp * 33];
        }
    }
}

// F1: local forward sweep of every chunk.  grid = (pair tiles, chunks), one warp per CTA.
template <int K>
__global__ void __launch_bounds__(32)
density_fwd_local_kernel(const double *__restrict__ fwt, i64 n, i64 nchunk, double *__restrict__ X, i64 ldx, i64 P,
                         double *__restrict__ E, double *__restrict__ T) {
    constexpr int NB = 2, S = K - 1;
    __shared__ DsRing<K, NB> R;
    const int lane = threadIdx.x;
    const i64 p0 = (i64)blockIdx.x * 32, c = blockIdx.y;
    const i64 a = ds_chunk_begin(c), b = ds_chunk_end(c, nchunk, n);
    const i64 ntile = (b - a + 31) / 32;
    double yw[K];
#pragma unroll
    for (int p = 0; p < K; ++p) yw[p] = 0.0;
    ds_issue(R, 0, a, b, X, ldx, p0, P, fwt);
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (i64 t = 0; t < ntile; ++t) {
        const int bb = (int)(t & 1);
        if (t + 1 < ntile) ds_issue(R, bb ^ 1, a + 32 * (t + 1), b, X, ldx, p0, P, fwt);
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();
        const i64 row0 = a + 32 * t;
        const int rows = b - row0 < 32 ? (int)(b - row0) : 32;
        double xr[32];
#pragma unroll
        for (int r = 0; r < 32; ++r) xr[r] = R.sx[bb][lane][r];
        auto row_step = [&](const int r) {
            const double *cf = &R.sc[bb][r * K];
            double acc = xr[r] * cf[0], acc2 = 0.0;
#pragma unroll
            for (int p = K - 1; p >= 2; --p) {
                if (p & 1) acc2 = fma(-cf[p], yw[p], acc2); else acc = fma(-cf[p], yw[p], acc);
            }
            acc += acc2;
            if (K > 1) acc = fma(-cf[1], yw[1], acc);
#pragma unroll
            for (int p = K - 1; p >= 2; --p) yw[p] = yw[p - 1];
            if (K > 1) yw[1] = acc;
            xr[r] = acc;
        };
        if (rows == 32) {                           // full tile: no per-row bound test, the window shifts are renamed away
#pragma unroll
            for (int r = 0; r < 32; ++r) row_step(r);
        } else {
#pragma unroll
            for (int r = 0; r < 32; ++r)
                if (r < rows) row_step(r);
        }
        // the swept rows go to the tiled intermediate of the solve: a row of the 32 pairs is one coalesced store
        {
            double *trow = T + ((i64)blockIdx.x * n + row0) * 32 + lane;
#pragma unroll
            for (int r = 0; r < 32; ++r)
                if (r < rows) trow[r * 32] = xr[r];
        }
        __syncwarp();
    }
    if (p0 + lane < P) {
#pragma unroll
        for (int p = 1; p < K; ++p) E[(c * S + (p - 1)) * P + p0 + lane] = yw[p];
    }
}

// F2 / B2: chunk-to-chunk propagation of the state, one thread per pair.  dir = +1: S_0 = 0, S_{c+1} = E_c + T_c S_c;
// dir = -1: S_{M-1} = 0, S_{c-1} = E_c + T_c S_c.  The chain of nchunk-1 dependent (k-1)x(k-1) products is latency
// bound, so nothing on it may wait for memory: the transfer matrices of DS_BATCH chunks are staged in shared memory
// at a time and the next chunk's E is fetched while the current one is multiplied.
constexpr int DS_STATE_THREADS = 32;
#ifndef DS_STATE_CP
#define DS_STATE_CP 1                             // component-parallel state propagation (density_state_cp_kernel)
#endif

template <int K>
__global__ void __launch_bounds__(DS_STATE_THREADS)
density_state_kernel(const double *__restrict__ T, i64 nchunk, i64 P, const double *__restrict__ E,
                     double *__restrict__ St, int dir) {
    constexpr int S = K > 1 ? K - 1 : 1;
    constexpr int BATCH = K <= 8 ? 16 : 8;          // chunks staged at a time (E and T together < 48 KB)
    __shared__ double sT[BATCH * S * S];
    __shared__ double sE[BATCH * S][DS_STATE_THREADS];
    const i64 p = blockIdx.x * (i64)DS_STATE_THREADS + threadIdx.x;
    const bool live = p < P;
    double s[S];
#pragma unroll
    for (int i = 0; i < S; ++i) s[i] = 0.0;
    const i64 c0 = dir > 0 ? 0 : nchunk - 1, steps = nchunk - 1;
    if (live) {
#pragma unroll
        for (int i = 0; i < S; ++i) St[(c0 * S + i) * P + p] = 0.0;
    }
    for (i64 base = 0; base < steps; base += BATCH) {
        const int nb = steps - base < BATCH ? (int)(steps - base) : BATCH;
        __syncthreads();
        for (int e = threadIdx.x; e < nb * S * S; e += DS_STATE_THREADS) {
            const int bq = e / (S * S);
            ds_cp8(&sT[e], T + (c0 + dir * (base + bq)) * (S * S) + (e - bq * S * S));
        }
        if (live) {
            for (int e = 0; e < nb * S; ++e) {
                const int bq = e / S, o = e - bq * S;
                ds_cp8(&sE[e][threadIdx.x], E + ((c0 + dir * (base + bq)) * S + o) * P + p);
            }
        }
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        if (!live) continue;
        for (int bq = 0; bq < nb; ++bq) {
            const i64 c = c0 + dir * (base + bq);
            double sn[S];
#pragma unroll
            for (int o = 0; o < S; ++o) {
                double v = sE[bq * S + o][threadIdx.x];
#pragma unroll
                for (int i = 0; i < S; ++i) v = fma(sT[(bq * S + o) * S + i], s[i], v);
                sn[o] = v;
            }
#pragma unroll
            for (int o = 0; o < S; ++o) { s[o] = sn[o]; St[((c + dir) * S + o) * P + p] = sn[o]; }
        }
    }
}

// F2 / B2, component-parallel form: one thread per (pair, state component).  A warp holds 32 / G pairs (G = 8 for
// K <= 9, else 16: the K-1 components of a pair padded to a power of two); a step is K-1 FMAs per thread with the old
// state gathered by shuffles, instead of (K-1)^2 FMAs and as many shared-memory reads in one thread.  The chain was
// issue-bound on that single warp per SM (93 us per sweep at C5: ~470 cycles per chunk); here it is K-1 times shorter
// and spread over K-1 times as many warps.
template <int K>
__global__ void __launch_bounds__(32)
density_state_cp_kernel(const double *__restrict__ T, i64 nchunk, i64 P, const double *__restrict__ E,
                        double *__restrict__ St, int dir) {
    constexpr int S = K > 1 ? K - 1 : 1;
    constexpr int G = S <= 8 ? 8 : 16;              // lanes per pair
    constexpr int PPW = 32 / G;                     // pairs per warp
    constexpr int BATCH = 16;                       // chunks of transfer matrices staged at a time
    __shared__ double sT[BATCH * S * S];
    const int lane = threadIdx.x, o = lane % G, pl = lane / G;
    const i64 p = (i64)blockIdx.x * PPW + pl;
    const bool live = p < P && o < S;
    const int gbase = pl * G;
    double s = 0.0;                                 // component o of the state of pair p
    const i64 c0 = dir > 0 ? 0 : nchunk - 1, steps = nchunk - 1;
    if (live) St[(c0 * S + o) * P + p] = 0.0;
    for (i64 base = 0; base < steps; base += BATCH) {
        const int nb = steps - base < BATCH ? (int)(steps - base) : BATCH;
        __syncwarp();
        for (int e = lane; e < nb * S * S; e += 32) {
            const int bq = e / (S * S);
            ds_cp8(&sT[e], T + (c0 + dir * (base + bq)) * (S * S) + (e - bq * S * S));
        }
        // this lane's E of the batch: independent loads, issued before the chain needs them
        double ev[BATCH];
#pragma unroll
        for (int bq = 0; bq < BATCH; ++bq)
            ev[bq] = (live && bq < nb) ? E[((c0 + dir * (base + bq)) * S + o) * P + p] : 0.0;
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int bq = 0; bq < BATCH; ++bq) {
            if (bq < nb) {
                const i64 c = c0 + dir * (base + bq);
                double v = ev[bq];
#pragma unroll
                for (int i = 0; i < S; ++i) {
                    const double si = __shfl_sync(0xffffffffu, s, gbase + i);
                    if (o < S) v = fma(sT[(bq * S + o) * S + i], si, v);
                }
                s = v;
                if (live) St[((c + dir) * S + o) * P + p] = v;
            }
        }
    }
}

// FB: spike correction of the forward sweep fused with the local backward sweep of every chunk, on the tiled
// intermediate T[(pair group * n + row) * 32 + pair in group].  Two pairs per lane (64 pairs per warp): the 2K-1
// coefficients of a row are the same for every pair (8 LDS.128 per row and warp), so a lane takes two independent pairs
// per coefficient row — half the coefficient loads per pair, two dependency chains to interleave.  Rows move in pieces
// of DB_R = 16 (16-byte cp.async, a piece of a pair group is 4 KB contiguous, no transposition) through two buffers:
// the next piece is in flight while the current one is swept, the swept piece goes back with plain stores.
// History at C5 (job 17.7 ms): on the caller's column-major layout (256-byte pieces of columns 400 KB apart, transposing
// shared-memory tiles) this sweep took 2.5 ms whatever was tried — 9 or 16 resident warps, half the instructions, plain
// loads instead of 8-byte cp.async, 512-byte pieces, L2 prefetches; the tiled layout brought 2.2 ms (3.6 TB/s), the double
// buffer nothing measurable on top.
constexpr int DB_R = 16;
template <int K>
struct DsBwdTile {
    double x[2][2][DB_R * 32];                     // [buffer][pair set][row][lane]
    double c[2][DB_R * (2 * K - 1)];               // [buffer][row][2K-1]
};

template <int K>
__global__ void __launch_bounds__(32, 9)
density_bwd_local2_kernel(const double *__restrict__ cfb, i64 n, i64 nchunk, double *__restrict__ T, i64 P,
                          const double *__restrict__ St, double *__restrict__ E) {
    constexpr int S = K - 1, CW = 2 * K - 1;
    __shared__ __align__(16) DsBwdTile<K> R;
    const int lane = threadIdx.x;
    const i64 p0 = (i64)blockIdx.x * 64, c = blockIdx.y;
    const i64 a = ds_chunk_begin(c), b = ds_chunk_end(c, nchunk, n);
    const i64 npc = (b - a + DB_R - 1) / DB_R;     // pieces of the chunk
    double sin[2][K], zw[2][K];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int p = 0; p < K; ++p) { zw[h][p] = 0.0; sin[h][p] = 0.0; }
        if (p0 + 32 * h + lane < P) {
#pragma unroll
            for (int i = 0; i < S; ++i) sin[h][i] = St[(c * S + i) * P + p0 + 32 * h + lane];
        }
    }
    // requests piece t (rows a + DB_R t ..) into buffer bf
    auto issue = [&](i64 t, int bf) {
        const i64 row0 = a + DB_R * t;
        const int rows = b - row0 < DB_R ? (int)(b - row0) : DB_R;
        const int nch = rows * 16;                  // 16-byte pieces of a pair set
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const double *tg = T + ((2 * (i64)blockIdx.x + h) * n + row0) * 32;
            if (p0 + 32 * h < P) {
#pragma unroll 4
                for (int e = lane; e < nch; e += 32) ds_cp16(&R.x[bf][h][2 * e], tg + 2 * e);
            }
        }
        const i64 c0 = row0 * CW;
        const int ne = rows * CW;
        for (int e = lane; e < ne; e += 32) ds_cp8(&R.c[bf][e], cfb + c0 + e);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    issue(npc - 1, 0);
    for (i64 t = npc - 1, it = 0; t >= 0; --t, ++it) {
        const int bf = (int)(it & 1);
        if (t > 0) issue(t - 1, bf ^ 1); else asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncwarp();
        const i64 row0 = a + DB_R * t;
        const int rows = b - row0 < DB_R ? (int)(b - row0) : DB_R;
        auto row_step = [&](const int r) {
            double cf[CW];
            {
                const double *cs = &R.c[bf][r * CW];
                const int odd = (r * CW) & 1;
                if (odd) cf[0] = cs[0];
#pragma unroll
                for (int i = odd; i < CW; i += 2) {
                    if (i + 1 < CW) {
                        const double2 t2 = *reinterpret_cast<const double2 *>(cs + i);
                        cf[i] = t2.x; cf[i + 1] = t2.y;
                    } else {
                        cf[i] = cs[i];
                    }
                }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                double y = R.x[bf][h][r * 32 + lane], y2 = 0.0;          // y_j = yloc_j + SF[j] . s_c
#pragma unroll
                for (int i = 0; i < S; ++i) {
                    if (i & 1) y2 = fma(cf[K + i], sin[h][i], y2); else y = fma(cf[K + i], sin[h][i], y);
                }
                y += y2;
                double acc = y * cf[0], acc2 = 0.0;
#pragma unroll
                for (int p = K - 1; p >= 2; --p) {
                    if (p & 1) acc2 = fma(-cf[p], zw[h][p], acc2); else acc = fma(-cf[p], zw[h][p], acc);
                }
                acc += acc2;
                if (K > 1) acc = fma(-cf[1], zw[h][1], acc);
#pragma unroll
                for (int p = K - 1; p >= 2; --p) zw[h][p] = zw[h][p - 1];
                if (K > 1) zw[h][1] = acc;
                R.x[bf][h][r * 32 + lane] = acc;
            }
        };
        if (rows == DB_R) {                         // full piece: no per-row bound test, the window shifts are renamed away
#pragma unroll
            for (int r = DB_R - 1; r >= 0; --r) row_step(r);
        } else {
#pragma unroll
            for (int r = DB_R - 1; r >= 0; --r)
                if (r < rows) row_step(r);
        }
        __syncwarp();
        const int nch = rows * 16;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            double *tg = T + ((2 * (i64)blockIdx.x + h) * n + row0) * 32;
            if (p0 + 32 * h < P) {
#pragma unroll 4
                for (int e = lane; e < nch; e += 32)
                    *reinterpret_cast<double2 *>(tg + 2 * e) = *reinterpret_cast<const double2 *>(&R.x[bf][h][2 * e]);
            }
        }
        __syncwarp();                               // buffer bf is requested again by the next trip's issue
    }
#pragma unroll
    for (int h = 0; h < 2; ++h)
        if (p0 + 32 * h + lane < P) {
#pragma unroll
            for (int p = 1; p < K; ++p) E[(c * S + (p - 1)) * P + p0 + 32 * h + lane] = zw[h][p];
        }
}

// B3: x_j = zloc_j + SB[j] . sb_c (+ addto): no recurrence.  A CTA takes 256 rows of one chunk and a tile of DS_FIXP pairs:
// the chunk's incoming states of those pairs sit in shared memory ([pair][8]: two 16-byte broadcast reads fetch four
// components), the spike row of a thread's row stays in registers, X is read and written along the rows.  (One thread
// per row with the states as warp-uniform global loads issued nine memory instructions per element: 2.5 ms at C5.)
constexpr int DS_FIXP = 32;
constexpr int DS_FIXU = DS_FIX_UNROLL;       // loads in flight per thread (4: 41 KB per SM, short of the latency-bandwidth product)

template <int K, bool ADD>
__global__ void __launch_bounds__(256, 3)
density_bwd_fix_kernel(const double *__restrict__ sb, i64 n, i64 nchunk, const double *__restrict__ T, double *__restrict__ X,
                       i64 ldx, i64 P, const double *__restrict__ St, const double *__restrict__ addto, i64 lda) {
    constexpr int S = K > 1 ? K - 1 : 1, SP = (S + 1) / 2 * 2;
    __shared__ __align__(16) double sst[DS_FIXP * SP];
    const i64 npg = (P + DS_FIXP - 1) / DS_FIXP;
    const i64 nrb = (n + 255) / 256;                // row blocks: 256 rows, never across a chunk boundary (DS_CH = 256) except in the last chunk
    for (i64 item = blockIdx.x; item < nrb * npg; item += gridDim.x) {
        const i64 rb = item % nrb, pg = item / nrb;
        const i64 pa = pg * DS_FIXP;
        const int np = (int)(pa + DS_FIXP < P ? DS_FIXP : P - pa);
        i64 c = rb * 256 / DS_CH; if (c > nchunk - 1) c = nchunk - 1;
        __syncthreads();
        for (int e = threadIdx.x; e < DS_FIXP * SP; e += 256) {
            const int pp = e / SP, i = e - pp * SP;
            sst[e] = (pp < np && i < S) ? St[(c * S + i) * P + pa + pp] : 0.0;
        }
        __syncthreads();
        const i64 j = rb * 256 + threadIdx.x;
        if (j >= n) continue;
        double sr[SP];
#pragma unroll
        for (int i = 0; i < SP; ++i) sr[i] = i < S ? sb[j * S + i] : 0.0;
        // DS_FIXU columns at a time, all their loads first.  The swept values come from the tiled intermediate — a
        // thread's 32 pairs of row j are 256 contiguous bytes — and leave in the caller's column-major layout.  (When
        // this kernel still read and wrote X in place the compiler kept every load behind the previous column's store:
        // one load per thread in flight, long-scoreboard stalls of 55 warps per issue, 0.5 of the HBM rate.)
        const double *trow = T + (pg * n + j) * 32;
        for (int pb = 0; pb < np; pb += DS_FIXU) {
            double v[DS_FIXU], ad[DS_FIXU];
            static_assert(DS_FIXU % 2 == 0 && DS_FIXP == 32, "tiled intermediate: 32 pairs per group, double2 loads");
#pragma unroll
            for (int u = 0; u < DS_FIXU; u += 2) {
                const double2 t2 = *reinterpret_cast<const double2 *>(trow + pb + u);
                v[u] = t2.x; v[u + 1] = t2.y;
            }
#pragma unroll
            for (int u = 0; u < DS_FIXU; ++u) {
                const int pp = pb + u;
                ad[u] = (ADD && pp < np) ? addto[j + lda * (pa + pp)] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < DS_FIXU; ++u) {
                const double2 *s2 = reinterpret_cast<const double2 *>(sst + (pb + u) * SP);
                double w = v[u];
#pragma unroll
                for (int h = 0; h < SP / 2; ++h) { const double2 t = s2[h]; w = fma(sr[2 * h], t.x, w); w = fma(sr[2 * h + 1], t.y, w); }
                if (ADD) w += ad[u];
                v[u] = w;
            }
#pragma unroll
            for (int u = 0; u < DS_FIXU; ++u)
                if (pb + u < np) X[j + ldx * (pa + pb + u)] = v[u];
        }
    }
}

// r = V' fvals for B \ f: one thread per (function, column).  table[(c*q + l)*K + a] is the value at node l of the
// non-empty interval with ordinal c of its a-th live function.
template <int K>
__global__ void density_project_rhs_kernel(const double *__restrict__ table, const int32_t *__restrict__ ord, i64 n_tt, int ml,
                                           int q, i64 col0, i64 ncols, const double *__restrict__ fvals, i64 ldf, i64 P,
                                           double *__restrict__ out, i64 ldo) {
    for (i64 p = blockIdx.y; p < P; p += gridDim.y) {
        const double *fc = fvals + ldf * p;
        for (i64 j = blockIdx.x * (i64)blockDim.x + threadIdx.x; j < ncols; j += (i64)gridDim.x * blockDim.x) {
            const i64 fj = col0 + 1 + j;            // 1-based function of the parent basis
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < K; ++r) {
                const i64 sv = fj - ml + r;         // 0-based interval of t.t; expanded 1-based index sv + ml
                if (sv < 0 || sv > n_tt - 2) continue;
                const int c = ord[sv + ml - 1];
                if (c < 0) continue;
                const int a = K - 1 - r;            // slot of function fj among the K live functions of the interval
                const double *tb = table + (i64)c * q * K + a;
                const double *fv = fc + (i64)c * q;
                for (int l = 0; l < q; ++l) s = fma(tb[l * K], fv[l], s);
            }
            out[j + ldo * p] = s;
        }
    }
}

// Orthogonal branches (src/densities.jl:166-185): rho = c .* (conj(f) .* g).
__global__ void density_orthogonal_kernel(const double *__restrict__ c, i64 n,
                                          const double *__restrict__ f, i64 ldf, i64 Pf,
                                          const double *__restrict__ g, i64 ldg, i64 Pg,
                                          double *__restrict__ rho, i64 ldr, i64 P) {
    for (i64 p = blockIdx.y; p < P; p += gridDim.y) {
        const double *fc = f + ldf * (Pf == 1 ? 0 : p), *gc = g + ldg * (Pg == 1 ? 0 : p);
        for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
            double v = fc[i] * gc[i];
            rho[i + ldr * p] = c ? c[i] * v : v;
        }
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
#define CB_DISPATCH_KD(kval, ...)                                                \
    switch (kval) {                                                              \
        case 1: { constexpr int K = 1; __VA_ARGS__; } break;                     \
        case 2: { constexpr int K = 2; __VA_ARGS__; } break;                     \
        case 3: { constexpr int K = 3; __VA_ARGS__; } break;                     \
        case 4: { constexpr int K = 4; __VA_ARGS__; } break;                     \
        case 5: { constexpr int K = 5; __VA_ARGS__; } break;                     \
        case 6: { constexpr int K = 6; __VA_ARGS__; } break;                     \
        case 7: { constexpr int K = 7; __VA_ARGS__; } break;                     \
        case 8: { constexpr int K = 8; __VA_ARGS__; } break;                     \
        case 9: { constexpr int K = 9; __VA_ARGS__; } break;                     \
        case 10: { constexpr int K = 10; __VA_ARGS__; } break;                   \
        case 11: { constexpr int K = 11; __VA_ARGS__; } break;                   \
        case 12: { constexpr int K = 12; __VA_ARGS__; } break;                   \
        case 13: { constexpr int K = 13; __VA_ARGS__; } break;                   \
        case 14: { constexpr int K = 14; __VA_ARGS__; } break;                   \
        case 15: { constexpr int K = 15; __VA_ARGS__; } break;                   \
        case 16: { constexpr int K = 16; __VA_ARGS__; } break;                   \
        default:                                                                 \
            cb_set_error("B-spline order k=%d outside compiled range 1..%d", kval, CB_KMAX); \
            return CB_ERR_UNSUPPORTED;                                           \
    }

static int check_pairs(const char *who, const cb_density *D, const double *f, i64 ldf, i64 Pf,
                       const double *g, i64 ldg, i64 Pg) {
    CB_REQUIRE(D != nullptr, "%s: NULL plan", who);
    CB_REQUIRE(f && g, "%s: NULL array", who);
    CB_REQUIRE(Pf >= 0 && Pg >= 0, "%s: negative batch", who);
    // src/densities.jl:57-60
    CB_REQUIRE(Pf == Pg || Pf == 1 || Pg == 1, "DimensionMismatch: Cannot broadcast %lld LHS and %lld RHS to common size",
               (long long)Pf, (long long)Pg);
    CB_REQUIRE(ldf >= D->ncols && ldg >= D->ncols, "%s: leading dimension smaller than the number of functions", who);
    return CB_OK;
}

static DensArgs make_args(const cb_density *D, const cb_bspline *, const double *f, i64 ldf, i64 Pf,
                          const double *g, i64 ldg, i64 Pg) {
    DensArgs A;
    A.ord = D->d_ord; A.table = D->d_table; A.wv = D->d_wv;
    A.n_tt = D->n_tt; A.nf = D->nf; A.col0 = D->col0; A.ncols = D->ncols; A.ml = D->ml; A.q = D->q;
    A.f = f; A.ldf = ldf; A.Pf = Pf; A.g = g; A.ldg = ldg; A.Pg = Pg;
    A.rho_in = nullptr; A.ldri = 0; A.out = nullptr; A.ldo = 0;
    A.P = Pf > Pg ? Pf : Pg;
    A.fwt = nullptr; A.E = nullptr; A.nchunk_ds = 0; A.out_tiled = nullptr;
    return A;
}

template <int K, int MODE>
static int launch_rhs(const DensArgs &A, cudaStream_t st) {
    const bool fuse = MODE == 0 && A.fwt != nullptr && A.E != nullptr && A.nchunk_ds >= 2 && K > 1;
    i64 items = (fuse ? A.nchunk_ds : (A.ncols + DR_CH - 1) / DR_CH) * ((A.P + 31) / 32);
    i64 blocks = (items + DR_WARPS - 1) / DR_WARPS;
    int grid = (int)(blocks < (i64)CB_NUM_SMS * 64 ? blocks : (i64)CB_NUM_SMS * 64);
    const int sord = (fuse ? 2 : 1) * DR_CH + K + 2;
    if (fuse) CB_REQUIRE(A.out_tiled != nullptr, "cb_density_apply: the fused right-hand side writes the tiled intermediate");
    size_t smem = sizeof(double) * DR_WARPS * ((fuse ? 2 : (MODE == 1 ? 4 : 3)) * DR_TILE + 2 * ((A.q * K + A.q + 1) & ~1) + ((sord + 31) / 32 + 3) / 4 * 2);
    auto launch = [&](auto kern) -> int {
        CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, DR_WARPS * 32, smem, st>>>(A);
        return CB_OK;
    };
    int rc;
    if (fuse) rc = A.q == K + 1 ? launch(density_rhs_kernel<K, 0, K + 1, true>) : launch(density_rhs_kernel<K, 0, 0, true>);
    else rc = A.q == K + 1 ? launch(density_rhs_kernel<K, MODE, K + 1>) : launch(density_rhs_kernel<K, MODE, 0>);
    if (rc) return rc;
    CB_LAUNCH_CHECK();
    return CB_OK;
}

// ws_fwd != NULL: workspace [2][nchunk][K-1][P] whose first half already holds the chunk states E of a forward sweep
// fused into the producer of X (density_rhs_kernel<FUSE>): X holds yloc, F1 is skipped; the caller frees ws_fwd.
template <int K>
static int launch_solve(const cb_density *D, double *X, i64 ldx, i64 P, const double *addto, i64 lda, cudaStream_t st,
                        double *ws_fwd = nullptr, double *T_fwd = nullptr) {
    const i64 n = D->ncols, ptiles = (P + 31) / 32;
    if (D->d_cfb == nullptr || K == 1) {            // short problems: one warp per 32 pairs walks all rows
        density_solve_kernel<K><<<(int)ptiles, 32, 0, st>>>(D->d_fw, D->d_bw, n, X, ldx, P, addto, lda);
        CB_LAUNCH_CHECK();
        return CB_OK;
    }
    const i64 M = D->nchunk;
    constexpr int S = K > 1 ? K - 1 : 1;
    double *ws = ws_fwd;                            // E and the propagated states, [M][K-1][P] each
    if (!ws) {
        int prc = cb_keep_scratch_pool(); if (prc) return prc;
        CB_CUDA(cudaMallocAsync(&ws, sizeof(double) * 2 * M * S * P, st));
    }
    double *E = ws, *St = ws + M * S * P;
    // the tiled intermediate T[(pair group * n + row) * 32 + pair in group] between the sweeps (cb_density_tiled_bytes)
    double *T = T_fwd;
    if (!T) {
        int prc = cb_keep_scratch_pool(); if (prc) return prc;
        CB_CUDA(cudaMallocAsync(&T, sizeof(double) * (size_t)ptiles * n * 32, st));
    }
    // grid.y = chunks <= 65535 is checked at plan time
    dim3 grid((unsigned)ptiles, (unsigned)M);
    if (!ws_fwd) density_fwd_local_kernel<K><<<grid, 32, 0, st>>>(D->d_fw, n, M, X, ldx, P, E, T);
#if DS_STATE_CP
    constexpr int PPW = 32 / ((K > 1 ? K - 1 : 1) <= 8 ? 8 : 16);
    const unsigned sblocks = (unsigned)((P + PPW - 1) / PPW);
    density_state_cp_kernel<K><<<sblocks, 32, 0, st>>>(D->d_tf, M, P, E, St, +1);
    density_bwd_local2_kernel<K><<<dim3((unsigned)((P + 63) / 64), (unsigned)M), 32, 0, st>>>(D->d_cfb, n, M, T, P, St, E);
    density_state_cp_kernel<K><<<sblocks, 32, 0, st>>>(D->d_tb, M, P, E, St, -1);
#else
    const unsigned sblocks = (unsigned)((P + DS_STATE_THREADS - 1) / DS_STATE_THREADS);
    density_state_kernel<K><<<sblocks, DS_STATE_THREADS, 0, st>>>(D->d_tf, M, P, E, St, +1);
    density_bwd_local2_kernel<K><<<dim3((unsigned)((P + 63) / 64), (unsigned)M), 32, 0, st>>>(D->d_cfb, n, M, T, P, St, E);
    density_state_kernel<K><<<sblocks, DS_STATE_THREADS, 0, st>>>(D->d_tb, M, P, E, St, -1);
#endif
    const i64 nfix = ((n + 255) / 256) * ((P + DS_FIXP - 1) / DS_FIXP);
    const unsigned gfix = (unsigned)(nfix < (i64)CB_NUM_SMS * 32 ? nfix : (i64)CB_NUM_SMS * 32);
    if (addto) density_bwd_fix_kernel<K, true><<<gfix, 256, 0, st>>>(D->d_sb, n, M, T, X, ldx, P, St, addto, lda);
    else density_bwd_fix_kernel<K, false><<<gfix, 256, 0, st>>>(D->d_sb, n, M, T, X, ldx, P, St, addto, lda);
    cudaError_t e = cudaGetLastError();
    if (!ws_fwd) cudaFreeAsync(ws, st);
    if (!T_fwd) cudaFreeAsync(T, st);
    if (e != cudaSuccess) { cb_set_error("cb_density_apply: %s", cudaGetErrorString(e)); return CB_ERR_CUDA; }
    return CB_OK;
}

// Spike tables of the partitioned substitution (after the Cholesky factor is known).
template <int K>
static int launch_spikes(cb_density_s *D, cudaStream_t st) {
    const i64 n = D->ncols;
    D->nchunk = n / DS_CH;
    if (D->nchunk < 2 || D->nchunk > 65535 || K == 1) { D->nchunk = 0; return CB_OK; }
    constexpr int S = K > 1 ? K - 1 : 1;
    CB_CUDA(cudaMalloc(&D->d_cfb, sizeof(double) * n * (2 * K - 1)));
    CB_CUDA(cudaMalloc(&D->d_sb, sizeof(double) * n * S));
    CB_CUDA(cudaMalloc(&D->d_tf, sizeof(double) * D->nchunk * S * S));
    CB_CUDA(cudaMalloc(&D->d_tb, sizeof(double) * D->nchunk * S * S));
    density_spike_kernel<K><<<(unsigned)D->nchunk, 32, 0, st>>>(D->d_fw, D->d_bw, n, D->nchunk, D->d_cfb, D->d_sb, D->d_tf, D->d_tb);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

static int launch_cholesky(int k, const double *g, i64 n, double *fw, double *bw, int *status, cudaStream_t st, int indef = 0) {
    CB_DISPATCH_KD(k, (density_cholesky_kernel<K><<<1, 32, 0, st>>>(g, n, fw, bw, status, indef)));
    return CB_OK;
}

extern "C" {

int cb_density_destroy(cb_density *D) {
    if (!D) return CB_OK;
    cudaFree(D->d_table); cudaFree(D->d_wv); cudaFree(D->d_ord); cudaFree(D->d_ivlist);
    cudaFree(D->d_fw); cudaFree(D->d_bw);
    cudaFree(D->d_cfb); cudaFree(D->d_sb); cudaFree(D->d_tf); cudaFree(D->d_tb);
    delete D;
    return CB_OK;
}

int cb_density_bspline_create(const cb_bspline *B, i64 col0, i64 ncols, const double *wvals_host, cb_density **out,
                              void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    CB_REQUIRE(out != nullptr, "cb_density_bspline_create: out is NULL");
    *out = nullptr;
    CB_REQUIRE(B != nullptr, "cb_density_bspline_create: NULL basis");
    CB_REQUIRE(col0 >= 0 && ncols >= 1 && col0 + ncols <= B->nf, "cb_density_bspline_create: restriction outside 1..nf");
    cb_density_s *D = new cb_density_s();
    D->k = B->k; D->q = B->q; D->ml = B->ml; D->n_tt = B->n_tt; D->nf = B->nf; D->ni = B->ni;
    D->col0 = col0; D->ncols = ncols; D->device = B->device; D->refine = 0; D->chol_status = 0;
    D->d_table = D->d_wv = D->d_fw = D->d_bw = nullptr; D->d_ord = nullptr; D->d_ivlist = nullptr;
    D->d_cfb = D->d_sb = D->d_tf = D->d_tb = nullptr; D->nchunk = 0;
    const i64 nn = B->ni * B->q;
    double *d_g = nullptr; int *d_status = nullptr;
    int rc = CB_OK;
    cudaError_t e = cudaSuccess;
    auto step = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    step(cudaMalloc(&D->d_table, sizeof(double) * (nn * B->k ? nn * B->k : 1)));
    step(cudaMalloc(&D->d_ord, sizeof(int32_t) * B->nt));
    step(cudaMalloc(&D->d_ivlist, sizeof(i64) * (B->ni ? B->ni : 1)));
    if (e == cudaSuccess && B->ni) step(cudaMemcpyAsync(D->d_ivlist, B->d_ivlist, sizeof(i64) * B->ni, cudaMemcpyDeviceToDevice, st));
    step(cudaMalloc(&D->d_fw, sizeof(double) * ncols * B->k));
    step(cudaMalloc(&D->d_bw, sizeof(double) * ncols * B->k));
    step(cudaMalloc(&d_g, sizeof(double) * ncols * (2 * B->k - 1)));
    step(cudaMalloc(&d_status, sizeof(int)));
    if (e == cudaSuccess) step(cudaMemsetAsync(d_status, 0, sizeof(int), st));
    if (e == cudaSuccess) step(cudaMemcpyAsync(D->d_ord, B->d_ord, sizeof(int32_t) * B->nt, cudaMemcpyDeviceToDevice, st));
    if (e == cudaSuccess && wvals_host) {
        step(cudaMalloc(&D->d_wv, sizeof(double) * (nn ? nn : 1)));
        if (e == cudaSuccess) step(cudaMemcpyAsync(D->d_wv, wvals_host, sizeof(double) * nn, cudaMemcpyHostToDevice, st));
        if (e == cudaSuccess) step(cudaStreamSynchronize(st));          // wvals_host may be pageable / freed by the caller
    }
    if (e == cudaSuccess) rc = cb_bspline_quad_table(B, 0, D->d_table, stream);
    if (e == cudaSuccess && rc == CB_OK) rc = cb_bspline_gram_internal(B, col0, ncols, d_g, stream);
    if (e == cudaSuccess && rc == CB_OK) {
        rc = launch_cholesky(B->k, d_g, ncols, D->d_fw, D->d_bw, d_status, st);
        step(cudaGetLastError());
        step(cudaMemcpyAsync(&D->chol_status, d_status, sizeof(int), cudaMemcpyDeviceToHost, st));
        step(cudaStreamSynchronize(st));
    }
    if (e == cudaSuccess && rc == CB_OK && D->chol_status == 0) {
        const int kk = B->k;
        rc = [&]() -> int { CB_DISPATCH_KD(kk, return launch_spikes<K>(D, st)); return CB_OK; }();
        step(cudaStreamSynchronize(st));
    }
    cudaFree(d_g); cudaFree(d_status);
    if (e != cudaSuccess) { cb_set_error("cb_density_bspline_create: %s", cudaGetErrorString(e)); cb_density_destroy(D); return CB_ERR_CUDA; }
    if (rc != CB_OK) { cb_density_destroy(D); return rc; }
    if (D->chol_status != 0) {
        cb_set_error("cb_density_bspline_create: Vandermonde matrix is rank deficient at function %d", D->chol_status);
        cb_density_destroy(D);
        return CB_ERR_ARG;
    }
    *out = D;
    return CB_OK;
}

int cb_density_set_refine(cb_density *D, int steps) {
    CB_REQUIRE(D != nullptr && steps >= 0 && steps <= 4, "cb_density_set_refine: bad argument");
    D->refine = steps;
    return CB_OK;
}

int cb_density_apply(const cb_density *D, const double *f, i64 ldf, i64 Pf,
                     const double *g, i64 ldg, i64 Pg, int conj, double *rho, i64 ldr, void *stream) {
    (void)conj;                                    // real Float64: conj is the identity
    int rc = check_pairs("cb_density_apply", D, f, ldf, Pf, g, ldg, Pg);
    if (rc) return rc;
    CB_REQUIRE(rho != nullptr && ldr >= D->ncols, "cb_density_apply: bad rho array");
    const i64 P = Pf > Pg ? Pf : Pg;
    if (P == 0) return CB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    DensArgs A = make_args(D, nullptr, f, ldf, Pf, g, ldg, Pg);
    A.out = rho; A.ldo = ldr;
    // partitioned plans: the right-hand side kernel runs the solver's local forward sweep on the rows it produces (one
    // pass over the batch less: 2.5 of 21.8 ms at C5)
    double *ws = nullptr, *Tt = nullptr;
    if (D->d_cfb != nullptr && D->nchunk >= 2 && D->k > 1) {
        { int prc = cb_keep_scratch_pool(); if (prc) return prc; }
        CB_CUDA(cudaMallocAsync(&ws, sizeof(double) * 2 * D->nchunk * (D->k - 1) * P, st));
        CB_CUDA(cudaMallocAsync(&Tt, sizeof(double) * (size_t)((P + 31) / 32) * D->ncols * 32, st));
        A.fwt = D->d_fw; A.E = ws; A.nchunk_ds = D->nchunk; A.out_tiled = Tt;
    }
    CB_DISPATCH_KD(D->k, rc = (launch_rhs<K, 0>(A, st)));
    if (rc == CB_OK) CB_DISPATCH_KD(D->k, rc = launch_solve<K>(D, rho, ldr, P, nullptr, 0, st, ws, Tt));
    if (ws) cudaFreeAsync(ws, st);
    if (Tt) cudaFreeAsync(Tt, st);
    A.fwt = nullptr; A.E = nullptr; A.nchunk_ds = 0; A.out_tiled = nullptr;
    if (rc) return rc;
    if (D->refine > 0) {
        double *tmp = nullptr;
        { int prc = cb_keep_scratch_pool(); if (prc) return prc; }
        CB_CUDA(cudaMallocAsync(&tmp, sizeof(double) * D->ncols * P, st));
        for (int it = 0; it < D->refine && rc == CB_OK; ++it) {
            DensArgs R = A;
            R.rho_in = rho; R.ldri = ldr; R.out = tmp; R.ldo = D->ncols;
            CB_DISPATCH_KD(D->k, rc = (launch_rhs<K, 1>(R, st)));
            if (rc) break;
            // tmp <- G^-1 tmp + rho, written to tmp; then copied back
            CB_DISPATCH_KD(D->k, rc = launch_solve<K>(D, tmp, D->ncols, P, rho, ldr, st));
            if (rc) break;
            cudaError_t e = cudaMemcpy2DAsync(rho, sizeof(double) * ldr, tmp, sizeof(double) * D->ncols,
                                              sizeof(double) * D->ncols, P, cudaMemcpyDeviceToDevice, st);
            if (e != cudaSuccess) { cb_set_error("cb_density_apply: %s", cudaGetErrorString(e)); rc = CB_ERR_CUDA; }
        }
        cudaFreeAsync(tmp, st);
    }
    return rc;
}

// Host-pointer form of copyto!(rho, f, g): the pair batch is streamed through the device in tiles on three private
// streams (H2D of tile t+1, the kernels of tile t and D2H of tile t-1 overlap).  f / g with a single column are
// uploaded once.  `stream`: producer stream of the plan (see the *_apply_host entries).  Returns when rho_host is done.
int cb_density_apply_host(const cb_density *D, const double *f_host, i64 ldf, i64 Pf, const double *g_host, i64 ldg, i64 Pg,
                          int conj, double *rho_host, i64 ldr, void *stream) {
    int rc = check_pairs("cb_density_apply_host", D, f_host, ldf, Pf, g_host, ldg, Pg);
    if (rc) return rc;
    CB_REQUIRE(rho_host != nullptr && ldr >= D->ncols, "cb_density_apply_host: bad rho array");
    const i64 P = Pf > Pg ? Pf : Pg, n = D->ncols;
    if (P == 0) return CB_OK;
    constexpr int NB = 3;
    i64 tp = (i64)(48 << 20) / (n * (i64)sizeof(double));           // ~48 MB per array and buffer
    tp = tp < 32 ? 32 : tp / 32 * 32;
    if (tp > P) tp = P;
    cudaStream_t st[NB] = {nullptr, nullptr, nullptr};
    cudaEvent_t ready = nullptr;
    double *df[NB] = {nullptr, nullptr, nullptr}, *dg[NB] = {nullptr, nullptr, nullptr}, *dr[NB] = {nullptr, nullptr, nullptr};
    double *f1 = nullptr, *g1 = nullptr;
    cudaError_t e = cudaSuccess;
    auto step = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    const size_t tile_b = sizeof(double) * (size_t)n * (size_t)tp, col_b = sizeof(double) * (size_t)n;
    // staging buffers come from the stream-ordered pool on the producer stream (kept between calls: no cudaMalloc /
    // cudaFree round trips, which cost more than the transfers of a small batch); the private streams wait for `ready`
    { int prc = cb_keep_scratch_pool(); if (prc) return prc; }
    const cudaStream_t ps = (cudaStream_t)stream;
    for (int b = 0; b < NB; ++b) {
        step(cudaStreamCreateWithFlags(&st[b], cudaStreamNonBlocking));
        if (Pf > 1) step(cudaMallocAsync(&df[b], tile_b, ps));
        if (Pg > 1) step(cudaMallocAsync(&dg[b], tile_b, ps));
        step(cudaMallocAsync(&dr[b], tile_b, ps));
    }
    step(cudaEventCreateWithFlags(&ready, cudaEventDisableTiming));
    if (Pf == 1) { step(cudaMallocAsync(&f1, col_b, ps)); if (e == cudaSuccess) step(cudaMemcpyAsync(f1, f_host, col_b, cudaMemcpyHostToDevice, ps)); }
    if (Pg == 1) { step(cudaMallocAsync(&g1, col_b, ps)); if (e == cudaSuccess) step(cudaMemcpyAsync(g1, g_host, col_b, cudaMemcpyHostToDevice, ps)); }
    if (e == cudaSuccess) step(cudaEventRecord(ready, (cudaStream_t)stream));
    for (int b = 0; b < NB && e == cudaSuccess; ++b) step(cudaStreamWaitEvent(st[b], ready, 0));
    auto copy2d = [&](double *dst, i64 ldd, const double *src, i64 lds, i64 nv, cudaMemcpyKind kind, cudaStream_t s_) {
        if (ldd == n && lds == n) return cudaMemcpyAsync(dst, src, sizeof(double) * n * nv, kind, s_);
        return cudaMemcpy2DAsync(dst, sizeof(double) * ldd, src, sizeof(double) * lds, sizeof(double) * n, nv, kind, s_);
    };
    i64 t = 0;
    for (i64 p0 = 0; p0 < P && e == cudaSuccess && rc == CB_OK; p0 += tp, ++t) {
        const int b = (int)(t % NB);
        const i64 np = p0 + tp <= P ? tp : P - p0;
        if (Pf > 1) step(copy2d(df[b], n, f_host + p0 * ldf, ldf, np, cudaMemcpyHostToDevice, st[b]));
        if (Pg > 1 && e == cudaSuccess) step(copy2d(dg[b], n, g_host + p0 * ldg, ldg, np, cudaMemcpyHostToDevice, st[b]));
        if (e == cudaSuccess)
            rc = cb_density_apply(D, Pf > 1 ? df[b] : f1, n, Pf > 1 ? np : 1, Pg > 1 ? dg[b] : g1, n, Pg > 1 ? np : 1, conj, dr[b], n, st[b]);
        if (e == cudaSuccess && rc == CB_OK) step(copy2d(rho_host + p0 * ldr, ldr, dr[b], n, np, cudaMemcpyDeviceToHost, st[b]));
    }
    for (int b = 0; b < NB; ++b) {
        if (st[b]) { cudaError_t e2 = cudaStreamSynchronize(st[b]); if (e == cudaSuccess) e = e2; cudaStreamDestroy(st[b]); }
        if (df[b]) cudaFreeAsync(df[b], ps);
        if (dg[b]) cudaFreeAsync(dg[b], ps);
        if (dr[b]) cudaFreeAsync(dr[b], ps);
    }
    if (f1) cudaFreeAsync(f1, ps);
    if (g1) cudaFreeAsync(g1, ps);
    if (ready) cudaEventDestroy(ready);
    if (e != cudaSuccess) { cb_set_error("cb_density_apply_host: %s", cudaGetErrorString(e)); return CB_ERR_CUDA; }
    return rc;
}

// B \ f (src/bsplines.jl:329-343): the reference solves the sparse least-squares problem  V c = f(x)  at the
// quadrature nodes (unweighted), i.e. c = (V'V)^-1 V' f(x) — the plan's Gram factor with a plain V' f right-hand side.
int cb_density_project(const cb_density *D, const double *fvals, i64 ldf, i64 P, double *coef, i64 ldc, void *stream) {
    CB_REQUIRE(D != nullptr && fvals != nullptr && coef != nullptr, "cb_density_project: NULL argument");
    CB_REQUIRE(P >= 0 && ldf >= D->ni * D->q && ldc >= D->ncols, "cb_density_project: leading dimension too small");
    if (P == 0) return CB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = CB_OK;
    dim3 grid((unsigned)cb_grid_for(D->ncols, 128, 16), (unsigned)(P < 16384 ? P : 16384));
    CB_DISPATCH_KD(D->k, (density_project_rhs_kernel<K><<<grid, 128, 0, st>>>(D->d_table, D->d_ord, D->n_tt, D->ml, D->q, D->col0,
                                                                                D->ncols, fvals, ldf, P, coef, ldc)));
    CB_LAUNCH_CHECK();
    CB_DISPATCH_KD(D->k, rc = launch_solve<K>(D, coef, ldc, P, nullptr, 0, st));
    return rc;
}

// ---- SPD banded factor / batched solve: the metric inverse of LinearOperator (src/linear_operators.jl:38-48,
// factorize(B'B) + ldiv!) ------------------------------------------------------------------------------------
static int bandchol_create(const double *band, i64 n, int l, int indef, cb_bandchol **out, cudaStream_t st);

int cb_bandchol_create(const double *band, i64 n, int l, cb_bandchol **out, void *stream) { return bandchol_create(band, n, l, 0, out, (cudaStream_t)stream); }
int cb_bandldl_create(const double *band, i64 n, int l, cb_bandchol **out, void *stream) { return bandchol_create(band, n, l, 1, out, (cudaStream_t)stream); }

}  // extern "C"

static int bandchol_create(const double *band, i64 n, int l, int indef, cb_bandchol **out, cudaStream_t st) {
    CB_REQUIRE(out != nullptr, "cb_bandchol_create: out is NULL");
    *out = nullptr;
    CB_REQUIRE(band != nullptr && n >= 1 && l >= 0, "cb_bandchol_create: bad argument");
    if (l + 1 > CB_KMAX) { cb_set_error("cb_bandchol_create: half-bandwidth %d outside compiled range 0..%d", l, CB_KMAX - 1); return CB_ERR_UNSUPPORTED; }
    cb_density_s *D = new cb_density_s();
    D->k = l + 1; D->q = 0; D->ml = 0; D->n_tt = 0; D->nf = n; D->ni = 0; D->col0 = 0; D->ncols = n; D->refine = 0; D->chol_status = 0;
    D->d_table = D->d_wv = D->d_fw = D->d_bw = nullptr; D->d_ord = nullptr; D->d_ivlist = nullptr;
    D->d_cfb = D->d_sb = D->d_tf = D->d_tb = nullptr; D->nchunk = 0;
    cudaGetDevice(&D->device);
    int *d_status = nullptr;
    cudaError_t e = cudaSuccess;
    auto step = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    step(cudaMalloc(&D->d_fw, sizeof(double) * n * D->k));
    step(cudaMalloc(&D->d_bw, sizeof(double) * n * D->k));
    step(cudaMalloc(&d_status, sizeof(int)));
    if (e == cudaSuccess) step(cudaMemsetAsync(d_status, 0, sizeof(int), st));
    int rc = CB_OK;
    if (e == cudaSuccess) {
        rc = launch_cholesky(D->k, band, n, D->d_fw, D->d_bw, d_status, st, indef);
        step(cudaGetLastError());
        step(cudaMemcpyAsync(&D->chol_status, d_status, sizeof(int), cudaMemcpyDeviceToHost, st));
        step(cudaStreamSynchronize(st));
    }
    if (e == cudaSuccess && rc == CB_OK && D->chol_status == 0) {
        const int kk = D->k;
        rc = [&]() -> int { CB_DISPATCH_KD(kk, return launch_spikes<K>(D, st)); return CB_OK; }();
        step(cudaStreamSynchronize(st));
    }
    cudaFree(d_status);
    if (e != cudaSuccess) { cb_set_error("cb_bandchol_create: %s", cudaGetErrorString(e)); cb_density_destroy(D); return CB_ERR_CUDA; }
    if (rc != CB_OK) { cb_density_destroy(D); return rc; }
    if (D->chol_status != 0) {
        // PosDefException of cholesky in the reference / breakdown of the unpivoted L D L'
        cb_set_error(indef ? "cb_bandldl_create: pivot breakdown at column %d (use the pivoted band LU)"
                           : "cb_bandchol_create: matrix is not positive definite (leading minor %d)", D->chol_status);
        cb_density_destroy(D);
        return CB_ERR_ARG;
    }
    *out = reinterpret_cast<cb_bandchol *>(D);
    return CB_OK;
}

extern "C" {

int cb_bandchol_solve(const cb_bandchol *F, double *X, i64 ldx, i64 nvec, void *stream) {
    const cb_density_s *D = reinterpret_cast<const cb_density_s *>(F);
    CB_REQUIRE(D != nullptr && X != nullptr, "cb_bandchol_solve: NULL argument");
    CB_REQUIRE(nvec >= 0 && ldx >= D->ncols, "DimensionMismatch: cb_bandchol_solve: leading dimension %lld smaller than the matrix order %lld",
               (long long)ldx, (long long)D->ncols);
    if (nvec == 0) return CB_OK;
    int rc = CB_OK;
    CB_DISPATCH_KD(D->k, rc = launch_solve<K>(D, X, ldx, nvec, nullptr, 0, (cudaStream_t)stream));
    return rc;
}

int cb_bandchol_destroy(cb_bandchol *F) { return cb_density_destroy(reinterpret_cast<cb_density_s *>(F)); }

int cb_density_nodal_product(const cb_density *D, const double *f, i64 ldf, i64 Pf,
                             const double *g, i64 ldg, i64 Pg, int conj, double *tmp, i64 ldt, void *stream) {
    (void)conj;
    int rc = check_pairs("cb_density_nodal_product", D, f, ldf, Pf, g, ldg, Pg);
    if (rc) return rc;
    CB_REQUIRE(tmp != nullptr && ldt >= D->ni * D->q, "cb_density_nodal_product: bad tmp array");
    const i64 P = Pf > Pg ? Pf : Pg;
    if (P == 0) return CB_OK;
    DensArgs A = make_args(D, nullptr, f, ldf, Pf, g, ldg, Pg);
    A.out = tmp; A.ldo = ldt;
    const i64 nn = D->ni * D->q;
    if (nn == 0) return CB_OK;
    dim3 grid((unsigned)cb_grid_for(nn, 256, 8), (unsigned)(P < 4096 ? P : 4096));
    CB_DISPATCH_KD(D->k, (density_nodal_kernel<K><<<grid, 256, 0, (cudaStream_t)stream>>>(A, D->d_ivlist, D->ni)));
    CB_LAUNCH_CHECK();
    return rc;
}

int cb_density_orthogonal(const double *c, i64 n, const double *f, i64 ldf, i64 Pf,
                          const double *g, i64 ldg, i64 Pg, int conj, double *rho, i64 ldr, void *stream) {
    (void)conj;
    CB_REQUIRE(n >= 0 && f && g && rho, "cb_density_orthogonal: bad argument");
    CB_REQUIRE(Pf == Pg || Pf == 1 || Pg == 1, "DimensionMismatch: Cannot broadcast %lld LHS and %lld RHS to common size",
               (long long)Pf, (long long)Pg);
    CB_REQUIRE(ldf >= n && ldg >= n && ldr >= n, "cb_density_orthogonal: leading dimension smaller than n");
    const i64 P = Pf > Pg ? Pf : Pg;
    if (P == 0 || n == 0) return CB_OK;
    dim3 grid((unsigned)cb_grid_for(n, 256, 4), (unsigned)(P < 8192 ? P : 8192));
    density_orthogonal_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(c, n, f, ldf, Pf, g, ldg, Pg, rho, ldr, P);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

}  // extern "C"
