// bandlu_part.cu — partitioned batched solve with the pivoted band LU factors of bandlu.cu (gbtrs, NoTrans).
//
// bandlu_solve_kernel walks all n rows of 32 vectors in one warp: 1024 vectors keep 32 warps of the whole GPU busy
// (28 ms at 50 007 x 1024, l = u = 7).  Both phases of gbtrs are linear recurrences along the rows, so the scheme of
// density.cu §3b carries over, with the row interchanges inside the chunk-local sweeps:
//
//   forward   x <- L_{n-2} P_{n-2} ... L_0 P_0 x.  Step j touches rows j .. j+l.  Rows are cut into chunks [a, b); the
//             steps of a chunk read rows [a, b + l).  With M_c the (linear) action of a chunk's steps and d_c the
//             modification that earlier chunks have made to its first l rows ("incoming state"),
//                 rows [a, b) after the chunk   =  (M_c x_orig)[a, b)   +  SF_c d_c
//                 d_{c+1}                       =  E_c + TF_c d_c,   E_c = (M_c x_orig)[b, b+l) - x_orig[b, b+l)
//             SF_c (rows x l) and TF_c (l x l) are the chunk's response to unit states: properties of the factors,
//             tabulated at factorisation time.
//   backward  U x = y, bandwidth kv = l + u, rows descending; step j touches rows j-kv .. j.  Same decomposition with
//             states of kv components, tables SB_c, TB_c.
//
//   P0  the first l rows of every chunk are saved (a chunk's local sweep writes its own rows in place, its upper
//       neighbour still needs the original values)
//   F1  every (chunk, 32 vectors): local forward sweep in place, E_c
//   F2  one thread per vector, chunk after chunk: d_{c+1} = E_c + TF_c d_c
//   H2  forward-final values of the last kv rows of every chunk (the lower halo of the backward sweeps)
//   FB  every (chunk, 32 vectors): forward correction SF_c d_c on the fly, local backward sweep in place, Eb_c
//   B2  backward states, as F2 downwards
//   B3  x_j = zloc_j + SB_c[j] . db_c   (no recurrence)
// Three passes over the batch with n / LP_CH x nvec / 32 independent warps each; exact (no truncation).
#include "common.cuh"

constexpr int LP_CH = 64;                          // rows per chunk (the last one may be shorter)
constexpr int LP_MAXS = 32;                        // state components (kv <= 30)

struct LuPart {
    i64 n, nchunk;
    int l, u, kv, ld;
    const double *ab; const int *ipiv;
    const double *SF, *TF, *SB, *TB;               // [n][l], [nchunk][l][l], [n][kv], [nchunk][kv][kv]  (out-major: T[c][out][in])
};

__host__ __device__ inline i64 lp_begin(i64 c) { return c * LP_CH; }
__host__ __device__ inline i64 lp_end(i64 c, i64 nchunk, i64 n) { (void)nchunk; return (c + 1) * LP_CH < n ? (c + 1) * LP_CH : n; }

__device__ __forceinline__ void lp_cp8(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void lp_cp4(void *dst, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
// The factor entries of a chunk's steps are staged in shared memory with all copies in flight at once: read from global
// memory inside the recurrence they put an L2 round trip on every step (15.8 ms at 50 007 x 1024 instead of ~1).
//   fl[s * l + r - 1] = multiplier of row j + r, pv[s] = pivot row of step j = a + s
__device__ __forceinline__ void lp_stage_L(const LuPart &F, double *fl, int *pv, i64 a, int nb, int lane) {
    const int l = F.l;
    for (int e = lane; e < nb * l; e += 32) {
        const int s = e / l, r = e - s * l;
        lp_cp8(fl + e, F.ab + (i64)F.ld * (a + s) + F.kv + 1 + r);
    }
    for (int s = lane; s < nb; s += 32) lp_cp4(pv + s, F.ipiv + a + s);
}
//   fu[s * (kv + 1) + d] = U(j - d, j), j = a + s;  lp_invert_diag turns the diagonal entries into reciprocals
__device__ __forceinline__ void lp_invert_diag(const LuPart &F, double *fu, int nb, int lane) {
    for (int s = lane; s < nb; s += 32) fu[s * (F.kv + 1)] = 1.0 / fu[s * (F.kv + 1)];
}
__device__ __forceinline__ void lp_stage_U(const LuPart &F, double *fu, i64 a, int nb, int lane) {
    const int kv1 = F.kv + 1;
    for (int e = lane; e < nb * kv1; e += 32) {
        const int s = e / kv1, d = e - s * kv1;
        lp_cp8(fu + e, F.ab + (i64)F.ld * (a + s) + F.kv - d);
    }
}

// Steps j = a .. b-1 of the forward phase on the tile xs (row r of the tile <-> global row a + r, lane <-> vector).
// Rows beyond n do not exist (lm clips them).
// LT > 0 (l == u known at compile time): the rows a step can touch live in a register window that slides along the
// chunk — a step is a warp-uniform swap, LT FMAs and one shared-memory load / store off the dependency chain, instead
// of a load -> FMA -> store round trip through shared memory per touched row (~300 cycles per step).
template <int LT>
__device__ __forceinline__ void lp_forward_local(const LuPart &F, double *xs, const double *fl, const int *pv, i64 a, i64 b, int lane) {
    const int l = LT > 0 ? LT : F.l;
    const i64 jend = b < F.n - 1 ? b : F.n - 1;             // steps a .. jend-1
    if constexpr (LT > 0) {
        const int rows = (int)(b - a) + LT;                 // tile rows (rows beyond the matrix hold zeros / are never stored)
        double w[LT + 1];
#pragma unroll
        for (int r = 0; r <= LT; ++r) w[r] = r < rows ? xs[r * 33 + lane] : 0.0;
        const int nst = (int)(jend - a);
        for (int s = 0; s < nst; ++s) {
            const int p = pv[s] - (int)(a + s);
            const double *lj = fl + s * LT - 1;
            const double nxt = s + LT + 1 < rows ? xs[(s + LT + 1) * 33 + lane] : 0.0;
#pragma unroll
            for (int r = 1; r <= LT; ++r)
                if (r == p) { const double t = w[0]; w[0] = w[r]; w[r] = t; }
            const double xj = w[0];
            xs[s * 33 + lane] = xj;
            if (F.n - 1 - (a + s) >= LT) {                   // the usual row: every multiplier exists
#pragma unroll
                for (int r = 1; r <= LT; ++r) w[r - 1] = fma(-lj[r], xj, w[r]);
            } else {
                const int lm = (int)(F.n - 1 - (a + s));
#pragma unroll
                for (int r = 1; r <= LT; ++r) w[r - 1] = r <= lm ? fma(-lj[r], xj, w[r]) : w[r];
            }
            w[LT] = nxt;
        }
#pragma unroll
        for (int r = 0; r <= LT; ++r)
            if (nst + r < rows) xs[(nst + r) * 33 + lane] = w[r];
    } else {
        for (i64 j = a; j < jend; ++j) {
            const int s = (int)(j - a);
            const int lm = (int)(l < F.n - 1 - j ? l : F.n - 1 - j);
            const int p = pv[s] - (int)j;
            const double *lj = fl + s * l - 1;                     // lj[r] = multiplier of row j + r
            double xj = xs[s * 33 + lane];
            if (p != 0) {
                const double xp = xs[(s + p) * 33 + lane];
                xs[(s + p) * 33 + lane] = xj;
                xs[s * 33 + lane] = xp;
                xj = xp;
            }
            for (int r = 1; r <= lm; ++r) xs[(s + r) * 33 + lane] -= lj[r] * xj;
        }
    }
}

// Steps j = b-1 .. a of the backward phase; tile row r <-> global row a - kv + r (rows below 0 do not exist and hold
// zeros).  fu[s][0] holds 1 / U(j, j) (one division per row of the chunk, shared by the 32 vectors).
template <int KVT>
__device__ __forceinline__ void lp_backward_local(const LuPart &F, double *xs, const double *fu, i64 a, i64 b, int lane) {
    const int kv = KVT > 0 ? KVT : F.kv;
    const int nb = (int)(b - a);
    if constexpr (KVT > 0) {
        double w[KVT + 1];                                   // w[d] = row j - d
        const int top = nb - 1 + KVT;
#pragma unroll
        for (int d = 0; d <= KVT; ++d) w[d] = top - d >= 0 ? xs[(top - d) * 33 + lane] : 0.0;
        for (int sj = nb - 1; sj >= 0; --sj) {               // j = a + sj
            const int s = sj + KVT;
            const double *uj = fu + sj * (KVT + 1);
            const double nxt = s - KVT - 1 >= 0 ? xs[(s - KVT - 1) * 33 + lane] : 0.0;
            const double xj = w[0] * uj[0];
            xs[s * 33 + lane] = xj;
            const i64 j = a + sj;
            if (j >= KVT) {                                  // the usual row: no per-entry bound test (two instructions per entry)
#pragma unroll
                for (int d = 1; d <= KVT; ++d) w[d - 1] = fma(-uj[d], xj, w[d]);
            } else {
                const int um = (int)j;                       // rows above the matrix do not exist
#pragma unroll
                for (int d = 1; d <= KVT; ++d) w[d - 1] = d <= um ? fma(-uj[d], xj, w[d]) : w[d];
            }
            w[KVT] = nxt;
        }
#pragma unroll
        for (int d = 0; d < KVT; ++d) xs[(KVT - 1 - d) * 33 + lane] = w[d];
    } else {
        for (i64 j = b - 1; j >= a; --j) {
            const int s = (int)(j - a) + kv;
            const double *uj = fu + (int)(j - a) * (kv + 1);       // uj[d] = U(j - d, j)
            const double xj = xs[s * 33 + lane] * uj[0];
            xs[s * 33 + lane] = xj;
            const int um = (int)(kv < j ? kv : j);
            for (int d = 1; d <= um; ++d) xs[(s - d) * 33 + lane] -= uj[d] * xj;
        }
    }
}

// ---- tables: one warp per chunk, lane i <-> unit state component i --------------------------------------------------
template <int LT>
__global__ void __launch_bounds__(32)
lp_spike_kernel(LuPart F, double *__restrict__ SF, double *__restrict__ TF, double *__restrict__ SB, double *__restrict__ TB) {
    extern __shared__ double xs[];                  // [rows][33], then the staged factor entries
    const int lane = threadIdx.x, l = F.l, kv = F.kv;
    const i64 c = blockIdx.x, a = lp_begin(c), b = lp_end(c, F.nchunk, F.n);
    const int nb = (int)(b - a);
    double *fac = xs + (size_t)(LP_CH + (kv > l ? kv : l)) * 33;
    int *pv = reinterpret_cast<int *>(fac + (size_t)LP_CH * (kv + 1));
    lp_stage_L(F, fac, pv, a, nb, lane);
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    // forward: unit modification of row a + i
    for (int r = 0; r < nb + l; ++r) xs[r * 33 + lane] = (r == lane && lane < l) ? 1.0 : 0.0;
    __syncwarp();
    lp_forward_local<LT>(F, xs, fac, pv, a, b, lane);
    __syncwarp();
    if (lane < l) {
        for (int r = 0; r < nb; ++r) SF[(a + r) * l + lane] = xs[r * 33 + lane];
        for (int r = 0; r < l; ++r) TF[(c * l + r) * l + lane] = (b + r < F.n) ? xs[(nb + r) * 33 + lane] : 0.0;
    }
    __syncwarp();
    lp_stage_U(F, fac, a, nb, lane);
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    lp_invert_diag(F, fac, nb, lane);
    // backward: unit modification of row b - kv + i (the top kv rows of the chunk)
    for (int r = 0; r < nb + kv; ++r) xs[r * 33 + lane] = (r == nb + lane && lane < kv) ? 1.0 : 0.0;
    __syncwarp();
    lp_backward_local<(LT > 0 ? 2 * LT : 0)>(F, xs, fac, a, b, lane);
    __syncwarp();
    if (lane < kv) {
        for (int r = 0; r < nb; ++r) SB[(a + r) * kv + lane] = xs[(kv + r) * 33 + lane];
        for (int r = 0; r < kv; ++r) TB[(c * kv + r) * kv + lane] = (a - kv + r >= 0) ? xs[r * 33 + lane] : 0.0;
    }
}

// rows [g0, g0 + cnt) of vectors v0 .. v0+nv-1 <-> tile rows [s0, s0 + cnt): lanes run along the rows (coalesced)
__device__ __forceinline__ void lp_load_rows(double *xs, const double *X, i64 ldx, i64 v0, int nv, i64 g0, int s0, int cnt, int lane) {
    for (int v = 0; v < nv; ++v) {
        const double *xc = X + (v0 + v) * ldx + g0;
        for (int r = lane; r < cnt; r += 32) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(&xs[(s0 + r) * 33 + v]);
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(xc + r) : "memory");
        }
    }
}
__device__ __forceinline__ void lp_store_rows(const double *xs, double *X, i64 ldx, i64 v0, int nv, i64 g0, int s0, int cnt, int lane) {
    for (int v = 0; v < nv; ++v) {
        double *xc = X + (v0 + v) * ldx + g0;
#pragma unroll 4
        for (int r = lane; r < cnt; r += 32) xc[r] = xs[(s0 + r) * 33 + v];
    }
}

// P0: H[c][r][v] = X[v][a_c + r], r < l   (c >= 1)
__global__ void lp_save_heads_kernel(LuPart F, const double *__restrict__ X, i64 ldx, i64 nvec, double *__restrict__ H) {
    const i64 tot = F.nchunk * F.l * nvec;
    for (i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x; e < tot; e += (i64)gridDim.x * blockDim.x) {
        const i64 v = e % nvec, cr = e / nvec, r = cr % F.l, c = cr / F.l;
        const i64 row = lp_begin(c) + r;
        H[e] = row < F.n ? X[v * ldx + row] : 0.0;
    }
}

// F1: grid = (vector groups, chunks), one warp per CTA
template <int LT>
__global__ void __launch_bounds__(32)
lp_fwd_local_kernel(LuPart F, double *__restrict__ X, i64 ldx, i64 nvec, const double *__restrict__ H, double *__restrict__ E) {
    extern __shared__ double xs[];
    const int lane = threadIdx.x, l = F.l;
    const i64 v0 = (i64)blockIdx.x * 32, c = blockIdx.y, a = lp_begin(c), b = lp_end(c, F.nchunk, F.n);
    const int nv = (int)(nvec - v0 < 32 ? nvec - v0 : 32), nb = (int)(b - a);
    double *fac = xs + (size_t)(LP_CH + (F.kv > l ? F.kv : l)) * 33;
    int *pv = reinterpret_cast<int *>(fac + (size_t)LP_CH * (F.kv + 1));
    lp_load_rows(xs, X, ldx, v0, nv, a, 0, nb, lane);
    lp_stage_L(F, fac, pv, a, nb, lane);
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    const bool live = lane < nv;
    // original values of the next chunk's first l rows (zero beyond the matrix)
    double h[LP_MAXS];
    for (int r = 0; r < l; ++r) {
        h[r] = (c + 1 < F.nchunk && live) ? H[((c + 1) * l + r) * nvec + v0 + lane] : 0.0;
        xs[(nb + r) * 33 + lane] = h[r];
    }
    __syncwarp();
    lp_forward_local<LT>(F, xs, fac, pv, a, b, lane);
    __syncwarp();
    if (live && c + 1 < F.nchunk)
        for (int r = 0; r < l; ++r) E[(c * l + r) * nvec + v0 + lane] = xs[(nb + r) * 33 + lane] - h[r];
    lp_store_rows(xs, X, ldx, v0, nv, a, 0, nb, lane);
}

// F2 / B2: S <- states.  dir = +1: S_0 = 0, S_{c+1} = E_c + T_c S_c;  dir = -1: S_{M-1} = 0, S_{c-1} = E_c + T_c S_c.
// E[c] is read before S overwrites a slot: S_{c+dir} is stored in E's slot c.  The chain of nchunk - 1 dependent
// products is latency bound, so it is spread over the components: one thread per (vector, state component), G lanes
// per vector (S padded to a power of two), the old state gathered by shuffles; the next chunk's E is fetched while
// the current product is formed.  (One thread per vector with the state in local memory took 4.3 + 12 ms at C5.)
constexpr int LP_SB = 8;                            // chunks staged per batch
template <int G>
__global__ void __launch_bounds__(128)
lp_state_kernel(const double *__restrict__ T, i64 nchunk, int S, i64 nvec, double *__restrict__ E, int dir) {
    extern __shared__ double ssm[];                 // [2][LP_SB][S*S] transfer matrices, [2][LP_SB][128] this block's E entries
    const int tid = threadIdx.x, lane = tid & 31, g = lane % G, vi = lane / G;
    const i64 v = ((i64)blockIdx.x * 4 + (tid >> 5)) * (32 / G) + vi;
    const bool ok = g < S && v < nvec;
    const i64 c0 = dir > 0 ? 0 : nchunk - 1, steps = nchunk - 1;
    const int SS = S * S;
    double *sT = ssm, *sE = ssm + 2 * LP_SB * SS;
    // a step of the chain takes ~150 cycles, an L2 round trip ~800: E and T of the next LP_SB chunks are copied to
    // shared memory (cp.async) while the current batch is multiplied
    auto issue = [&](i64 base, int buf) {
        const int nb = (int)(steps - base < LP_SB ? steps - base : LP_SB);
        for (int e = tid; e < nb * SS; e += 128) {
            const int q = e / SS;
            lp_cp8(sT + (buf * LP_SB + q) * SS + (e - q * SS), T + (c0 + dir * (base + q)) * SS + (e - q * SS));
        }
        if (ok)
            for (int q = 0; q < nb; ++q) lp_cp8(sE + (buf * LP_SB + q) * 128 + tid, E + ((c0 + dir * (base + q)) * S + g) * nvec + v);
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    double s = 0.0;
    if (steps > 0) issue(0, 0);
    int buf = 0;
    for (i64 base = 0; base < steps; base += LP_SB, buf ^= 1) {
        if (base + LP_SB < steps) issue(base + LP_SB, buf ^ 1);
        else asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        __syncthreads();
        const int nb = (int)(steps - base < LP_SB ? steps - base : LP_SB);
        for (int q = 0; q < nb; ++q) {
            const i64 c = c0 + dir * (base + q);
            const double *tr = sT + (buf * LP_SB + q) * SS + (g < S ? g : 0) * S;
            double a0 = ok ? sE[(buf * LP_SB + q) * 128 + tid] : 0.0, a1 = 0.0;
#pragma unroll
            for (int i = 0; i < G; i += 2) {
                const double s0 = __shfl_sync(0xffffffffu, s, vi * G + i), s1 = __shfl_sync(0xffffffffu, s, vi * G + i + 1);
                if (i < S) a0 = fma(tr[i], s0, a0);
                if (i + 1 < S) a1 = fma(tr[i + 1], s1, a1);
            }
            s = g < S ? a0 + a1 : 0.0;
            if (ok) E[(c * S + g) * nvec + v] = s;          // slot c now holds the state of chunk c + dir
        }
        __syncthreads();
    }
}
// state of chunk c after lp_state_kernel: forward (dir +1) -> slot c - 1 (chunk 0: zero); backward -> slot c + 1 (last: zero)

// H2: forward-final values of the last kv rows of every chunk: H2[c][r][v] = X[v][b_c - kv + r] + SF[row] . d_c
__global__ void lp_tails_kernel(LuPart F, const double *__restrict__ X, i64 ldx, i64 nvec, const double *__restrict__ St,
                                double *__restrict__ H2) {
    const i64 tot = F.nchunk * F.kv * nvec;
    const int l = F.l, kv = F.kv;
    for (i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x; e < tot; e += (i64)gridDim.x * blockDim.x) {
        const i64 v = e % nvec, cr = e / nvec, r = cr % kv, c = cr / kv;
        const i64 row = lp_end(c, F.nchunk, F.n) - kv + r;
        double y = 0.0;
        if (row >= 0) {
            y = X[v * ldx + row];
            if (c > 0)
                for (int i = 0; i < l; ++i) y = fma(__ldg(F.SF + row * l + i), St[((c - 1) * l + i) * nvec + v], y);
        }
        H2[e] = y;
    }
}

// FB: forward correction + local backward sweep.  grid = (vector groups, chunks)
template <int LT>
__global__ void __launch_bounds__(32)
lp_bwd_local_kernel(LuPart F, double *__restrict__ X, i64 ldx, i64 nvec, const double *__restrict__ St,
                    const double *__restrict__ H2, double *__restrict__ Eb) {
    extern __shared__ double xs[];
    const int lane = threadIdx.x, l = F.l, kv = F.kv;
    const i64 v0 = (i64)blockIdx.x * 32, c = blockIdx.y, a = lp_begin(c), b = lp_end(c, F.nchunk, F.n);
    const int nv = (int)(nvec - v0 < 32 ? nvec - v0 : 32), nb = (int)(b - a);
    double *fac = xs + (size_t)(LP_CH + (kv > l ? kv : l)) * 33;
    double *sfs = fac + (size_t)LP_CH * (kv + 1) + LP_CH / 2;      // (after the pivot area of the forward kernel)
    lp_load_rows(xs, X, ldx, v0, nv, a, kv, nb, lane);
    lp_stage_U(F, fac, a, nb, lane);
    if (c > 0)
        for (int e = lane; e < nb * l; e += 32) lp_cp8(sfs + e, F.SF + a * l + e);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const bool live = lane < nv;
    double d[LP_MAXS], h[LP_MAXS];
    for (int i = 0; i < l; ++i) d[i] = (c > 0 && live) ? St[((c - 1) * l + i) * nvec + v0 + lane] : 0.0;
    for (int r = 0; r < kv; ++r) {
        h[r] = (c > 0 && live) ? H2[((c - 1) * kv + r) * nvec + v0 + lane] : 0.0;
        xs[r * 33 + lane] = h[r];
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    lp_invert_diag(F, fac, nb, lane);
    if (c > 0) {
        if constexpr (LT > 0) {                              // l = LT: the state stays in registers, the sum unrolls
            double dr[LT];
#pragma unroll
            for (int i = 0; i < LT; ++i) dr[i] = d[i];
            for (int r = 0; r < nb; ++r) {
                const double *sf = sfs + r * LT;
                double y = xs[(kv + r) * 33 + lane];
#pragma unroll
                for (int i = 0; i < LT; ++i) y = fma(sf[i], dr[i], y);
                xs[(kv + r) * 33 + lane] = y;
            }
        } else {
            for (int r = 0; r < nb; ++r) {
                const double *sf = sfs + r * l;
                double y = xs[(kv + r) * 33 + lane];
                for (int i = 0; i < l; ++i) y = fma(sf[i], d[i], y);
                xs[(kv + r) * 33 + lane] = y;
            }
        }
    }
    __syncwarp();
    lp_backward_local<(LT > 0 ? 2 * LT : 0)>(F, xs, fac, a, b, lane);
    __syncwarp();
    if (live && c > 0)
        for (int r = 0; r < kv; ++r) Eb[(c * kv + r) * nvec + v0 + lane] = xs[r * 33 + lane] - h[r];
    lp_store_rows(xs, X, ldx, v0, nv, a, kv, nb, lane);
}

// B3: x_j = zloc_j + SB[j] . db_c.  A CTA takes 256 rows (four chunks) and 32 vectors: the chunks' incoming states of
// those vectors sit in shared memory ([chunk][vector][KVP], zero-padded: 16-byte broadcast reads), a thread keeps the
// spike row of its row in registers (KVP is a compile-time bound: with the run-time kv the row lived in local memory
// and every FMA fetched its state component with a global load — 179 M warp instructions for 0.8 GB), and reads eight
// vectors before it stores any (X is read and written through the same pointer).
constexpr int LF_V = 32, LF_U = 8;
template <int KVP>
__global__ void __launch_bounds__(256, 2)
lp_final_kernel(LuPart F, double *__restrict__ X, i64 ldx, i64 nvec, const double *__restrict__ Sb) {
    __shared__ __align__(16) double sst[(256 / LP_CH) * LF_V * KVP];
    const int kv = F.kv;
    const i64 nrb = (F.n + 255) / 256, nvg = (nvec + LF_V - 1) / LF_V;
    for (i64 item = blockIdx.x; item < nrb * nvg; item += gridDim.x) {
        const i64 rb = item % nrb, vg = item / nrb;
        const i64 va = vg * LF_V;
        const int nv = (int)(va + LF_V < nvec ? LF_V : nvec - va);
        const i64 c0 = rb * (256 / LP_CH);          // first chunk of the row block
        __syncthreads();
        for (int e = threadIdx.x; e < (256 / LP_CH) * LF_V * KVP; e += 256) {
            const int lc = e / (LF_V * KVP), v = (e / KVP) % LF_V, i = e % KVP;
            const i64 c = c0 + lc;                  // the state that enters chunk c is that of chunk c + 1
            sst[e] = (i < kv && v < nv && c + 1 < F.nchunk) ? Sb[((c + 1) * kv + i) * nvec + va + v] : 0.0;
        }
        __syncthreads();
        const i64 row = rb * 256 + threadIdx.x;
        if (row >= F.n) continue;
        const int lc = threadIdx.x / LP_CH;
        if (c0 + lc + 1 >= F.nchunk) continue;      // the last chunk has no incoming state
        double sb[KVP];
#pragma unroll
        for (int i = 0; i < KVP; ++i) sb[i] = i < kv ? __ldg(F.SB + row * kv + i) : 0.0;
        const double *sc = sst + (size_t)lc * LF_V * KVP;
        for (int vb = 0; vb < nv; vb += LF_U) {
            double x[LF_U];
#pragma unroll
            for (int u = 0; u < LF_U; ++u) x[u] = vb + u < nv ? X[(va + vb + u) * ldx + row] : 0.0;
#pragma unroll
            for (int u = 0; u < LF_U; ++u) {
                const double2 *s2 = reinterpret_cast<const double2 *>(sc + (vb + u) * KVP);
#pragma unroll
                for (int h = 0; h < KVP / 2; ++h) { const double2 t = s2[h]; x[u] = fma(sb[2 * h], t.x, x[u]); x[u] = fma(sb[2 * h + 1], t.y, x[u]); }
            }
#pragma unroll
            for (int u = 0; u < LF_U; ++u)
                if (vb + u < nv) X[(va + vb + u) * ldx + row] = x[u];
        }
    }
}

// ---- host -------------------------------------------------------------------------------------------------------------
struct cb_bandlu_part {
    i64 n, nchunk;
    int l, u;
    double *SF, *TF, *SB, *TB;
};

// tile [LP_CH + reach][33] | factor entries [LP_CH][kv + 1] | pivots [LP_CH] ints (LP_CH / 2 doubles) | spikes [LP_CH][l]
static size_t lp_smem(int l, int kv) {
    return sizeof(double) * ((size_t)(LP_CH + (kv > l ? kv : l)) * 33 + (size_t)LP_CH * (kv + 1) + LP_CH / 2 + (size_t)LP_CH * l);
}

bool cb_bandlu_part_applicable(i64 n, int l, int u) { return l >= 1 && l + u <= 30 && l + u <= LP_CH && n >= 2 * (i64)LP_CH; }

void cb_bandlu_part_destroy(cb_bandlu_part *P) {
    if (!P) return;
    cudaFree(P->SF); cudaFree(P->TF); cudaFree(P->SB); cudaFree(P->TB);
    delete P;
}

static LuPart lp_view(const cb_bandlu_part *P, const double *ab, const int *ipiv) {
    LuPart F;
    F.n = P->n; F.nchunk = P->nchunk; F.l = P->l; F.u = P->u; F.kv = P->l + P->u; F.ld = 2 * P->l + P->u + 1;
    F.ab = ab; F.ipiv = ipiv; F.SF = P->SF; F.TF = P->TF; F.SB = P->SB; F.TB = P->TB;
    return F;
}

#define CB_LP_DISPATCH(L_, U_, ...)                                          \
    switch ((L_) == (U_) && (L_) <= 15 ? (L_) : 0) {                          \
        case 1: { constexpr int LT = 1; __VA_ARGS__; } break;                        \
        case 2: { constexpr int LT = 2; __VA_ARGS__; } break;                        \
        case 3: { constexpr int LT = 3; __VA_ARGS__; } break;                        \
        case 4: { constexpr int LT = 4; __VA_ARGS__; } break;                        \
        case 5: { constexpr int LT = 5; __VA_ARGS__; } break;                        \
        case 6: { constexpr int LT = 6; __VA_ARGS__; } break;                        \
        case 7: { constexpr int LT = 7; __VA_ARGS__; } break;                        \
        case 8: { constexpr int LT = 8; __VA_ARGS__; } break;                        \
        case 9: { constexpr int LT = 9; __VA_ARGS__; } break;                        \
        case 10: { constexpr int LT = 10; __VA_ARGS__; } break;                      \
        default: { constexpr int LT = 0; __VA_ARGS__; } break;                       \
    }

static void lp_launch_state(const double *T, i64 M, int S, i64 nvec, double *E, int dir, cudaStream_t st) {
    const int G = S <= 4 ? 4 : S <= 8 ? 8 : S <= 16 ? 16 : 32;
    const i64 vpb = 4 * (32 / G);                              // vectors per block of 4 warps
    const unsigned grid = (unsigned)((nvec + vpb - 1) / vpb);
    const size_t smem = sizeof(double) * (2 * LP_SB * (size_t)S * S + 2 * LP_SB * 128);       // <= 131 KB at S = 30
    switch (G) {
#define CB_LP_STATE(GG) case GG: cudaFuncSetAttribute(lp_state_kernel<GG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
                                 lp_state_kernel<GG><<<grid, 128, smem, st>>>(T, M, S, nvec, E, dir); break;
        CB_LP_STATE(4) CB_LP_STATE(8) CB_LP_STATE(16)
        default: cudaFuncSetAttribute(lp_state_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
                 lp_state_kernel<32><<<grid, 128, smem, st>>>(T, M, S, nvec, E, dir); break;
#undef CB_LP_STATE
    }
}

// Tables of the partitioned solve for factors ab / ipiv (device).  Synchronises st.
int cb_bandlu_part_create(const double *ab, const int *ipiv, i64 n, int l, int u, cb_bandlu_part **out, cudaStream_t st) {
    *out = nullptr;
    cb_bandlu_part *P = new cb_bandlu_part();
    P->n = n; P->l = l; P->u = u; P->nchunk = (n + LP_CH - 1) / LP_CH;
    P->SF = P->TF = P->SB = P->TB = nullptr;
    const int kv = l + u;
    cudaError_t e = cudaMalloc(&P->SF, sizeof(double) * (size_t)n * l);
    if (e == cudaSuccess) e = cudaMalloc(&P->TF, sizeof(double) * (size_t)P->nchunk * l * l);
    if (e == cudaSuccess) e = cudaMalloc(&P->SB, sizeof(double) * (size_t)n * kv);
    if (e == cudaSuccess) e = cudaMalloc(&P->TB, sizeof(double) * (size_t)P->nchunk * kv * kv);
    if (e != cudaSuccess) { cb_set_error("cb_bandlu_create: %s", cudaGetErrorString(e)); cb_bandlu_part_destroy(P); return CB_ERR_CUDA; }
    const LuPart F = lp_view(P, ab, ipiv);
    const size_t smem = lp_smem(l, kv);
    CB_LP_DISPATCH(l, u, {
        auto kern = lp_spike_kernel<LT>;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) kern<<<(unsigned)P->nchunk, 32, smem, st>>>(F, P->SF, P->TF, P->SB, P->TB);
    });
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { cb_set_error("cb_bandlu_create: %s", cudaGetErrorString(e)); cb_bandlu_part_destroy(P); return CB_ERR_CUDA; }
    *out = P;
    return CB_OK;
}

// X <- A^-1 X, n x nvec column-major, in place; asynchronous on st.
int cb_bandlu_part_solve(const cb_bandlu_part *P, const double *ab, const int *ipiv, double *X, i64 ldx, i64 nvec, cudaStream_t st) {
    const LuPart F = lp_view(P, ab, ipiv);
    const int l = P->l, kv = l + P->u;
    const i64 M = P->nchunk;
    { int prc = cb_keep_scratch_pool(); if (prc) return prc; }
    // scratch: H [M][l][nvec], E / forward states [M][l][nvec], H2 [M][kv][nvec], Eb / backward states [M][kv][nvec]
    const size_t per = (size_t)M * nvec;
    double *scr = nullptr;
    CB_CUDA(cudaMallocAsync(&scr, sizeof(double) * per * (size_t)(2 * l + 2 * kv), st));
    double *H = scr, *E = H + per * l, *H2 = E + per * l, *Eb = H2 + per * kv;
    const size_t smem = lp_smem(l, kv);
    const unsigned ngroups = (unsigned)((nvec + 31) / 32);
    const dim3 grid(ngroups, (unsigned)M);
    int rc = CB_OK;
    lp_save_heads_kernel<<<cb_grid_for((i64)per * l, 256, 16), 256, 0, st>>>(F, X, ldx, nvec, H);
    CB_LP_DISPATCH(l, P->u, {
        auto k1 = lp_fwd_local_kernel<LT>;
        auto k2 = lp_bwd_local_kernel<LT>;
        cudaError_t e = cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { cb_set_error("cb_bandlu_solve: %s", cudaGetErrorString(e)); rc = CB_ERR_CUDA; }
        else {
            k1<<<grid, 32, smem, st>>>(F, X, ldx, nvec, H, E);
            lp_launch_state(P->TF, M, l, nvec, E, +1, st);
            lp_tails_kernel<<<cb_grid_for((i64)per * kv, 256, 16), 256, 0, st>>>(F, X, ldx, nvec, E, H2);
            k2<<<grid, 32, smem, st>>>(F, X, ldx, nvec, E, H2, Eb);
            lp_launch_state(P->TB, M, kv, nvec, Eb, -1, st);
            const i64 nitem = ((P->n + 255) / 256) * ((nvec + LF_V - 1) / LF_V);
            const unsigned gf = (unsigned)(nitem < (i64)CB_NUM_SMS * 16 ? nitem : (i64)CB_NUM_SMS * 16);
            if (kv <= 8) lp_final_kernel<8><<<gf, 256, 0, st>>>(F, X, ldx, nvec, Eb);
            else if (kv <= 16) lp_final_kernel<16><<<gf, 256, 0, st>>>(F, X, ldx, nvec, Eb);
            else lp_final_kernel<32><<<gf, 256, 0, st>>>(F, X, ldx, nvec, Eb);
        }
    });
    cudaFreeAsync(scr, st);
    if (rc) return rc;
    CB_LAUNCH_CHECK();
    return CB_OK;
}
