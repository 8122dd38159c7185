// bandlu.cu — factorize(A) and ldiv! for a general BandedMatrix: banded LU with partial pivoting.
//
// Reference: ShiftAndInvert's constructor, `A⁻¹ = factorize(A - σ*B)` (src/linear_operators.jl:203, :214), and its
// `mul!` (`ldiv!(y, M.A⁻¹, M.temp)`, :229-236).  For a BandedMatrix the reference's `factorize` is BandedMatrices'
// banded LU, i.e. LAPACK gbtrf / gbtrs (un-vendored: BandedMatrices 0.16 -> LAPACK); this file restates those two
// routines (unblocked dgbtf2, dgbtrs with NoTrans) for the device:
//   * the factorisation is one warp walking the columns with the active window of l+u+1 columns in shared memory
//     (a one-time cost per operator, like the banded Cholesky of density.cu);
//   * the solve is batched over the vectors of an n x nvec column-major batch: one warp per 32 vectors (lane <->
//     vector), rows staged through shared memory in chunks so that global memory is read and written coalesced along
//     the rows, the substitution itself running on the chunk in shared memory.
//   * matrices of at least 256 rows are solved by the partitioned form of the same substitution (bandlu_part.cu): chunks
//     of rows x 32 vectors are independent warps, the states between chunks travel through small tabulated transfer
//     matrices.
// This path serves the indefinite shifts of interior-eigenvalue problems (docs/src/bsplines/eigenproblems.md:96-110)
// and the non-symmetric left-hand sides of the non-uniform compact finite-difference scheme.
#include "common.cuh"

struct cb_bandlu_part;                              // bandlu_part.cu: tables of the partitioned solve
bool cb_bandlu_part_applicable(i64 n, int l, int u);
int cb_bandlu_part_create(const double *ab, const int *ipiv, i64 n, int l, int u, cb_bandlu_part **out, cudaStream_t st);
int cb_bandlu_part_solve(const cb_bandlu_part *P, const double *ab, const int *ipiv, double *X, i64 ldx, i64 nvec, cudaStream_t st);
void cb_bandlu_part_destroy(cb_bandlu_part *P);

struct cb_bandlu_s {
    i64 n;
    int l, u, ld;                                   // ld = 2l + u + 1 rows per column: [l fill | u super | diag | l sub]
    double *d_ab;                                   // factors, LAPACK band layout: AB[kv + i - j + ld * j], kv = l + u
    int *d_ipiv;                                    // 0-based pivot rows
    int info;                                       // 0, or j+1 of the first exactly zero pivot
    int device;
    cb_bandlu_part *part;                           // non-NULL: solves run partitioned (long matrices)
};

constexpr int LU_NB = 64;                           // columns finished per slide of the factorisation window

// One warp.  band: BandedMatrices layout, band[(u + i - j) + (l+u+1) * j].
__global__ void bandlu_factor_kernel(const double *__restrict__ band, i64 n, int l, int u,
                                     double *__restrict__ ab, int *__restrict__ ipiv, int *__restrict__ info) {
    extern __shared__ double win[];                 // [LU_NB + kv][ld]
    const int lane = threadIdx.x;
    const int kv = l + u, ld = 2 * l + u + 1, W = l + u + 1, WC = LU_NB + kv;
    auto load_cols = [&](i64 c_first, int w_first, int count) {      // global columns c_first.. into window slots w_first..
        for (int e = lane; e < count * ld; e += 32) {
            const int wc = e / ld, r = e - wc * ld;
            const i64 c = c_first + wc;
            double v = 0.0;
            if (c < n && r >= l) v = band[(r - l) + (i64)W * c];     // rows l.. of AB are the band rows 0..l+u
            win[(w_first + wc) * ld + r] = v;
        }
    };
    auto store_cols = [&](i64 c_first, int w_first, int count) {
        for (int e = lane; e < count * ld; e += 32) {
            const int wc = e / ld, r = e - wc * ld;
            const i64 c = c_first + wc;
            if (c < n) ab[(i64)ld * c + r] = win[(w_first + wc) * ld + r];
        }
    };
    i64 c0 = 0;                                     // global column of window slot 0
    load_cols(0, 0, WC);
    __syncwarp();
    i64 ju = 0;                                     // last column touched by the row interchanges so far (LAPACK's JU)
    int first_zero = 0;
    for (i64 j = 0; j < n; ++j) {
        if (j - c0 == LU_NB) {                      // slide the window by LU_NB columns
            store_cols(c0, 0, LU_NB);
            __syncwarp();
            for (int e = lane; e < kv * ld; e += 32) win[e] = win[LU_NB * ld + e];   // kv < LU_NB: source and target are disjoint
            __syncwarp();
            c0 += LU_NB;
            load_cols(c0 + kv, kv, LU_NB);
            __syncwarp();
        }
        const int wj = (int)(j - c0);
        double *cj = win + wj * ld;
        const int km = (int)(l < n - 1 - j ? l : n - 1 - j);
        // pivot: first row of maximal modulus among rows j .. j+km (idamax)
        double av = lane <= km ? fabs(cj[kv + lane]) : -1.0;
        int ai = lane;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, av, o);
            const int oi = __shfl_xor_sync(0xffffffffu, ai, o);
            if (ov > av || (ov == av && oi < ai)) { av = ov; ai = oi; }
        }
        const int jp = ai;
        if (lane == 0) ipiv[j] = (int)(j + jp);
        if (av == 0.0) {                            // exactly singular: LAPACK records the column and goes on
            if (first_zero == 0) first_zero = (int)(j + 1 < 0x7fffffff ? j + 1 : 0x7fffffff);
            continue;
        }
        {
            const i64 reach = j + u + jp < n - 1 ? j + u + jp : n - 1;
            if (reach > ju) ju = reach;
        }
        const int nc = (int)(ju - j);               // columns j+1 .. ju take part in the update (nc <= kv)
        if (jp != 0 && lane <= nc) {                // interchange rows j and j+jp over columns j .. ju
            double *cc = win + (wj + lane) * ld;
            const double a = cc[kv - lane], b = cc[kv + jp - lane];
            cc[kv - lane] = b; cc[kv + jp - lane] = a;
        }
        __syncwarp();
        const double rp = 1.0 / cj[kv];
        if (lane >= 1 && lane <= km) cj[kv + lane] *= rp;            // multipliers
        __syncwarp();
        if (lane >= 1 && lane <= nc) {              // rank-one update of the trailing columns, lane <-> column j + lane
            double *cc = win + (wj + lane) * ld;
            const double ujc = cc[kv - lane];       // U(j, j + lane)
            for (int r = 1; r <= km; ++r) cc[kv + r - lane] -= cj[kv + r] * ujc;
        }
        __syncwarp();
    }
    store_cols(c0, 0, WC);
    if (lane == 0) *info = first_zero;
}

// Batched gbtrs (NoTrans): X <- A^-1 X in place.  One warp per 32 vectors, lane <-> vector; chunk of LU_R rows (+ the
// rows a step can reach ahead) staged in shared memory as xs[row][33] (lane-contiguous: conflict-free for the
// substitution, and for the row-wise movers up to a 2-way conflict).
constexpr int LU_R = 96;                            // rows per chunk
constexpr int LU_WARPS = 4;                         // independent warps (vector groups) per CTA

// LT > 0: l = LT and kv = KVT at compile time (symmetric bandwidths l = u, kv = 2l): the updates of a row are then
// unrolled with all their shared-memory loads issued before the first FMA, instead of one load -> FMA -> store round
// trip after the other (58 ms -> a few ms at 50 007 x 1024, l = u = 7).
template <int LT, int KVT>
__global__ void __launch_bounds__(32 * LU_WARPS)
bandlu_solve_kernel(const double *__restrict__ ab, const int *__restrict__ ipiv, i64 n, int l_rt, int u,
                    double *__restrict__ X, i64 ldx, i64 nvec) {
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int l = LT > 0 ? LT : l_rt;
    const int kv = LT > 0 ? KVT : l + u, ld = 2 * l + u + 1;
    const int reach = kv > l ? kv : l;              // rows beyond the chunk a step touches (forward: l, backward: kv)
    const int rows_s = LU_R + reach;
    const int fw = kv + 1 > l ? kv + 1 : l;         // factor entries staged per row: l multipliers or kv + 1 entries of U
    const size_t per_warp = (size_t)rows_s * 33 + (size_t)LU_R * fw + LU_R / 2;
    double *xs = sm + (size_t)warp * per_warp;
    // the factor columns and pivots of the chunk's rows, staged per chunk with coalesced, independent loads: read per row
    // from global memory they put two L2 round trips on every step of the recurrence (130 ms for 50 007 x 1024)
    double *fac = xs + (size_t)rows_s * 33;
    int *piv = reinterpret_cast<int *>(fac + (size_t)LU_R * fw);
    const i64 ngroups = (nvec + 31) / 32;
    for (i64 grp = (i64)blockIdx.x * LU_WARPS + warp; grp < ngroups; grp += (i64)gridDim.x * LU_WARPS) {
        const i64 v0 = grp * 32;
        const int nv = (int)(nvec - v0 < 32 ? nvec - v0 : 32);
        // rows [g0, g0 + cnt) of the group's vectors <-> xs rows [s0, s0 + cnt); lanes run along the rows
        // (cp.async: all of a chunk's loads are in flight at once; a plain load-then-store loop had one load per lane
        // in flight and spent ~700 cycles of memory latency per row of the recurrence)
        auto load_rows = [&](i64 g0, int s0, int cnt) {
            for (int v = 0; v < nv; ++v) {
                const double *xc = X + (v0 + v) * ldx + g0;
                for (int r = lane; r < cnt; r += 32) {
                    const unsigned dst = (unsigned)__cvta_generic_to_shared(&xs[(s0 + r) * 33 + v]);
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(xc + r) : "memory");
                }
            }
            asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
        };
        auto store_rows = [&](i64 g0, int s0, int cnt) {
            for (int v = 0; v < nv; ++v) {
                double *xc = X + (v0 + v) * ldx + g0;
#pragma unroll 4
                for (int r = lane; r < cnt; r += 32) xc[r] = xs[(s0 + r) * 33 + v];
            }
        };
        // ---- forward: L with the row interchanges, top to bottom --------------------------------
        // xs row 0 <-> global row c0; rows [c0, c0 + have) are resident
        if (l > 0) {
            i64 c0 = 0;
            int have = (int)(n < rows_s ? n : rows_s);
            load_rows(0, 0, have);
            __syncwarp();
            auto stage_L = [&](i64 r0) {            // multipliers and (absolute) pivots of rows r0 .. r0 + LU_R - 1, all in flight
                for (int e = lane; e < LU_R * l; e += 32) {
                    const int rr = e / l, i = e - rr * l;
                    const i64 jj = r0 + rr;
                    if (jj < n) {
                        const unsigned dst = (unsigned)__cvta_generic_to_shared(&fac[rr * l + i]);
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(ab + (i64)ld * jj + kv + 1 + i) : "memory");
                    } else {
                        fac[rr * l + i] = 0.0;
                    }
                }
                for (int rr = lane; rr < LU_R; rr += 32) {
                    if (r0 + rr < n) {
                        const unsigned dst = (unsigned)__cvta_generic_to_shared(&piv[rr]);
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(ipiv + r0 + rr) : "memory");
                    } else {
                        piv[rr] = (int)(r0 + rr);
                    }
                }
                asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
            };
            stage_L(0);
            __syncwarp();
            for (i64 j = 0; j < n - 1; ++j) {
                if (j - c0 == LU_R) {               // rows [c0, c0 + LU_R) are final: write them out, move the rest up
                    store_rows(c0, 0, LU_R);
                    __syncwarp();
                    const int keep = have - LU_R;   // = reach (or fewer at the end)
                    for (int r = 0; r < keep; ++r) xs[r * 33 + lane] = xs[(r + LU_R) * 33 + lane];
                    c0 += LU_R;
                    const i64 next = c0 + keep;
                    const int more = (int)(n - next < LU_R ? (n - next > 0 ? n - next : 0) : LU_R);
                    __syncwarp();
                    load_rows(next, keep, more);
                    stage_L(c0);
                    have = keep + more;
                    __syncwarp();
                }
                const int s = (int)(j - c0);
                const int lm = (int)(l < n - 1 - j ? l : n - 1 - j);
                const int p = piv[s] - (int)j;      // 0 .. l
                double xj = xs[s * 33 + lane];
                if (p != 0) {
                    const double xp = xs[(s + p) * 33 + lane];
                    xs[(s + p) * 33 + lane] = xj;
                    xs[s * 33 + lane] = xp;
                    xj = xp;
                }
                const double *lj = fac + s * l - 1;
                if (LT > 0) {
                    double xv[LT > 0 ? LT : 1];
#pragma unroll
                    for (int r = 1; r <= LT; ++r) xv[r - 1] = r <= lm ? xs[(s + r) * 33 + lane] : 0.0;
#pragma unroll
                    for (int r = 1; r <= LT; ++r) xv[r - 1] = fma(-lj[r], xj, xv[r - 1]);
#pragma unroll
                    for (int r = 1; r <= LT; ++r) if (r <= lm) xs[(s + r) * 33 + lane] = xv[r - 1];
                } else {
                    for (int r = 1; r <= lm; ++r) xs[(s + r) * 33 + lane] -= lj[r] * xj;
                }
            }
            __syncwarp();
            store_rows(c0, 0, have);
            __syncwarp();
        }
        // ---- backward: U (bandwidth kv), bottom to top ------------------------------------------
        // xs row (rows_s - 1) <-> global row hi; rows (hi - have, hi] are resident
        {
            i64 hi = n - 1;
            int have = (int)(n < rows_s ? n : rows_s);
            load_rows(hi - have + 1, rows_s - have, have);
            __syncwarp();
            auto stage_U = [&](i64 top) {           // U columns of rows top, top - 1, ..., top - LU_R + 1: fac[(top - j)][d] = U(j - d, j)
                for (int e = lane; e < LU_R * (kv + 1); e += 32) {
                    const int rr = e / (kv + 1), dd = e - rr * (kv + 1);
                    const i64 jj = top - rr;
                    if (jj >= 0) {
                        const unsigned dst = (unsigned)__cvta_generic_to_shared(&fac[rr * (kv + 1) + dd]);
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(ab + (i64)ld * jj + kv - dd) : "memory");
                    } else {
                        fac[rr * (kv + 1) + dd] = 1.0;
                    }
                }
                asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
            };
            __syncwarp();
            stage_U(hi);
            __syncwarp();
            for (i64 j = n - 1; j >= 0; --j) {
                if (hi - j == LU_R) {               // rows (hi - LU_R, hi] are final
                    store_rows(hi - LU_R + 1, rows_s - LU_R, LU_R);
                    __syncwarp();
                    const int keep = have - LU_R;
                    for (int r = keep - 1; r >= 0; --r)
                        xs[(rows_s - 1 - r) * 33 + lane] = xs[(rows_s - 1 - r - LU_R) * 33 + lane];
                    hi -= LU_R;
                    const i64 top = hi - keep;      // last row not yet resident
                    const int more = (int)(top + 1 < LU_R ? (top + 1 > 0 ? top + 1 : 0) : LU_R);
                    __syncwarp();
                    load_rows(top - more + 1, rows_s - keep - more, more);
                    stage_U(hi);
                    have = keep + more;
                    __syncwarp();
                }
                const int s = rows_s - 1 - (int)(hi - j);
                const double *uj = fac + (int)(hi - j) * (kv + 1);   // uj[d] = U(j - d, j)
                const double xj = xs[s * 33 + lane] / uj[0];
                xs[s * 33 + lane] = xj;
                const int um = (int)(kv < j ? kv : j);
                if (LT > 0) {
                    double xv[KVT > 0 ? KVT : 1];
#pragma unroll
                    for (int d = 1; d <= KVT; ++d) xv[d - 1] = d <= um ? xs[(s - d) * 33 + lane] : 0.0;
#pragma unroll
                    for (int d = 1; d <= KVT; ++d) xv[d - 1] = fma(-uj[d], xj, xv[d - 1]);
#pragma unroll
                    for (int d = 1; d <= KVT; ++d) if (d <= um) xs[(s - d) * 33 + lane] = xv[d - 1];
                } else {
                    for (int d = 1; d <= um; ++d) xs[(s - d) * 33 + lane] -= uj[d] * xj;
                }
            }
            __syncwarp();
            store_rows(hi - have + 1, rows_s - have, have);
            __syncwarp();
        }
    }
}

extern "C" {

int cb_bandlu_destroy(cb_bandlu *F) {
    if (!F) return CB_OK;
    cb_bandlu_part_destroy(F->part);
    cudaFree(F->d_ab); cudaFree(F->d_ipiv);
    delete F;
    return CB_OK;
}

static size_t bandlu_solve_smem(int l, int u) {
    const int kv = l + u, reach = kv > l ? kv : l;
    const int fw = kv + 1 > l ? kv + 1 : l;
    return sizeof(double) * (size_t)LU_WARPS * ((size_t)(LU_R + reach) * 33 + (size_t)LU_R * fw + LU_R / 2);
}

int cb_bandlu_create(const double *band, i64 n, int l, int u, cb_bandlu **out, void *stream) {
    CB_REQUIRE(out != nullptr, "cb_bandlu_create: out is NULL");
    *out = nullptr;
    CB_REQUIRE(band != nullptr && n >= 1 && l >= 0 && u >= 0, "cb_bandlu_create: bad argument");
    // the factorisation's lanes cover l + 1 pivot candidates and l + u trailing columns; the solve's row chunk must fit
    // the 227 KB of shared memory a CTA can opt into (l + u = 31 does not) — refuse here, not at the first solve
    if (l + u > 31 || bandlu_solve_smem(l, u) > 227 * 1024) {
        cb_set_error("cb_bandlu_create: l + u = %d outside the supported range 0..30", l + u);
        return CB_ERR_UNSUPPORTED;
    }
    cudaStream_t st = (cudaStream_t)stream;
    cb_bandlu_s *F = new cb_bandlu_s();
    F->n = n; F->l = l; F->u = u; F->ld = 2 * l + u + 1; F->d_ab = nullptr; F->d_ipiv = nullptr; F->info = 0; F->part = nullptr;
    cudaGetDevice(&F->device);
    int *d_info = nullptr;
    cudaError_t e = cudaSuccess;
    auto step = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    step(cudaMalloc(&F->d_ab, sizeof(double) * (size_t)F->ld * (size_t)n));
    step(cudaMalloc(&F->d_ipiv, sizeof(int) * (size_t)n));
    step(cudaMalloc(&d_info, sizeof(int)));
    if (e == cudaSuccess) {
        const size_t smem = sizeof(double) * (size_t)(LU_NB + l + u) * F->ld;
        step(cudaFuncSetAttribute(bandlu_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        if (e == cudaSuccess) bandlu_factor_kernel<<<1, 32, smem, st>>>(band, n, l, u, F->d_ab, F->d_ipiv, d_info);
        step(cudaGetLastError());
        step(cudaMemcpyAsync(&F->info, d_info, sizeof(int), cudaMemcpyDeviceToHost, st));
        step(cudaStreamSynchronize(st));
    }
    cudaFree(d_info);
    if (e != cudaSuccess) { cb_set_error("cb_bandlu_create: %s", cudaGetErrorString(e)); cb_bandlu_destroy(F); return CB_ERR_CUDA; }
    if (F->info != 0) {
        // SingularException of lu / factorize in the reference
        cb_set_error("cb_bandlu_create: matrix is singular (zero pivot in column %d)", F->info);
        cb_bandlu_destroy(F);
        return CB_ERR_ARG;
    }
    if (cb_bandlu_part_applicable(n, l, u)) {
        const int prc = cb_bandlu_part_create(F->d_ab, F->d_ipiv, n, l, u, &F->part, st);
        if (prc) { cb_bandlu_destroy(F); return prc; }
    }
    *out = F;
    return CB_OK;
}

int cb_bandlu_solve(const cb_bandlu *F, double *X, i64 ldx, i64 nvec, void *stream) {
    CB_REQUIRE(F != nullptr && X != nullptr, "cb_bandlu_solve: NULL argument");
    CB_REQUIRE(nvec >= 0 && ldx >= F->n, "DimensionMismatch: cb_bandlu_solve: leading dimension %lld smaller than the matrix order %lld",
               (long long)ldx, (long long)F->n);
    if (nvec == 0) return CB_OK;
    if (F->part) return cb_bandlu_part_solve(F->part, F->d_ab, F->d_ipiv, X, ldx, nvec, (cudaStream_t)stream);
    const size_t smem = bandlu_solve_smem(F->l, F->u);
    const i64 ngroups = (nvec + 31) / 32;
    i64 grid = (ngroups + LU_WARPS - 1) / LU_WARPS;
    if (grid > (i64)CB_NUM_SMS * 4) grid = (i64)CB_NUM_SMS * 4;
    auto launch = [&](auto kern) -> int {
        CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)grid, 32 * LU_WARPS, smem, (cudaStream_t)stream>>>(F->d_ab, F->d_ipiv, F->n, F->l, F->u, X, ldx, nvec);
        return CB_OK;
    };
    int lrc = CB_OK;
    switch (F->l == F->u ? F->l : 0) {              // symmetric bandwidths (the operators of the path): compile-time loops
#define CB_LU(LL) case LL: lrc = launch(bandlu_solve_kernel<LL, 2 * LL>); break;
        CB_LU(1) CB_LU(2) CB_LU(3) CB_LU(4) CB_LU(5) CB_LU(6) CB_LU(7) CB_LU(8) CB_LU(9) CB_LU(10) CB_LU(11) CB_LU(12)
        CB_LU(13) CB_LU(14) CB_LU(15)
#undef CB_LU
        default: lrc = launch(bandlu_solve_kernel<0, 0>); break;
    }
    if (lrc) return lrc;
    CB_LAUNCH_CHECK();
    return CB_OK;
}

}  // extern "C"
