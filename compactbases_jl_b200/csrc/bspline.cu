// bspline.cu — B-spline basis handle, lgwt, batched evaluation (K1), CSC export
// (K1b), fused spline evaluation and quadrature tables (K2).
#include <cub/cub.cuh>
#include <mutex>

#include "bspline_core.cuh"

#define CB_DISPATCH_K(kval, ...)                                                 \
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

// ---------------------------------------------------------------------------
// lgwt (src/quadrature.jl:60-77, change_interval! :18-23, lerp :4)
// ---------------------------------------------------------------------------
__global__ void lgwt_kernel(KnotView kv, const i64 *__restrict__ ivlist, i64 ni,
                            const double *__restrict__ xi, const double *__restrict__ wr, int q,
                            double *__restrict__ x, double *__restrict__ w) {
    i64 tot = ni * q;
    for (i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x; e < tot; e += (i64)gridDim.x * blockDim.x) {
        i64 c = e / q;
        int l = (int)(e - c * q);
        i64 I = ivlist[c];
        double a = kv.at(I - 1), b = kv.at(I);
        double t = (xi[l] + 1.0) / 2.0;
        x[e] = fma(t, b, fma(-t, a, a));
        w[e] = (b - a) * wr[l] / 2.0;
    }
}

// ---------------------------------------------------------------------------
// K1: batched evaluation, one thread per point.
//
// Knots are staged in shared memory once per (persistent) CTA when they fit: with unsorted
// points every lane of a warp looks at a different interval, and 24 scattered 8-byte reads per
// point made the first version of this kernel L1-wavefront bound (ncu, round 1: 2.0 TB/s of 6.5).
// The interval comes from an interpolation guess that is then checked against the actual knots, so
// the result is exactly "last i with t[i] <= x" (find_interval, src/knot_sets.jl:209-216) for any
// knot distribution; if the guess is off, a binary search takes over.  The K-wide rows are staged
// through shared memory so that they leave the SM as coalesced 16-byte stores.
// ---------------------------------------------------------------------------
constexpr int EVAL_THREADS = 256;
constexpr int EVAL_SMEM_KNOTS = 6144;              // largest t.t staged in shared memory (48 KB)
#ifndef EVAL_REP
#define EVAL_REP 16                                // bank-private copies of a small knot table
#endif

// Knot table view: entry j of the (padded) table.  With REP > 1 the table is replicated REP times in shared memory,
// copy c at the addresses = c (mod REP) doubles, and a lane reads copy lane % REP: the 16 lanes of a half-warp then
// hit 16 different 8-byte banks whatever their intervals are (REP = 16), where a single copy read at 32 unrelated
// indices costs ~4.6 wavefronts per LDS.64 instead of 2 — the profile of the single-copy kernel
// (profiles/r1u_bspline_eval_summary.txt) had the shared-memory pipe at 80 %, 35.8 M of its 65 M wavefronts conflicts.
template <int REP> struct KnotTab {
    const double *base;                            // REP == 1: plain array;  else base already includes the lane's copy
    __device__ __forceinline__ double operator[](int j) const { return base[j * REP]; }
};

// first index in [0, n_tt) with tt[idx] > x (n_tt if none), x inside [tt[0], tt[n_tt-1])
template <class Tab>
__device__ __forceinline__ int upper_bound_guess(const Tab &tt, int n_tt, double x, double first, double scale) {
    int g = (int)((x - first) * scale);
    g = g < 0 ? 0 : (g > n_tt - 2 ? n_tt - 2 : g);
    const double a = tt[g], b = tt[g + 1];
    if (a <= x && b > x) return g + 1;
    int lo, hi;
    if (a > x) { lo = 0; hi = g; } else { lo = g + 2; hi = n_tt; }   // tt[g+1] <= x
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (tt[mid] > x) hi = mid; else lo = mid + 1;
    }
    return lo;
}

template <int K, bool SMEM_KNOTS, int MT, int REP, int NT>
__global__ void __launch_bounds__(NT, 1024 / NT)
bspline_eval_kernel(KnotView kv, i64 nf, const double *__restrict__ x, i64 nx, int m,
                    int32_t *__restrict__ interval, double *__restrict__ vals, int vec_ok) {
    constexpr int KP = K | 1;                      // odd row pitch: conflict-free staging
    extern __shared__ double esm[];                // [NT/32][32*KP] row staging, then the knot table(s)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double *sw = esm + warp * (32 * KP);
    double *sknots = esm + (NT / 32) * (32 * KP);
    const int n_tt = (int)kv.n_tt;
    KnotTab<SMEM_KNOTS ? REP : 1> tt;
    tt.base = kv.tt;
    if (SMEM_KNOTS) {
        // K-1 copies of the end knots on either side: the knot window of any interval is then read without clamping
        for (int i = threadIdx.x; i < (n_tt + 2 * (K - 1)) * REP; i += NT) {
            int j = i / REP - (K - 1);
            j = j < 0 ? 0 : (j > n_tt - 1 ? n_tt - 1 : j);
            sknots[i] = kv.tt[j];
        }
        __syncthreads();
        tt.base = sknots + (K - 1) * REP + (lane & (REP - 1));
    }
    const double first = kv.tt[0], last = kv.tt[n_tt - 1];
    const double scale = (double)(n_tt - 1) / (last - first);
    const i64 nblk = (nx + NT - 1) / NT;
    // the point of the next iteration is loaded before this one is evaluated (one load in flight per thread)
    double xnext = 0.0;
    if ((i64)blockIdx.x * NT + threadIdx.x < nx) xnext = __ldcs(x + (i64)blockIdx.x * NT + threadIdx.x);
    for (i64 blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const i64 wbase = blk * NT + warp * 32;     // first point of this warp
        const i64 p = wbase + lane;
        const double xp = xnext;
        {
            const i64 pn = p + (i64)gridDim.x * NT;
            if (pn < nx) xnext = __ldcs(x + pn);
        }
        double N[K];
#pragma unroll
        for (int r = 0; r < K; ++r) N[r] = 0.0;
        if (p < nx) {
            int I = 0;
            if (SMEM_KNOTS) {
                if (xp >= first && xp <= last)     // also rejects NaN
                    I = xp == last ? (int)(kv.nt - kv.mr) : upper_bound_guess(tt, n_tt, xp, first, scale) + (kv.ml - 1);
            } else {
                I = cb_find_interval(kv.tt, kv.n_tt, kv.ml, kv.mr, kv.nt, xp);
            }
            if (I > 0) {
                double tw[2 * K];
                const int jb = I - K - (kv.ml - 1);        // 0-based index into t.t of expanded knot I-K+1
#pragma unroll
                for (int s = 0; s < 2 * K; ++s) {
                    int j = jb + s;
                    if (!SMEM_KNOTS) j = j < 0 ? 0 : (j > n_tt - 1 ? n_tt - 1 : j);
                    tw[s] = tt[j];
                }
                cb_bspline_all<K, MT>(tw, xp, m, N);
#pragma unroll
                for (int r = 0; r < K; ++r) {
                    const i64 j = (i64)I - K + 1 + r;
                    if (j < 1 || j > nf) N[r] = 0.0;
                }
            }
            __stcs(interval + p, I);
        }
#pragma unroll
        for (int r = 0; r < K; ++r) sw[lane * KP + r] = N[r];
        __syncwarp();
        // the warp's 32 rows are 32*K contiguous doubles in vals
        i64 npts = nx - wbase; npts = npts > 32 ? 32 : (npts < 0 ? 0 : npts);
        const int tot = (int)npts * K;
        double *out = vals + wbase * K;             // wbase*K*8 bytes: 16-byte aligned (wbase % 32 == 0)
        if (vec_ok && npts == 32) {
#pragma unroll
            for (int it = 0; it < (32 * K + 63) / 64; ++it) {
                const int e2 = lane * 2 + 64 * it;
                if (32 * K % 64 == 0 || e2 < 32 * K) {
                    const int r0 = e2 / K, c0 = e2 - r0 * K, r1 = (e2 + 1) / K, c1 = (e2 + 1) - r1 * K;
                    __stcs(reinterpret_cast<double2 *>(out + e2), make_double2(sw[r0 * KP + c0], sw[r1 * KP + c1]));
                }
            }
        } else {
            for (int e = lane; e < tot; e += 32) {
                const int r0 = e / K, c0 = e - r0 * K;
                out[e] = sw[r0 * KP + c0];
            }
        }
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------
// Fused spline evaluation (B*C)[x]: thread per point, loop over vectors.
// ---------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(256)
bspline_eval_spline_kernel(KnotView kv, i64 nf, const double *__restrict__ x, i64 nx, int m,
                           const double *__restrict__ C, i64 ldc, i64 nvec,
                           double *__restrict__ out, i64 ldo) {
    for (i64 p = blockIdx.x * (i64)blockDim.x + threadIdx.x; p < nx; p += (i64)gridDim.x * blockDim.x) {
        double xp = x[p];
        int I = cb_find_interval(kv.tt, kv.n_tt, kv.ml, kv.mr, kv.nt, xp);
        double N[K];
        if (I > 0) {
            double tw[2 * K];
            cb_knot_window<K>(kv.tt, kv.n_tt, kv.ml, I, tw);
            cb_bspline_all<K>(tw, xp, m, N);
        }
        for (i64 v = 0; v < nvec; ++v) {
            double s = 0.0;
            if (I > 0) {
#pragma unroll
                for (int r = 0; r < K; ++r) {
                    i64 j = (i64)I - K + 1 + r;
                    if (j >= 1 && j <= nf) s = fma(N[r], C[(j - 1) + ldc * v], s);
                }
            }
            out[p + ldo * v] = s;
        }
    }
}

// ---------------------------------------------------------------------------
// K2: basis_functions(t, x, m) tables at the quadrature nodes
// ---------------------------------------------------------------------------
template <int K>
__global__ void __launch_bounds__(256)
bspline_table_kernel(KnotView kv, i64 nf, const i64 *__restrict__ ivlist, i64 ni, int q,
                     const double *__restrict__ x, int m, double *__restrict__ table) {
    i64 tot = ni * q;
    for (i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x; e < tot; e += (i64)gridDim.x * blockDim.x) {
        i64 c = e / q;
        int I = (int)ivlist[c];
        double tw[2 * K], N[K];
        cb_knot_window<K>(kv.tt, kv.n_tt, kv.ml, I, tw);
        cb_bspline_all<K>(tw, x[e], m, N);
#pragma unroll
        for (int r = 0; r < K; ++r) {
            i64 j = (i64)I - K + 1 + r;
            table[e * K + r] = (j >= 1 && j <= nf) ? N[r] : 0.0;
        }
    }
}

// ---------------------------------------------------------------------------
// K1b: ELL rows -> SparseMatrixCSC (1-based, zeros dropped, rows ascending)
// ---------------------------------------------------------------------------
__global__ void csc_keys_kernel(const int32_t *__restrict__ interval, const double *__restrict__ vals,
                                i64 nx, int k, i64 nf, i64 col0, i64 ncols,
                                uint32_t *__restrict__ keys, uint32_t *__restrict__ idx) {
    i64 tot = nx * k;
    for (i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x; e < tot; e += (i64)gridDim.x * blockDim.x) {
        i64 p = e / k;
        int a = (int)(e - p * k);
        int I = interval[p];
        uint32_t key = (uint32_t)ncols;
        if (I > 0) {
            i64 j = (i64)I - k + 1 + a;
            i64 c = j - 1 - col0;
            if (j >= 1 && j <= nf && c >= 0 && c < ncols && vals[e] != 0.0) key = (uint32_t)c;
        }
        keys[e] = key;
        idx[e] = (uint32_t)e;
    }
}

__global__ void csc_colptr_kernel(const uint32_t *__restrict__ keys_sorted, i64 tot, i64 ncols,
                                  i64 *__restrict__ colptr) {
    for (i64 c = blockIdx.x * (i64)blockDim.x + threadIdx.x; c <= ncols; c += (i64)gridDim.x * blockDim.x) {
        i64 lo = 0, hi = tot;                      // first position with key >= c
        while (lo < hi) {
            i64 mid = (lo + hi) >> 1;
            if (keys_sorted[mid] >= (uint32_t)c) hi = mid; else lo = mid + 1;
        }
        colptr[c] = lo + 1;
    }
}

__global__ void csc_fill_kernel(const uint32_t *__restrict__ idx_sorted, const double *__restrict__ vals,
                                const i64 *__restrict__ colptr, i64 ncols, int k,
                                i64 *__restrict__ rowval, double *__restrict__ nzval) {
    i64 nnz = colptr[ncols] - 1;
    for (i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x; e < nnz; e += (i64)gridDim.x * blockDim.x) {
        uint32_t src = idx_sorted[e];
        rowval[e] = (i64)(src / (uint32_t)k) + 1;
        nzval[e] = vals[src];
    }
}

// hand-written counting-sort path (csc.cu)
bool cb_csc_fast_applicable(int k, i64 nx, i64 ncols);
size_t cb_csc_fast_scratch_bytes(i64 nx, i64 ncols);
int cb_csc_fast(const int32_t *interval, const double *vals, i64 nx, int k, i64 nf, i64 col0, i64 ncols,
                void *scratch, i64 *colptr, i64 *rowval, double *nzval, cudaStream_t st);

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" {

int cb_bspline_create(const double *tt, i64 n_tt, int k, int ml, int mr,
                      const double *xi, const double *wr, int q, cb_bspline **out) {
    CB_REQUIRE(out != nullptr, "cb_bspline_create: out is NULL");
    *out = nullptr;
    // assert_multiplicities, src/knot_sets.jl:26-33
    CB_REQUIRE(k > 0, "Polynomial order must be a positive integer, got k = %d", k);
    CB_REQUIRE(ml >= 1 && ml <= k, "Left multiplicity has to be in range 1..%d, got %d", k, ml);
    CB_REQUIRE(mr >= 1 && mr <= k, "Right multiplicity has to be in range 1..%d, got %d", k, mr);
    i64 nreq = 2 > k + 3 - ml - mr ? 2 : k + 3 - ml - mr;
    CB_REQUIRE(tt != nullptr && n_tt >= nreq,
               "Polynomial order k = %d and left,right multiplicities %d,%d require a knot sequence of at least %lld points",
               k, ml, mr, (long long)nreq);
    CB_REQUIRE(q >= 1 && q <= 64 && xi != nullptr && wr != nullptr, "cb_bspline_create: need 1..64 reference nodes");
    if (k > CB_KMAX) { cb_set_error("B-spline order k=%d outside compiled range 1..%d", k, CB_KMAX); return CB_ERR_UNSUPPORTED; }
    for (i64 i = 1; i < n_tt; ++i)
        CB_REQUIRE(tt[i] >= tt[i - 1], "cb_bspline_create: knots must be non-decreasing (index %lld)", (long long)i);

    cb_bspline_s *B = new cb_bspline_s();
    B->k = k; B->ml = ml; B->mr = mr; B->q = q;
    B->n_tt = n_tt; B->nt = n_tt + ml + mr - 2; B->nf = B->nt - k;
    B->h_tt.assign(tt, tt + n_tt); B->h_xi.assign(xi, xi + q); B->h_wr.assign(wr, wr + q);
    B->d_tt = B->d_xi = B->d_wr = B->d_x = B->d_w = nullptr; B->d_ivlist = nullptr; B->d_ord = nullptr;
    // nonempty_intervals (src/knot_sets.jl:196-198) in expanded 1-based numbering:
    // expanded interval I <-> t.t interval I-(ml-1); the repeated end knots give empty ones.
    std::vector<int32_t> ord((size_t)B->nt, -1);
    for (i64 s = 0; s + 1 < n_tt; ++s)
        if (tt[s] != tt[s + 1]) {
            i64 I = s + 1 + (ml - 1);
            ord[(size_t)(I - 1)] = (int32_t)B->h_ivlist.size();
            B->h_ivlist.push_back(I);
        }
    B->ni = (i64)B->h_ivlist.size();
    int rc = CB_OK;
    auto fail = [&](int code) { cb_bspline_destroy(B); return code; };
#define CB_TRY(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { cb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); return fail(CB_ERR_CUDA); } } while (0)
    CB_TRY(cudaGetDevice(&B->device));
    CB_TRY(cudaMalloc(&B->d_tt, sizeof(double) * n_tt));
    CB_TRY(cudaMalloc(&B->d_xi, sizeof(double) * q));
    CB_TRY(cudaMalloc(&B->d_wr, sizeof(double) * q));
    CB_TRY(cudaMalloc(&B->d_ivlist, sizeof(i64) * (B->ni ? B->ni : 1)));
    CB_TRY(cudaMalloc(&B->d_ord, sizeof(int32_t) * B->nt));
    CB_TRY(cudaMalloc(&B->d_x, sizeof(double) * (B->ni * q ? B->ni * q : 1)));
    CB_TRY(cudaMalloc(&B->d_w, sizeof(double) * (B->ni * q ? B->ni * q : 1)));
    CB_TRY(cudaMemcpy(B->d_tt, tt, sizeof(double) * n_tt, cudaMemcpyHostToDevice));
    CB_TRY(cudaMemcpy(B->d_xi, xi, sizeof(double) * q, cudaMemcpyHostToDevice));
    CB_TRY(cudaMemcpy(B->d_wr, wr, sizeof(double) * q, cudaMemcpyHostToDevice));
    if (B->ni) CB_TRY(cudaMemcpy(B->d_ivlist, B->h_ivlist.data(), sizeof(i64) * B->ni, cudaMemcpyHostToDevice));
    CB_TRY(cudaMemcpy(B->d_ord, ord.data(), sizeof(int32_t) * B->nt, cudaMemcpyHostToDevice));
    if (B->ni) {
        lgwt_kernel<<<cb_grid_for(B->ni * q, 256), 256>>>(make_knot_view(B), B->d_ivlist, B->ni, B->d_xi, B->d_wr, q, B->d_x, B->d_w);
        CB_TRY(cudaGetLastError());
        CB_TRY(cudaDeviceSynchronize());
    }
#undef CB_TRY
    *out = B;
    return rc;
}

int cb_bspline_destroy(cb_bspline *B) {
    if (!B) return CB_OK;
    cudaFree(B->d_tt); cudaFree(B->d_xi); cudaFree(B->d_wr); cudaFree(B->d_ivlist);
    cudaFree(B->d_ord); cudaFree(B->d_x); cudaFree(B->d_w);
    delete B;
    return CB_OK;
}

i64 cb_bspline_numfunctions(const cb_bspline *B) { return B ? B->nf : -1; }
i64 cb_bspline_numknots(const cb_bspline *B) { return B ? B->nt : -1; }
i64 cb_bspline_num_nonempty(const cb_bspline *B) { return B ? B->ni : -1; }
i64 cb_bspline_numnodes(const cb_bspline *B) { return B ? B->ni * B->q : -1; }

int cb_bspline_nodes(const cb_bspline *B, const double **x, const double **w) {
    CB_REQUIRE(B != nullptr, "cb_bspline_nodes: NULL basis");
    if (x) *x = B->d_x;
    if (w) *w = B->d_w;
    return CB_OK;
}

int cb_bspline_nodes_copy(const cb_bspline *B, double *x, double *w, void *stream) {
    CB_REQUIRE(B != nullptr, "cb_bspline_nodes_copy: NULL basis");
    const size_t bytes = sizeof(double) * (size_t)(B->ni * B->q);
    if (bytes == 0) return CB_OK;
    if (x) CB_CUDA(cudaMemcpyAsync(x, B->d_x, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    if (w) CB_CUDA(cudaMemcpyAsync(w, B->d_w, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return CB_OK;
}

int cb_bspline_eval(const cb_bspline *B, const double *x, i64 nx, int m,
                    int32_t *interval, double *vals, void *stream) {
    CB_REQUIRE(B != nullptr, "cb_bspline_eval: NULL basis");
    CB_REQUIRE(nx >= 0 && m >= 0 && m < (B->k > 1 ? B->k : 2), "cb_bspline_eval: need nx >= 0 and 0 <= m < k");
    if (nx == 0) return CB_OK;
    CB_REQUIRE(x && interval && vals, "cb_bspline_eval: NULL array");
    cudaStream_t st = (cudaStream_t)stream;
    int vec_ok = ((uintptr_t)vals % 16) == 0;
    const size_t ktab = (size_t)B->n_tt + 2 * (size_t)(B->k - 1);   // + the clamped copies of the end knots
    const size_t kp = (size_t)(B->k | 1);
    if (ktab * EVAL_REP * 8 + (1024 / 32) * 32 * kp * 8 <= 200 * 1024) {
        // small knot sets (<= ~1100 knots at k = 7): one 1024-thread CTA per SM with bank-private knot replicas
        const size_t smem = sizeof(double) * (ktab * EVAL_REP + (1024 / 32) * 32 * kp);
        const i64 nblk = (nx + 1023) / 1024;
        const int grid = (int)(nblk < CB_NUM_SMS ? nblk : CB_NUM_SMS);
        CB_DISPATCH_K(B->k, {
            auto kern = m == 0 ? bspline_eval_kernel<K, true, 0, EVAL_REP, 1024> : bspline_eval_kernel<K, true, -1, EVAL_REP, 1024>;
            CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid, 1024, smem, st>>>(make_knot_view(B), B->nf, x, nx, m, interval, vals, vec_ok);
        });
    } else if (B->n_tt <= EVAL_SMEM_KNOTS) {
        const size_t smem = sizeof(double) * (ktab + (EVAL_THREADS / 32) * 32 * kp);
#ifndef EVAL_WAVES
#define EVAL_WAVES 8
#endif
        int grid = cb_grid_for(nx, EVAL_THREADS, EVAL_WAVES);   // a whole number of waves of the 4 resident CTAs per SM; the knot copy is amortised over many blocks
        CB_DISPATCH_K(B->k, {
            auto kern = m == 0 ? bspline_eval_kernel<K, true, 0, 1, EVAL_THREADS> : bspline_eval_kernel<K, true, -1, 1, EVAL_THREADS>;
            CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid, EVAL_THREADS, smem, st>>>(make_knot_view(B), B->nf, x, nx, m, interval, vals, vec_ok);
        });
    } else {
        const size_t smem = sizeof(double) * (EVAL_THREADS / 32) * 32 * kp;
        int grid = cb_grid_for(nx, EVAL_THREADS, 16);
        CB_DISPATCH_K(B->k, (bspline_eval_kernel<K, false, -1, 1, EVAL_THREADS><<<grid, EVAL_THREADS, smem, st>>>(make_knot_view(B), B->nf, x, nx, m, interval, vals, vec_ok)));
    }
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int cb_bspline_eval_host(const cb_bspline *B, const double *xh, i64 nx, int m,
                         int32_t *ih, double *vh) {
    CB_REQUIRE(B != nullptr, "cb_bspline_eval_host: NULL basis");
    if (nx == 0) return CB_OK;
    // one scratch block from the stream-ordered pool, asynchronous copies (PCIe speed from pinned arrays), points in
    // chunks on two streams so that the upload of one chunk overlaps the evaluation / download of the other
    { int prc = cb_keep_scratch_pool(); if (prc) return prc; }
    static cudaStream_t hs[64][2] = {};
    static std::mutex hs_mu;                         // the copy streams are shared: host-buffer calls are serialised
    std::lock_guard<std::mutex> hs_lock(hs_mu);
    int dev = 0;
    CB_CUDA(cudaGetDevice(&dev));
    CB_REQUIRE(dev >= 0 && dev < 64, "cb_bspline_eval_host: device index %d out of range", dev);
    for (int i = 0; i < 2; ++i)
        if (!hs[dev][i]) CB_CUDA(cudaStreamCreateWithFlags(&hs[dev][i], cudaStreamNonBlocking));
    const i64 chunk = nx > (i64)1 << 18 ? (nx + 7) / 8 : nx;          // ~8 chunks for large batches
    const size_t b_x = sizeof(double) * (size_t)chunk, b_v = sizeof(double) * (size_t)chunk * B->k, b_i = (sizeof(int32_t) * (size_t)chunk + 15) & ~(size_t)15;
    char *scr[2] = {nullptr, nullptr};
    for (int i = 0; i < 2; ++i) CB_CUDA(cudaMallocAsync(&scr[i], b_x + b_v + b_i, hs[dev][i]));
    int rc = CB_OK;
    cudaError_t e = cudaSuccess;
    i64 c = 0;
    for (i64 p0 = 0; p0 < nx && e == cudaSuccess && rc == CB_OK; p0 += chunk, ++c) {
        const int i = (int)(c & 1);
        const i64 np = p0 + chunk <= nx ? chunk : nx - p0;
        double *dx = reinterpret_cast<double *>(scr[i]), *dv = reinterpret_cast<double *>(scr[i] + b_x);
        int32_t *di = reinterpret_cast<int32_t *>(scr[i] + b_x + b_v);
        e = cudaMemcpyAsync(dx, xh + p0, sizeof(double) * np, cudaMemcpyHostToDevice, hs[dev][i]);
        if (e == cudaSuccess) rc = cb_bspline_eval(B, dx, np, m, di, dv, hs[dev][i]);
        if (e == cudaSuccess && rc == CB_OK) e = cudaMemcpyAsync(ih + p0, di, sizeof(int32_t) * np, cudaMemcpyDeviceToHost, hs[dev][i]);
        if (e == cudaSuccess && rc == CB_OK) e = cudaMemcpyAsync(vh + p0 * B->k, dv, sizeof(double) * np * B->k, cudaMemcpyDeviceToHost, hs[dev][i]);
    }
    for (int i = 0; i < 2; ++i) {
        cudaFreeAsync(scr[i], hs[dev][i]);
        const cudaError_t e2 = cudaStreamSynchronize(hs[dev][i]);
        if (e == cudaSuccess) e = e2;
    }
    if (e != cudaSuccess) { cb_set_error("cb_bspline_eval_host: %s", cudaGetErrorString(e)); return CB_ERR_CUDA; }
    return rc;
}

int cb_bspline_eval_spline(const cb_bspline *B, const double *x, i64 nx, int m,
                           const double *C, i64 ldc, i64 nvec, double *out, i64 ldo, void *stream) {
    CB_REQUIRE(B != nullptr, "cb_bspline_eval_spline: NULL basis");
    CB_REQUIRE(nx >= 0 && nvec >= 0 && m >= 0 && m < (B->k > 1 ? B->k : 2), "cb_bspline_eval_spline: bad sizes");
    CB_REQUIRE(ldc >= B->nf && ldo >= nx, "cb_bspline_eval_spline: leading dimensions too small");
    if (nx == 0 || nvec == 0) return CB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int grid = cb_grid_for(nx, 256, 16);
    CB_DISPATCH_K(B->k, (bspline_eval_spline_kernel<K><<<grid, 256, 0, st>>>(make_knot_view(B), B->nf, x, nx, m, C, ldc, nvec, out, ldo)));
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int cb_bspline_quad_table(const cb_bspline *B, int m, double *table, void *stream) {
    CB_REQUIRE(B != nullptr && table != nullptr, "cb_bspline_quad_table: NULL argument");
    CB_REQUIRE(m >= 0 && m < (B->k > 1 ? B->k : 2), "cb_bspline_quad_table: need 0 <= m < k");
    if (B->ni == 0) return CB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int grid = cb_grid_for(B->ni * B->q, 256, 16);
    CB_DISPATCH_K(B->k, (bspline_table_kernel<K><<<grid, 256, 0, st>>>(make_knot_view(B), B->nf, B->d_ivlist, B->ni, B->q, B->d_x, m, table)));
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int cb_bspline_eval_csc(const cb_bspline *B, const double *x, i64 nx, i64 col0, i64 ncols,
                        i64 *colptr, i64 *rowval, double *nzval, i64 *nnz_host, void *stream) {
    CB_REQUIRE(B != nullptr && colptr != nullptr, "cb_bspline_eval_csc: NULL argument");
    CB_REQUIRE(col0 >= 0 && ncols >= 0 && col0 + ncols <= B->nf, "cb_bspline_eval_csc: column range outside 1..nf");
    i64 tot = nx * B->k;
    if (tot >= ((i64)1 << 31)) { cb_set_error("cb_bspline_eval_csc: nx*k >= 2^31, split the point batch"); return CB_ERR_UNSUPPORTED; }
    cudaStream_t st = (cudaStream_t)stream;
    { int prc = cb_keep_scratch_pool(); if (prc) return prc; }
    if (cb_csc_fast_applicable(B->k, nx, ncols)) {
        // one scratch block from the stream-ordered pool: rows of K1 (vals, intervals) + the counters of the counting sort
        const size_t b_vals = sizeof(double) * (size_t)tot, b_iv = (sizeof(int32_t) * (size_t)nx + 15) & ~(size_t)15;
        char *scr = nullptr;
        CB_CUDA(cudaMallocAsync(&scr, b_vals + b_iv + cb_csc_fast_scratch_bytes(nx, ncols), st));
        double *f_vals = reinterpret_cast<double *>(scr);
        int32_t *f_iv = reinterpret_cast<int32_t *>(scr + b_vals);
        int frc = cb_bspline_eval(B, x, nx, 0, f_iv, f_vals, stream);
        if (frc == CB_OK) frc = cb_csc_fast(f_iv, f_vals, nx, B->k, B->nf, col0, ncols, scr + b_vals + b_iv, colptr, rowval, nzval, st);
        i64 lastf = 1;
        cudaError_t fe = frc == CB_OK ? cudaMemcpyAsync(&lastf, colptr + ncols, sizeof(i64), cudaMemcpyDeviceToHost, st) : cudaSuccess;
        cudaFreeAsync(scr, st);
        const cudaError_t fe2 = cudaStreamSynchronize(st);
        if (frc != CB_OK) return frc;
        if (fe == cudaSuccess) fe = fe2;
        if (fe != cudaSuccess) { cb_set_error("cb_bspline_eval_csc: %s", cudaGetErrorString(fe)); return CB_ERR_CUDA; }
        if (nnz_host) *nnz_host = lastf - 1;
        return CB_OK;
    }
    // large column counts: radix sort of (column, entry) keys (library sort)
    int32_t *d_iv = nullptr; double *d_vals = nullptr; uint32_t *d_keys = nullptr, *d_idx = nullptr; void *d_tmp = nullptr;
    i64 n1 = tot ? tot : 1;
    CB_CUDA(cudaMallocAsync(&d_iv, sizeof(int32_t) * (nx ? nx : 1), st));
    CB_CUDA(cudaMallocAsync(&d_vals, sizeof(double) * n1, st));
    CB_CUDA(cudaMallocAsync(&d_keys, sizeof(uint32_t) * 2 * n1, st));
    CB_CUDA(cudaMallocAsync(&d_idx, sizeof(uint32_t) * 2 * n1, st));
    int rc = cb_bspline_eval(B, x, nx, 0, d_iv, d_vals, stream);
    if (rc != CB_OK) return rc;
    if (tot) {
        csc_keys_kernel<<<cb_grid_for(tot, 256, 16), 256, 0, st>>>(d_iv, d_vals, nx, B->k, B->nf, col0, ncols, d_keys, d_idx);
        CB_LAUNCH_CHECK();
        int end_bit = 1; while (((i64)1 << end_bit) <= ncols) ++end_bit;
        size_t tmp_bytes = 0;
        CB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys, d_keys + n1, d_idx, d_idx + n1, (int)tot, 0, end_bit, st));
        CB_CUDA(cudaMallocAsync(&d_tmp, tmp_bytes ? tmp_bytes : 1, st));
        CB_CUDA(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_keys, d_keys + n1, d_idx, d_idx + n1, (int)tot, 0, end_bit, st));
    }
    csc_colptr_kernel<<<cb_grid_for(ncols + 1, 256), 256, 0, st>>>(d_keys + n1, tot, ncols, colptr);
    CB_LAUNCH_CHECK();
    if (tot && rowval && nzval) {
        csc_fill_kernel<<<cb_grid_for(tot, 256, 16), 256, 0, st>>>(d_idx + n1, d_vals, colptr, ncols, B->k, rowval, nzval);
        CB_LAUNCH_CHECK();
    }
    i64 last = 1;
    CB_CUDA(cudaMemcpyAsync(&last, colptr + ncols, sizeof(i64), cudaMemcpyDeviceToHost, st));
    CB_CUDA(cudaFreeAsync(d_iv, st)); CB_CUDA(cudaFreeAsync(d_vals, st));
    CB_CUDA(cudaFreeAsync(d_keys, st)); CB_CUDA(cudaFreeAsync(d_idx, st));
    if (d_tmp) CB_CUDA(cudaFreeAsync(d_tmp, st));
    CB_CUDA(cudaStreamSynchronize(st));
    if (nnz_host) *nnz_host = last - 1;
    return CB_OK;
}

}  // extern "C"
