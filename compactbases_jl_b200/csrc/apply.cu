// apply.cu — Y = alpha*A*X + beta*Y for the operator containers of the hot path
// (SURVEY §8 a17): Tridiagonal/SymTridiagonal (K9), BandedMatrix (K7), Diagonal,
// and the FE-DVR element-block operator (K8); plus the BLAS-1 helpers around them (axpby, batched dot).
// The arithmetic of mul! lives in LinearAlgebra / BandedMatrices / BlockBandedMatrices in the reference (call
// sites src/linear_operators.jl:40,50-52, src/inner_products.jl:30,46, test/derivative_accuracy_utils.jl:26-37).
//
// All kernels are HBM-bound stencils over X (n x nvec, column-major).  Each of K7, K8, K9 exists twice:
//   * a warp-specialised TMA pipeline (one persistent CTA per SM: a loader warp and a storer warp move tiles with
//     cp.async.bulk — one 1-D bulk copy per vector column — through a ring of shared-memory stages handed over by
//     mbarriers; the other warps compute in shared memory: DMMA for K7 / K8, plain FMAs for K9).  Taken when
//     beta == 0 and X, Y agree in 16-byte phase per column;
//   * the earlier kernel whose warps load, compute and store in turn (cp.async / register staging), for every
//     other call.  Both give the same results; the parity tests run both.
#include <stdlib.h>
#include <mutex>

#include "common.cuh"

// ---------------------------------------------------------------------------
// K9: tridiagonal.  thread <-> row pair, loop over the vectors of the CTA's
// vector group; neighbours come from warp shuffles (one 16-byte load per lane).
// ---------------------------------------------------------------------------
constexpr int TRI_THREADS = 256;
constexpr int TRI_VUNROLL = 4;

template <bool VEC2>
__global__ void __launch_bounds__(TRI_THREADS)
tridiag_apply_kernel(const double *__restrict__ dl, const double *__restrict__ d,
                     const double *__restrict__ du, i64 n,
                     const double *__restrict__ X, i64 ldx, i64 nvec, i64 vgroup,
                     double alpha, double beta, double *__restrict__ Y, i64 ldy) {
    constexpr int RPT = VEC2 ? 2 : 1;              // rows per thread
    const int lane = threadIdx.x & 31;
    const i64 rows_per_blk = (i64)TRI_THREADS * RPT;
    const i64 nrblk = (n + rows_per_blk - 1) / rows_per_blk;
    const i64 v0 = (i64)blockIdx.y * vgroup;
    const i64 v1 = v0 + vgroup < nvec ? v0 + vgroup : nvec;
    for (i64 rb = blockIdx.x; rb < nrblk; rb += gridDim.x) {
        const i64 i0 = rb * rows_per_blk + (i64)threadIdx.x * RPT;   // first row of this thread
        double cl[RPT], cd[RPT], cu[RPT];
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            i64 i = i0 + r;
            cd[r] = i < n ? d[i] : 0.0;
            cl[r] = (i > 0 && i < n) ? dl[i - 1] : 0.0;
            cu[r] = (i < n - 1) ? du[i] : 0.0;
        }
        for (i64 vb = v0; vb < v1; vb += TRI_VUNROLL) {
            double xa[TRI_VUNROLL][RPT], xm[TRI_VUNROLL], xp[TRI_VUNROLL];
#pragma unroll
            for (int s = 0; s < TRI_VUNROLL; ++s) {
                const i64 v = vb + s;
                const double *xc = X + v * ldx;
                xm[s] = 0.0; xp[s] = 0.0;
#pragma unroll
                for (int r = 0; r < RPT; ++r) xa[s][r] = 0.0;
                if (v < v1) {
                    if (VEC2) {
                        if (i0 + 1 < n) {
                            double2 t = *reinterpret_cast<const double2 *>(xc + i0);
                            xa[s][0] = t.x; xa[s][RPT - 1] = t.y;
                        } else if (i0 < n) {
                            xa[s][0] = xc[i0];
                        }
                    } else if (i0 < n) {
                        xa[s][0] = xc[i0];
                    }
                    // halo rows of the warp: lane 0 needs x[i0-1], lane 31 needs x[i0+RPT]
                    if (lane == 0 && i0 > 0 && i0 - 1 < n) xm[s] = xc[i0 - 1];
                    if (lane == 31 && i0 + RPT < n) xp[s] = xc[i0 + RPT];
                }
            }
#pragma unroll
            for (int s = 0; s < TRI_VUNROLL; ++s) {
                const i64 v = vb + s;
                double up = __shfl_up_sync(0xffffffffu, xa[s][RPT - 1], 1);
                double dn = __shfl_down_sync(0xffffffffu, xa[s][0], 1);
                if (lane != 0) xm[s] = up;
                if (lane != 31) xp[s] = dn;
                if (v < v1) {
                    double *yc = Y + v * ldy;
                    double yv[RPT];
#pragma unroll
                    for (int r = 0; r < RPT; ++r) {
                        double left = r == 0 ? xm[s] : xa[s][r - 1];
                        double right = r == RPT - 1 ? xp[s] : xa[s][r + 1];
                        yv[r] = fma(cu[r], right, fma(cd[r], xa[s][r], cl[r] * left));
                    }
                    if (VEC2 && i0 + 1 < n) {
                        double2 o;
                        if (beta == 0.0) {
                            o.x = alpha * yv[0]; o.y = alpha * yv[RPT - 1];
                        } else {
                            double2 y0 = *reinterpret_cast<const double2 *>(yc + i0);
                            o.x = fma(alpha, yv[0], beta * y0.x); o.y = fma(alpha, yv[RPT - 1], beta * y0.y);
                        }
                        *reinterpret_cast<double2 *>(yc + i0) = o;
                    } else if (i0 < n) {
                        yc[i0] = axpby(yv[0], alpha, beta, yc + i0);
                    }
                }
            }
        }
    }
}

// Diagonal operator.
__global__ void diag_apply_kernel(const double *__restrict__ d, i64 n, const double *__restrict__ X, i64 ldx,
                                  i64 nvec, double alpha, double beta, double *__restrict__ Y, i64 ldy) {
    for (i64 v = blockIdx.y; v < nvec; v += gridDim.y)
        for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
            Y[i + ldy * v] = axpby(d[i] * X[i + ldx * v], alpha, beta, Y + i + ldy * v);
}

// Y = alpha X + beta Y (DiagonalOperator's update of y with the density coefficients).
__global__ void axpby_kernel(i64 n, i64 nvec, double alpha, const double *__restrict__ X, i64 ldx,
                             double beta, double *__restrict__ Y, i64 ldy) {
    for (i64 v = blockIdx.y; v < nvec; v += gridDim.y)
        for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x)
            Y[i + ldy * v] = axpby(X[i + ldx * v], alpha, beta, Y + i + ldy * v);
}

// out[p] = scale * sum_i U[i, p] * W[i, p]: the pairwise inner products u_p' (S v_p) once W = S V has been applied
// (src/inner_products.jl:22-46; FE-DVR / finite differences: W = V, scale = 1 or the grid step, :3-4).  One CTA per
// pair; every thread sums a strided slice in row order, then a fixed shuffle / shared-memory tree: the result does not
// depend on the launch geometry of other pairs and is reproducible run to run.
constexpr int DOT_THREADS = 512;
__global__ void __launch_bounds__(DOT_THREADS)
batched_dot_kernel(const double *__restrict__ U, i64 ldu, const double *__restrict__ Wm, i64 ldw, i64 n, i64 P,
                   double scale, double *__restrict__ out) {
    __shared__ double part[DOT_THREADS / 32];
    for (i64 p = blockIdx.x; p < P; p += gridDim.x) {
        const double *u = U + p * ldu, *w = Wm + p * ldw;
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;           // four chains: loads of four rows in flight per thread
        i64 i = threadIdx.x;
        for (; i + 3 * DOT_THREADS < n; i += 4 * DOT_THREADS) {
            s0 = fma(u[i], w[i], s0);
            s1 = fma(u[i + DOT_THREADS], w[i + DOT_THREADS], s1);
            s2 = fma(u[i + 2 * DOT_THREADS], w[i + 2 * DOT_THREADS], s2);
            s3 = fma(u[i + 3 * DOT_THREADS], w[i + 3 * DOT_THREADS], s3);
        }
        for (; i < n; i += DOT_THREADS) s0 = fma(u[i], w[i], s0);
        double s = (s0 + s1) + (s2 + s3);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        __syncthreads();                            // part[] of the previous pair has been read
        if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x < 32) {
            double t = threadIdx.x < DOT_THREADS / 32 ? part[threadIdx.x] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
            if (threadIdx.x == 0) out[p] = scale * t;
        }
    }
}

// ---------------------------------------------------------------------------
// Shared-memory tile helpers (rows x TV vectors, xs[v*pitch + row]).
// ---------------------------------------------------------------------------
// Loads global rows [g0, g0+rows) (zero outside [0, n)) of vectors [v0, v0+TV).
template <int THREADS>
__device__ __forceinline__ void load_tile(double *xs, int pitch, const double *__restrict__ X, i64 ldx, i64 n,
                                          i64 g0, int rows, i64 v0, int tv, i64 nvec) {
    for (int v = 0; v < tv; ++v) {
        const i64 gv = v0 + v;
        const double *xc = X + gv * ldx;
        for (int r = threadIdx.x; r < rows; r += THREADS) {
            i64 g = g0 + r;
            xs[v * pitch + r] = (gv < nvec && g >= 0 && g < n) ? xc[g] : 0.0;
        }
    }
}

// Stores tile rows [r_lo, r_hi) (tile-local) to global rows g0+r with alpha/beta.
template <int THREADS>
__device__ __forceinline__ void store_tile(const double *ys, int pitch, double *__restrict__ Y, i64 ldy, i64 n,
                                           i64 g0, int r_lo, int r_hi, i64 v0, int tv, i64 nvec,
                                           double alpha, double beta) {
    for (int v = 0; v < tv; ++v) {
        const i64 gv = v0 + v;
        if (gv >= nvec) break;
        double *yc = Y + gv * ldy;
        for (int r = r_lo + threadIdx.x; r < r_hi; r += THREADS) {
            i64 g = g0 + r;
            if (g >= 0 && g < n) yc[g] = axpby(ys[v * pitch + r], alpha, beta, yc + g);
        }
    }
}

// ---------------------------------------------------------------------------
// Asynchronous tile movers for interior tiles (every row and vector in range).
//
// cp.async (LDGSTS) copies global -> shared memory without a register round trip, so a CTA can
// keep several tiles in flight while it computes on another one: on B200 an HBM-bound kernel needs
// ~80 KB of loads in flight per SM at all times (6.5 TB/s x ~1.7 us / 148 SMs); a CTA that loads,
// computes and stores one tile at a time measured 3.2 TB/s here.  8-byte copies keep the only
// alignment requirement that of a double: Julia's column-major batches have arbitrary (often odd)
// leading dimensions, e.g. 90 001, which also rules out TMA tensor maps (global strides must be
// multiples of 16 bytes).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void cp_async8(double *smem_dst, const double *gsrc) {
    unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int ROWS, int TV, int NWARPS>
__device__ __forceinline__ void load_tile_async(double *xs, int pitch, const double *__restrict__ X, i64 ldx,
                                                i64 g0, i64 v0) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NIT = (ROWS + 31) / 32;
#pragma unroll 4
    for (int v = warp; v < TV; v += NWARPS) {
        const double *xc = X + (v0 + v) * ldx + g0;
        double *xd = xs + v * pitch;
#pragma unroll
        for (int i = 0; i < NIT; ++i) {
            const int r = lane + 32 * i;
            if (32 * i + 31 < ROWS || r < ROWS) cp_async8(xd + r, xc + r);
        }
    }
}

// Stores alpha * tile rows [r_lo, r_hi) (beta == 0: Y is not read); lanes run along the rows.
template <int TV, int NWARPS>
__device__ __forceinline__ void store_tile_fast(const double *ys, int pitch, double *__restrict__ Y, i64 ldy,
                                                i64 g0, int r_lo, int r_hi, i64 v0, double alpha) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll 2
    for (int v = warp; v < TV; v += NWARPS) {
        double *yc = Y + (v0 + v) * ldy + g0;
        for (int r = r_lo + lane; r < r_hi; r += 32) __stcs(yc + r, alpha * ys[v * pitch + r]);
    }
}

// Runtime-row-count variant of load_tile_async with clipping: tile rows whose global row g0 + r
// lies outside [0, n) and vectors >= nvec are zero-filled.
template <int TV, int NWARPS>
__device__ __forceinline__ void load_tile_async_rt(double *xs, int pitch, const double *__restrict__ X, i64 ldx, i64 n,
                                                   i64 g0, int rows, i64 v0, i64 nvec) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rlo = g0 < 0 ? (int)(-g0 < rows ? -g0 : rows) : 0;
    const int rhi = g0 + rows > n ? (int)(n - g0 > 0 ? n - g0 : 0) : rows;
#pragma unroll 2
    for (int v = warp; v < TV; v += NWARPS) {
        double *xd = xs + v * pitch;
        if (v0 + v < nvec) {
            const double *xc = X + (v0 + v) * ldx + g0;
            for (int r = rlo + lane; r < rhi; r += 32) cp_async8(xd + r, xc + r);
            for (int r = lane; r < rlo; r += 32) xd[r] = 0.0;
            for (int r = rhi + lane; r < rows; r += 32) xd[r] = 0.0;
        } else {
            for (int r = lane; r < rows; r += 32) xd[r] = 0.0;
        }
    }
}

// Clipped variant of store_tile_fast: rows [r_lo, r_hi) (already clipped by the caller), vectors < nvec.
template <int TV, int NWARPS>
__device__ __forceinline__ void store_tile_fast_clip(const double *ys, int pitch, double *__restrict__ Y, i64 ldy,
                                                     i64 g0, int r_lo, int r_hi, i64 v0, i64 nvec, double alpha) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll 2
    for (int v = warp; v < TV; v += NWARPS) {
        if (v0 + v >= nvec) break;
        double *yc = Y + (v0 + v) * ldy + g0;
        for (int r = r_lo + lane; r < r_hi; r += 32) __stcs(yc + r, alpha * ys[v * pitch + r]);
    }
}

// D = A(8x4, row-major fragment) * B(4x8, column fragment) + C(8x8) in FP64 on the tensor cores
// (SASS: DMMA.8x8x4; measured 37.1 TFLOP/s on B200 against 36.2 for DFMA, at 1/8 of the issue slots).
// Fragments, with g = lane >> 2 and t = lane & 3:  a = A[g][t],  b = B[t][g],  c0/c1 = C[g][2t], C[g][2t+1].
__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ---------------------------------------------------------------------------
// mbarrier / TMA bulk-copy primitives of the warp-specialised pipelines (K7, K8).
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE;\n"
        "bra LAB_WAIT;\n"
        "LAB_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// global -> shared bulk copy (TMA unit), completion counted in bytes on the mbarrier
__device__ __forceinline__ void bulk_load(void *sdst, const void *gsrc, unsigned bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(sdst)), "l"(__cvta_generic_to_global(gsrc)), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// shared -> global bulk copy
__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(__cvta_generic_to_global(gdst)), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// orders this thread's shared-memory writes before later accesses by the TMA unit (async proxy)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Column swizzle of the pipelines' tiles: col(v) = v * P0 + fe_swz(v) with P0 = 0 (mod 16).
__host__ __device__ constexpr int fe_swz(int v) { return 4 * ((v & 3) ^ ((v >> 2) & 1)); }

// One column of a pipeline tile through the TMA unit, executed by the loader / storer warp with lane <-> column.
// scol[r] / gcol[r] address row r of the column in shared / global memory; shared offset and global address have the
// same parity (the tile is laid out that way), so the 16-byte aligned part of rows [ra, rb) is one bulk copy and at
// most one element at either end moves through the lane.  Rows outside [ra, rb) (outside the matrix, or a vector
// beyond the batch: ra == rb) are zero-filled up to nrows_tile.
//   phase 1 (tma_col_load_prepare): zero fill, end elements, returns the bytes the bulk copy will bring;
//   the caller sums them over the warp, arms the mbarrier (release: the plain stores above become visible with it);
//   phase 2 (tma_col_load_issue): the bulk copy.
struct ColPlan { int a, cnt; };
__device__ __forceinline__ ColPlan col_plan(int par0, int ra, int rb) {
    ColPlan p;
    p.a = ra + ((par0 + ra) & 1);                   // first 16-byte aligned row
    p.cnt = rb > p.a ? (rb - p.a) & ~1 : 0;
    return p;
}
__device__ __forceinline__ ColPlan tma_col_load_prepare(double *scol, const double *gcol, int par0, int ra, int rb,
                                                        int nrows_tile) {
    const ColPlan p = col_plan(par0, ra, rb);
    for (int r = 0; r < ra; ++r) scol[r] = 0.0;
    for (int r = rb; r < nrows_tile; ++r) scol[r] = 0.0;
    if (p.a > ra && ra < rb) scol[ra] = gcol[ra];
    for (int r = p.a + p.cnt; r < rb; ++r) scol[r] = gcol[r];       // at most one row (two when the range is short)
    return p;
}
__device__ __forceinline__ void tma_col_store(double *gcol, const double *scol, int par0, int ra, int rb) {
    const ColPlan p = col_plan(par0, ra, rb);
    if (p.cnt) bulk_store(gcol + p.a, scol + p.a, 8u * (unsigned)p.cnt);
    bulk_commit();
    if (p.a > ra && ra < rb) gcol[ra] = scol[ra];
    for (int r = p.a + p.cnt; r < rb; ++r) gcol[r] = scol[r];
    bulk_wait_read0();
}

// ---------------------------------------------------------------------------
// K7: BandedMatrix (l,u).  CTA tile: BT_ROWS rows x 32 vectors, x window staged in shared memory.
// The product runs on the FP64 tensor cores: an 8-row block of A against its window of
// 8 + l + u columns is KS = ceil((8+l+u)/4) DMMA k-steps per 8-vector tile.  The band tile is
// stored in shared memory directly as A fragments (one double per lane per (row block, k-step),
// zero outside band and matrix), staged once per work item and reused for all its vector tiles.
// Every warp keeps its results in registers until all warps have read the x window, then writes
// them over the tile in place.  (A zero of the padded fragment multiplies x entries of the same
// 8-row block that lie outside a row's own band: non-finite X entries therefore also reach the
// other rows of their block, which a scalar banded product would not do.)
// ---------------------------------------------------------------------------
constexpr int BT_VEC = 32;                          // vectors per CTA tile
constexpr int BT_NG = 8;                            // warps per CTA
constexpr int BT_RB = 2;                            // 8-row blocks per warp
constexpr int BT_THREADS = 32 * BT_NG;
constexpr int BT_ROWS = BT_NG * BT_RB * 8;          // 128 rows per CTA tile
#ifndef BT_PITCH_MOD
#define BT_PITCH_MOD 12
#endif
#ifndef BT_TAIL_PCT
#define BT_TAIL_PCT 0                               // share of the vector tiles handed out as single-tile items at the end
#endif
#ifndef BT_PIPE
#define BT_PIPE 1
#endif
#ifndef TP_PIPE
#define TP_PIPE 1
#endif
constexpr int BT_KSMAX = 8;                         // k-steps the pipelined kernel supports (l + u + 1 <= 25)
#ifndef BTP_NG
#define BTP_NG 8                                    // compute warps of the pipelined kernel
#endif
#ifndef BTP_NROWS
#define BTP_NROWS 128
#endif
#ifndef BTP_NVEC
#define BTP_NVEC 32
#endif
constexpr int BTP_ROWS = BTP_NROWS;                 // rows per tile of the pipelined kernel
constexpr int BTP_VEC = BTP_NVEC;                   // vectors per tile (<= 32: one loader / storer lane per column)
#ifndef BTP_GROUPS
#define BTP_GROUPS 1                                // independent compute groups working on consecutive pipeline stages (see below)
#endif
constexpr int BTP_NGG = BTP_NG / BTP_GROUPS;        // warps per compute group
constexpr int BTP_RB = BTP_ROWS / 8 / BTP_NGG;      // 8-row blocks per compute warp
constexpr int BTP_NAV = (BTP_ROWS / 8) * BT_KSMAX * 32 / (32 * BTP_NGG);   // A-fragment entries per compute thread
#ifndef BT_GRIDCAP
#define BT_GRIDCAP 1024                             // CTAs per SM in the grid at most (items beyond that are strided)
#endif
// x-tile pitch = BT_PITCH_MOD (mod 16): the B-fragment reads (lane (gq, tq) reads column gq, row 4ks + tq) are then
// bank-conflict free; the movers run along the rows and do not care
__host__ __device__ inline int bt_pitch(int rows) { return BT_PITCH_MOD < 0 ? (rows | 1) : rows + ((BT_PITCH_MOD - rows) % 16 + 16) % 16; }

__global__ void __launch_bounds__(BT_THREADS)
banded_apply_kernel(const double *__restrict__ band, i64 m, i64 n, int l, int u,
                    const double *__restrict__ X, i64 ldx, i64 nvec, int vchunk, i64 nvc_big,
                    double alpha, double beta, double *__restrict__ Y, i64 ldy) {
    extern __shared__ __align__(16) double sm[];
    const int W = l + u + 1;
    const int KS = (8 + W - 1 + 3) / 4;             // k-steps per 8-row block
    const int xrows = BT_ROWS + l + u;
    const int pitch = bt_pitch(xrows + 4);              // + room for the k padding of the last block
    double *as = sm;                                // [BT_ROWS/8][KS][32] A fragments
    double *xs = sm + (BT_ROWS / 8) * KS * 32;      // [BT_VEC][pitch]; tile row r <-> matrix column i0 - l + r
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gq = lane >> 2, tq = lane & 3;
    const i64 nrt = (m + BT_ROWS - 1) / BT_ROWS;
    const i64 nvt = (nvec + BT_VEC - 1) / BT_VEC;
    // work item = (row tile, chunk of vchunk vector tiles): the band tile is staged once per item
    // the first nvc_big chunks take vchunk vector tiles each, the remaining vector tiles are items of their own: the
    // hardware hands the small items out last, so the SMs run dry within one small item of each other
    const i64 nitems = nrt * (nvc_big + (nvt - nvc_big * vchunk));
    for (i64 item = blockIdx.x; item < nitems; item += gridDim.x) {
      const i64 vc = item / nrt, rt = item - vc * nrt;              // row tiles vary fastest (contiguous streams)
      const i64 vt_begin = vc < nvc_big ? vc * vchunk : nvc_big * vchunk + (vc - nvc_big);
      const i64 vt_end = vc < nvc_big ? vt_begin + vchunk : vt_begin + 1;
      const i64 i0 = rt * BT_ROWS;                  // first row of the tile
      __syncthreads();                              // previous item fully consumed
      // A fragments: as[rb][ks][lane] = A(i0 + 8rb + g, i0 + 8rb - l + 4ks + t), 0 outside band/matrix
      for (int e = threadIdx.x; e < (BT_ROWS / 8) * KS * 32; e += BT_THREADS) {
          const int ln = e & 31, ks = (e >> 5) % KS, rb = (e >> 5) / KS;
          const i64 i = i0 + 8 * rb + (ln >> 2), j = i0 + 8 * rb - l + 4 * ks + (ln & 3);
          double a = 0.0;
          if (i < m && j >= 0 && j < n && j - i <= u && i - j <= l) a = __ldg(band + (u + i - j) + (i64)W * j);
          as[e] = a;
      }
      for (i64 vt = vt_begin; vt < vt_end; ++vt) {
        const i64 v0 = vt * BT_VEC;
        __syncthreads();                            // previous tile fully consumed
        // x window rows i0-l .. i0+BT_ROWS-1+u (+ k padding), zero outside the matrix
        load_tile_async_rt<BT_VEC, BT_NG>(xs, pitch, X, ldx, n, i0 - l, xrows + 3, v0, nvec);
        cp_async_commit();
        cp_async_wait<0>();
        __syncthreads();
        // A dependent DMMA has ~180 cycles of latency (scripts/micro/dmma_lat.cu): the k-steps of the warp's BT_RB row
        // blocks advance together, BT_RB * 4 independent accumulator chains in flight
        double c[BT_RB][BT_VEC / 8][2];
#pragma unroll
        for (int q = 0; q < BT_RB; ++q)
#pragma unroll
            for (int nt = 0; nt < BT_VEC / 8; ++nt) { c[q][nt][0] = 0.0; c[q][nt][1] = 0.0; }
        {
            const double *af = as + warp * BT_RB * KS * 32 + lane;
            const double *xb = xs + gq * pitch + 8 * warp * BT_RB + tq;   // + q*8 + nt*8*pitch + 4*ks
            for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
                for (int q = 0; q < BT_RB; ++q) {
                    const double a = af[(q * KS + ks) * 32];
#pragma unroll
                    for (int nt = 0; nt < BT_VEC / 8; ++nt)
                        dmma884(c[q][nt][0], c[q][nt][1], a, xb[q * 8 + nt * 8 * pitch + 4 * ks]);
                }
            }
        }
        __syncthreads();                            // everyone done reading xs
#pragma unroll
        for (int q = 0; q < BT_RB; ++q) {
            const int row = (warp * BT_RB + q) * 8 + gq;             // tile-local output row
#pragma unroll
            for (int nt = 0; nt < BT_VEC / 8; ++nt) {
                const int v = nt * 8 + 2 * tq;
                xs[v * pitch + row] = c[q][nt][0];
                xs[(v + 1) * pitch + row] = c[q][nt][1];
            }
        }
        __syncthreads();
        const int r_hi = m - i0 < BT_ROWS ? (int)(m - i0) : BT_ROWS;
        if (beta == 0.0) store_tile_fast_clip<BT_VEC, BT_NG>(xs, pitch, Y, ldy, i0, 0, r_hi, v0, nvec, alpha);
        else store_tile<BT_THREADS>(xs, pitch, Y, ldy, m, i0, 0, BT_ROWS, v0, BT_VEC, nvec, alpha, beta);
      }
    }
}

#ifdef FEDVR_PHASE_PROF
__device__ unsigned long long fe_prof[8];
#define FE_PROF_T(i) do { if (threadIdx.x == 0) { long long t__ = clock64(); pf[i] += t__ - tprev; tprev = t__; } } while (0)
#define FE_PROF_L(i) do { if (lane == 0 && (w == 0 || w == PROF_NG || w == PROF_NG + 1)) { long long t__ = clock64(); pf[i] += t__ - tprev; tprev = t__; } } while (0)
#else
#define FE_PROF_T(i) do { } while (0)
#define FE_PROF_L(i) do { } while (0)
#endif
// ---------------------------------------------------------------------------
// K7, warp-specialised TMA pipeline (same structure as the K8 pipeline below, which has the measurements): one
// persistent CTA per SM; a loader warp keeps a ring of x windows filled with one bulk copy per vector column, the
// compute warps multiply a window by the band tile's A fragments (staged once per work item) and write the 128 result
// rows over the window in place (rows l .. l+127 of the window are the tile's own rows), a storer warp streams them
// out with bulk copies.  The compute warps can be split into BTP_GROUPS independent groups that take the pipeline
// stages in turn (group g: every BTP_GROUPS-th vector tile, own copy of the A fragments, own named barrier), so that one
// group multiplies the next window while the other sits in the barrier + write-back of its tile.  Measured (round 2,
// gpurun_out/bench_r2v.json): two groups of four warps are SLOWER than one group of eight — k = 7, 1006 x 65 536:
// 211 vs 193 us; C3, 200 009 x 256: 206 vs 182 us — a tile then takes twice as long per group and the two groups
// contend for the same DMMA pipe, while 2 x 8 warps do not fit the register file at 158 registers per thread.  The
// default stays one group.  Work item = (row tile, chunk of vector tiles) as in banded_apply_kernel; items are dealt
// round-robin to the CTAs.  Columns whose window starts on an odd element are copied from one element earlier
// (16-byte alignment of bulk copies): window row r of column v sits at xs[col(v) + par_v + r].  Requires l, u >= 0,
// l + u + 1 <= 25, beta == 0 and X, Y of equal 16-byte phase per column; the launcher falls back to
// banded_apply_kernel otherwise.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * (BTP_NG + 2), 1)
banded_apply_pipe_kernel(const double *__restrict__ band, i64 m, i64 n, int l, int u,
                         const double *__restrict__ X, i64 ldx, i64 nvec, int vchunk, int P0, int nstg,
                         double alpha, double beta, double *__restrict__ Y, i64 ldy) {
    extern __shared__ __align__(128) double sm[];
    const int W = l + u + 1;
    const int KS = (8 + W - 1 + 3) / 4;             // k-steps per 8-row block
    const int xrows = BTP_ROWS + l + u;
    const int wrows = xrows + 3;                    // + the k padding of the last row block
    const int STAGE = BTP_VEC * P0;
    const int ASZ = (BTP_ROWS / 8) * KS * 32;
    double *as_all = sm + (size_t)nstg * STAGE;     // [BTP_GROUPS][BTP_ROWS/8][KS][32] A fragments of the groups' current items
    uint64_t *bars = reinterpret_cast<uint64_t *>(as_all + BTP_GROUPS * ASZ);
    uint64_t *full = bars, *done = bars + nstg, *empty = bars + 2 * nstg;
    const int w = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < nstg; ++s) { mbar_init(full + s, 1); mbar_init(done + s, BTP_NGG); mbar_init(empty + s, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const i64 nrt = (m + BTP_ROWS - 1) / BTP_ROWS;
    const i64 nvt = (nvec + BTP_VEC - 1) / BTP_VEC;
    const i64 nvc = (nvt + vchunk - 1) / vchunk;
    const i64 nitems = nrt * nvc;
    const i64 xbase = (i64)(reinterpret_cast<uintptr_t>(X) >> 3);
    int stage = 0;
    unsigned phase = 0;
    constexpr int PROF_NG = BTP_NG; (void)PROF_NG;
#ifdef FEDVR_PHASE_PROF
    long long pf[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
#endif
    auto next_stage = [&]() { if (++stage == nstg) { stage = 0; phase ^= 1; } };
    // parity shift of column v: window row r (matrix column g0 + r) of vector v sits at xs[col(v) + par + r]
    auto col_parity = [&](i64 g0, i64 v) { return (int)((xbase + v * ldx + g0) & 1); };
    // item -> (vector chunk, row tile): row tiles vary fastest (CTAs running at one time stream contiguous rows of the
    // same columns); the row tile is rotated by the chunk index so that a CTA's items (stride gridDim.x) do not keep
    // hitting the same residue of row tiles — the clipped first / last row tiles then spread evenly over the CTAs
    auto item_coords = [&](i64 item, i64 &vc, i64 &rt) { vc = item / nrt; rt = (item - vc * nrt + vc) % nrt; };
    if (w < BTP_NG) {
        // ================= compute warps: group grp takes the tiles with (running tile count) % BTP_GROUPS == grp ====
        const int grp = w / BTP_NGG, wg = w - grp * BTP_NGG;          // group, warp inside the group
        const int tg = (int)threadIdx.x - grp * 32 * BTP_NGG;         // thread inside the group
        double *as = as_all + grp * ASZ;
        const int gq = lane >> 2, tq = lane & 3;
        const double wscale = alpha;                // beta == 0 (launcher): alpha is folded into the tile
        // A fragments of a row tile: as[rb][ks][lane] = A(i0 + 8rb + g, i0 + 8rb - l + 4ks + t), 0 outside band/matrix.
        // Each thread fetches its BTP_NAV entries of the NEXT item into registers while the last vector tile of the
        // current item is in flight, so the band's memory latency is off the tile loop.
        double av[BTP_NAV];
        auto fetch_frags = [&](i64 item_) {
            i64 vc_, rt_; item_coords(item_, vc_, rt_);
            const i64 i0_ = rt_ * BTP_ROWS;
#pragma unroll
            for (int t = 0; t < BTP_NAV; ++t) {
                const int e = tg + t * 32 * BTP_NGG;
                const int ln = e & 31, ks = (e >> 5) % KS, rb = (e >> 5) / KS;
                const i64 i = i0_ + 8 * rb + (ln >> 2), j = i0_ + 8 * rb - l + 4 * ks + (ln & 3);
                av[t] = (e < ASZ && i < m && j >= 0 && j < n && j - i <= u && i - j <= l)
                            ? __ldg(band + (u + i - j) + (i64)W * j) : 0.0;
            }
        };
        if ((i64)blockIdx.x < nitems) fetch_frags(blockIdx.x);
        unsigned tcount = 0;                         // tiles seen so far (all groups see the same sequence)
        for (i64 item = blockIdx.x; item < nitems; item += gridDim.x) {
            i64 vc, rt; item_coords(item, vc, rt);
            const i64 i0 = rt * BTP_ROWS, g0 = i0 - l;
            named_bar_sync(1 + grp, 32 * BTP_NGG);   // the group's previous item's fragments are no longer read
#pragma unroll
            for (int t = 0; t < BTP_NAV; ++t) {
                const int e = tg + t * 32 * BTP_NGG;
                if (e < ASZ) as[e] = av[t];
            }
            named_bar_sync(1 + grp, 32 * BTP_NGG);
            FE_PROF_L(7);
            const i64 vt_end = (vc + 1) * vchunk < nvt ? (vc + 1) * vchunk : nvt;
            for (i64 vt = vc * vchunk; vt < vt_end; ++vt, next_stage(), ++tcount) {
                if (vt + 1 == vt_end && item + gridDim.x < nitems) fetch_frags(item + gridDim.x);
                if ((int)(tcount % BTP_GROUPS) != grp) continue;      // the other group's tile
                const i64 v0 = vt * BTP_VEC;
                double *xs = sm + (size_t)stage * STAGE;
                const int par_r = col_parity(g0, v0 + gq);
                const int par_c0 = col_parity(g0, v0 + 2 * tq), par_c1 = col_parity(g0, v0 + 2 * tq + 1);
                mbar_wait(full + stage, phase);
                FE_PROF_L(2);
                // A dependent DMMA has ~180 cycles of latency: the k-steps of the warp's BTP_RB row blocks advance
                // together, BTP_RB * 4 independent accumulator chains in flight
                double c[BTP_RB][BTP_VEC / 8][2];
#pragma unroll
                for (int q = 0; q < BTP_RB; ++q)
#pragma unroll
                    for (int nt = 0; nt < BTP_VEC / 8; ++nt) { c[q][nt][0] = 0.0; c[q][nt][1] = 0.0; }
                {
                    const double *af = as + wg * BTP_RB * KS * 32 + lane;
                    const double *xb = xs + gq * P0 + fe_swz(gq) + par_r + 8 * wg * BTP_RB + tq;   // + q*8 + nt*8*P0 + 4*ks
                    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
                        for (int q = 0; q < BTP_RB; ++q) {
                            const double a = af[(q * KS + ks) * 32];
#pragma unroll
                            for (int nt = 0; nt < BTP_VEC / 8; ++nt)
                                dmma884(c[q][nt][0], c[q][nt][1], a, xb[q * 8 + nt * 8 * P0 + 4 * ks]);
                        }
                    }
                }
                FE_PROF_L(3);
                named_bar_sync(1 + grp, 32 * BTP_NGG);   // every warp of the group has read the window
                FE_PROF_L(4);
#pragma unroll
                for (int q = 0; q < BTP_RB; ++q) {
                    const int row = l + (wg * BTP_RB + q) * 8 + gq;        // window row of the output row
                    double *y0 = xs + (2 * tq) * P0 + fe_swz(2 * tq) + par_c0 + row;
                    double *y1 = xs + (2 * tq + 1) * P0 + fe_swz(2 * tq + 1) + par_c1 + row;
#pragma unroll
                    for (int nt = 0; nt < BTP_VEC / 8; ++nt) {
                        y0[nt * 8 * P0] = wscale * c[q][nt][0];
                        y1[nt * 8 * P0] = wscale * c[q][nt][1];
                    }
                }
                fence_proxy_async();                // the storer's bulk copies read what this thread wrote
                __syncwarp();
                if (lane == 0) mbar_arrive(done + stage);
                FE_PROF_L(7);
            }
        }
    } else if (w == BTP_NG) {
        // ================= loader warp: lane <-> vector column ================================
        for (i64 item = blockIdx.x; item < nitems; item += gridDim.x) {
            i64 vc, rt; item_coords(item, vc, rt);
            const i64 g0 = rt * BTP_ROWS - l;        // matrix column of window row 0
            const i64 vt_end = (vc + 1) * vchunk < nvt ? (vc + 1) * vchunk : nvt;
            for (i64 vt = vc * vchunk; vt < vt_end; ++vt, next_stage()) {
                const i64 v0 = vt * BTP_VEC;
                double *xs = sm + (size_t)stage * STAGE;
                mbar_wait(empty + stage, phase ^ 1);
                FE_PROF_L(0);
                const int cl = lane < BTP_VEC ? lane : 0;             // lanes beyond the tile's columns idle
                if (g0 >= 1 && g0 + wrows + 1 <= n && v0 + BTP_VEC <= nvec) {
                    const int par = col_parity(g0, v0 + cl);
                    const unsigned bytes = lane < BTP_VEC ? 8u * (unsigned)((wrows + par + 1) & ~1) : 0u;
                    const unsigned total = __reduce_add_sync(0xffffffffu, bytes);
                    if (lane == 0) mbar_arrive_expect_tx(full + stage, total);
                    __syncwarp();
                    if (lane < BTP_VEC) bulk_load(xs + cl * P0 + fe_swz(cl), X + (v0 + cl) * ldx + g0 - par, bytes, full + stage);
                } else {
                    // clipped window (first / last row tiles, partial vector tile): valid rows through the TMA unit as
                    // well, the rest zero
                    const bool vok = lane < BTP_VEC && v0 + lane < nvec;
                    const int par = col_parity(g0, v0 + cl);
                    int ra = g0 < 0 ? (int)(-g0 < wrows ? -g0 : wrows) : 0;
                    int rb = g0 + wrows > n ? (int)(n - g0 > ra ? n - g0 : ra) : wrows;
                    if (!vok) { ra = 0; rb = 0; }
                    double *scol = xs + cl * P0 + fe_swz(cl) + par;
                    const double *gcol = X + (vok ? v0 + lane : 0) * ldx + g0;
                    const ColPlan cp = tma_col_load_prepare(scol, gcol, par, ra, rb, lane < BTP_VEC ? wrows : 0);
                    const unsigned total = __reduce_add_sync(0xffffffffu, 8u * (unsigned)cp.cnt);
                    __syncwarp();
                    if (lane == 0) { if (total) mbar_arrive_expect_tx(full + stage, total); else mbar_arrive(full + stage); }
                    __syncwarp();
                    if (cp.cnt) bulk_load(scol + cp.a, gcol + cp.a, 8u * (unsigned)cp.cnt, full + stage);
                }
                FE_PROF_L(1);
            }
        }
    } else {
        // ================= storer warp: lane <-> vector column ================================
        for (i64 item = blockIdx.x; item < nitems; item += gridDim.x) {
            i64 vc, rt; item_coords(item, vc, rt);
            const i64 i0 = rt * BTP_ROWS, g0 = i0 - l;
            const i64 vt_end = (vc + 1) * vchunk < nvt ? (vc + 1) * vchunk : nvt;
            for (i64 vt = vc * vchunk; vt < vt_end; ++vt, next_stage()) {
                const i64 v0 = vt * BTP_VEC;
                const double *xs = sm + (size_t)stage * STAGE;
                mbar_wait(done + stage, phase);
                FE_PROF_L(5);
                {
                    // rows i0 .. of column lane: the 16-byte aligned part as a bulk copy, an odd end element through the
                    // lane (X, Y and the window have the same parity per column, checked by the launcher)
                    const int cl = lane < BTP_VEC ? lane : 0;
                    const bool vok = lane < BTP_VEC && v0 + lane < nvec;
                    const int par = col_parity(g0, v0 + cl);
                    const int r_hi = !vok ? 0 : m - i0 < BTP_ROWS ? (int)(m - i0) : BTP_ROWS;
                    const double *ys = xs + cl * P0 + fe_swz(cl) + par + l;
                    double *yc = Y + (vok ? v0 + lane : 0) * ldy + i0;
                    tma_col_store(yc, ys, par + l, 0, r_hi);
                }
                __syncwarp();                       // every lane has read its part of the tile
                if (lane == 0) mbar_arrive(empty + stage);
                FE_PROF_L(6);
            }
        }
        bulk_wait0();
    }
#ifdef FEDVR_PHASE_PROF
    if (lane == 0 && (w == 0 || w == PROF_NG || w == PROF_NG + 1))
        for (int i = 0; i < 8; ++i) atomicAdd(&fe_prof[i], (unsigned long long)pf[i]);
#endif
}

// ---------------------------------------------------------------------------
// K9, warp-specialised TMA pipeline (same roles as the K7 / K8 pipelines): a loader warp fills a ring of x windows
// (TP_ROWS + 2 rows x 32 vectors, one bulk copy per column), TP_CW compute warps apply the three-point stencil with
// lane <-> row (the coefficients of a row stay in registers for all the vector tiles of a work item) into a separate
// y tile, a storer warp streams the y tiles out.  x buffer of a stage: free again once the compute warps are done
// (the loader waits on `done`); y buffer: free once the storer has read it (the compute warps wait on `empty`).
// Same arithmetic, in the same order, as tridiag_apply_kernel.  Requires beta == 0 and X, Y of equal 16-byte phase.
// ---------------------------------------------------------------------------
// Measured (N = 1e7 x 128 vectors, device time): rows x vectors per tile 64x32: 5.96 ms | 96x32: 4.24 | 128x32: 3.45 |
// 160x32: 3.59 | 192x32: 3.30 | 256x16: 3.28 | 512x8: 3.25 | 384x16: 3.13 (previous kernel: 3.71) — the longer the bulk copies
// (3 KB per column at 384 rows) the better the TMA unit does, as long as two stages still fit.
#ifndef TP_CWARPS
#define TP_CWARPS 12
#endif
constexpr int TP_CW = TP_CWARPS;                    // compute warps
constexpr int TP_ROWS = 32 * TP_CW;                 // rows per tile
#ifndef TP_NVEC
#define TP_NVEC 16
#endif
constexpr int TP_VEC = TP_NVEC;                     // vectors per tile (<= 32: one loader / storer lane per column)
constexpr int TP_PX = TP_ROWS + 4;                  // x column pitch: halo rows + parity shift, even
constexpr int TP_PY = TP_ROWS + 2;                  // y column pitch: parity shift, even
constexpr int TP_STAGE = TP_VEC * (TP_PX + TP_PY);  // doubles per stage
constexpr int TP_NSTG = (227 * 1024 - 256) / (TP_STAGE * 8) > 6 ? 6 : (227 * 1024 - 256) / (TP_STAGE * 8);

__global__ void __launch_bounds__(32 * (TP_CW + 2), 1)
tridiag_apply_pipe_kernel(const double *__restrict__ dl, const double *__restrict__ d, const double *__restrict__ du,
                          i64 n, const double *__restrict__ X, i64 ldx, i64 nvec, int vchunk,
                          double alpha, double *__restrict__ Y, i64 ldy) {
    extern __shared__ __align__(128) double sm[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(sm + (size_t)TP_NSTG * TP_STAGE);
    uint64_t *full = bars, *done = bars + TP_NSTG, *empty = bars + 2 * TP_NSTG;
    const int w = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < TP_NSTG; ++s) { mbar_init(full + s, 1); mbar_init(done + s, TP_CW); mbar_init(empty + s, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const i64 nrt = (n + TP_ROWS - 1) / TP_ROWS;
    const i64 nvt = (nvec + TP_VEC - 1) / TP_VEC;
    const i64 nvc = (nvt + vchunk - 1) / vchunk;
    const i64 nitems = nrt * nvc;
    const i64 xbase = (i64)(reinterpret_cast<uintptr_t>(X) >> 3);
    const int ldodd = (int)(ldx & 1);
    int stage = 0;
    unsigned phase = 0;
    auto next_stage = [&]() { if (++stage == TP_NSTG) { stage = 0; phase ^= 1; } };
    // window row 0 of a tile is global row g0 = i0 - 1; it sits at xs[c * TP_PX + par_c], par_c = parity of its address
    auto col_parity = [&](i64 g0, i64 v) { return (int)((xbase + v * ldx + g0) & 1); };
    if (w < TP_CW) {
        // ================= compute warps: lane <-> row =========================================
        const int r = 32 * w + lane;
        for (i64 item = blockIdx.x; item < nitems; item += gridDim.x) {
            const i64 vc = item / nrt, rt = item - vc * nrt;             // row tiles vary fastest
            const i64 i0 = rt * TP_ROWS, i = i0 + r;
            const double cd = i < n ? __ldg(d + i) : 0.0;
            const double cl = (i > 0 && i < n) ? __ldg(dl + i - 1) : 0.0;
            const double cu = (i < n - 1) ? __ldg(du + i) : 0.0;
            const i64 vt_end = (vc + 1) * vchunk < nvt ? (vc + 1) * vchunk : nvt;
            for (i64 vt = vc * vchunk; vt < vt_end; ++vt, next_stage()) {
                const double *xs = sm + (size_t)stage * TP_STAGE;
                double *ys = sm + (size_t)stage * TP_STAGE + TP_VEC * TP_PX;
                const int par0 = col_parity(i0 - 1, vt * TP_VEC);
                mbar_wait(full + stage, phase);
                mbar_wait(empty + stage, phase ^ 1);                      // the y buffer's previous tile has left
#pragma unroll 8
                for (int c = 0; c < TP_VEC; ++c) {
                    const int par = (par0 + c * ldodd) & 1;
                    const double *xc = xs + c * TP_PX + par + r;
                    const double left = xc[0], mid = xc[1], right = xc[2];
                    ys[c * TP_PY + (par ^ 1) + r] = alpha * fma(cu, right, fma(cd, mid, cl * left));
                }
                fence_proxy_async();                // the storer's bulk copies read what this thread wrote
                __syncwarp();
                if (lane == 0) mbar_arrive(done + stage);
            }
        }
    } else if (w == TP_CW) {
        // ================= loader warp: lane <-> vector column ================================
        for (i64 item = blockIdx.x; item < nitems; item += gridDim.x) {
            const i64 vc = item / nrt, rt = item - vc * nrt;
            const i64 g0 = rt * TP_ROWS - 1;
            const i64 vt_end = (vc + 1) * vchunk < nvt ? (vc + 1) * vchunk : nvt;
            for (i64 vt = vc * vchunk; vt < vt_end; ++vt, next_stage()) {
                const i64 v0 = vt * TP_VEC;
                double *xs = sm + (size_t)stage * TP_STAGE;
                mbar_wait(done + stage, phase ^ 1);                       // the x buffer's previous tile has been consumed
                const bool vok = lane < TP_VEC && v0 + lane < nvec;
                const int par = col_parity(g0, v0 + lane);
                double *scol = xs + (lane < TP_VEC ? lane : 0) * TP_PX + par;
                const double *gcol = X + (vok ? v0 + lane : 0) * ldx + g0;
                if (g0 >= 1 && g0 + TP_ROWS + 2 + 1 <= n && v0 + TP_VEC <= nvec) {
                    const unsigned bytes = lane < TP_VEC ? 8u * (unsigned)((TP_ROWS + 2 + par + 1) & ~1) : 0u;
                    const unsigned total = __reduce_add_sync(0xffffffffu, bytes);
                    if (lane == 0) mbar_arrive_expect_tx(full + stage, total);
                    __syncwarp();
                    if (lane < TP_VEC) bulk_load(scol - par, gcol - par, bytes, full + stage);
                } else {
                    int ra = g0 < 0 ? (int)(-g0 < TP_ROWS + 2 ? -g0 : TP_ROWS + 2) : 0;
                    int rb = g0 + TP_ROWS + 2 > n ? (int)(n - g0 > ra ? n - g0 : ra) : TP_ROWS + 2;
                    if (!vok) { ra = 0; rb = 0; }
                    const ColPlan cp = tma_col_load_prepare(scol, gcol, par, ra, rb, lane < TP_VEC ? TP_ROWS + 2 : 0);
                    const unsigned total = __reduce_add_sync(0xffffffffu, 8u * (unsigned)cp.cnt);
                    __syncwarp();
                    if (lane == 0) { if (total) mbar_arrive_expect_tx(full + stage, total); else mbar_arrive(full + stage); }
                    __syncwarp();
                    if (cp.cnt) bulk_load(scol + cp.a, gcol + cp.a, 8u * (unsigned)cp.cnt, full + stage);
                }
            }
        }
    } else {
        // ================= storer warp: lane <-> vector column ================================
        for (i64 item = blockIdx.x; item < nitems; item += gridDim.x) {
            const i64 vc = item / nrt, rt = item - vc * nrt;
            const i64 i0 = rt * TP_ROWS;
            const i64 vt_end = (vc + 1) * vchunk < nvt ? (vc + 1) * vchunk : nvt;
            for (i64 vt = vc * vchunk; vt < vt_end; ++vt, next_stage()) {
                const i64 v0 = vt * TP_VEC;
                const double *ys = sm + (size_t)stage * TP_STAGE + TP_VEC * TP_PX;
                mbar_wait(done + stage, phase);
                const bool vok = lane < TP_VEC && v0 + lane < nvec;
                const int pary = col_parity(i0 - 1, v0 + lane) ^ 1;       // parity of row i0 in X, Y and the y tile
                const int r_hi = !vok ? 0 : n - i0 < TP_ROWS ? (int)(n - i0) : TP_ROWS;
                tma_col_store(Y + (vok ? v0 + lane : 0) * ldy + i0, ys + (lane < TP_VEC ? lane : 0) * TP_PY + pary, pary, 0, r_hi);
                __syncwarp();
                if (lane == 0) mbar_arrive(empty + stage);
            }
        }
        bulk_wait0();
    }
}

// ---------------------------------------------------------------------------
// K8: FE-DVR element-block operator, uniform order O (compile time).
//
// A tile is ET = NG - 1 elements x 32 vectors; tile rows (local r <-> global node eA*(O-1) + r): elements
// eA..eA+ET-1 occupy rows 0..ET*(O-1), the halo element eA+ET — needed for the last bridge row — adds O-1 more.
// Warp g owns element eA + g (the last warp: row 0 of the halo element).  Per tile the warp fetches its O x O block
// straight from global memory (L2-resident, 8 MB at C2) into FP64 tensor-core A fragments, multiplies in place
// (DMMA 8x8x4: block fragments x four 8-vector tiles); the bridge rows shared by neighbouring elements are summed
// from two side buffers.  Two kernels share this arithmetic: the warp-specialised TMA pipeline (default) and the
// "lean" kernel whose warps load, compute and store in turn (fallback when X and Y differ in 16-byte phase).
//
// History of the design, device time of one C2 apply (order 10, 1e4 elements, 1024 vectors) on B200:
//   DFMA compute, blocks staged in shared memory per tile: 458 us;  DMMA compute, one element per warp, 16 warps: 386;
//   block fragments from global, even x-tile pitch: 368 (movers only: 289);  fragments kept for 2/4/8 vector tiles:
//   398/432/483;  half of the DMMAs removed: 379;  accumulators stored straight to global: 476;  cp.async tile ring of
//   2/3 stages per CTA: 390/520;  grid of 1/2/3/6 waves of resident CTAs: 381/371/363/362;  instruction count halved
//   (lean kernel): 364;  warp-specialised with cp.async loader and LDS/STG storer warps: 378-409 (the movers fill the
//   LSU queue the compute warps' LDS/STS need);  warp-specialised with TMA bulk copies: 267, clipped tiles through the
//   TMA unit too: 254;  two element slots per warp x 16 vectors (2.3 KB bulk copies instead of 1.2 KB, half the halo
//   share, same registers and DMMA count): 274 — the longer copies that carry K9 do not pay here.
// ---------------------------------------------------------------------------
constexpr int FE_TV = 32;                           // vectors per tile (one warp wide)
constexpr int FE_BRP = FE_TV + 1;                   // pitch of the bridge buffers
#ifndef FEDVR_NG
#define FEDVR_NG 16
#endif
#ifndef FE_PITCH_MOD
#define FE_PITCH_MOD 12
#endif
// Pitch of the lean kernel's x tile: = FE_PITCH_MOD (mod 16) doubles.  The B-fragment reads (lane (gq, tq) reads column
// gq, row 4ks + tq) are bank-conflict free for 12 (mod 16); the movers run along the rows and do not care.
__host__ __device__ constexpr int fe_pitch(int rows) { return FE_PITCH_MOD < 0 ? (rows | 1) : rows + ((FE_PITCH_MOD - rows) % 16 + 16) % 16; }
#ifndef FEDVR_MINB
#define FEDVR_MINB 3
#endif
#ifndef FEDVR_PIPE
#define FEDVR_PIPE 1
#endif
#ifndef FEDVR_WAVES
#define FEDVR_WAVES 4                               // waves of resident CTAs in the lean kernel's grid
#endif

// K8, lean form.  Its predecessor spent most of its issue slots on bookkeeping:
// the ncu source page (profiles/r1t) showed 10.5 k warp instructions per tile of which ~1.4 k were the
// copies, DMMAs and fragment traffic — the rest was 64-bit tile arithmetic done twice per tile, divergence bookkeeping
// around every predicated DMMA (the warp index was not known to be warp-uniform), runtime-bounded mover loops and
// three-way branches in the write-back.  Here interior tiles (every element, row and vector present, beta == 0, not
// the first tile of a vector group) take a straight-line path: warp index through a shuffle (uniform), tile
// coordinates carried incrementally, compile-time mover loops with immediate offsets, per-lane write-back
// destinations computed once per tile.  Every other tile runs the clipped path of the original kernel.
template <int O, int NG>
__global__ void __launch_bounds__(32 * NG, (NG <= 8 ? 4 : FEDVR_MINB))
fedvr_apply_lean_kernel(const double *__restrict__ blocks, i64 nel, i64 first0 /*0-based first function*/, i64 nrows,
                        const double *__restrict__ X, i64 ldx, i64 nvec,
                        double alpha, double beta, double *__restrict__ Y, i64 ldy) {
    constexpr int THREADS = 32 * NG;
    constexpr int ET = NG - 1;                      // elements per tile
    constexpr int ROWS = (ET + 1) * (O - 1) + 1;    // incl. halo element
    constexpr int PITCH = fe_pitch(ROWS);
    constexpr int STAGE = (FE_TV * PITCH + 1) / 2 * 2;
    constexpr int MT = (O + 7) / 8, KS = (O + 3) / 4;
    constexpr int VPW = FE_TV / NG;                 // vectors per warp in the movers
    constexpr int NIT = (ROWS + 31) / 32;
    constexpr int NR = ET * (O - 1);                // rows an interior tile owns: tile rows 1 .. NR
    constexpr int NSI = (NR + 31) / 32;
    static_assert(FE_TV % NG == 0, "movers assign whole vectors to warps");
    extern __shared__ __align__(16) double sm[];
    double *xs = sm;
    double *br0 = sm + STAGE;                       // [(ET+1)][33] first-row products (bridge contributions)
    double *br9 = br0 + (ET + 1) * FE_BRP;          // [ET][33]     last-row products
    const int g = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // warp-uniform for the compiler as well
    const int lane = threadIdx.x & 31, gq = lane >> 2, tq = lane & 3;
    const i64 net = (nel + ET - 1) / ET;
    const i64 nvt = (nvec + FE_TV - 1) / FE_TV;
    // tile = vt * net + et: element tiles vary fastest so that the CTAs running at one time stream long contiguous row
    // ranges of a few vector groups; (et, vt) advance by gridDim.x tiles without a division
    i64 vt = (i64)blockIdx.x / net, et = (i64)blockIdx.x - vt * net;
#ifdef FEDVR_PHASE_PROF
    long long pf[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
#endif
    for (; vt < nvt;) {
        const i64 eA = et * ET, v0 = vt * FE_TV;
        const i64 x0row = eA * (O - 1) - first0;    // X/Y row of tile row 0 (rows are relative to the restriction)
        const bool interior = x0row >= 0 && x0row + ROWS <= nrows && v0 + FE_TV <= nvec && eA + ET + 1 <= nel;
        if (interior && eA > 0 && beta == 0.0) {
            // ---- x tile: warp g copies vectors g, g + NG, ... ----------------------------------
            {
                const double *xc = X + (v0 + g) * ldx + x0row + lane;
                double *xd = xs + g * PITCH + lane;
#pragma unroll
                for (int h = 0; h < VPW; ++h)
#pragma unroll
                    for (int i = 0; i < NIT; ++i)
                        if (32 * i + 31 < ROWS || lane + 32 * i < ROWS)
                            cp_async8(xd + h * NG * PITCH + 32 * i, xc + (i64)h * NG * ldx + 32 * i);
                cp_async_commit();
            }
            // ---- block fragments a = B_e[gq + 8 mt][4 ks + tq], zero outside the block ---------
            double af[MT][KS];
            {
                const double *b = blocks + (eA + g) * (O * O) + tq * O + gq;
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks)
                        af[mt][ks] = (gq + 8 * mt < O && 4 * ks + tq < O) ? __ldg(b + 4 * ks * O + 8 * mt) : 0.0;
            }
            FE_PROF_T(0);
            cp_async_wait<0>();
            FE_PROF_T(1);
            __syncthreads();
            FE_PROF_T(2);
            if (g < ET) {
                // write-back destination of this lane's accumulator rows gq + 8 mt: row 0 and row O-1 are bridge
                // contributions (side buffers), rows 1..O-2 go back into the tile in place
                int dst[MT], str[MT], nst[MT];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    const int mm = gq + 8 * mt;
                    const bool side = (mm == 0) || (mm == O - 1);
                    dst[mt] = mm == 0 ? STAGE + g * FE_BRP + 2 * tq
                            : mm == O - 1 ? STAGE + (ET + 1) * FE_BRP + g * FE_BRP + 2 * tq
                            : 2 * tq * PITCH + g * (O - 1) + mm;
                    str[mt] = side ? 1 : PITCH;
                    nst[mt] = side ? 8 : 8 * PITCH;
                }
#pragma unroll
                for (int np = 0; np < FE_TV / 8; ++np) {
                    double c[MT][2];
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) { c[mt][0] = 0.0; c[mt][1] = 0.0; }
                    const double *xe = xs + (8 * np + gq) * PITCH + g * (O - 1) + tq;
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) {
                        const double bf = (4 * ks + 3 < O || 4 * ks + tq < O) ? xe[4 * ks] : 0.0;
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) dmma884(c[mt][0], c[mt][1], af[mt][ks], bf);
                    }
                    __syncwarp();                   // all lanes have read this vector tile's x rows
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt)
                        if (8 * mt + 7 < O || gq + 8 * mt < O) {
                            sm[dst[mt] + np * nst[mt]] = c[mt][0];
                            sm[dst[mt] + np * nst[mt] + str[mt]] = c[mt][1];
                        }
                }
            } else {
                // halo element: only its first row (the contribution to the tile's last bridge row)
#pragma unroll
                for (int np = 0; np < FE_TV / 8; ++np) {
                    double c0 = 0.0, c1 = 0.0;
                    const double *xe = xs + (8 * np + gq) * PITCH + ET * (O - 1) + tq;
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) {
                        const double bf = (4 * ks + 3 < O || 4 * ks + tq < O) ? xe[4 * ks] : 0.0;
                        dmma884(c0, c1, af[0][ks], bf);
                    }
                    if (gq == 0) { br0[ET * FE_BRP + 8 * np + 2 * tq] = c0; br0[ET * FE_BRP + 8 * np + 2 * tq + 1] = c1; }
                }
            }
            FE_PROF_T(3);
            __syncthreads();
            FE_PROF_T(4);
            // bridge rows: last row of element g-1 plus first row of element g (tile row 0 belongs to the previous tile)
            if (g >= 1) xs[lane * PITCH + g * (O - 1)] = br9[(g - 1) * FE_BRP + lane] + br0[g * FE_BRP + lane];
            __syncthreads();
            FE_PROF_T(5);
            {
                double *yc = Y + (v0 + g) * ldy + x0row + 1 + lane;
                const double *ys = xs + g * PITCH + 1 + lane;
#pragma unroll
                for (int h = 0; h < VPW; ++h)
#pragma unroll
                    for (int i = 0; i < NSI; ++i)
                        if (32 * i + 31 < NR || lane + 32 * i < NR)
                            __stcs(yc + (i64)h * NG * ldy + 32 * i, alpha * ys[h * NG * PITCH + 32 * i]);
            }
            FE_PROF_T(6);
            __syncthreads();                        // the tile may be refilled
            FE_PROF_T(7);
        } else {
            // ---- clipped path: first tile of a vector group, partial tiles, beta != 0 ----------
            load_tile<THREADS>(xs, PITCH, X, ldx, nrows, x0row, ROWS, v0, FE_TV, nvec);
            double af[MT][KS];
            {
                const double *b = blocks + (eA + g) * (O * O);
                const bool have = eA + g < nel;
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int ks = 0; ks < KS; ++ks) {
                        const int mm = gq + 8 * mt, mp = 4 * ks + tq;
                        af[mt][ks] = (have && mm < O && mp < O && (g < ET || mt == 0)) ? __ldg(b + mp * O + mm) : 0.0;
                    }
            }
            __syncthreads();
#pragma unroll
            for (int np = 0; np < FE_TV / 8; ++np) {
                double c[MT][2];
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) { c[mt][0] = 0.0; c[mt][1] = 0.0; }
                const double *xe = xs + (8 * np + gq) * PITCH + g * (O - 1) + tq;
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    const double bf = (4 * ks + tq < O) ? xe[4 * ks] : 0.0;
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) dmma884(c[mt][0], c[mt][1], af[mt][ks], bf);
                }
                __syncwarp();
                const int v = 8 * np + 2 * tq;
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    const int mm = gq + 8 * mt;
                    if (mm == 0) { br0[g * FE_BRP + v] = c[mt][0]; br0[g * FE_BRP + v + 1] = c[mt][1]; }
                    else if (g < ET && mm == O - 1) { br9[g * FE_BRP + v] = c[mt][0]; br9[g * FE_BRP + v + 1] = c[mt][1]; }
                    else if (g < ET && mm < O - 1) {
                        xs[v * PITCH + g * (O - 1) + mm] = c[mt][0];
                        xs[(v + 1) * PITCH + g * (O - 1) + mm] = c[mt][1];
                    }
                }
            }
            __syncthreads();
            // global row 0 has no previous element
            if (g == 0) { if (eA == 0) xs[lane * PITCH] = br0[lane]; }
            else xs[lane * PITCH + g * (O - 1)] = br9[(g - 1) * FE_BRP + lane] + br0[g * FE_BRP + lane];
            __syncthreads();
            // rows owned by this tile: 1 .. ET*(O-1), plus row 0 for the first tile
            const int r_lo = (eA == 0) ? 0 : 1;
            const i64 last_el = eA + ET < nel ? eA + ET : nel;            // elements actually present
            const int r_hi = (int)((last_el - eA) * (O - 1)) + 1;
            store_tile<THREADS>(xs, PITCH, Y, ldy, nrows, x0row, r_lo, r_hi, v0, FE_TV, nvec, alpha, beta);
            __syncthreads();
        }
        et += gridDim.x;
        while (et >= net) { et -= net; ++vt; }
    }
#ifdef FEDVR_PHASE_PROF
    if (threadIdx.x == 0)
        for (int i = 0; i < 8; ++i) atomicAdd(&fe_prof[i], (unsigned long long)pf[i]);
#endif
}

// ---------------------------------------------------------------------------
// K8, warp-specialised TMA pipeline.
//
// The phase clocks of the kernels above (scripts/fe_phase.py) showed where a tile's 14.6 k cycles go when every
// warp loads, computes and stores in turn: 2.6 k issuing 10 cp.async + 6 LDG per warp against a backed-up LSU,
// 2.6 k waiting for them, 3.0 k in the dependent DMMA chains, 2.9 k issuing 10 LDS + 10 STG, 3.5 k in four CTA
// barriers.  Splitting the roles over warps with LSU movers (cp.async loaders, LDS/STG storers) did not help: the
// movers' instructions share the LSU queue with the compute warps' LDS/STS, which then stall behind global-memory
// back-pressure.  Here the tiles move through the TMA unit instead (cp.async.bulk, one 1-D bulk copy per vector
// column issued by one lane each), so the LSU only sees the compute warps:
//   * one loader warp keeps a ring of x tiles filled (completion: expect_tx / complete_tx on an mbarrier),
//   * NG compute warps multiply a tile in place (all four 8-vector tiles interleaved: one DMMA latency chain per
//     tile instead of four; a thread may use ~110 registers now) and sum the bridge rows,
//   * one storer warp streams finished tiles out (bulk stores) and frees the stage.
// Stages hand over through mbarriers (full -> done -> empty); the only wider construct left is a named barrier among
// the compute warps between the products and the sums of the bridge rows.
//
// Bulk copies need 16-byte aligned addresses and sizes, Julia's columns are 8-byte aligned (odd leading dimensions):
// a column whose tile starts on an odd element is copied from one element earlier, i.e. tile row r of column v sits
// at xs[col(v) + p_v + r] with p_v the parity of the element address; the storer writes the one unaligned element at
// the head or tail of each column with an ordinary store.  X and Y must have the same parity per column (same
// alignment class and leading dimensions of equal parity) and beta must be 0 (bulk stores do not read Y), otherwise
// the launcher takes the lean kernel.
//
// Tile columns are swizzled, col(v) = v * P0 + 4 * ((v & 3) ^ ((v >> 2) & 1)) with P0 = 0 (mod 16): the B-fragment
// reads (lane (gq, tq): column gq, row 4 ks + tq) and the accumulator write-back (rows gq, columns 2 tq, 2 tq + 1)
// are then free of bank conflicts per half-warp up to the parity shift, which no plain pitch achieves.
// ---------------------------------------------------------------------------
#ifndef FEDVR_PNTI
#define FEDVR_PNTI 4                                // 8-vector tiles whose DMMA chains are interleaved
#endif
constexpr int FP_NTI = FEDVR_PNTI;
template <int O, int NG> struct FePipe {
    static constexpr int ET = NG - 1;
    static constexpr int ROWS = (ET + 1) * (O - 1) + 1;
    static constexpr int KS = (O + 3) / 4, MT = (O + 7) / 8;
    // rows a column holds: the tile (+ parity shift, rounded to 16 bytes) or the k padding read by the halo warp; + skew
    static constexpr int NEED = (ROWS + 2 > ET * (O - 1) + 4 * KS + 1 ? ROWS + 2 : ET * (O - 1) + 4 * KS + 1) + 12;
    static constexpr int P0 = (NEED + 15) / 16 * 16;
    static constexpr int STAGE = FE_TV * P0;                             // doubles per stage
    static constexpr int BR = (2 * ET + 1) * FE_BRP;                     // one bridge buffer set (br0 + br9)
    static constexpr int THREADS = 32 * (NG + 2);
    static constexpr int max_stages() {
        int n = (int)((227 * 1024 - 2 * BR * 8 - 256) / (STAGE * 8));
        return n > 6 ? 6 : n;
    }
    static constexpr int NSTG = max_stages();
    static constexpr size_t SMEM = (size_t)NSTG * STAGE * 8 + 2 * BR * 8 + 256;
};

template <int O, int NG>
__global__ void __launch_bounds__(FePipe<O, NG>::THREADS, 1)
fedvr_apply_pipe_kernel(const double *__restrict__ blocks, i64 nel, i64 first0 /*0-based first function*/, i64 nrows,
                        const double *__restrict__ X, i64 ldx, i64 nvec,
                        double alpha, double beta, double *__restrict__ Y, i64 ldy) {
    using P = FePipe<O, NG>;
    constexpr int ET = P::ET, ROWS = P::ROWS, P0 = P::P0, STAGE = P::STAGE, NSTG = P::NSTG, MT = P::MT, KS = P::KS;
    constexpr int NR = ET * (O - 1);                // rows an interior tile owns: tile rows 1 .. NR
    constexpr int PROF_NG = NG; (void)PROF_NG;
    extern __shared__ __align__(128) double sm[];
    double *brbuf = sm + (size_t)NSTG * STAGE;      // [2][br0 (ET+1) x 33 | br9 ET x 33]
    uint64_t *bars = reinterpret_cast<uint64_t *>(brbuf + 2 * P::BR);    // full[NSTG], done[NSTG], empty[NSTG]
    uint64_t *full = bars, *done = bars + NSTG, *empty = bars + 2 * NSTG;
    const int w = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);  // warp-uniform for the compiler as well
    const int lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < NSTG; ++s) {
            mbar_init(full + s, 1);                 // the loader's lane 0 (+ the bytes of the bulk copies)
            mbar_init(done + s, NG);                // one lane per compute warp
            mbar_init(empty + s, 1);                // the storer's lane 0
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const i64 net = (nel + ET - 1) / ET;
    const i64 nvt = (nvec + FE_TV - 1) / FE_TV;
    const i64 xbase = (i64)(reinterpret_cast<uintptr_t>(X) >> 3);        // element address of X[0] (parity is what counts)
    // tile = vt * net + et, element tiles fastest; every role walks the same sequence blockIdx.x, + gridDim.x, ...
    i64 vt = (i64)blockIdx.x / net, et = (i64)blockIdx.x - vt * net;
    int stage = 0;
    unsigned phase = 0;
#ifdef FEDVR_PHASE_PROF
    long long pf[8] = {0, 0, 0, 0, 0, 0, 0, 0}, tprev = clock64();
#endif
    auto advance = [&]() {
        et += gridDim.x;
        while (et >= net) { et -= net; ++vt; }
        if (++stage == NSTG) { stage = 0; phase ^= 1; }
    };
    // parity shift of column v of the current tile: tile row r of column v sits at xs[col(v) + par(v) + r]
    auto col_parity = [&](i64 x0row, i64 v0, int v) { return (int)((xbase + (v0 + v) * ldx + x0row) & 1); };
    if (w < NG) {
        // ================= compute warps: y_e = B_e x_e in place ==============================
        const int g = w, gq = lane >> 2, tq = lane & 3;
        const int swz_r = fe_swz(gq);               // column skew of the B-fragment reads (vector 8 np + gq)
        const int swz_c0 = fe_swz(2 * tq), swz_c1 = fe_swz(2 * tq + 1);   // of the accumulator columns 8 np + 2 tq (+ 1)
        const double wscale = alpha;                // beta == 0 (launcher): alpha is folded into the tile
        int parity = 0;                             // bridge buffer set
        // block fragments a = B_e[gq + 8 mt][4 ks + tq], zero outside the block / beyond the last element; the
        // fragments of the next tile are requested as soon as the last DMMA of this one has been issued, so their L2
        // latency hides behind the write-back, the barrier and the bridge sums
        double af[MT][KS];
        auto load_frags = [&](i64 eA_) {
            const bool have = eA_ + g < nel;
            const double *b = blocks + (eA_ + g) * (O * O) + tq * O + gq;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int ks = 0; ks < KS; ++ks)
                    af[mt][ks] = (have && gq + 8 * mt < O && 4 * ks + tq < O && (mt == 0 || g < ET))
                                     ? __ldg(b + 4 * ks * O + 8 * mt) : 0.0;
        };
        // launched behind cb_fedvr_assemble with programmatic stream serialisation (cb_fedvr_assemble_apply): the loader
        // already fills the ring with x tiles while the assembly grid drains; the blocks are read after it has completed
        asm volatile("griddepcontrol.wait;" ::: "memory");
        if (vt < nvt) load_frags(et * ET);
        for (; vt < nvt; parity ^= 1) {
            const i64 eA = et * ET;
            const i64 x0row = eA * (O - 1) - first0;
            double *xs = sm + (size_t)stage * STAGE;
            double *br0 = brbuf + parity * P::BR, *br9 = br0 + (ET + 1) * FE_BRP;
            const int cur_stage = stage;
            const int par_r = col_parity(x0row, vt * FE_TV, gq);          // columns 8 np + gq
            const int par_c0 = col_parity(x0row, vt * FE_TV, 2 * tq), par_c1 = col_parity(x0row, vt * FE_TV, 2 * tq + 1);
            const int par_l = col_parity(x0row, vt * FE_TV, lane);
            FE_PROF_L(7);
            mbar_wait(full + stage, phase);
            FE_PROF_L(2);
            advance();                              // (et, vt, stage, phase) now describe the next tile
#pragma unroll
            for (int np0 = 0; np0 < FE_TV / 8; np0 += FP_NTI) {
                double c[FP_NTI][MT][2];
#pragma unroll
                for (int i = 0; i < FP_NTI; ++i)
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt) { c[i][mt][0] = 0.0; c[i][mt][1] = 0.0; }
                const double *xe = xs + (8 * np0 + gq) * P0 + swz_r + par_r + g * (O - 1) + tq;
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    double bf[FP_NTI];
#pragma unroll
                    for (int i = 0; i < FP_NTI; ++i)
                        bf[i] = (4 * ks + 3 < O || 4 * ks + tq < O) ? xe[i * 8 * P0 + 4 * ks] : 0.0;
#pragma unroll
                    for (int mt = 0; mt < MT; ++mt)
                        if (mt == 0 || g < ET) {
#pragma unroll
                            for (int i = 0; i < FP_NTI; ++i) dmma884(c[i][mt][0], c[i][mt][1], af[mt][ks], bf[i]);
                        }
                }
                if (np0 + FP_NTI >= FE_TV / 8 && vt < nvt) load_frags(et * ET);   // next tile's fragments
                if (wscale != 1.0) {
#pragma unroll
                    for (int i = 0; i < FP_NTI; ++i)
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) { c[i][mt][0] *= wscale; c[i][mt][1] *= wscale; }
                }
                __syncwarp();                       // all lanes have read these vector tiles' x rows
                // rows 1..O-2 go back in place (no other warp reads them); rows 0 and O-1 are contributions to the
                // bridge rows shared with the neighbouring elements and go to the side buffers
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    const int mm = gq + 8 * mt;
                    if (mm == 0) {
#pragma unroll
                        for (int i = 0; i < FP_NTI; ++i) {
                            br0[g * FE_BRP + 8 * (np0 + i) + 2 * tq] = c[i][mt][0];
                            br0[g * FE_BRP + 8 * (np0 + i) + 2 * tq + 1] = c[i][mt][1];
                        }
                    } else if (g < ET && mm == O - 1) {
#pragma unroll
                        for (int i = 0; i < FP_NTI; ++i) {
                            br9[g * FE_BRP + 8 * (np0 + i) + 2 * tq] = c[i][mt][0];
                            br9[g * FE_BRP + 8 * (np0 + i) + 2 * tq + 1] = c[i][mt][1];
                        }
                    } else if (g < ET && mm < O - 1) {
                        double *y0 = xs + (8 * np0 + 2 * tq) * P0 + swz_c0 + par_c0 + g * (O - 1) + mm;
                        double *y1 = xs + (8 * np0 + 2 * tq + 1) * P0 + swz_c1 + par_c1 + g * (O - 1) + mm;
#pragma unroll
                        for (int i = 0; i < FP_NTI; ++i) {
                            y0[i * 8 * P0] = c[i][mt][0];
                            y1[i * 8 * P0] = c[i][mt][1];
                        }
                    }
                }
            }
            FE_PROF_L(3);
            named_bar_sync(1, 32 * NG);             // every product of the tile is in place
            FE_PROF_L(4);
            // bridge rows: last row of element g-1 plus first row of element g; global row 0 has no previous element
            {
                double *col = xs + lane * P0 + fe_swz(lane) + par_l;
                if (g == 0) { if (eA == 0) col[0] = br0[lane]; }
                else col[g * (O - 1)] = br9[(g - 1) * FE_BRP + lane] + br0[g * FE_BRP + lane];
            }
            fence_proxy_async();                    // the storer's bulk copies read what this thread wrote
            __syncwarp();
            if (lane == 0) mbar_arrive(done + cur_stage);
        }
    } else if (w == NG) {
        // ================= loader warp: x tiles into the ring, lane <-> vector column ==========
        for (; vt < nvt; advance()) {
            const i64 eA = et * ET, v0 = vt * FE_TV;
            const i64 x0row = eA * (O - 1) - first0;    // X/Y row of tile row 0 (rows are relative to the restriction)
            double *xs = sm + (size_t)stage * STAGE;
            mbar_wait(empty + stage, phase ^ 1);
            FE_PROF_L(0);
            if (x0row >= 1 && x0row + ROWS + 1 <= nrows && v0 + FE_TV <= nvec) {
                // copy rows x0row - par .. of column lane: 16-byte aligned start, an even number of elements
                const int par = col_parity(x0row, v0, lane);
                const unsigned bytes = 8u * (unsigned)((ROWS + par + 1) & ~1);
                const unsigned total = __reduce_add_sync(0xffffffffu, bytes);
                if (lane == 0) mbar_arrive_expect_tx(full + stage, total);
                __syncwarp();
                bulk_load(xs + lane * P0 + fe_swz(lane), X + (v0 + lane) * ldx + x0row - par, bytes, full + stage);
            } else {
                // clipped tile (the first and last of a vector group, partial vector groups): valid rows through the TMA
                // unit as well, the rest zero
                const bool vok = v0 + lane < nvec;
                const int par = col_parity(x0row, v0, lane);
                int ra = x0row < 0 ? (int)(-x0row < ROWS ? -x0row : ROWS) : 0;
                int rb = x0row + ROWS > nrows ? (int)(nrows - x0row > ra ? nrows - x0row : ra) : ROWS;
                if (!vok) { ra = 0; rb = 0; }
                double *scol = xs + lane * P0 + fe_swz(lane) + par;
                const double *gcol = X + (vok ? v0 + lane : 0) * ldx + x0row;
                const ColPlan cp = tma_col_load_prepare(scol, gcol, par, ra, rb, ROWS);
                const unsigned total = __reduce_add_sync(0xffffffffu, 8u * (unsigned)cp.cnt);
                __syncwarp();
                if (lane == 0) { if (total) mbar_arrive_expect_tx(full + stage, total); else mbar_arrive(full + stage); }
                __syncwarp();
                if (cp.cnt) bulk_load(scol + cp.a, gcol + cp.a, 8u * (unsigned)cp.cnt, full + stage);
            }
            FE_PROF_L(1);
        }
    } else {
        // ================= storer warp: finished tiles out, lane <-> vector column ============
        for (; vt < nvt; advance()) {
            const i64 eA = et * ET, v0 = vt * FE_TV;
            const i64 x0row = eA * (O - 1) - first0;
            const double *xs = sm + (size_t)stage * STAGE;
            mbar_wait(done + stage, phase);
            FE_PROF_L(5);
            {
                // rows owned by this tile: 1 .. (elements present)*(O-1), plus row 0 for the first tile, clipped to the
                // restriction; the 16-byte aligned part of column lane as a bulk copy, an odd end element through the
                // lane (X and Y columns have the same parity, checked by the launcher)
                const bool vok = v0 + lane < nvec;
                const int par = col_parity(x0row, v0, lane);
                const i64 last_el = eA + ET < nel ? eA + ET : nel;
                int ra = (eA == 0) ? 0 : 1, rb = (int)((last_el - eA) * (O - 1)) + 1;
                if (x0row + ra < 0) ra = (int)(-x0row);
                if (x0row + rb > nrows) rb = (int)(nrows - x0row);
                if (!vok || rb < ra) rb = ra;
                const double *ys = xs + lane * P0 + fe_swz(lane) + par;
                double *yc = Y + (vok ? v0 + lane : 0) * ldy + x0row;
                tma_col_store(yc, ys, par, ra, rb);
            }
            __syncwarp();                           // every lane has read its part of the tile
            if (lane == 0) mbar_arrive(empty + stage);
            FE_PROF_L(6);
        }
        bulk_wait0();
    }
#ifdef FEDVR_PHASE_PROF
    if (lane == 0 && (w == 0 || w == PROF_NG || w == PROF_NG + 1))
        for (int i = 0; i < 8; ++i) atomicAdd(&fe_prof[i], (unsigned long long)pf[i]);
#endif
}

// Generic FE-DVR apply (per-element orders): thread per (row, vector); each row
// gathers from the one or two elements that contain it.  rowmap[i] = element of
// node i as owner (the element where it is NOT the first node, except node 0).
__global__ void fedvr_apply_generic_kernel(const double *__restrict__ blocks, const i64 *__restrict__ estart,
                                           const int32_t *__restrict__ order, const i64 *__restrict__ boff,
                                           i64 nel, i64 first0, i64 nrows,
                                           const double *__restrict__ X, i64 ldx, i64 nvec,
                                           double alpha, double beta, double *__restrict__ Y, i64 ldy) {
    for (i64 v = blockIdx.y; v < nvec; v += gridDim.y) {
        const double *xc = X + v * ldx;
        for (i64 r = blockIdx.x * (i64)blockDim.x + threadIdx.x; r < nrows; r += (i64)gridDim.x * blockDim.x) {
            const i64 node = r + first0;
            // last element whose first node is <= node
            i64 lo = 0, hi = nel;
            while (lo < hi) { i64 mid = (lo + hi) >> 1; if (estart[mid] <= node) lo = mid + 1; else hi = mid; }
            i64 e = lo - 1;
            double s = 0.0;
            for (int pass = 0; pass < 2; ++pass) {
                // pass 0: element e (node is local row node-estart[e]);
                // pass 1: element e-1 if node is also its last node (bridge)
                i64 el = pass == 0 ? e : e - 1;
                if (el < 0 || el >= nel) continue;
                int o = order[el];
                i64 loc = node - estart[el];
                if (loc < 0 || loc >= o) continue;
                if (pass == 1 && loc != o - 1) continue;
                const double *b = blocks + boff[el];
                for (int mp = 0; mp < o; ++mp) {
                    i64 xr = estart[el] + mp - first0;
                    double xval = (xr >= 0 && xr < nrows) ? xc[xr] : 0.0;
                    s = fma(b[loc + (i64)o * mp], xval, s);
                }
            }
            Y[r + ldy * v] = axpby(s, alpha, beta, Y + r + ldy * v);
        }
    }
}

// Dense export of the (restricted) operator.
__global__ void fedvr_to_dense_kernel(const double *__restrict__ blocks, const i64 *__restrict__ estart,
                                      const int32_t *__restrict__ order, const i64 *__restrict__ boff,
                                      i64 nel, i64 first0, i64 nrows, double *__restrict__ A) {
    // one block per element; A must be zero-initialised; bridge corners via atomicAdd
    for (i64 e = blockIdx.x; e < nel; e += gridDim.x) {
        int o = order[e];
        const double *b = blocks + boff[e];
        for (int t = threadIdx.x; t < o * o; t += blockDim.x) {
            int mm = t % o, mp = t / o;
            i64 gi = estart[e] + mm - first0, gj = estart[e] + mp - first0;
            if (gi < 0 || gi >= nrows || gj < 0 || gj >= nrows) continue;
            bool shared_corner = (mm == mp) && ((mm == 0 && e > 0) || (mm == o - 1 && e < nel - 1));
            if (shared_corner) atomicAdd(&A[gi + nrows * gj], b[t]);
            else A[gi + nrows * gj] = b[t];
        }
    }
}

// Band export of the (restricted) operator: BandedMatrices storage band[(u + i - j) + (l + u + 1) j] with
// l = u = w >= max order - 1 (an element couples its own functions only, so the operator is banded with half
// bandwidth o - 1).  band must be zero-initialised; the two corners that meet on a bridge function are added atomically
// (two addends: the sum does not depend on their order).
__global__ void fedvr_to_band_kernel(const double *__restrict__ blocks, const i64 *__restrict__ estart,
                                     const int32_t *__restrict__ order, const i64 *__restrict__ boff,
                                     i64 nel, i64 first0, i64 nrows, int w, double *__restrict__ band) {
    for (i64 e = blockIdx.x; e < nel; e += gridDim.x) {
        const int o = order[e];
        const double *b = blocks + boff[e];
        for (int t = threadIdx.x; t < o * o; t += blockDim.x) {
            const int mm = t % o, mp = t / o;
            const i64 gi = estart[e] + mm - first0, gj = estart[e] + mp - first0;
            if (gi < 0 || gi >= nrows || gj < 0 || gj >= nrows) continue;
            double *dst = band + (w + gi - gj) + (i64)(2 * w + 1) * gj;
            const bool shared_corner = (mm == mp) && ((mm == 0 && e > 0) || (mm == o - 1 && e < nel - 1));
            if (shared_corner) atomicAdd(dst, b[t]);
            else *dst = b[t];
        }
    }
}

// Export into BlockBandedMatrices' BlockSkylineMatrix storage (Matrix(undef, B) + set_elements!,
// src/fedvr_operators.jl:17-30,87-134,148-154).  The skyline keeps, per block column J, the block rows kfirst[J]..klast[J]
// stacked into one dense column-major panel of cstride[J] rows that begins at data[cstart[J]]; blk_cum are the
// cumulative block sizes (block b = functions blk_cum[b]..blk_cum[b+1]-1 of the restriction).  data must be zero; one
// CTA per element scatters its block, the two corners that meet on a bridge function are added atomically (two
// addends: the sum does not depend on their order).  Entries outside the skyline can only be structural zeros of
// the block pattern; a non-zero one raises the error flag.
__global__ void fedvr_to_skyline_kernel(const double *__restrict__ blocks, const i64 *__restrict__ estart,
                                        const int32_t *__restrict__ order, const i64 *__restrict__ boff,
                                        i64 nel, i64 first0, i64 nrows,
                                        const i64 *__restrict__ blk_cum, i64 nblk, const i64 *__restrict__ kfirst,
                                        const i64 *__restrict__ klast, const i64 *__restrict__ cstart,
                                        const i64 *__restrict__ cstride, double *__restrict__ data, i64 ndata,
                                        int *__restrict__ err) {
    auto block_of = [&](i64 i) {                    // last b with blk_cum[b] <= i
        i64 lo = 0, hi = nblk;
        while (lo < hi) { const i64 mid = (lo + hi) >> 1; if (blk_cum[mid] <= i) lo = mid + 1; else hi = mid; }
        return lo - 1;
    };
    for (i64 e = blockIdx.x; e < nel; e += gridDim.x) {
        const int o = order[e];
        const double *b = blocks + boff[e];
        for (int t = threadIdx.x; t < o * o; t += blockDim.x) {
            const int mm = t % o, mp = t / o;
            const i64 gi = estart[e] + mm - first0, gj = estart[e] + mp - first0;
            if (gi < 0 || gi >= nrows || gj < 0 || gj >= nrows) continue;
            const i64 K = block_of(gi), J = block_of(gj);
            if (K < kfirst[J] || K > klast[J]) { if (b[t] != 0.0) atomicOr(err, 1); continue; }
            const i64 idx = cstart[J] + (gi - blk_cum[kfirst[J]]) + (gj - blk_cum[J]) * cstride[J];
            if (idx < 0 || idx >= ndata) { atomicOr(err, 2); continue; }
            const bool shared_corner = (mm == mp) && ((mm == 0 && e > 0) || (mm == o - 1 && e < nel - 1));
            if (shared_corner) atomicAdd(&data[idx], b[t]);
            else data[idx] = b[t];
        }
    }
}

// All orders 2: Matrix(undef, B) is a Tridiagonal (src/fedvr_operators.jl:18-23), filled by set_element!(::Tridiagonal)
// (:136-146).  d must be zero.
__global__ void fedvr_to_tridiag_kernel(const double *__restrict__ blocks, const i64 *__restrict__ estart,
                                        const i64 *__restrict__ boff, i64 nel, i64 first0, i64 nrows,
                                        double *__restrict__ dl, double *__restrict__ d, double *__restrict__ du) {
    for (i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x; e < nel; e += (i64)gridDim.x * blockDim.x) {
        const double *b = blocks + boff[e];         // 2 x 2, column-major
        const i64 g0 = estart[e] - first0, g1 = g0 + 1;
        if (g0 >= 0 && g0 < nrows) atomicAdd(&d[g0], b[0]);
        if (g1 >= 0 && g1 < nrows) atomicAdd(&d[g1], b[3]);
        if (g0 >= 0 && g1 < nrows) { dl[g0] = b[1]; du[g0] = b[2]; }
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
static int check_apply_args(const char *who, i64 m, i64 n, const double *X, i64 ldx, i64 nvec, double *Y, i64 ldy) {
    CB_REQUIRE(m >= 0 && n >= 0 && nvec >= 0, "%s: negative size", who);
    CB_REQUIRE(X != nullptr && Y != nullptr, "%s: NULL array", who);
    // DimensionMismatch in the reference's mul!
    CB_REQUIRE(ldx >= n && ldy >= m, "%s: leading dimension smaller than the number of rows (ldx=%lld n=%lld ldy=%lld m=%lld)",
               who, (long long)ldx, (long long)n, (long long)ldy, (long long)m);
    return CB_OK;
}

// set by cb_fedvr_assemble_apply around its apply launch: the pipeline kernel may start while the assembly kernel drains
static thread_local bool g_fedvr_after_assembly = false;

template <int O>
static int launch_fedvr(const double *blocks, i64 nel, i64 first0, i64 nrows, const double *X, i64 ldx, i64 nvec,
                        double alpha, double beta, double *Y, i64 ldy, cudaStream_t st) {
    constexpr int NG = (O <= 12) ? FEDVR_NG : 8;
    constexpr int ET = NG - 1;
    constexpr int ROWS = (ET + 1) * (O - 1) + 1;
    constexpr int PITCH = fe_pitch(ROWS);
    constexpr size_t smem = sizeof(double) * ((size_t)((FE_TV * PITCH + 1) / 2 * 2) + (size_t)(2 * ET + 1) * FE_BRP);
    static_assert(smem <= 227 * 1024, "tile ring does not fit in shared memory");
#if FEDVR_PIPE
    // the TMA pipeline needs X and Y columns of equal 16-byte phase (see the kernel comment)
    if (beta == 0.0 && ((reinterpret_cast<uintptr_t>(X) ^ reinterpret_cast<uintptr_t>(Y)) & 15) == 0 &&
        ((ldx ^ ldy) & 1) == 0) {
        using P = FePipe<O, NG>;
        static_assert(P::NSTG >= 2, "tile ring does not fit in shared memory");
        auto pk = fedvr_apply_pipe_kernel<O, NG>;
        CB_CUDA(cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)P::SMEM));
        const i64 ntiles = ((nel + P::ET - 1) / P::ET) * ((nvec + FE_TV - 1) / FE_TV);
        const int pgrid = (int)(ntiles < CB_NUM_SMS ? ntiles : CB_NUM_SMS);
        if (g_fedvr_after_assembly) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)pgrid); cfg.blockDim = dim3((unsigned)P::THREADS);
            cfg.dynamicSmemBytes = P::SMEM; cfg.stream = st;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            CB_CUDA(cudaLaunchKernelEx(&cfg, pk, blocks, nel, first0, nrows, X, ldx, nvec, alpha, beta, Y, ldy));
        } else {
            pk<<<pgrid, P::THREADS, P::SMEM, st>>>(blocks, nel, first0, nrows, X, ldx, nvec, alpha, beta, Y, ldy);
        }
        CB_LAUNCH_CHECK();
        return CB_OK;
    }
#endif
    auto kern = fedvr_apply_lean_kernel<O, NG>;
    CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const i64 tiles = ((nel + ET - 1) / ET) * ((nvec + FE_TV - 1) / FE_TV);
    // CTAs the machine holds at once (registers limit this kernel to 3 per SM at NG = 16, not shared memory); the grid is
    // a whole number of such waves so that no SM idles while the last CTAs finish
    static int per_sm = 0;
    if (per_sm == 0) {
        int occ = 0;
        CB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 32 * NG, smem));
        per_sm = occ < 1 ? 1 : occ;
    }
    const i64 cap = (i64)CB_NUM_SMS * per_sm * FEDVR_WAVES;
    int grid = (int)(tiles < cap ? tiles : cap);
    kern<<<grid, 32 * NG, smem, st>>>(blocks, nel, first0, nrows, X, ldx, nvec, alpha, beta, Y, ldy);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

#ifdef FEDVR_PHASE_PROF
extern "C" int cb_debug_fedvr_prof(unsigned long long *out, int reset) {
    cudaMemcpyFromSymbol(out, fe_prof, sizeof(unsigned long long) * 8);
    if (reset) { unsigned long long z[8] = {0}; cudaMemcpyToSymbol(fe_prof, z, sizeof(z)); }
    return 0;
}
#endif

extern "C" {

int cb_tridiag_apply(const double *dl, const double *d, const double *du, i64 n,
                     const double *X, i64 ldx, i64 nvec, double alpha, double beta,
                     double *Y, i64 ldy, void *stream) {
    int rc = check_apply_args("cb_tridiag_apply", n, n, X, ldx, nvec, Y, ldy);
    if (rc) return rc;
    if (n == 0 || nvec == 0) return CB_OK;
    CB_REQUIRE(d != nullptr && (n == 1 || (dl != nullptr && du != nullptr)), "cb_tridiag_apply: NULL diagonal");
    cudaStream_t st = (cudaStream_t)stream;
#if TP_PIPE
    // warp-specialised TMA pipeline: beta == 0 (bulk stores do not read Y), X and Y columns of equal 16-byte phase
    if (beta == 0.0 && n >= 2 * TP_ROWS && ((reinterpret_cast<uintptr_t>(X) ^ reinterpret_cast<uintptr_t>(Y)) & 15) == 0 &&
        ((ldx ^ ldy) & 1) == 0) {
        const size_t psmem = sizeof(double) * (size_t)TP_NSTG * TP_STAGE + 256;
        CB_CUDA(cudaFuncSetAttribute(tridiag_apply_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
        const i64 nrt = (n + TP_ROWS - 1) / TP_ROWS, nvt = (nvec + TP_VEC - 1) / TP_VEC;
        i64 vchunk = nrt * nvt / ((i64)CB_NUM_SMS * 8);
        vchunk = vchunk < 1 ? 1 : (vchunk > 8 ? 8 : vchunk);
        const i64 pitems = nrt * ((nvt + vchunk - 1) / vchunk);
        const int pgrid = (int)(pitems < CB_NUM_SMS ? pitems : CB_NUM_SMS);
        tridiag_apply_pipe_kernel<<<pgrid, 32 * (TP_CW + 2), psmem, st>>>(dl ? dl : d, d, du ? du : d, n, X, ldx, nvec, (int)vchunk, alpha, Y, ldy);
        CB_LAUNCH_CHECK();
        return CB_OK;
    }
#endif
    bool vec2 = (ldx % 2 == 0) && (ldy % 2 == 0) && ((uintptr_t)X % 16 == 0) && ((uintptr_t)Y % 16 == 0);
    i64 rows_per_blk = (i64)TRI_THREADS * (vec2 ? 2 : 1);
    i64 nrblk = (n + rows_per_blk - 1) / rows_per_blk;
    // vector groups: enough CTAs to fill the machine, coefficients reused across the group
    i64 vgroup = 32;
    while (vgroup > 4 && nrblk * ((nvec + vgroup - 1) / vgroup) < (i64)CB_NUM_SMS * 8) vgroup /= 2;
    i64 gy = (nvec + vgroup - 1) / vgroup;
    if (gy > 65535) { vgroup = (nvec + 65534) / 65535; vgroup = (vgroup + TRI_VUNROLL - 1) / TRI_VUNROLL * TRI_VUNROLL; gy = (nvec + vgroup - 1) / vgroup; }
    i64 gx = nrblk;
    i64 cap = ((i64)CB_NUM_SMS * 16 + gy - 1) / gy; if (cap < 1) cap = 1;
    if (gx > cap) gx = cap;
    dim3 grid((unsigned)gx, (unsigned)gy);
    if (vec2)
        tridiag_apply_kernel<true><<<grid, TRI_THREADS, 0, st>>>(dl, d, du, n, X, ldx, nvec, vgroup, alpha, beta, Y, ldy);
    else
        tridiag_apply_kernel<false><<<grid, TRI_THREADS, 0, st>>>(dl, d, du, n, X, ldx, nvec, vgroup, alpha, beta, Y, ldy);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int cb_diag_apply(const double *d, i64 n, const double *X, i64 ldx, i64 nvec,
                  double alpha, double beta, double *Y, i64 ldy, void *stream) {
    int rc = check_apply_args("cb_diag_apply", n, n, X, ldx, nvec, Y, ldy);
    if (rc) return rc;
    if (n == 0 || nvec == 0) return CB_OK;
    CB_REQUIRE(d != nullptr, "cb_diag_apply: NULL diagonal");
    dim3 grid((unsigned)cb_grid_for(n, 256, 4), (unsigned)(nvec < 4096 ? nvec : 4096));
    diag_apply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d, n, X, ldx, nvec, alpha, beta, Y, ldy);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int cb_axpby(i64 n, i64 nvec, double alpha, const double *X, i64 ldx, double beta, double *Y, i64 ldy, void *stream) {
    int rc = check_apply_args("cb_axpby", n, n, X, ldx, nvec, Y, ldy);
    if (rc) return rc;
    if (n == 0 || nvec == 0) return CB_OK;
    dim3 grid((unsigned)cb_grid_for(n, 256, 4), (unsigned)(nvec < 4096 ? nvec : 4096));
    axpby_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(n, nvec, alpha, X, ldx, beta, Y, ldy);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int cb_batched_dot(const double *U, i64 ldu, const double *W, i64 ldw, i64 n, i64 P, double scale, double *out, void *stream) {
    CB_REQUIRE(n >= 0 && P >= 0, "cb_batched_dot: negative size");
    if (P == 0) return CB_OK;
    CB_REQUIRE(U != nullptr && W != nullptr && out != nullptr, "cb_batched_dot: NULL array");
    CB_REQUIRE(ldu >= n && ldw >= n, "DimensionMismatch: cb_batched_dot: leading dimension smaller than the number of rows");
    const i64 grid = P < (i64)CB_NUM_SMS * 16 ? P : (i64)CB_NUM_SMS * 16;
    batched_dot_kernel<<<(unsigned)grid, DOT_THREADS, 0, (cudaStream_t)stream>>>(U, ldu, W, ldw, n, P, scale, out);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int cb_banded_apply(const double *band, i64 m, i64 n, int l, int u,
                    const double *X, i64 ldx, i64 nvec, double alpha, double beta,
                    double *Y, i64 ldy, void *stream) {
    int rc = check_apply_args("cb_banded_apply", m, n, X, ldx, nvec, Y, ldy);
    if (rc) return rc;
    CB_REQUIRE(band != nullptr, "cb_banded_apply: NULL band");
    CB_REQUIRE(l + u + 1 >= 1, "cb_banded_apply: invalid bandwidths (%d,%d)", l, u);
    if (m == 0 || nvec == 0) return CB_OK;
    if (-l > n - 1 || -u > m - 1) {
        // every stored band lies outside the m x n matrix (BandedMatrices allows it: operators between restrictions
        // that do not overlap, test/fd/derivatives.jl:36-59): A = 0, Y <- beta Y
        if (beta == 0.0) CB_CUDA(cudaMemset2DAsync(Y, sizeof(double) * ldy, 0, sizeof(double) * m, nvec, (cudaStream_t)stream));
        else if (beta != 1.0) return cb_axpby(m, nvec, 0.0, Y, ldy, beta, Y, ldy, stream);
        return CB_OK;
    }
    // BandedMatrices allows negative l or u (bands entirely above/below the diagonal, e.g.
    // src/fd_operators.jl:16); the window arithmetic below is valid for them as long as l+u >= 0.
    const int W = l + u + 1;
    const int KS = (8 + W - 1 + 3) / 4;
    const int xrows = BT_ROWS + l + u;
    const int pitch = bt_pitch(xrows + 4);
    size_t smem = sizeof(double) * ((size_t)BT_VEC * pitch + (size_t)(BT_ROWS / 8) * KS * 32);
    if (smem > 200 * 1024) { cb_set_error("cb_banded_apply: bandwidth %d too large for the shared-memory tile", W); return CB_ERR_UNSUPPORTED; }
    CB_CUDA(cudaFuncSetAttribute(banded_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const i64 nrt = (m + BT_ROWS - 1) / BT_ROWS, nvt = (nvec + BT_VEC - 1) / BT_VEC;
#if BT_PIPE
    // warp-specialised TMA pipeline: needs non-negative bandwidths and X, Y columns of equal 16-byte phase
    if (l >= 0 && u >= 0 && KS <= BT_KSMAX && beta == 0.0 &&
        ((reinterpret_cast<uintptr_t>(X) ^ reinterpret_cast<uintptr_t>(Y)) & 15) == 0 && ((ldx ^ ldy) & 1) == 0) {
        const int P0 = (BTP_ROWS + l + u + 3 + 2 + 12 + 15) / 16 * 16;      // window + k padding + parity shift / rounding + skew
        const size_t stage_b = sizeof(double) * BTP_VEC * P0, as_b = sizeof(double) * BTP_GROUPS * (BTP_ROWS / 8) * KS * 32;
        int nstg = (int)((227 * 1024 - as_b - 256) / stage_b);
        if (nstg > 6) nstg = 6;
        if (nstg >= 2) {
            const size_t psmem = nstg * stage_b + as_b + 256;
            CB_CUDA(cudaFuncSetAttribute(banded_apply_pipe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)psmem));
            // vector tiles per work item: the one of 8..4 (1 if there are few) that deals the items most evenly
            // (measured, C3 / k = 7: 1 tile per item 301 / 350 us, 2: 227 / 243, 4: 194 / 197, 8: 180 / 191, 16: - / 220,
            // 32: - / 249 — few tiles per item expose the staging of the A fragments, many unbalance the CTAs)
            const i64 nrt = (m + BTP_ROWS - 1) / BTP_ROWS, nvt = (nvec + BTP_VEC - 1) / BTP_VEC;
            i64 best = 1;
            if (nrt * nvt >= (i64)CB_NUM_SMS * 8) {
                double best_eff = 0.0;
                for (i64 vcn = 8; vcn >= 4; --vcn) {
                    const i64 it = nrt * ((nvt + vcn - 1) / vcn);
                    const i64 per = (it + CB_NUM_SMS - 1) / CB_NUM_SMS;
                    // tiles the busiest CTA runs against the mean (the last chunk of a row tile may be short)
                    const double eff = (double)(nrt * nvt) / ((double)per * vcn * CB_NUM_SMS);
                    if (eff > best_eff + 1e-9) { best_eff = eff; best = vcn; }
                }
            }
#ifdef BTP_VCHUNK
            best = BTP_VCHUNK;
#endif
            const i64 pitems = nrt * ((nvt + best - 1) / best);
            const int pgrid = (int)(pitems < CB_NUM_SMS ? pitems : CB_NUM_SMS);
            banded_apply_pipe_kernel<<<pgrid, 32 * (BTP_NG + 2), psmem, (cudaStream_t)stream>>>(
                band, m, n, l, u, X, ldx, nvec, (int)best, P0, nstg, alpha, beta, Y, ldy);
            CB_LAUNCH_CHECK();
            return CB_OK;
        }
    }
#endif
    // vector tiles per work item: amortise the band-tile staging while keeping >= ~8 items per SM
    i64 vchunk = nrt * nvt / ((i64)CB_NUM_SMS * 8);
    vchunk = vchunk < 1 ? 1 : (vchunk > 8 ? 8 : vchunk);          // 3..8 measured within 2% of each other, 1-2 slower
    // big items first, then single-tile items for the last ~BT_TAIL_PCT % of the vector tiles (tail balancing: one CTA
    // per item, handed out in order by the hardware)
    i64 nvc_big = nvt / vchunk;
    if (vchunk > 1) {
        const i64 tail = (nvt * BT_TAIL_PCT + 99) / 100;
        const i64 big = nvt - tail;                                // vector tiles covered by the big items
        if (big < vchunk) vchunk = big < 1 ? 1 : big;
        nvc_big = big < 1 ? 0 : big / vchunk;
    }
    const i64 items = nrt * (nvc_big + (nvt - nvc_big * vchunk));
    int grid = (int)(items < (i64)CB_NUM_SMS * BT_GRIDCAP ? items : (i64)CB_NUM_SMS * BT_GRIDCAP);
    banded_apply_kernel<<<grid, BT_THREADS, smem, (cudaStream_t)stream>>>(band, m, n, l, u, X, ldx, nvec, (int)vchunk, nvc_big, alpha, beta, Y, ldy);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int cb_fedvr_apply(const double *blocks, const i64 *estart, const int32_t *order, const i64 *boff,
                   i64 nel, int uniform_order, i64 first, i64 last,
                   const double *X, i64 ldx, i64 nvec, double alpha, double beta,
                   double *Y, i64 ldy, void *stream) {
    CB_REQUIRE(blocks != nullptr && nel >= 1, "cb_fedvr_apply: NULL blocks / no elements");
    CB_REQUIRE(first >= 1 && last >= first, "cb_fedvr_apply: invalid restriction %lld..%lld", (long long)first, (long long)last);
    const i64 nrows = last - first + 1;
    int rc = check_apply_args("cb_fedvr_apply", nrows, nrows, X, ldx, nvec, Y, ldy);
    if (rc) return rc;
    if (nvec == 0) return CB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const i64 first0 = first - 1;
    if (uniform_order > 0) {
        CB_REQUIRE(uniform_order >= 2, "cb_fedvr_apply: element order must be >= 2");
        CB_REQUIRE(last <= nel * (uniform_order - 1) + 1, "cb_fedvr_apply: restriction exceeds the basis");
        switch (uniform_order) {
#define CB_FE(OO) case OO: return launch_fedvr<OO>(blocks, nel, first0, nrows, X, ldx, nvec, alpha, beta, Y, ldy, st);
            CB_FE(2) CB_FE(3) CB_FE(4) CB_FE(5) CB_FE(6) CB_FE(7) CB_FE(8) CB_FE(9) CB_FE(10)
            CB_FE(11) CB_FE(12) CB_FE(13) CB_FE(14) CB_FE(15) CB_FE(16)
#undef CB_FE
            default: break;                         // fall through to the generic kernel
        }
    }
    CB_REQUIRE(estart != nullptr && order != nullptr && boff != nullptr,
               "cb_fedvr_apply: per-element arrays required for non-uniform / large orders");
    dim3 grid((unsigned)cb_grid_for(nrows, 128, 8), (unsigned)(nvec < 2048 ? nvec : 2048));
    fedvr_apply_generic_kernel<<<grid, 128, 0, st>>>(blocks, estart, order, boff, nel, first0, nrows, X, ldx, nvec, alpha, beta, Y, ldy);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

// derop! followed by mul! in one call: the two launches are enqueued back to back from C (the host side of a step is
// then one foreign call instead of the operator algebra of the caller's language).
int cb_fedvr_assemble_apply(const double *x, const double *wcat, const double *nrm,
                            const i64 *estart, const int32_t *order, const i64 *boff, const i64 *woff,
                            i64 nel, int uniform_order, int max_order, int nder, double *blocks,
                            i64 first, i64 last, const double *X, i64 ldx, i64 nvec,
                            double alpha, double beta, double *Y, i64 ldy, void *stream) {
    int rc = cb_fedvr_assemble(x, wcat, nrm, estart, order, boff, woff, nel, uniform_order, max_order, nder, blocks, stream);
    if (rc) return rc;
    static const bool serial = getenv("CB_FEDVR_NO_PDL") != nullptr;
    g_fedvr_after_assembly = !serial;
    rc = cb_fedvr_apply(blocks, estart, order, boff, nel, uniform_order, first, last, X, ldx, nvec, alpha, beta, Y, ldy, stream);
    g_fedvr_after_assembly = false;
    return rc;
}

int cb_fedvr_to_dense(const double *blocks, const i64 *estart, const int32_t *order, const i64 *boff,
                      i64 nel, i64 first, i64 last, double *A, void *stream) {
    CB_REQUIRE(blocks && estart && order && boff && A, "cb_fedvr_to_dense: NULL argument");
    CB_REQUIRE(first >= 1 && last >= first, "cb_fedvr_to_dense: invalid restriction");
    const i64 nrows = last - first + 1;
    cudaStream_t st = (cudaStream_t)stream;
    CB_CUDA(cudaMemsetAsync(A, 0, sizeof(double) * nrows * nrows, st));
    fedvr_to_dense_kernel<<<(unsigned)(nel < 4096 ? nel : 4096), 128, 0, st>>>(blocks, estart, order, boff, nel, first - 1, nrows, A);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

// FE-DVR basis functions at points (src/fedvr.jl:211-222, 272-295): the functions of the element that contains x,
//   chi_{i,m}(x) = n_m prod_{j != m} (x - x_j) / (x_m - x_j),
// evaluated in the reference's order (one quotient per factor, round-to-nearest: bit-identical to the scalar loop).  A point
// on an interior element boundary belongs to the element to its right — the one whose values the reference's column
// loop writes last; both give n for the bridge function and exact zeros for the others.  One thread per point.
__global__ void fedvr_eval_kernel(const double *__restrict__ xn, const double *__restrict__ nrm, const double *__restrict__ t,
                                  const i64 *__restrict__ estart, const int32_t *__restrict__ order, i64 nel, int omax,
                                  const double *__restrict__ xq, i64 nx, int32_t *__restrict__ elem, double *__restrict__ vals) {
    for (i64 p = blockIdx.x * (i64)blockDim.x + threadIdx.x; p < nx; p += (i64)gridDim.x * blockDim.x) {
        const double x = xq[p];
        double *v = vals + p * omax;
        if (!(x >= t[0] && x <= t[nel])) {              // outside the basis (or NaN): no entries
            elem[p] = 0;
            for (int m = 0; m < omax; ++m) v[m] = 0.0;
            continue;
        }
        i64 lo = 0, hi = nel;                           // last element e with t[e] <= x (the last one at x = t[nel])
        while (hi - lo > 1) {
            const i64 mid = (lo + hi) >> 1;
            if (t[mid] <= x) lo = mid; else hi = mid;
        }
        const i64 e = lo;
        const int o = order[e];
        const double *xe = xn + estart[e], *ne = nrm + estart[e];
        elem[p] = (int32_t)(e + 1);
        for (int m = 0; m < omax; ++m) {
            double chi = 0.0;
            if (m < o) {
                chi = ne[m];
                const double xm = xe[m];
                for (int j = 0; j < o; ++j) {
                    if (j == m) continue;
                    chi = __dmul_rn(chi, __ddiv_rn(__dsub_rn(x, xe[j]), __dsub_rn(xm, xe[j])));
                }
            }
            v[m] = chi;
        }
    }
}

int cb_fedvr_eval(const double *xn, const double *nrm, const double *t, const i64 *estart, const int32_t *order,
                  i64 nel, int max_order, const double *xq, i64 nx, int32_t *elem, double *vals, void *stream) {
    CB_REQUIRE(nx >= 0 && nel >= 1, "cb_fedvr_eval: bad sizes");
    if (nx == 0) return CB_OK;
    CB_REQUIRE(xn && nrm && t && estart && order && xq && elem && vals, "cb_fedvr_eval: NULL argument");
    CB_REQUIRE(max_order >= 2, "cb_fedvr_eval: element order must be >= 2");
    if (max_order > CB_FEDVR_OMAX) { cb_set_error("cb_fedvr_eval: element order %d > %d", max_order, CB_FEDVR_OMAX); return CB_ERR_UNSUPPORTED; }
    fedvr_eval_kernel<<<cb_grid_for(nx, 128, 16), 128, 0, (cudaStream_t)stream>>>(xn, nrm, t, estart, order, nel, max_order, xq, nx, elem, vals);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

// Finite-difference basis functions at points (getindex, src/finite_differences.jl:63-93): function j is the tent over
// (l_{j-1}, l_j, l_{j+1}) times weight(B, j), tent(x, a, b, c) = 1 - clamp((b-x)/(b-a), 0, 1) - clamp((b-x)/(b-c), 0, 1),
// with the mirrored neighbours 2 l_1 - l_2 and 2 l_n - l_{n-1} beyond the ends.  At most two functions are non-zero at
// a point: first[p] = j (1-based; 0: none) and vals[2 p], vals[2 p + 1] = functions j and j + 1.  One thread per point.
__device__ __forceinline__ double fd_clamp01(double v) { return v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v); }
__device__ __forceinline__ double fd_tent(double x, double a, double b, double c) {
    return __dsub_rn(__dsub_rn(1.0, fd_clamp01(__ddiv_rn(__dsub_rn(b, x), __dsub_rn(b, a)))),
                     fd_clamp01(__ddiv_rn(__dsub_rn(b, x), __dsub_rn(b, c))));
}
__global__ void fd_eval_kernel(const double *__restrict__ l, const double *__restrict__ w, i64 n,
                               const double *__restrict__ xq, i64 nx, int32_t *__restrict__ first, double *__restrict__ vals) {
    for (i64 p = blockIdx.x * (i64)blockDim.x + threadIdx.x; p < nx; p += (i64)gridDim.x * blockDim.x) {
        const double x = xq[p];
        const double lo_ext = __dsub_rn(__dmul_rn(2.0, l[0]), l[1]), hi_ext = __dsub_rn(__dmul_rn(2.0, l[n - 1]), l[n - 2]);
        first[p] = 0; vals[2 * p] = 0.0; vals[2 * p + 1] = 0.0;
        if (!(x >= lo_ext && x <= hi_ext)) continue;
        // j0: last node with l[j0] <= x (0-based; -1 left of the first node)
        i64 lo = -1, hi = n;
        while (hi - lo > 1) {
            const i64 mid = (lo + hi) >> 1;
            if (l[mid] <= x) lo = mid; else hi = mid;
        }
        auto node = [&](i64 j) { return j < 0 ? lo_ext : (j >= n ? hi_ext : l[j]); };
        const i64 ja = lo < 0 ? 0 : lo;                 // first function that can be non-zero at x
        first[p] = (int32_t)(ja + 1);
        for (int u = 0; u < 2; ++u) {
            const i64 j = ja + u;
            if (j >= n || (lo < 0 && u == 1)) break;
            const double v = __dmul_rn(w ? w[j] : 1.0, fd_tent(x, node(j - 1), node(j), node(j + 1)));
            vals[2 * p + u] = v;
        }
    }
}

int cb_fd_eval(const double *nodes, const double *weights, i64 n, const double *xq, i64 nx, int32_t *first, double *vals,
               void *stream) {
    CB_REQUIRE(nx >= 0 && n >= 2, "cb_fd_eval: a grid needs at least two nodes");
    if (nx == 0) return CB_OK;
    CB_REQUIRE(nodes && xq && first && vals, "cb_fd_eval: NULL argument");
    fd_eval_kernel<<<cb_grid_for(nx, 256, 16), 256, 0, (cudaStream_t)stream>>>(nodes, weights, n, xq, nx, first, vals);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int cb_fedvr_to_band(const double *blocks, const i64 *estart, const int32_t *order, const i64 *boff,
                     i64 nel, i64 first, i64 last, int w, double *band, void *stream) {
    CB_REQUIRE(blocks && estart && order && boff && band, "cb_fedvr_to_band: NULL argument");
    CB_REQUIRE(first >= 1 && last >= first, "cb_fedvr_to_band: invalid restriction");
    CB_REQUIRE(w >= 1 && w < CB_FEDVR_OMAX, "cb_fedvr_to_band: half bandwidth must be >= max order - 1");
    const i64 nrows = last - first + 1;
    cudaStream_t st = (cudaStream_t)stream;
    CB_CUDA(cudaMemsetAsync(band, 0, sizeof(double) * (size_t)nrows * (2 * w + 1), st));
    fedvr_to_band_kernel<<<(unsigned)(nel < 8192 ? nel : 8192), 128, 0, st>>>(blocks, estart, order, boff, nel, first - 1, nrows, w, band);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int cb_fedvr_to_skyline(const double *blocks, const i64 *estart, const int32_t *order, const i64 *boff,
                        i64 nel, i64 first, i64 last, const i64 *blk_cum, i64 nblk, const i64 *col_kfirst,
                        const i64 *col_klast, const i64 *col_start, const i64 *col_stride, double *data, i64 ndata,
                        void *stream) {
    CB_REQUIRE(blocks && estart && order && boff && data, "cb_fedvr_to_skyline: NULL argument");
    CB_REQUIRE(blk_cum && col_kfirst && col_klast && col_start && col_stride && nblk >= 1, "cb_fedvr_to_skyline: NULL block structure");
    CB_REQUIRE(first >= 1 && last >= first && ndata >= 0, "cb_fedvr_to_skyline: invalid restriction");
    const i64 nrows = last - first + 1;
    cudaStream_t st = (cudaStream_t)stream;
    { int prc = cb_keep_scratch_pool(); if (prc) return prc; }
    int *d_err = nullptr;
    CB_CUDA(cudaMallocAsync(&d_err, sizeof(int), st));
    CB_CUDA(cudaMemsetAsync(d_err, 0, sizeof(int), st));
    CB_CUDA(cudaMemsetAsync(data, 0, sizeof(double) * (size_t)ndata, st));      // A .= false, src/fedvr_operators.jl:149
    fedvr_to_skyline_kernel<<<(unsigned)(nel < 8192 ? nel : 8192), 128, 0, st>>>(blocks, estart, order, boff, nel, first - 1, nrows,
                                                                                 blk_cum, nblk, col_kfirst, col_klast, col_start,
                                                                                 col_stride, data, ndata, d_err);
    cudaError_t e = cudaGetLastError();
    int herr = 0;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&herr, d_err, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFreeAsync(d_err, st);
    if (e != cudaSuccess) { cb_set_error("cb_fedvr_to_skyline: %s", cudaGetErrorString(e)); return CB_ERR_CUDA; }
    CB_REQUIRE(herr == 0, "cb_fedvr_to_skyline: the block structure does not hold the element blocks (%s)",
               (herr & 2) ? "index outside data" : "non-zero entry outside the skyline");
    return CB_OK;
}

int cb_fedvr_to_tridiag(const double *blocks, const i64 *estart, const int32_t *order, const i64 *boff,
                        i64 nel, i64 first, i64 last, double *dl, double *d, double *du, void *stream) {
    (void)order;                                    // all orders are 2 (the caller's Matrix(undef, B) is a Tridiagonal)
    CB_REQUIRE(blocks && estart && boff && d, "cb_fedvr_to_tridiag: NULL argument");
    CB_REQUIRE(first >= 1 && last >= first, "cb_fedvr_to_tridiag: invalid restriction");
    const i64 nrows = last - first + 1;
    CB_REQUIRE(nrows == 1 || (dl && du), "cb_fedvr_to_tridiag: NULL off-diagonal");
    cudaStream_t st = (cudaStream_t)stream;
    CB_CUDA(cudaMemsetAsync(d, 0, sizeof(double) * (size_t)nrows, st));
    fedvr_to_tridiag_kernel<<<cb_grid_for(nel, 256), 256, 0, st>>>(blocks, estart, boff, nel, first - 1, nrows, dl, d, du);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

// ---- host-buffer forms: stream the vector batch through the device ----------
typedef int (*apply_dev_fn)(void *ctx, const double *Xd, i64 ld, i64 nv, double alpha, double beta, double *Yd, cudaStream_t st);

// Device staging buffers and streams of the host-buffer entry points, kept per device between calls
// (allocation and stream creation cost milliseconds, a 737 MB batch over PCIe ~15 ms each way).
// The entry points are serialised per device by `mu` (the staging ring is shared state).
struct HostPipe {
    static constexpr int NBUF = 3;
    std::mutex mu;
    bool init = false;
    size_t bytes = 0;                               // capacity of each buffer
    double *dX[NBUF] = {nullptr, nullptr, nullptr}, *dY[NBUF] = {nullptr, nullptr, nullptr};
    cudaStream_t st[NBUF] = {nullptr, nullptr, nullptr};
    cudaEvent_t producer = nullptr;                 // "the operator arrays are ready" on the caller's stream
};
static HostPipe g_pipes[64];

// `producer`: the stream on which the caller enqueued the work that produces the operator arrays (assembly kernels);
// the staging streams wait for everything enqueued on it so far before their first kernel.
static int apply_host_pipeline(apply_dev_fn fn, void *ctx, i64 nrows_in, i64 nrows_out, const double *Xh, i64 ldx, i64 nvec,
                               double alpha, double beta, double *Yh, i64 ldy, cudaStream_t producer) {
    if (nvec == 0 || nrows_out == 0) return CB_OK;
    int dev = 0;
    CB_CUDA(cudaGetDevice(&dev));
    CB_REQUIRE(dev >= 0 && dev < 64, "host pipeline: device index out of range");
    HostPipe &P = g_pipes[dev];
    std::lock_guard<std::mutex> lock(P.mu);
    // tile of vectors: ~32 MB per buffer (CB_HOST_CHUNK_MB overrides, for tuning); the device copy is dense (ld = rows)
    static i64 chunk_mb = 0;
    if (chunk_mb == 0) {
        const char *env = getenv("CB_HOST_CHUNK_MB");
        chunk_mb = env ? atoll(env) : 32;
        if (chunk_mb < 1 || chunk_mb > 1024) chunk_mb = 32;
    }
    const i64 nrmax = nrows_in > nrows_out ? nrows_in : nrows_out;
    i64 tv = (chunk_mb << 20) / (nrmax * (i64)sizeof(double));
    if (tv < 1) tv = 1;
    if (tv > nvec) tv = nvec;
    const size_t need = sizeof(double) * (size_t)nrmax * (size_t)tv;
    cudaError_t e = cudaSuccess;
    if (!P.init) {
        for (int b = 0; b < HostPipe::NBUF && e == cudaSuccess; ++b) e = cudaStreamCreateWithFlags(&P.st[b], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&P.producer, cudaEventDisableTiming);
        P.init = e == cudaSuccess;
    }
    if (e == cudaSuccess && P.bytes < need) {
        for (int b = 0; b < HostPipe::NBUF; ++b) { cudaFree(P.dX[b]); cudaFree(P.dY[b]); P.dX[b] = P.dY[b] = nullptr; }
        P.bytes = 0;
        for (int b = 0; b < HostPipe::NBUF && e == cudaSuccess; ++b) {
            e = cudaMalloc(&P.dX[b], need);
            if (e == cudaSuccess) e = cudaMalloc(&P.dY[b], need);
        }
        if (e == cudaSuccess) P.bytes = need;
    }
    // order the staging streams after the producer of the operator arrays (they are non-blocking streams: not even the
    // legacy default stream orders them implicitly)
    if (e == cudaSuccess) e = cudaEventRecord(P.producer, producer);
    for (int b = 0; b < HostPipe::NBUF && e == cudaSuccess; ++b) e = cudaStreamWaitEvent(P.st[b], P.producer, 0);
    int rc = CB_OK;
    auto copy2d = [&](double *dst, i64 ldd, const double *src, i64 lds, i64 nr, i64 nv, cudaMemcpyKind kind, cudaStream_t st) {
        if (ldd == nr && lds == nr) return cudaMemcpyAsync(dst, src, sizeof(double) * nr * nv, kind, st);
        return cudaMemcpy2DAsync(dst, sizeof(double) * ldd, src, sizeof(double) * lds, sizeof(double) * nr, nv, kind, st);
    };
    i64 t = 0;
    for (i64 v0 = 0; v0 < nvec && e == cudaSuccess && rc == CB_OK; v0 += tv, ++t) {
        const int b = (int)(t % HostPipe::NBUF);
        const i64 nv = v0 + tv <= nvec ? tv : nvec - v0;
        e = copy2d(P.dX[b], nrows_in, Xh + v0 * ldx, ldx, nrows_in, nv, cudaMemcpyHostToDevice, P.st[b]);
        if (e == cudaSuccess && beta != 0.0) e = copy2d(P.dY[b], nrows_out, Yh + v0 * ldy, ldy, nrows_out, nv, cudaMemcpyHostToDevice, P.st[b]);
        if (e == cudaSuccess) rc = fn(ctx, P.dX[b], nrows_in, nv, alpha, beta, P.dY[b], P.st[b]);
        if (e == cudaSuccess && rc == CB_OK) e = copy2d(Yh + v0 * ldy, ldy, P.dY[b], nrows_out, nrows_out, nv, cudaMemcpyDeviceToHost, P.st[b]);
    }
    for (int b = 0; b < HostPipe::NBUF; ++b) {
        if (P.st[b]) { cudaError_t e2 = cudaStreamSynchronize(P.st[b]); if (e == cudaSuccess) e = e2; }
    }
    if (e != cudaSuccess) { cb_set_error("host pipeline: %s", cudaGetErrorString(e)); return CB_ERR_CUDA; }
    return rc;
}

struct fedvr_ctx { const double *blocks; const i64 *estart; const int32_t *order; const i64 *boff; i64 nel; int uo; i64 first, last; };
static int fedvr_dev(void *c, const double *Xd, i64 ld, i64 nv, double alpha, double beta, double *Yd, cudaStream_t st) {
    fedvr_ctx *f = (fedvr_ctx *)c;
    return cb_fedvr_apply(f->blocks, f->estart, f->order, f->boff, f->nel, f->uo, f->first, f->last, Xd, ld, nv, alpha, beta, Yd, ld, st);
}

int cb_fedvr_apply_host(const double *blocks, const i64 *estart, const int32_t *order, const i64 *boff,
                        i64 nel, int uniform_order, i64 first, i64 last,
                        const double *Xh, i64 ldx, i64 nvec, double alpha, double beta, double *Yh, i64 ldy, void *stream) {
    CB_REQUIRE(first >= 1 && last >= first, "cb_fedvr_apply_host: invalid restriction");
    CB_REQUIRE(Xh && Yh && ldx >= last - first + 1 && ldy >= last - first + 1, "cb_fedvr_apply_host: bad host arrays");
    fedvr_ctx c{blocks, estart, order, boff, nel, uniform_order, first, last};
    return apply_host_pipeline(fedvr_dev, &c, last - first + 1, last - first + 1, Xh, ldx, nvec, alpha, beta, Yh, ldy, (cudaStream_t)stream);
}

struct tri_ctx { const double *dl, *d, *du; i64 n; };
static int tri_dev(void *c, const double *Xd, i64 ld, i64 nv, double alpha, double beta, double *Yd, cudaStream_t st) {
    tri_ctx *t = (tri_ctx *)c;
    return cb_tridiag_apply(t->dl, t->d, t->du, t->n, Xd, ld, nv, alpha, beta, Yd, ld, st);
}

int cb_tridiag_apply_host(const double *dl, const double *d, const double *du, i64 n,
                          const double *Xh, i64 ldx, i64 nvec, double alpha, double beta, double *Yh, i64 ldy, void *stream) {
    CB_REQUIRE(Xh && Yh && ldx >= n && ldy >= n, "cb_tridiag_apply_host: bad host arrays");
    tri_ctx c{dl, d, du, n};
    return apply_host_pipeline(tri_dev, &c, n, n, Xh, ldx, nvec, alpha, beta, Yh, ldy, (cudaStream_t)stream);
}

struct band_ctx { const double *band; i64 m, n; int l, u; };
static int band_dev(void *c, const double *Xd, i64 ld, i64 nv, double alpha, double beta, double *Yd, cudaStream_t st) {
    band_ctx *b = (band_ctx *)c;
    return cb_banded_apply(b->band, b->m, b->n, b->l, b->u, Xd, b->n, nv, alpha, beta, Yd, b->m, st);
}

int cb_banded_apply_host(const double *band, i64 m, i64 n, int l, int u,
                         const double *Xh, i64 ldx, i64 nvec, double alpha, double beta, double *Yh, i64 ldy, void *stream) {
    CB_REQUIRE(band && Xh && Yh && m >= 0 && n >= 0 && ldx >= n && ldy >= m, "cb_banded_apply_host: bad arguments");
    band_ctx c{band, m, n, l, u};
    return apply_host_pipeline(band_dev, &c, n, m, Xh, ldx, nvec, alpha, beta, Yh, ldy, (cudaStream_t)stream);
}

// cudaHostRegister / cudaHostUnregister for caller-owned host arrays (a Julia `Array` is pageable: registering it once
// lets the host-buffer entries copy at PCIe speed; CUDA.jl's `CUDA.pin` does the same).
int cb_host_register(void *ptr, size_t bytes) {
    CB_REQUIRE(ptr != nullptr && bytes > 0, "cb_host_register: bad argument");
    CB_CUDA(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return CB_OK;
}
int cb_host_unregister(void *ptr) {
    CB_REQUIRE(ptr != nullptr, "cb_host_unregister: NULL");
    CB_CUDA(cudaHostUnregister(ptr));
    return CB_OK;
}

}  // extern "C"
