// csc.cu — K1b: B[x, cols] as Julia's SparseMatrixCSC (src/bsplines.jl:217-224) from the row ("ELL") form of K1,
// without a library sort.
//
// The CSC order is (column, row) with rows = point indices ascending, i.e. a STABLE counting sort of the k entries per
// point by column.  Columns are few (nf, ~1e3) and every point touches k consecutive ones, so one counting pass is
// enough:
//
//   count    a block of CSC_P consecutive points builds the histogram of its valid entries (column inside the
//            restriction, value != 0.0 — Julia's setindex! on spzeros drops exact zeros) over the columns in shared
//            memory and writes it, transposed, to cnt[column][block]
//   scan     exclusive prefix sum over cnt in (column, block) order = the position of the first entry of every
//            (column, block) run in rowval / nzval; three small kernels (tile sums, scan of the tile sums, rescan);
//            colptr[c] = 1 + offs[c][0]
//   scatter  the block ranks its entries: inside a warp lane after lane bumps the warp's private per-column counter
//            (lanes are consecutive points, so the rank order is the row order), a per-column scan over the warps
//            turns the counters into offsets; the entries are placed in shared memory in (column, row) order and
//            leave the SM slot by slot, so that the entries of a run — consecutive in rowval / nzval — are written by
//            consecutive threads
//
// Cost at C1 (1e6 points, k = 7, 1006 columns): 60 MB of rows read twice (second time from L2), 112 MB of CSC
// written, 8 MB of counters.
#include "bspline_core.cuh"

constexpr int CSC_P = 512;                     // points per block (16 warps)
constexpr int CSC_WARPS = CSC_P / 32;
constexpr int SCAN_T = 256, SCAN_E = 8, SCAN_TILE = SCAN_T * SCAN_E;

__device__ __forceinline__ int csc_entry_col(int I, int a, int k, i64 nf, i64 col0, i64 ncols, double v) {
    if (I <= 0) return -1;
    const i64 j = (i64)I - k + 1 + a;           // 1-based function
    const i64 c = j - 1 - col0;
    return (j >= 1 && j <= nf && c >= 0 && c < ncols && v != 0.0) ? (int)c : -1;
}

__global__ void __launch_bounds__(CSC_P)
csc_count_kernel(const int32_t *__restrict__ interval, const double *__restrict__ vals, i64 nx, int k, i64 nf,
                 i64 col0, int ncols, i64 nblk, uint32_t *__restrict__ cnt /*[ncols][nblk]*/) {
    extern __shared__ uint32_t hist[];
    for (i64 b = blockIdx.x; b < nblk; b += gridDim.x) {
        for (int c = threadIdx.x; c < ncols; c += CSC_P) hist[c] = 0;
        __syncthreads();
        const i64 p = b * CSC_P + threadIdx.x;
        if (p < nx) {
            const int I = interval[p];
            if (I > 0)
                for (int a = 0; a < k; ++a) {
                    const int c = csc_entry_col(I, a, k, nf, col0, ncols, vals[p * k + a]);
                    if (c >= 0) atomicAdd(&hist[c], 1u);
                }
        }
        __syncthreads();
        for (int c = threadIdx.x; c < ncols; c += CSC_P) cnt[(size_t)c * nblk + b] = hist[c];
        __syncthreads();
    }
}

// ---- exclusive scan of n uint32 counters (in place), 64-bit tile sums ----
__global__ void __launch_bounds__(SCAN_T)
scan_tile_sums_kernel(const uint32_t *__restrict__ v, i64 n, unsigned long long *__restrict__ tile_sum) {
    __shared__ unsigned long long ws[SCAN_T / 32];
    const i64 base = (i64)blockIdx.x * SCAN_TILE;
    unsigned long long s = 0;
    for (int e = 0; e < SCAN_E; ++e) {
        const i64 i = base + (i64)e * SCAN_T + threadIdx.x;
        if (i < n) s += v[i];
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int w = 0; w < SCAN_T / 32; ++w) t += ws[w];
        tile_sum[blockIdx.x] = t;
    }
}

// one block: exclusive scan of the tile sums (any count), total -> *total
__global__ void __launch_bounds__(1024)
scan_tiles_kernel(unsigned long long *__restrict__ tile_sum, i64 ntiles, unsigned long long *__restrict__ total) {
    __shared__ unsigned long long ws[32];
    __shared__ unsigned long long carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (i64 base = 0; base < ntiles; base += 1024) {
        const i64 i = base + threadIdx.x;
        const unsigned long long x = i < ntiles ? tile_sum[i] : 0;
        unsigned long long inc = x;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xffffffffu, inc, o);
            if ((threadIdx.x & 31) >= o) inc += y;
        }
        if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) {
            unsigned long long w = ws[threadIdx.x], winc = w;
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long y = __shfl_up_sync(0xffffffffu, winc, o);
                if (threadIdx.x >= o) winc += y;
            }
            ws[threadIdx.x] = winc - w;             // exclusive over warps
        }
        __syncthreads();
        const unsigned long long c0 = carry;
        if (i < ntiles) tile_sum[i] = c0 + ws[threadIdx.x >> 5] + inc - x;
        __syncthreads();
        if (threadIdx.x == 1023) carry = c0 + ws[31] + inc;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

// rescan of every tile with its prefix: v[i] <- exclusive prefix (as uint32: nnz < 2^31); colptr from the first
// block's entry of every column
__global__ void __launch_bounds__(SCAN_T)
scan_apply_kernel(uint32_t *__restrict__ v, i64 n, const unsigned long long *__restrict__ tile_pre, i64 nblk, i64 ncols,
                  const unsigned long long *__restrict__ total, i64 *__restrict__ colptr) {
    __shared__ unsigned long long ws[SCAN_T / 32];
    const i64 base = (i64)blockIdx.x * SCAN_TILE + (i64)threadIdx.x * SCAN_E;     // blocked: 8 consecutive per thread
    uint32_t x[SCAN_E];
    unsigned long long s = 0;
#pragma unroll
    for (int e = 0; e < SCAN_E; ++e) { x[e] = base + e < n ? v[base + e] : 0u; s += x[e]; }
    unsigned long long inc = s;
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xffffffffu, inc, o);
        if ((threadIdx.x & 31) >= o) inc += y;
    }
    if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = inc;
    __syncthreads();
    unsigned long long wpre = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) wpre += ws[w];
    unsigned long long run = tile_pre[blockIdx.x] + wpre + inc - s;
    i64 col = base / nblk, rem = base - col * nblk;           // element base + e is (column col, block rem)
#pragma unroll
    for (int e = 0; e < SCAN_E; ++e) {
        const i64 i = base + e;
        if (i < n) {
            v[i] = (uint32_t)run;
            if (rem == 0) colptr[col] = (i64)run + 1;
        }
        run += x[e];
        if (++rem == nblk) { rem = 0; ++col; }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) colptr[ncols] = (i64)(*total) + 1;
}

// ---- scatter ----
// shared memory: sval[CSC_P * k] staged values, dbase[ncols] (u32: position of the block's run in rowval / nzval),
// cstart[ncols + 1] (u32: first slot of the column's run inside the block), wcnt[CSC_WARPS][ncols] (u16: entries of a
// warp in a column, then the warp's offset inside the column's run), srow / scol[CSC_P * k] (u16)
template <int K>
__global__ void __launch_bounds__(CSC_P, 2)
csc_scatter_kernel(const int32_t *__restrict__ interval, const double *__restrict__ vals, i64 nx, i64 nf,
                   i64 col0, int ncols, i64 nblk, const uint32_t *__restrict__ offs /*[ncols][nblk]*/,
                   i64 *__restrict__ rowval, double *__restrict__ nzval) {
    extern __shared__ __align__(16) unsigned char smraw[];
    __shared__ unsigned wtot[CSC_WARPS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int nslot = CSC_P * K;
    double *sval = reinterpret_cast<double *>(smraw);
    uint32_t *dbase = reinterpret_cast<uint32_t *>(sval + nslot);
    uint32_t *cstart = dbase + ncols;
    uint16_t *wcnt = reinterpret_cast<uint16_t *>(cstart + ncols + 1 + ((ncols + 1) & 1));   // 4-byte words below: even count
    uint16_t *srow = wcnt + (size_t)CSC_WARPS * (ncols + (ncols & 1));
    uint16_t *scol = srow + nslot;
    const int ncp = ncols + (ncols & 1);            // u16 row pitch of wcnt (even: rows start on a word)
    const int cpt = (ncols + CSC_P - 1) / CSC_P;    // columns per thread in the block-wide scan (consecutive)
    for (i64 b = blockIdx.x; b < nblk; b += gridDim.x) {
        {
            uint32_t *wz = reinterpret_cast<uint32_t *>(wcnt);
            for (int i = threadIdx.x; i < CSC_WARPS * ncp / 2; i += CSC_P) wz[i] = 0u;
        }
        // positions of the block's runs: one 4-byte gather per column (stride nblk), asynchronous — they are needed last
        for (int c = threadIdx.x; c < ncols; c += CSC_P)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(dbase + c)),
                         "l"(offs + (size_t)c * nblk + b) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
        // this thread's point and its K entries
        const i64 p = b * CSC_P + threadIdx.x;
        int I = 0;
        if (p < nx) I = interval[p];
        double v[K];
        int colv[K];
        unsigned vm = 0;                             // valid entries (column inside the restriction, value != 0)
#pragma unroll
        for (int a = 0; a < K; ++a) {
            v[a] = 0.0; colv[a] = -1;
            if (I > 0) {
                v[a] = vals[p * K + a];
                colv[a] = csc_entry_col(I, a, K, nf, col0, ncols, v[a]);
                if (colv[a] >= 0) vm |= 1u << a;
            }
        }
        // rank inside the warp = number of lower lanes (earlier rows) with an entry in the same column.  Lane L' holds
        // the columns I' - K + 1 + a'; its entry a' = a - (I' - I) is in my column a, so its valid mask shifted by
        // I' - I lines up with mine.  Lanes further than K-1 intervals away (the usual case for unsorted points) are
        // skipped by a warp-uniform test.
        unsigned char rk[K];
#pragma unroll
        for (int a = 0; a < K; ++a) rk[a] = 0;
        unsigned above = 0;                          // my columns that a higher lane also holds
        for (int Lp = 0; Lp < 32; ++Lp) {
            const int Ip = __shfl_sync(0xffffffffu, I, Lp);
            const int d = Ip - I;
            const bool near = Lp != lane && d > -K && d < K && Ip > 0 && vm != 0u;
            if (__any_sync(0xffffffffu, near)) {     // warp-uniform: unsorted points rarely share columns inside a warp
                const unsigned vp = __shfl_sync(0xffffffffu, vm, Lp);
                if (near && vp != 0u) {
                    const unsigned sh = (d >= 0 ? (vp << d) : (vp >> (-d))) & vm;
                    if (Lp < lane) {
#pragma unroll
                        for (int a = 0; a < K; ++a) rk[a] += (sh >> a) & 1u;
                    } else {
                        above |= sh;
                    }
                }
            }
        }
        __syncthreads();                             // wcnt zeroed
        // the highest lane of a (warp, column) knows the count
        uint16_t *wc = wcnt + (size_t)warp * ncp;
#pragma unroll
        for (int a = 0; a < K; ++a)
            if (((vm >> a) & 1u) && !((above >> a) & 1u)) wc[colv[a]] = (uint16_t)(rk[a] + 1);
        __syncthreads();
        // per column: counts of the warps -> offsets of the warps inside the column's run, run length -> block scan
        unsigned mine[8];                            // cpt <= 8 (ncols <= 4096)
        unsigned tsum = 0;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            mine[e] = 0;
            const int c = threadIdx.x * cpt + e;
            if (e < cpt && c < ncols) {
                unsigned run = 0;
                for (int w = 0; w < CSC_WARPS; ++w) {
                    const unsigned t = wcnt[(size_t)w * ncp + c];
                    wcnt[(size_t)w * ncp + c] = (uint16_t)run;
                    run += t;
                }
                mine[e] = run;
                tsum += run;
            }
        }
        unsigned inc = tsum;
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += y;
        }
        if (lane == 31) wtot[warp] = inc;
        __syncthreads();
        unsigned pre = inc - tsum;
        for (int w = 0; w < warp; ++w) pre += wtot[w];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int c = threadIdx.x * cpt + e;
            if (e < cpt && c < ncols) { cstart[c] = pre; pre += mine[e]; }
        }
        if (threadIdx.x == CSC_P - 1) cstart[ncols] = pre;
        __syncthreads();
        // place the entries in (column, row) order
#pragma unroll
        for (int a = 0; a < K; ++a)
            if ((vm >> a) & 1u) {
                const unsigned s = cstart[colv[a]] + wc[colv[a]] + rk[a];
                sval[s] = v[a];
                srow[s] = (uint16_t)threadIdx.x;
                scol[s] = (uint16_t)colv[a];
            }
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();
        const unsigned tot = cstart[ncols];
        const i64 row0 = b * CSC_P + 1;             // 1-based row of the block's first point
        for (unsigned s = threadIdx.x; s < tot; s += CSC_P) {
            const unsigned c = scol[s];
            const size_t dst = (size_t)dbase[c] + (s - cstart[c]);
            rowval[dst] = row0 + srow[s];
            nzval[dst] = sval[s];
        }
        __syncthreads();
    }
}

static size_t csc_scatter_smem(int k, int ncols) {
    const size_t nslot = (size_t)CSC_P * k, ncp = (size_t)ncols + (ncols & 1);
    return nslot * 8 + (size_t)ncols * 4 + ((size_t)ncols + 2) * 4 + (size_t)CSC_WARPS * ncp * 2 + nslot * 2 + nslot * 2 + 32;
}

// True if the hand-written path applies (else the caller falls back to the radix-sort path of bspline.cu).
bool cb_csc_fast_applicable(int k, i64 nx, i64 ncols) {
    if (ncols < 1 || ncols > 8 * CSC_P || nx < 1 || k < 1 || k > CB_KMAX) return false;
    const i64 nblk = (nx + CSC_P - 1) / CSC_P;
    if (nblk * ncols > ((i64)1 << 27)) return false;          // 512 MB of counters
    return csc_scatter_smem(k, (int)ncols) <= 200 * 1024;
}

size_t cb_csc_fast_scratch_bytes(i64 nx, i64 ncols) {
    const i64 nblk = (nx + CSC_P - 1) / CSC_P, n = nblk * ncols, ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    return (size_t)n * 4 + (size_t)(ntiles + 2) * 8 + 64;
}

// interval / vals: the rows of K1 (device); scratch: cb_csc_fast_scratch_bytes bytes (device, 8-byte aligned).
// rowval / nzval may be NULL (structure only: colptr).  Asynchronous on st.
int cb_csc_fast(const int32_t *interval, const double *vals, i64 nx, int k, i64 nf, i64 col0, i64 ncols,
                void *scratch, i64 *colptr, i64 *rowval, double *nzval, cudaStream_t st) {
    const i64 nblk = (nx + CSC_P - 1) / CSC_P, n = nblk * ncols, ntiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    uint32_t *cnt = static_cast<uint32_t *>(scratch);
    unsigned long long *tile = reinterpret_cast<unsigned long long *>(static_cast<char *>(scratch) + (((size_t)n * 4 + 7) & ~(size_t)7));
    unsigned long long *total = tile + ntiles;
    const int grid = (int)(nblk < (i64)CB_NUM_SMS * 8 ? nblk : (i64)CB_NUM_SMS * 8);
    CB_CUDA(cudaFuncSetAttribute(csc_count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(uint32_t) * (size_t)ncols)));
    csc_count_kernel<<<grid, CSC_P, sizeof(uint32_t) * (size_t)ncols, st>>>(interval, vals, nx, k, nf, col0, (int)ncols, nblk, cnt);
    CB_LAUNCH_CHECK();
    scan_tile_sums_kernel<<<(unsigned)ntiles, SCAN_T, 0, st>>>(cnt, n, tile);
    scan_tiles_kernel<<<1, 1024, 0, st>>>(tile, ntiles, total);
    scan_apply_kernel<<<(unsigned)ntiles, SCAN_T, 0, st>>>(cnt, n, tile, nblk, ncols, total, colptr);
    CB_LAUNCH_CHECK();
    if (rowval && nzval) {
        const size_t smem = csc_scatter_smem(k, (int)ncols);
#define CB_CSC_CASE(KT)                                                                                              \
    case KT: {                                                                                                       \
        auto kern = csc_scatter_kernel<KT>;                                                                          \
        CB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));                 \
        int per_sm = 1;                                                                                              \
        CB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, CSC_P, smem));                          \
        if (per_sm < 1) per_sm = 1;                                                                                  \
        const i64 cap = (i64)CB_NUM_SMS * per_sm;                                                                    \
        kern<<<(int)(nblk < cap ? nblk : cap), CSC_P, smem, st>>>(interval, vals, nx, nf, col0, (int)ncols, nblk, cnt, rowval, nzval); \
    } break;
        switch (k) {
            CB_CSC_CASE(1) CB_CSC_CASE(2) CB_CSC_CASE(3) CB_CSC_CASE(4) CB_CSC_CASE(5) CB_CSC_CASE(6) CB_CSC_CASE(7) CB_CSC_CASE(8)
            CB_CSC_CASE(9) CB_CSC_CASE(10) CB_CSC_CASE(11) CB_CSC_CASE(12) CB_CSC_CASE(13) CB_CSC_CASE(14) CB_CSC_CASE(15) CB_CSC_CASE(16)
            default: cb_set_error("cb_bspline_eval_csc: order k=%d outside 1..%d", k, CB_KMAX); return CB_ERR_UNSUPPORTED;
        }
#undef CB_CSC_CASE
        CB_LAUNCH_CHECK();
    }
    return CB_OK;
}
