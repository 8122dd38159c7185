// complex.cu — ComplexF64 paths of the reference:
//   * coefficient batches of ComplexF64 with real or complex operators (test/linear_operators.jl:1-35 runs every basis
//     with T = ComplexF64: complex potential, complex coefficients);
//   * the complex-scaled FE-DVR basis (src/fedvr.jl:17-37,95-96): complex nodes make the element matrices of
//     src/fedvr_derivatives.jl:15-57 complex.
// A complex operator is kept as a PAIR of real device operators (A = Ar + i Ai — assembly is linear in the potential, and
// the FE-DVR blocks are written as two real block arrays), a complex batch Z (interleaved re / im, Julia's layout) is
// split once into a real batch of 2 nvec vectors [Re | Im], the real apply kernels run on that batch — twice as many
// vectors through the same HBM-bound pipelines — and one kernel recombines
//     Y = alpha (Ar Xr - Ai Xi + i (Ar Xi + Ai Xr)) + beta Y.
#include "common.cuh"

// Z: n x nvec complex, column-major, leading dimension ldz (complex elements) -> R: n x 2 nvec real (ldr):
// columns [0, nvec) = Re, [nvec, 2 nvec) = Im
__global__ void complex_split_kernel(const double2 *__restrict__ Z, i64 ldz, i64 n, i64 nvec, double *__restrict__ R, i64 ldr) {
    const i64 v = blockIdx.y;
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        const double2 z = Z[i + ldz * v];
        R[i + ldr * v] = z.x;
        R[i + ldr * (nvec + v)] = z.y;
    }
}

// T1 = Ar [Xr | Xi], T2 = Ai [Xr | Xi] (NULL: real operator), both n x 2 nvec real; Y complex
__global__ void complex_combine_kernel(const double *__restrict__ T1, const double *__restrict__ T2, i64 ldt, i64 n, i64 nvec,
                                       double ar, double ai, double br, double bi, double2 *__restrict__ Y, i64 ldy) {
    const i64 v = blockIdx.y;
    const bool useb = br != 0.0 || bi != 0.0;
    for (i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x; i < n; i += (i64)gridDim.x * blockDim.x) {
        double wr = T1[i + ldt * v], wi = T1[i + ldt * (nvec + v)];
        if (T2) { wr -= T2[i + ldt * (nvec + v)]; wi += T2[i + ldt * v]; }
        double yr = ar * wr - ai * wi, yi = ar * wi + ai * wr;
        if (useb) {
            const double2 y = Y[i + ldy * v];
            yr += br * y.x - bi * y.y;
            yi += br * y.y + bi * y.x;
        }
        Y[i + ldy * v] = make_double2(yr, yi);
    }
}

// ---- complex-scaled FE-DVR element blocks ---------------------------------------------------------------------------
struct cplx { double re, im; };
__device__ __forceinline__ cplx cmk(double a, double b) { cplx z; z.re = a; z.im = b; return z; }
__device__ __forceinline__ cplx operator+(cplx a, cplx b) { return cmk(a.re + b.re, a.im + b.im); }
__device__ __forceinline__ cplx operator-(cplx a, cplx b) { return cmk(a.re - b.re, a.im - b.im); }
__device__ __forceinline__ cplx operator*(cplx a, cplx b) { return cmk(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
__device__ __forceinline__ cplx operator*(double s, cplx a) { return cmk(s * a.re, s * a.im); }
__device__ __forceinline__ cplx cconj(cplx a) { return cmk(a.re, -a.im); }
__device__ __forceinline__ cplx crcp(cplx a) { const double d = 1.0 / (a.re * a.re + a.im * a.im); return cmk(a.re * d, -a.im * d); }

// One warp per element (as fedvr_assemble_kernel, csrc/assemble.cu):
//   L'[m,m'] = lagrangeder!                                   (src/fedvr_derivatives.jl:15-31, complex nodes, real weights)
//   D[m,m']  = n_m n_m' dot(Lt[m,:], w .* L'[m',:]),  Lt = I | -L'   (:33-57; Julia's dot conjugates its first argument)
//   D ./= sqrt(e^{i phi}) for the rotated elements                (:55)
// x, nrm: complex (interleaved); wcat: real.  isq = 1 / sqrt(e^{i phi}); elements e >= e_rot are rotated.
constexpr int CFEA_WARPS = 4;
__global__ void __launch_bounds__(CFEA_WARPS * 32)
fedvr_assemble_complex_kernel(const double2 *__restrict__ x, const double *__restrict__ wcat, const double2 *__restrict__ nrm,
                              const i64 *__restrict__ estart, const int32_t *__restrict__ order, const i64 *__restrict__ boff,
                              const i64 *__restrict__ woff, int uniform_order, i64 nel, int nder, int omax,
                              i64 e_rot, double isq_re, double isq_im, double *__restrict__ bre, double *__restrict__ bim) {
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    cplx *Lp = reinterpret_cast<cplx *>(sm + (size_t)warp * (4 * omax * omax + 5 * omax));   // [o][o] column-major L'[m + o m']
    cplx *rd = Lp + omax * omax;                    // [o][o] 1 / (x_m - x_k)
    cplx *xe = rd + omax * omax, *ne = xe + omax;
    double *we = reinterpret_cast<double *>(ne + omax);
    const cplx isq = cmk(isq_re, isq_im);
    for (i64 e = (i64)blockIdx.x * CFEA_WARPS + warp; e < nel; e += (i64)gridDim.x * CFEA_WARPS) {
        const int o = uniform_order > 0 ? uniform_order : order[e];
        const i64 es = uniform_order > 0 ? e * (o - 1) : estart[e];
        const i64 ws = uniform_order > 0 ? e * o : woff[e];
        const i64 bo = uniform_order > 0 ? e * (i64)o * o : boff[e];
        __syncwarp();
        for (int t = lane; t < o; t += 32) {
            const double2 xv = x[es + t], nv = nrm[es + t];
            xe[t] = cmk(xv.x, xv.y); ne[t] = cmk(nv.x, nv.y); we[t] = wcat[ws + t];
        }
        __syncwarp();
        for (int t = lane; t < o * o; t += 32) {
            const int m = t % o, kk = t / o;
            rd[m * o + kk] = m == kk ? cmk(0.0, 0.0) : crcp(xe[m] - xe[kk]);
        }
        __syncwarp();
        for (int t = lane; t < o * o; t += 32) {
            const int m = t % o, mp = t / o;
            cplx v;
            if (m == mp) {
                v = cmk((double)((m == o - 1) - (m == 0)) / (2 * we[m]), 0.0);
            } else {
                cplx f = cmk(1.0, 0.0);
                const cplx xmp = xe[mp];
                const cplx *rm = rd + m * o;
                for (int kk = 0; kk < o; ++kk) {
                    if (kk == m || kk == mp) continue;
                    f = f * ((xmp - xe[kk]) * rm[kk]);
                }
                v = f * rm[mp];
            }
            Lp[t] = v;
        }
        __syncwarp();
        for (int t = lane; t < o * o; t += 32) {
            const int m = t % o, mp = t / o;
            cplx s;
            if (nder == 1) {
                s = we[m] * Lp[mp + o * m];         // dot(e_m, w .* L'[m',:]) = w_m L'[m', m]
            } else {
                s = cmk(0.0, 0.0);
                for (int qd = 0; qd < o; ++qd) s = s + cconj(cmk(-Lp[m + o * qd].re, -Lp[m + o * qd].im)) * (we[qd] * Lp[mp + o * qd]);
            }
            cplx d = (ne[m] * ne[mp]) * s;
            if (e >= e_rot) d = d * isq;
            bre[bo + t] = d.re;
            bim[bo + t] = d.im;
        }
    }
}

extern "C" {

int cb_complex_split(const double *Z, i64 ldz, i64 n, i64 nvec, double *R, i64 ldr, void *stream) {
    CB_REQUIRE(n >= 0 && nvec >= 0 && ldz >= n && ldr >= n, "DimensionMismatch: cb_complex_split: leading dimension smaller than n");
    if (n == 0 || nvec == 0) return CB_OK;
    CB_REQUIRE(Z != nullptr && R != nullptr, "cb_complex_split: NULL array");
    CB_REQUIRE(nvec <= 65535, "cb_complex_split: more than 65535 vectors per call");
    dim3 grid((unsigned)cb_grid_for(n, 256, 8), (unsigned)nvec);
    complex_split_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const double2 *>(Z), ldz, n, nvec, R, ldr);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int cb_complex_combine(const double *T1, const double *T2, i64 ldt, i64 n, i64 nvec, double alpha_re, double alpha_im,
                       double beta_re, double beta_im, double *Y, i64 ldy, void *stream) {
    CB_REQUIRE(n >= 0 && nvec >= 0 && ldt >= n && ldy >= n, "DimensionMismatch: cb_complex_combine: leading dimension smaller than n");
    if (n == 0 || nvec == 0) return CB_OK;
    CB_REQUIRE(T1 != nullptr && Y != nullptr, "cb_complex_combine: NULL array");
    CB_REQUIRE(nvec <= 65535, "cb_complex_combine: more than 65535 vectors per call");
    dim3 grid((unsigned)cb_grid_for(n, 256, 8), (unsigned)nvec);
    complex_combine_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(T1, T2, ldt, n, nvec, alpha_re, alpha_im, beta_re, beta_im,
                                                                   reinterpret_cast<double2 *>(Y), ldy);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int cb_fedvr_assemble_complex(const double *x, const double *wcat, const double *nrm,
                              const i64 *estart, const int32_t *order, const i64 *boff, const i64 *woff,
                              i64 nel, int uniform_order, int max_order, int nder, i64 first_rotated, double phi,
                              double *blocks_re, double *blocks_im, void *stream) {
    CB_REQUIRE(x && wcat && nrm && blocks_re && blocks_im, "cb_fedvr_assemble_complex: NULL array");
    CB_REQUIRE(nel >= 1, "cb_fedvr_assemble_complex: no elements");
    CB_REQUIRE(nder == 1 || nder == 2, "cb_fedvr_assemble_complex: derivative order must be 1 or 2");
    const int omax = uniform_order > 0 ? uniform_order : max_order;
    CB_REQUIRE(omax >= 2, "cb_fedvr_assemble_complex: element order must be >= 2");
    if (omax > CB_FEDVR_OMAX) { cb_set_error("cb_fedvr_assemble_complex: element order %d > %d", omax, CB_FEDVR_OMAX); return CB_ERR_UNSUPPORTED; }
    if (uniform_order <= 0) CB_REQUIRE(estart && order && boff && woff, "cb_fedvr_assemble_complex: per-element arrays required");
    const size_t smem = sizeof(double) * CFEA_WARPS * ((size_t)4 * omax * omax + 5 * omax);
    const int grid = cb_grid_for(nel, CFEA_WARPS, 16);
    CB_CUDA(cudaFuncSetAttribute(fedvr_assemble_complex_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // 1 / sqrt(e^{i phi}) = e^{-i phi / 2}
    fedvr_assemble_complex_kernel<<<grid, CFEA_WARPS * 32, smem, (cudaStream_t)stream>>>(
        reinterpret_cast<const double2 *>(x), wcat, reinterpret_cast<const double2 *>(nrm), estart, order, boff, woff, uniform_order,
        nel, nder, omax, first_rotated, cos(-0.5 * phi), sin(-0.5 * phi), blocks_re, blocks_im);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

}  // extern "C"
