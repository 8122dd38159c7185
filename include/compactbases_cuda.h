/*
 * compactbases_cuda.h — C ABI of libcompactbases_cuda.so
 *
 * B200 (sm_100a) implementation of the data-parallel hot path of
 * CompactBases.jl v0.3.1.  Each entry point names the reference method(s)
 * (file:line in the reference tree) whose body it replaces; INTEGRATION.md
 * shows the Julia `ccall` glue.
 *
 * Conventions
 *  - Every function returns an int status: CB_OK (0) or a CB_ERR_* code;
 *    cb_last_error() gives a thread-local message.  The reference throws
 *    before doing work (ArgumentError / DimensionMismatch / BoundsError);
 *    the glue keeps those checks and maps a non-zero status to an exception.
 *  - Arrays are caller-owned.  Unless the name ends in `_host`, array
 *    arguments are DEVICE pointers (CUDA.jl CuPtr / torch data_ptr) valid on
 *    the current device; `_host` forms take host pointers and copy in/out
 *    internally (pinned or pageable).
 *  - Matrices are column-major.  Coefficient batches X, Y are n x nvec with
 *    leading dimension ldx/ldy (each coefficient vector contiguous).
 *  - BandedMatrix storage: data[(u + i - j) + (l+u+1)*j], 0-based i, j.
 *  - Index arguments documented as "1-based" follow Julia; `*0` offsets
 *    (rowL0, colR0, col0) are 0-based counts of skipped functions.
 *  - `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *    Calls are asynchronous w.r.t. the host unless stated otherwise.
 *  - No CPU fallback exists: without a CUDA device every compute entry
 *    returns CB_ERR_CUDA.
 */
#ifndef COMPACTBASES_CUDA_H
#define COMPACTBASES_CUDA_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int64_t cb_i64;

enum {
    CB_OK = 0,
    CB_ERR_ARG = 1,          /* invalid argument / incompatible bases / dimension mismatch */
    CB_ERR_CUDA = 2,         /* CUDA runtime error (message in cb_last_error) */
    CB_ERR_UNSUPPORTED = 3,  /* valid request outside the compiled range (e.g. order > CB_KMAX) */
    CB_ERR_ALLOC = 4
};

#define CB_KMAX 16           /* largest B-spline order with a compiled kernel */
#define CB_FEDVR_OMAX 32     /* largest FE-DVR element order */

const char *cb_version(void);
const char *cb_last_error(void);
int cb_device_count(int *count);

/* ------------------------------------------------------------------------
 * B-spline basis handle  ==  the reference's `BSpline(t, x, w, ...)` struct
 * (src/bsplines.jl:9-20,95-119) minus the cached sparse table and metric.
 * tt = t.t (host, n_tt = numintervals+1 non-decreasing knots); ml/mr are the
 * implicit end multiplicities (src/knot_sets.jl:14,138-151); xi_ref/w_ref are
 * the q Gauss-Legendre reference nodes/weights on [-1,1] (host;
 * FastGaussQuadrature.gausslegendre(q), src/quadrature.jl:62).  Creation
 * uploads the knots and runs the lgwt kernel (src/quadrature.jl:60-77).
 * ---------------------------------------------------------------------- */
typedef struct cb_bspline_s cb_bspline;

int cb_bspline_create(const double *tt_host, cb_i64 n_tt, int k, int ml, int mr,
                      const double *xi_ref_host, const double *w_ref_host, int q,
                      cb_bspline **out);
int cb_bspline_destroy(cb_bspline *B);
cb_i64 cb_bspline_numfunctions(const cb_bspline *B);      /* src/knot_sets.jl:71 */
cb_i64 cb_bspline_numknots(const cb_bspline *B);          /* src/knot_sets.jl:116 */
cb_i64 cb_bspline_num_nonempty(const cb_bspline *B);      /* length(nonempty_intervals(t)), :196 */
cb_i64 cb_bspline_numnodes(const cb_bspline *B);          /* length(B.x) */
/* Device pointers to locs(B) / weights(B) (src/bsplines.jl:136-137), owned by the handle. */
int cb_bspline_nodes(const cb_bspline *B, const double **x_dev, const double **w_dev);
/* Copies locs(B) / weights(B) into caller-owned device arrays of numnodes entries (either may be NULL). */
int cb_bspline_nodes_copy(const cb_bspline *B, double *x_dev, double *w_dev, void *stream);

/* B[x, :] in row ("ELL") form — replaces getindex(B::BSpline, x, sel) and the
 * scalar form (src/bsplines.jl:204-224), find_interval / within_support
 * (src/knot_sets.jl:209-238) and deBoor (src/bsplines.jl:159-202).
 * interval[p] (int32): 1-based knot interval containing x[p] (intervals are
 *   [t_i, t_{i+1}), the last one closed), 0 if x[p] is outside the support.
 * vals[p*k + a]: m-th derivative of function j = interval[p]-k+1+a at x[p]
 *   (0 where j is not in 1..nf or the point is outside). */
int cb_bspline_eval(const cb_bspline *B, const double *x, cb_i64 nx, int m,
                    int32_t *interval, double *vals, void *stream);
int cb_bspline_eval_host(const cb_bspline *B, const double *x_host, cb_i64 nx, int m,
                         int32_t *interval_host, double *vals_host);

/* B[x, col0+1 : col0+ncols] as Julia's SparseMatrixCSC{Float64,Int64}
 * (src/bsplines.jl:217-224): 1-based colptr[ncols+1] / rowval, rows ascending
 * within a column, exact zeros not stored.  rowval/nzval need room for
 * k*nx entries.  *nnz_host receives the count (the call synchronises). */
int cb_bspline_eval_csc(const cb_bspline *B, const double *x, cb_i64 nx,
                        cb_i64 col0, cb_i64 ncols,
                        cb_i64 *colptr, cb_i64 *rowval, double *nzval,
                        cb_i64 *nnz_host, void *stream);

/* Fused spline evaluation (B*C)[x] = B[x,:]*C for a batch of coefficient
 * vectors (test/bsplines/splines.jl:68-88): out[p + ldo*v]. */
int cb_bspline_eval_spline(const cb_bspline *B, const double *x, cb_i64 nx, int m,
                           const double *C, cb_i64 ldc, cb_i64 nvec,
                           double *out, cb_i64 ldo, void *stream);

/* basis_functions(t, x, m) (src/bsplines.jl:28-52) as dense tables
 * table[(c*q + l)*k + a] for the c-th non-empty interval. */
int cb_bspline_quad_table(const cb_bspline *B, int m, double *table, void *stream);

/* overlap_matrix! / diffop! and the copyto! bodies built on them
 * (src/bsplines.jl:58-93,282-324; src/bspline_derivatives.jl:3-80;
 * src/densities.jl:206-208):
 *   band[i,j] = < (-1)^a d^a L_i | op | d^b R_j >  by Gauss-Legendre quadrature.
 * L, R must share t.t and the nodes (ArgumentError in the reference,
 * src/bsplines.jl:75-82).  Rows are functions rowL0+1..rowL0+m of L, columns
 * colR0+1..colR0+n of R (restrictions B[:,a:b]); (l,u) must be the bandwidths
 * of Matrix(undef, A, B) (src/bsplines.jl:258-270), i.e.
 * (k-1+ij, k-1-ij), k = max(kL,kR), ij = colR0-rowL0.
 * opvals: NULL, or the diagonal operator at the quadrature nodes
 * (getindex.(Ref(D.diag), locs(B)), src/bsplines.jl:305). */
int cb_bspline_assemble(const cb_bspline *L, const cb_bspline *R, int a, int b,
                        const double *opvals,
                        cb_i64 rowL0, cb_i64 m, cb_i64 colR0, cb_i64 n,
                        int l, int u, double *band, void *stream);
/* Same with op(x) = -Z/x evaluated on chip (B'*QuasiDiagonal(-Z/r)*B,
 * test/bsplines/operators.jl:19), no opvals array read. */
int cb_bspline_assemble_coulomb(const cb_bspline *L, const cb_bspline *R, double Z,
                                cb_i64 rowL0, cb_i64 m, cb_i64 colR0, cb_i64 n,
                                int l, int u, double *band, void *stream);

/* Sharded form: writes only columns col_lo..col_hi-1 (0-based,
 * within the restriction) into band_cols (column col_lo first), summing only
 * over the t.t intervals s_lo..s_hi-1 (0-based).  coulomb: 0 = opvals/none,
 * 1 = op(x) = -Z/x.  An interval-sharded assembly gives each rank an interval
 * range and the columns that touch it; the k-1 boundary columns hold partial
 * sums that the neighbours add (halo exchange), or each rank recomputes its
 * own columns over the full interval range (no exchange). */
int cb_bspline_assemble_part(const cb_bspline *L, const cb_bspline *R, int a, int b,
                             const double *opvals, int coulomb, double Z,
                             cb_i64 rowL0, cb_i64 m, cb_i64 colR0, cb_i64 n, int l, int u,
                             cb_i64 col_lo, cb_i64 col_hi, cb_i64 s_lo, cb_i64 s_hi,
                             double *band_cols, void *stream);
/* Host-pointer form: band_host is the data of a host BandedMatrix ((l+u+1) x n), opvals_host NULL or numnodes
 * doubles; coulomb 0 / 1 as in cb_bspline_assemble_part.  Synchronous (returns when band_host is complete). */
int cb_bspline_assemble_host(const cb_bspline *L, const cb_bspline *R, int a, int b, const double *opvals_host,
                             int coulomb, double Z, cb_i64 rowL0, cb_i64 m, cb_i64 colR0, cb_i64 n, int l, int u,
                             double *band_host, void *stream);
/* The two operators of a radial Hamiltonian on the (restricted) basis R = B[:, col0+1 : col0+n] in one launch:
 *   bandV <- R' * QuasiDiagonal(op) * R   (src/bsplines.jl:300-310; opvals / coulomb / Z as in cb_bspline_assemble_part)
 *   bandT <- R' * D' * D * R              (src/bspline_derivatives.jl:48-63, diffop! with a = b = 1)
 * Both are (2k-1) x n BandedMatrix data arrays (l = u = k-1).  Values and first derivatives of the basis functions at
 * the quadrature nodes come from one Cox-de Boor sweep, so the pair costs about 1.3 single assemblies; the entries are
 * bit-identical to the two separate calls (test/bsplines/operators.jl:15-22 builds exactly this pair). */
int cb_bspline_assemble_pair(const cb_bspline *B, const double *opvals, int coulomb, double Z,
                             cb_i64 col0, cb_i64 n, double *bandV, double *bandT, void *stream);
/* BandedMatrix(l=u=1) -> Tridiagonal(dl,d,du): the k == 2 && m == n case of
 * Matrix(undef, A, B) (src/bsplines.jl:258-262). */
int cb_band_to_tridiag(const double *band, cb_i64 n, double *dl, double *d, double *du, void *stream);

/* ------------------------------------------------------------------------
 * FE-DVR (src/fedvr*.jl).  The basis is described by the fields of the FEDVR
 * struct (src/fedvr.jl:3-53), built on the host by the caller:
 *   x[n]      B.x, nodes (neighbouring elements share one node)
 *   wcat      the per-element weight vectors B.w^i concatenated (sum of orders)
 *   nrm[n]    B.n, inverse-sqrt weights with bridge normalisation
 *   estart[e] 0-based index in x of the first node of element e
 *   order[e]  number of nodes of element e
 *   boff[e]   offset of element e's o x o block in `blocks` (sum of o^2 before)
 * `blocks` is the device-native operator layout: one dense column-major
 * o x o matrix per element holding that element's own contribution
 * (diff(B,n,i), src/fedvr_derivatives.jl:43-57); entries on a shared bridge
 * node are the sum of the two neighbouring corners (src/fedvr_operators.jl:
 * 104-130) and are added when the operator is applied or exported.
 * ---------------------------------------------------------------------- */
/* derop!(A, B, nder) / set_elements!(A, B, difffun(B, nder))
 * (src/fedvr_derivatives.jl:15-63, src/fedvr_operators.jl:148-154).
 * woff[e]: offset of element e's weights in wcat (sum of orders before).
 * uniform_order > 0: every element has that order and the per-element arrays
 * may be NULL (estart = e*(o-1), woff = e*o, boff = e*o*o); otherwise
 * max_order = maximum(order). */
int cb_fedvr_assemble(const double *x, const double *wcat, const double *nrm,
                      const cb_i64 *estart, const int32_t *order, const cb_i64 *boff,
                      const cb_i64 *woff, cb_i64 nel, int uniform_order, int max_order,
                      int nder /*1: B'DB, 2: B'D'DB*/, double *blocks, void *stream);
/* Dense (last-first+1)^2 column-major copy of the operator restricted to
 * functions first..last (1-based), what set_elements! builds
 * (src/fedvr_operators.jl:148-154); for export/tests on small bases. */
int cb_fedvr_to_dense(const double *blocks, const cb_i64 *estart, const int32_t *order,
                      const cb_i64 *boff, cb_i64 nel, cb_i64 first, cb_i64 last,
                      double *A, void *stream);
/* B[x, :] for FE-DVR bases (getindex, src/fedvr.jl:211-222, 272-295), row form: for every point the (1-based) element
 * that contains it (0: outside the basis) and the values of that element's order[e] functions, vals[p * max_order + m]
 * (function estart[e] + m + 1 of the basis; zero-padded to max_order).  t: the nel + 1 element boundaries (device).  A
 * point on an interior boundary is reported with the element to its right. */
int cb_fedvr_eval(const double *x_nodes, const double *nrm, const double *t, const cb_i64 *estart,
                  const int32_t *order, cb_i64 nel, int max_order, const double *xq, cb_i64 nx,
                  int32_t *elem, double *vals, void *stream);
/* B[x, :] for finite-difference bases (getindex, src/finite_differences.jl:63-93), row form: first[p] = the (1-based)
 * first function that can be non-zero at x_p (0: none), vals[2p], vals[2p+1] = weight * tent of that function and the
 * next.  nodes: the n grid points (device); weights: weights(B) (device) or NULL on uniform grids. */
int cb_fd_eval(const double *nodes, const double *weights, cb_i64 n, const double *xq, cb_i64 nx,
               int32_t *first, double *vals, void *stream);
/* The same operator as a BandedMatrix with l = u = w (>= max order - 1), data (2w+1) x (last-first+1)
 * column-major: the form factorize(A - sigma*I) of ShiftAndInvert takes for FE-DVR operators
 * (src/linear_operators.jl:185-236, test/linear_operators.jl:72-88 "FE-DVR"); feeds cb_bandchol_* / cb_bandldl_* /
 * cb_bandlu_*. */
int cb_fedvr_to_band(const double *blocks, const cb_i64 *estart, const int32_t *order,
                     const cb_i64 *boff, cb_i64 nel, cb_i64 first, cb_i64 last, int w,
                     double *band, void *stream);
/* The same operator in the reference's own destination types (the body of derop!(A, B, n) =
 * set_elements!(difffun(B, n), A, B), src/fedvr_derivatives.jl:62-63, src/fedvr_operators.jl:40-154):
 *
 * BlockSkylineMatrix (BlockBandedMatrices 0.10), what Matrix(undef, B) allocates (src/fedvr_operators.jl:24-28) from
 * rows = block_structure(B), (l, u) = block_bandwidths(B, rows) (src/fedvr.jl:121-176).  Its storage is one flat
 * `data` vector; block column J keeps the block rows max(1, J-u[J]) : min(J+l[J], N) stacked into one dense
 * column-major panel of block_strides[J] rows whose first entry is data[block_starts[first K, J]].  Passed 0-based:
 *   blk_cum[b]     cumulative block sizes, b = 0..nblk (block b holds functions blk_cum[b]..blk_cum[b+1]-1)
 *   col_kfirst[J]  first / last stored block row of block column J
 *   col_klast[J]
 *   col_start[J]   index in data of the first entry of the panel of block column J
 *   col_stride[J]  rows of that panel
 * (device arrays of nblk (+1) entries; INTEGRATION.md derives them from A.block_sizes).  data (ndata doubles,
 * device) is zeroed (A .= false, :149) and filled; shared bridge entries are the sums of two corners (:108-130).
 * The call synchronises `stream`; a block structure that cannot hold the element blocks -> CB_ERR_ARG. */
int cb_fedvr_to_skyline(const double *blocks, const cb_i64 *estart, const int32_t *order, const cb_i64 *boff,
                        cb_i64 nel, cb_i64 first, cb_i64 last,
                        const cb_i64 *blk_cum, cb_i64 nblk, const cb_i64 *col_kfirst, const cb_i64 *col_klast,
                        const cb_i64 *col_start, const cb_i64 *col_stride,
                        double *data, cb_i64 ndata, void *stream);
/* Tridiagonal(dl, d, du): the destination when every element has order 2 (src/fedvr_operators.jl:18-23,136-146). */
int cb_fedvr_to_tridiag(const double *blocks, const cb_i64 *estart, const int32_t *order, const cb_i64 *boff,
                        cb_i64 nel, cb_i64 first, cb_i64 last, double *dl, double *d, double *du, void *stream);
/* Y = alpha*A*X + beta*Y for the (restricted) FE-DVR operator: mul! on the
 * BlockSkylineMatrix / Tridiagonal of src/fedvr_operators.jl:17-30.
 * X, Y have last-first+1 rows.  uniform_order > 0 asserts that every element
 * has that order (fast path); 0 = per-element orders are read. */
int cb_fedvr_apply(const double *blocks, const cb_i64 *estart, const int32_t *order,
                   const cb_i64 *boff, cb_i64 nel, int uniform_order,
                   cb_i64 first, cb_i64 last,
                   const double *X, cb_i64 ldx, cb_i64 nvec,
                   double alpha, double beta, double *Y, cb_i64 ldy, void *stream);
/* derop!(A, B, nder) followed by mul!(Y, A, X, alpha, beta): cb_fedvr_assemble + cb_fedvr_apply enqueued by one
 * call (`blocks` receives the operator).  The apply pipeline is launched behind the assembly kernel with
 * programmatic stream serialisation: it fetches X while the assembly drains and waits for the blocks before it
 * reads them; X and Y must not overlap the node / weight / block arrays. */
int cb_fedvr_assemble_apply(const double *x, const double *wcat, const double *nrm,
                            const cb_i64 *estart, const int32_t *order, const cb_i64 *boff, const cb_i64 *woff,
                            cb_i64 nel, int uniform_order, int max_order, int nder, double *blocks,
                            cb_i64 first, cb_i64 last, const double *X, cb_i64 ldx, cb_i64 nvec,
                            double alpha, double beta, double *Y, cb_i64 ldy, void *stream);
/* Host-buffer forms (`*_apply_host`): X_host/Y_host in host memory (pinned for PCIe speed, see
 * cb_host_register; pageable arrays work, staged by the driver); the vector batch is streamed through the
 * device in tiles on the library's own copy streams (H2D, kernel, D2H of consecutive tiles overlap).
 * `stream` is the PRODUCER stream: the one on which the caller enqueued the work that fills the operator
 * arrays (cb_fedvr_assemble, cb_fd_derivative, ...); the copy streams wait for everything enqueued on it so
 * far.  The call returns when Y_host is complete.  Calls on one device are serialised by the library (the
 * staging ring is shared), so the entries are thread-safe but not concurrent. */
int cb_fedvr_apply_host(const double *blocks_dev, const cb_i64 *estart_dev,
                        const int32_t *order_dev, const cb_i64 *boff_dev,
                        cb_i64 nel, int uniform_order, cb_i64 first, cb_i64 last,
                        const double *X_host, cb_i64 ldx, cb_i64 nvec,
                        double alpha, double beta, double *Y_host, cb_i64 ldy, void *stream);

/* ------------------------------------------------------------------------
 * Finite differences (src/fd_derivatives.jl).
 * scheme: 0 FiniteDifferences (:12-15), 1 StaggeredFiniteDifferences uniform
 * (:20-25), 2 StaggeredFiniteDifferences non-uniform (:27-92; r[N] required;
 * fhalf NULL or f(max(0,r_{j-1/2})), j=1..N+1, for the weighted forms).
 * Writes rows i0+1..i0+n (restriction) of
 *   nder=1: Tridiagonal(-delta/2h, gamma/2h, delta/2h)        (:233-237)
 *   nder=2: (Sym)Tridiagonal(alpha/h^2, -2beta/h^2, alpha/h^2) (:265-277)
 * into dl[n-1], d[n], du[n-1] (dl or du may be NULL).  For nder=2 with
 * i0==0, ell==0, Z!=0 and a staggered scheme the Coulomb correction of
 * laplacian_correction! (:542-551) is applied to d[0]. */
int cb_fd_derivative(int scheme, int nder, const double *r, const double *fhalf,
                     cb_i64 N, double step, cb_i64 i0, cb_i64 n,
                     double Z, int ell,
                     double *dl, double *d, double *du, void *stream);

/* The same operators between DIFFERENT restrictions A = R[:, rowi0+1 : rowi0+m] and B = R[:, colj0+1 : colj0+n]
 * (derivative_matrix, src/fd_derivatives.jl:214-219; band fills :244-249, :278-289): the destination is a
 * BandedMatrix{T}(undef, (m, n), (1 - offset, 1 + offset)), offset = rowi0 - colj0; band is its data, 3 x n
 * column-major (slots outside the matrix are zeroed).  The Coulomb correction is applied only when both
 * restrictions start at the first node (:620-622). */
int cb_fd_derivative_banded(int scheme, int nder, const double *r, const double *fhalf, cb_i64 N, double step,
                            cb_i64 rowi0, cb_i64 m, cb_i64 colj0, cb_i64 n, double Z, int ell,
                            double *band, void *stream);
/* ImplicitFiniteDifferences on a non-uniform grid (Patchkovskii & Muller 2016): implicit_stencil_common and its
 * callers (src/fd_derivatives.jl:105-143, implicit_lhs :410-424) for nodes i0+1 .. i0+n of r[N]:
 *   alpha_t[n-1], alpha[n-1], beta[n]   right-hand side  Tridiagonal(alpha_t, -2 beta, alpha)
 *   m_dl[n-1], m_d[n], m_du[n-1]        left-hand side   Tridiagonal(m_dl, m_d, m_du)
 * with the reference's order of operations (bit-identical to a scalar evaluation). */
int cb_fd_implicit_stencils(const double *r, cb_i64 N, cb_i64 i0, cb_i64 n,
                            double *alpha_t, double *alpha, double *beta,
                            double *m_dl, double *m_d, double *m_du, void *stream);

/* ------------------------------------------------------------------------
 * Operator application  Y = alpha*A*X + beta*Y  (mul!; call
 * sites src/linear_operators.jl:40,50-52, src/inner_products.jl:30,46).
 * ---------------------------------------------------------------------- */
int cb_banded_apply(const double *band, cb_i64 m, cb_i64 n, int l, int u,
                    const double *X, cb_i64 ldx, cb_i64 nvec,
                    double alpha, double beta, double *Y, cb_i64 ldy, void *stream);
/* Tridiagonal(dl,d,du); SymTridiagonal(dv,ev) is dl=du=ev. */
int cb_tridiag_apply(const double *dl, const double *d, const double *du, cb_i64 n,
                     const double *X, cb_i64 ldx, cb_i64 nvec,
                     double alpha, double beta, double *Y, cb_i64 ldy, void *stream);
int cb_tridiag_apply_host(const double *dl_dev, const double *d_dev, const double *du_dev,
                          cb_i64 n, const double *X_host, cb_i64 ldx, cb_i64 nvec,
                          double alpha, double beta, double *Y_host, cb_i64 ldy, void *stream);
int cb_banded_apply_host(const double *band_dev, cb_i64 m, cb_i64 n, int l, int u,
                         const double *X_host, cb_i64 ldx, cb_i64 nvec,
                         double alpha, double beta, double *Y_host, cb_i64 ldy, void *stream);
/* Page-locks / releases a caller-owned host array (cudaHostRegister): a Julia `Array` is pageable, and the
 * `_host` entries reach PCIe speed only from pinned memory (CUDA.jl: `CUDA.pin(A)`). */
int cb_host_register(void *ptr, size_t bytes);
int cb_host_unregister(void *ptr);
/* Diagonal (FE-DVR / FD scalar operators, src/fedvr_operators.jl:158-185,
 * src/fd_operators.jl:3-35). */
int cb_diag_apply(const double *d, cb_i64 n, const double *X, cb_i64 ldx, cb_i64 nvec,
                  double alpha, double beta, double *Y, cb_i64 ldy, void *stream);
/* Y = alpha*X + beta*Y for n x nvec column-major batches (beta == 0: Y is not read): the update
 * `y .= beta .* y .+ alpha .* rho` of DiagonalOperator's mul!, src/linear_operators.jl:129-138. */
int cb_axpby(cb_i64 n, cb_i64 nvec, double alpha, const double *X, cb_i64 ldx,
             double beta, double *Y, cb_i64 ldy, void *stream);
/* Pairwise inner products of coefficient batches, src/inner_products.jl:3-46 (`u' * (S*v)`, `dot(u, S, v)`, `norm`):
 * out[p] = scale * sum_i U[i,p] * W[i,p] with W = S*V already applied (cb_banded_apply for B-splines; W = V and
 * scale = step(B) for finite differences, scale = 1 for FE-DVR).  out: P doubles on the device; the summation order
 * is fixed (reproducible). */
int cb_batched_dot(const double *U, cb_i64 ldu, const double *W, cb_i64 ldw, cb_i64 n, cb_i64 P,
                   double scale, double *out, void *stream);

/* ------------------------------------------------------------------------
 * ComplexF64.  The reference runs every basis with T = ComplexF64 (test/linear_operators.jl:1-35:
 * complex potential, complex coefficients) and has a complex-scaled FE-DVR basis (src/fedvr.jl:17-37,95-96).  A complex
 * operator is a PAIR of real device operators A = Ar + i Ai (assembly is linear in the potential; FE-DVR blocks come as
 * two real block arrays); a complex batch Z (n x nvec ComplexF64, column-major, interleaved re / im — Julia's layout;
 * leading dimensions in complex elements) is split into the real batch R = [Re | Im] of 2 nvec vectors, the real apply
 * entries above run on R, and one call recombines:
 *   T1 = Ar R, T2 = Ai R (NULL for a real operator):   Y = alpha (T1[Re] - T2[Im] + i (T1[Im] + T2[Re])) + beta Y
 * with complex alpha, beta (beta == 0: Y is not read). */
int cb_complex_split(const double *Z, cb_i64 ldz, cb_i64 n, cb_i64 nvec, double *R, cb_i64 ldr, void *stream);
int cb_complex_combine(const double *T1, const double *T2, cb_i64 ldt, cb_i64 n, cb_i64 nvec,
                       double alpha_re, double alpha_im, double beta_re, double beta_im,
                       double *Y, cb_i64 ldy, void *stream);
/* derop!(A, B, nder) for the complex-scaled FE-DVR basis (src/fedvr_derivatives.jl:15-63 with complex nodes x and
 * normalisation nrm — interleaved re / im — and real weights wcat): elements e >= first_rotated (0-based; i0 - 1 of the
 * reference) are divided by sqrt(e^{i phi}) (:55); Julia's dot conjugates the transposed derivative in the second-order
 * form.  blocks_re / blocks_im: the two real block arrays of cb_fedvr_assemble's layout. */
int cb_fedvr_assemble_complex(const double *x, const double *wcat, const double *nrm,
                              const cb_i64 *estart, const int32_t *order, const cb_i64 *boff, const cb_i64 *woff,
                              cb_i64 nel, int uniform_order, int max_order, int nder, cb_i64 first_rotated, double phi,
                              double *blocks_re, double *blocks_im, void *stream);

/* ------------------------------------------------------------------------
 * Mutual densities (src/densities.jl).  rho = C * (conj(RV f) .* (LV g)).
 * ---------------------------------------------------------------------- */
typedef struct cb_density_s cb_density;

/* Non-orthogonal B-spline branch (src/densities.jl:92-99): one-time banded
 * factorisation that replaces pinv(Matrix(RV)) (:96).  wvals: NULL (w=one) or
 * the weight function at the quadrature nodes (host, numnodes).  Functions
 * col0+1..col0+ncols (restriction). */
int cb_density_bspline_create(const cb_bspline *B, cb_i64 col0, cb_i64 ncols,
                              const double *wvals_host, cb_density **out, void *stream);
int cb_density_destroy(cb_density *D);
/* Number of corrected-semi-normal-equation refinement steps per apply
 * (default 0: plain banded Cholesky solve of the normal equations, which
 * agrees with the dense pinv to ~cond(V)^2 eps, < 1e-12 for k <= 10). */
int cb_density_set_refine(cb_density *D, int steps);
/* copyto!(rho, f, g) (src/densities.jl:156-164).  f is nf x Pf, g is nf x Pg,
 * Pf==Pg or one of them 1 (broadcast, :57-72); rho is nf x max(Pf,Pg).
 * conj is accepted for ABI completeness (real Float64: no-op).
 * Stream-ordered scratch from the device's memory pool (released on the stream
 * before the call returns): 32 * ncols * ceil(P / 32) doubles for the tiled
 * intermediate of the partitioned solve (one more array of the batch's size)
 * plus 2 * (ncols / 256) * (k - 1) * P doubles of chunk states. */
int cb_density_apply(const cb_density *D, const double *f, cb_i64 ldf, cb_i64 Pf,
                     const double *g, cb_i64 ldg, cb_i64 Pg, int conj,
                     double *rho, cb_i64 ldr, void *stream);
/* Host-pointer form: f_host, g_host, rho_host in (preferably pinned) host memory; the pair batch is streamed through
 * the device in tiles (copies and kernels of consecutive tiles overlap).  `stream`: producer stream, see
 * cb_fedvr_apply_host.  Returns when rho_host is complete. */
int cb_density_apply_host(const cb_density *D, const double *f_host, cb_i64 ldf, cb_i64 Pf,
                          const double *g_host, cb_i64 ldg, cb_i64 Pg, int conj,
                          double *rho_host, cb_i64 ldr, void *stream);
/* ftmp .* gtmp at the quadrature nodes (src/densities.jl:158-161), the
 * rho.tmp that copyto!(M, rho) (:206-208) feeds to overlap_matrix! as the
 * diagonal operator: pass it as `opvals` to cb_bspline_assemble.
 * tmp is numnodes x max(Pf,Pg), leading dimension ldt. */
int cb_density_nodal_product(const cb_density *D, const double *f, cb_i64 ldf, cb_i64 Pf,
                             const double *g, cb_i64 ldg, cb_i64 Pg, int conj,
                             double *tmp, cb_i64 ldt, void *stream);
/* Orthogonal branches (src/densities.jl:166-185): rho = c .* (f .* g);
 * c NULL for the uniform case (C = I). */
int cb_density_orthogonal(const double *c, cb_i64 n,
                          const double *f, cb_i64 ldf, cb_i64 Pf,
                          const double *g, cb_i64 ldg, cb_i64 Pg, int conj,
                          double *rho, cb_i64 ldr, void *stream);

/* ---- "next" rows of the scope table --------------------------------------------------
 * B \ f for B-splines, Base.:(\)(B::BSpline, f) src/bsplines.jl:329-343: the least-squares fit V c = f(x) at the
 * quadrature nodes, c = (V'V)^-1 V' f(x).  fvals: f at the nodes of the PARENT basis, (ni*q) x P column-major
 * (device); coef: ncols x P (device).  Uses the Gram factor of a plan made by cb_density_bspline_create. */
int cb_density_project(const cb_density *D, const double *fvals, cb_i64 ldf, cb_i64 P,
                       double *coef, cb_i64 ldc, void *stream);

/* Metric inverse of LinearOperator, src/linear_operators.jl:26-48 (factorize(B'B) once, ldiv! per application):
 * Cholesky factor of a symmetric positive definite BandedMatrix with l = u ((2l+1) x n column-major, device) and
 * the in-place batched solve X <- A^-1 X for n x nvec column-major batches (partitioned substitution: every
 * chunk of rows x 32 vectors is an independent warp).  Not positive definite -> CB_ERR_ARG (PosDefException). */
typedef struct cb_bandchol_s cb_bandchol;
/* `stream` (all *_create entries that read device arrays): the factorisation kernels run on it, so that a band
 * produced by earlier work on that stream is complete before it is read; the call synchronises that stream. */
int cb_bandchol_create(const double *band, cb_i64 n, int l, cb_bandchol **out, void *stream);
/* Same handle and solve for a symmetric INDEFINITE band (the shifted operator A - sigma*B of ShiftAndInvert with an
 * interior shift): L D L' without pivoting, so that the partitioned batched solve applies.  A pivot breakdown ->
 * CB_ERR_ARG; callers should check a residual (no pivoting: no growth bound) and fall back to cb_bandlu_*. */
int cb_bandldl_create(const double *band, cb_i64 n, int l, cb_bandchol **out, void *stream);
int cb_bandchol_solve(const cb_bandchol *F, double *X, cb_i64 ldx, cb_i64 nvec, void *stream);
int cb_bandchol_destroy(cb_bandchol *F);

/* factorize(A - sigma*B) of ShiftAndInvert for a general (indefinite) BandedMatrix, src/linear_operators.jl:203-236:
 * banded LU with partial pivoting (the LAPACK gbtrf / gbtrs that BandedMatrices' `factorize` / `ldiv!` call) and the
 * in-place batched solve X <- A^-1 X.  band: (l+u+1) x n column-major (device), l + u <= 30.  Exactly singular ->
 * CB_ERR_ARG (SingularException).  Positive definite operators should prefer cb_bandchol_* (partitioned solve). */
typedef struct cb_bandlu_s cb_bandlu;
int cb_bandlu_create(const double *band, cb_i64 n, int l, int u, cb_bandlu **out, void *stream);
int cb_bandlu_solve(const cb_bandlu *F, double *X, cb_i64 ldx, cb_i64 nvec, void *stream);
int cb_bandlu_destroy(cb_bandlu *F);

/* ------------------------------------------------------------------------
 * Multi-GPU (the reference is single-process, so these entries have no reference counterpart).
 * Point, vector and pair batches are split into contiguous ranges with no communication (cb_batch_range).  Only the
 * interval-sharded band assembly exchanges data: the k-1 boundary columns of neighbouring interval ranges.  That
 * exchange is fused into the assembly kernel: a rank stores the partial columns of its left neighbour directly into
 * the neighbour's mailbox (peer GPU memory mapped into this process, NVLink / NVSwitch) and raises a flag; the owner's
 * last tiles add the mailbox.  No NCCL call, no host round trip.
 *
 * One process per GPU:   cb_comm_create -> cb_comm_export (64 bytes) -> host all-gather -> cb_comm_connect.
 * One process, N GPUs:   cb_comm_create per device -> cb_comm_connect_local.
 * ---------------------------------------------------------------------- */
typedef struct cb_comm_s cb_comm;
int cb_batch_range(cb_i64 n, int rank, int world, cb_i64 *lo, cb_i64 *hi);
int cb_comm_create(int rank, int world, cb_comm **out);          /* on the current device */
int cb_comm_export(const cb_comm *c, void *handle64);            /* cudaIpcMemHandle_t of this rank's mailbox */
int cb_comm_connect(cb_comm *c, const void *all_handles);        /* world x 64 bytes, rank order */
int cb_comm_connect_local(cb_comm **comms, int world);
int cb_comm_status(const cb_comm *c, int *status, void *stream); /* non-zero: a wait for a neighbour timed out */
int cb_comm_destroy(cb_comm *c);
/* The share of rank `rank`: t.t intervals [s_lo, s_hi), owned columns [own_lo, own_hi) (those whose first interval
 * lies in the range), and touch_lo <= own_lo, the first column with any support in the range (0-based, relative to
 * the restriction colR0+1 .. colR0+n of R). */
int cb_interval_shard(const cb_bspline *R, cb_i64 colR0, cb_i64 n, int rank, int world,
                      cb_i64 *s_lo, cb_i64 *s_hi, cb_i64 *own_lo, cb_i64 *own_hi, cb_i64 *touch_lo);
/* cb_bspline_assemble for this rank's share: own_cols = columns own_lo .. own_hi-1 of the BandedMatrix data.
 * halo 0: exchange fused into the kernel (every rank must make the same sequence of such calls; needs at least k
 * intervals per rank); halo 1: recompute the k-1 boundary intervals instead (no communication).
 * coulomb: 0 = opvals / none, 1 = op(x) = -Z/x. */
int cb_bspline_assemble_sharded(cb_comm *c, const cb_bspline *L, const cb_bspline *R, int a, int b,
                                const double *opvals, int coulomb, double Z,
                                cb_i64 rowL0, cb_i64 m, cb_i64 colR0, cb_i64 n, int l, int u,
                                int halo, double *own_cols, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* COMPACTBASES_CUDA_H */
