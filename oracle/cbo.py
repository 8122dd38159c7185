"""Python face of the CPU oracle (TEST INFRASTRUCTURE ONLY).

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` may import this module.  The product
package ``compactbases_jl_b200`` never does.

* ctypes bindings of ``oracle/libcbo.so`` (C restatement, see ``cbo.c``);
* numpy restatements of the pieces that are integer logic or tiny dense algebra
  (FE-DVR construction and block structure, finite-difference stencils, the
  dense-``pinv`` density projection).

Every function cites the reference file:line it follows (CompactBases.jl
v0.3.1, paths relative to the reference tree).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from fractions import Fraction

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

i64 = C.c_int64
dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int64)
i32p = C.POINTER(C.c_int32)


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libcbo.so")
    src = os.path.join(_HERE, "cbo.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        L.cbo_find_interval.restype = i64
        L.cbo_find_interval.argtypes = [dp, i64, C.c_int, C.c_double, i64]
        L.cbo_find_interval_bisect.restype = i64
        L.cbo_find_interval_bisect.argtypes = [dp, i64, C.c_int, C.c_double]
        L.cbo_nonempty_intervals.restype = i64
        L.cbo_nonempty_intervals.argtypes = [dp, i64, ip]
        L.cbo_deboor_unit.restype = C.c_double
        L.cbo_deboor_unit.argtypes = [dp, i64, C.c_int, i64, i64, C.c_double, i64, C.c_int]
        L.cbo_deboor.restype = C.c_double
        L.cbo_deboor.argtypes = [dp, i64, C.c_int, dp, i64, C.c_double, i64, C.c_int]
        L.cbo_lgwt.restype = i64
        L.cbo_max_threads.restype = C.c_int
        L.cbo_bspline_eval_csc.restype = i64
        L.cbo_num_quadrature_points.restype = C.c_int
    return _LIB


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(dp)


def _p(a, t=dp):
    return a.ctypes.data_as(t)


# ---------------------------------------------------------------------------
# Knot sets (src/knot_sets.jl)
# ---------------------------------------------------------------------------

def julia_range(a, b, n):
    """Correctly rounded ``range(a, stop=b, length=n)`` (Julia's StepRangeLen
    uses TwicePrecision, i.e. the exactly rounded a + i*(b-a)/(n-1))."""
    fa, fb = Fraction(float(a)), Fraction(float(b))
    if n == 1:
        return np.array([float(a)])
    return np.array([float(fa + (fb - fa) * i / (n - 1)) for i in range(n)])


def linear_knots(a, b, N):
    """t.t of LinearKnotSet(k, a, b, N): src/knot_sets.jl:301-306."""
    return julia_range(a, b, N + 1)


def exp_knots(a, b, N, base=10.0, include0=True):
    """t.t of ExpKnotSet(k, a, b, N): src/knot_sets.jl:350-357."""
    e = julia_range(a, b, N if include0 else N + 1)
    t = np.power(float(base), e)
    return np.concatenate([[0.0], t]) if include0 else t


def expand_knots(tt, k, ml=None, mr=None):
    """collect(t): src/knot_sets.jl:116,138-151."""
    ml = k if ml is None else ml
    mr = k if mr is None else mr
    tt = np.asarray(tt, dtype=np.float64)
    return np.concatenate([np.full(ml - 1, tt[0]), tt, np.full(mr - 1, tt[-1])])


def nonempty_intervals(t):
    """src/knot_sets.jl:196-198 (1-based)."""
    t = np.asarray(t)
    return np.nonzero(t[:-1] != t[1:])[0] + 1


def find_interval(t, mr, x, i=None, ml=1):
    """src/knot_sets.jl:209-216, literal linear scan; None for `nothing`."""
    t, tp = _d(t)
    r = lib().cbo_find_interval(tp, len(t), mr, float(x), ml if i is None else i)
    return None if r == 0 else int(r)


def find_interval_bisect(t, mr, x):
    t, tp = _d(t)
    r = lib().cbo_find_interval_bisect(tp, len(t), mr, float(x))
    return None if r == 0 else int(r)


def within_support(x, tt, k, ml, mr, j):
    """src/knot_sets.jl:228-238: list of (1-based point indices, interval)."""
    x = np.asarray(x, dtype=np.float64)
    if x.size == 0:
        return []
    t = expand_knots(tt, k, ml, mr)
    ni = len(tt) - 1
    out = []
    for i in range(j, j + k):
        lo, hi = t[i - 1], t[i]
        if i - ml + 1 == ni:
            sel = (x >= lo) & (x <= hi)
        else:
            sel = (x >= lo) & (x < hi)
        idx = np.nonzero(sel)[0] + 1
        if idx.size:
            out.append((idx, i))
    return out


# ---------------------------------------------------------------------------
# Quadrature (src/quadrature.jl; reference nodes are FastGaussQuadrature 0.4,
# not vendored: restated from the published definitions)
# ---------------------------------------------------------------------------

def num_quadrature_points(k, kp=3):
    """src/quadrature.jl:38-41."""
    return int(lib().cbo_num_quadrature_points(k, kp))


def gausslegendre(n):
    """Gauss-Legendre nodes/weights on [-1,1] (gausslegendre(n) of
    FastGaussQuadrature 0.4): Golub-Welsch start + Newton polish in long double."""
    x, _ = np.polynomial.legendre.leggauss(n)
    x = x.astype(np.longdouble)
    for _ in range(3):
        p0, p1 = np.ones_like(x), x.copy()
        for m in range(2, n + 1):
            p0, p1 = p1, ((2 * m - 1) * x * p1 - (m - 1) * p0) / m
        dp_ = n * (x * p1 - p0) / (x * x - 1)
        x = x - p1 / dp_
    p0, p1 = np.ones_like(x), x.copy()
    for m in range(2, n + 1):
        p0, p1 = p1, ((2 * m - 1) * x * p1 - (m - 1) * p0) / m
    dp_ = n * (x * p1 - p0) / (x * x - 1)
    w = 2 / ((1 - x * x) * dp_ * dp_)
    if n == 1:
        return np.array([0.0]), np.array([2.0])
    return np.asarray(x, dtype=np.float64), np.asarray(w, dtype=np.float64)


def gausslobatto(n):
    """Gauss-Lobatto nodes/weights on [-1,1] (gausslobatto(n)): endpoints plus
    the roots of P'_{n-1}; w = 2 / (n (n-1) P_{n-1}(x)^2)."""
    if n == 2:
        return np.array([-1.0, 1.0]), np.array([1.0, 1.0])
    N = n - 1
    # start: Chebyshev-Gauss-Lobatto, Newton on (1-x^2) P'_N
    x = -np.cos(np.pi * np.arange(n) / N).astype(np.longdouble)
    for _ in range(100):
        p0, p1 = np.ones_like(x), x.copy()
        for m in range(2, N + 1):
            p0, p1 = p1, ((2 * m - 1) * x * p1 - (m - 1) * p0) / m
        # p1 = P_N, p0 = P_{N-1};  q = (1-x^2)P'_N = N (P_{N-1} - x P_N), q' = -N(N+1) P_N
        q = N * (p0 - x * p1)
        dq = -N * (N + 1) * p1
        dx = q / dq
        dx[0] = dx[-1] = 0
        x = x - dx
        if np.max(np.abs(dx)) < 1e-19:
            break
    x[0], x[-1] = -1, 1
    p0, p1 = np.ones_like(x), x.copy()
    for m in range(2, N + 1):
        p0, p1 = p1, ((2 * m - 1) * x * p1 - (m - 1) * p0) / m
    w = 2 / (N * (N + 1) * p1 * p1)
    x = np.asarray(x, dtype=np.float64)
    x = 0.5 * (x - x[::-1])  # exact antisymmetry
    return x, np.asarray(w, dtype=np.float64)


def lgwt(t, q, xi=None, wr=None):
    """src/quadrature.jl:60-77: nodes/weights on every non-empty interval."""
    if xi is None:
        xi, wr = gausslegendre(q)
    t, tp = _d(t)
    xi, xip = _d(xi)
    wr, wrp = _d(wr)
    ni = len(nonempty_intervals(t))
    x = np.empty(ni * q)
    w = np.empty(ni * q)
    got = lib().cbo_lgwt(tp, i64(len(t)), xip, wrp, C.c_int(q), _p(x), _p(w))
    assert got == ni
    return x, w


# ---------------------------------------------------------------------------
# B-spline evaluation (src/bsplines.jl)
# ---------------------------------------------------------------------------

def deboor_unit(t, k, j, x, i, m=0):
    """deBoor(t, UnitVector(nf, j), x, i, m): src/bsplines.jl:159-202."""
    t, tp = _d(t)
    nf = len(t) - k
    return lib().cbo_deboor_unit(tp, len(t), k, nf, j, float(x), i or 0, m)


def deboor(t, k, c, x, i, m=0):
    t, tp = _d(t)
    c, cp = _d(c)
    return lib().cbo_deboor(tp, len(t), k, cp, len(c), float(x), i or 0, m)


def bspline_eval_ell(t, k, mr, x, m=0, threads=False):
    """Row form of B[x,:] (src/bsplines.jl:204-224): (interval[nx] int32,
    vals[nx,k])."""
    t, tp = _d(t)
    x, xp = _d(x)
    nx = len(x)
    iv = np.zeros(nx, dtype=np.int32)
    vals = np.zeros((nx, k))
    f = lib().cbo_bspline_eval_ell_mt if threads else lib().cbo_bspline_eval_ell
    f(tp, i64(len(t)), C.c_int(k), C.c_int(mr), xp, i64(nx), C.c_int(m), _p(iv, i32p), _p(vals))
    return iv, vals


def bspline_eval_csc(t, k, mr, x, col0=0, ncols=None):
    """B[x, col0+1:col0+ncols] as Julia's SparseMatrixCSC (1-based colptr,
    rowval; zeros dropped): src/bsplines.jl:217-224."""
    t, tp = _d(t)
    x, xp = _d(x)
    nf = len(t) - k
    ncols = nf - col0 if ncols is None else ncols
    nx = len(x)
    colptr = np.zeros(ncols + 1, dtype=np.int64)
    cap = max(1, nx * min(k, max(ncols, 1)))
    rowval = np.zeros(cap, dtype=np.int64)
    nzval = np.zeros(cap)
    nnz = lib().cbo_bspline_eval_csc(tp, i64(len(t)), C.c_int(k), C.c_int(mr), xp, i64(nx),
                                     i64(col0), i64(ncols), _p(colptr, ip), _p(rowval, ip), _p(nzval))
    return colptr, rowval[:nnz].copy(), nzval[:nnz].copy()


def bspline_eval_literal(tt, k, ml, mr, x, col0=0, ncols=None):
    """The reference's literal O(nf*k*nx) scan, dense nx x ncols result."""
    tt, ttp = _d(tt)
    x, xp = _d(x)
    nf = len(tt) + ml + mr - 2 - k
    ncols = nf - col0 if ncols is None else ncols
    chi = np.zeros((ncols, len(x)))
    lib().cbo_bspline_eval_literal(ttp, i64(len(tt)), C.c_int(k), C.c_int(ml), C.c_int(mr), xp,
                                   i64(len(x)), i64(col0), i64(ncols), _p(chi))
    return chi.T


def basis_table(t, k, x, q, m=0):
    """basis_functions(t, x, m) (src/bsplines.jl:28-52) as [ni, q, k] tables."""
    t, tp = _d(t)
    x, xp = _d(x)
    ni = len(x) // q
    tab = np.zeros((ni, q, k))
    lib().cbo_basis_table(tp, i64(len(t)), C.c_int(k), xp, C.c_int(q), C.c_int(m), _p(tab))
    return tab


def basis_matrix(t, k, x, q, m=0):
    """Dense (ni*q) x nf version of basis_functions(t, x, m)."""
    tab = basis_table(t, k, x, q, m)
    nf = len(t) - k
    nei = nonempty_intervals(t)
    B = np.zeros((len(x), nf))
    for c, i in enumerate(nei):
        for a in range(k):
            j = i - k + 1 + a
            if 1 <= j <= nf:
                B[c * q:(c + 1) * q, j - 1] = tab[c, :, a]
    return B


# ---------------------------------------------------------------------------
# B-spline operator assembly (src/bsplines.jl:58-93,258-324,
# src/bspline_derivatives.jl)
# ---------------------------------------------------------------------------

def band_shape(kL, kR, rowL0, colR0):
    """(l,u) of Matrix(undef, A, B): src/bsplines.jl:258-270."""
    k = max(kL, kR)
    ij = colR0 - rowL0
    return k - 1 + ij, k - 1 - ij


def bspline_assemble(tL, kL, tR, kR, x, w, q, a=0, b=0, op=None,
                     rowL0=0, m=None, colR0=0, n=None):
    """overlap_matrix!/diffop! into BandedMatrix data; returns (band, l, u)."""
    tL, tLp = _d(tL)
    tR, tRp = _d(tR)
    x, xp = _d(x)
    w, wp = _d(w)
    nfL, nfR = len(tL) - kL, len(tR) - kR
    m = nfL - rowL0 if m is None else m
    n = nfR - colR0 if n is None else n
    l, u = band_shape(kL, kR, rowL0, colR0)
    band = np.zeros((n, l + u + 1))  # column-major (l+u+1) x n  == C-order [n][l+u+1]
    opp = None
    if op is not None:
        op, opp = _d(op)
    lib().cbo_bspline_assemble(tLp, i64(len(tL)), C.c_int(kL), tRp, i64(len(tR)), C.c_int(kR),
                               xp, wp, C.c_int(q), C.c_int(a), C.c_int(b), opp,
                               i64(rowL0), i64(m), i64(colR0), i64(n), C.c_int(l), C.c_int(u), _p(band))
    return band, l, u


def bspline_assemble_literal(tL, kL, tR, kR, x, w, q, a=0, b=0, op=None, rowL0=0, m=None, colR0=0, n=None):
    """The quadratic-cost literal form (every band entry scans all intervals, like the reference's sparse column
    dot products on its whole table); reduced sizes only."""
    tL, tLp = _d(tL)
    tR, tRp = _d(tR)
    x, xp = _d(x)
    w, wp = _d(w)
    nfL, nfR = len(tL) - kL, len(tR) - kR
    m = nfL - rowL0 if m is None else m
    n = nfR - colR0 if n is None else n
    l, u = band_shape(kL, kR, rowL0, colR0)
    band = np.zeros((n, l + u + 1))
    opp = None
    if op is not None:
        op, opp = _d(op)
    lib().cbo_bspline_assemble_literal(tLp, i64(len(tL)), C.c_int(kL), tRp, i64(len(tR)), C.c_int(kR),
                                       xp, wp, C.c_int(q), C.c_int(a), C.c_int(b), opp,
                                       i64(rowL0), i64(m), i64(colR0), i64(n), C.c_int(l), C.c_int(u), _p(band))
    return band, l, u


def band_to_dense(band, m, n, l, u):
    """BandedMatrix data[(u+i-j), j] -> dense m x n."""
    A = np.zeros((m, n))
    for j in range(n):
        for i in range(max(0, j - u), min(m, j + l + 1)):
            A[i, j] = band[j, u + i - j]
    return A


# ---------------------------------------------------------------------------
# FE-DVR (src/fedvr.jl, src/fedvr_operators.jl, src/fedvr_derivatives.jl)
# ---------------------------------------------------------------------------

class FEDVROracle:
    """Fields of the FEDVR struct, real case (src/fedvr.jl:3-53)."""

    def __init__(self, t, order):
        t = np.asarray(t, dtype=np.float64)
        nel = len(t) - 1
        order = np.full(nel, order, dtype=np.int64) if np.isscalar(order) else np.asarray(order, dtype=np.int64)
        assert len(order) == nel and np.all(order > 1)
        t0 = t[0]
        xs, ws = [], []
        for i in range(nel):
            # element_grid (src/quadrature.jl:81-85) with c = t0, change_interval! :18-23
            xi, wr = gausslobatto(int(order[i]))
            a, b = t[i] - t0, t[i + 1] - t0
            tt_ = (xi + 1) / 2
            xloc = np.array([np.float64(_fma(tv, b, _fma(-tv, a, a))) for tv in tt_])
            xs.append(t0 + xloc)
            ws.append((b - a) * wr / 2)
        ns = [1.0 / np.sqrt(w_) for w_ in ws]                      # :34
        for i in range(nel - 1):                                    # :35-37
            v = 1.0 / np.sqrt(ws[i][-1] + ws[i + 1][0])
            ns[i][-1] = v
            ns[i + 1][0] = v
        self.t, self.order, self.nel = t, order, nel
        self.x = np.concatenate([xs[0]] + [x_[1:] for x_ in xs[1:]])   # :39
        self.n = np.concatenate([ns[0]] + [n_[1:] for n_ in ns[1:]])   # :40
        self.w = ws
        self.wcat = np.concatenate(ws)
        self.estart = np.concatenate([[0], np.cumsum(order[:-1] - 1)]).astype(np.int64)  # elems :42-49 (0-based start)
        self.boff = np.concatenate([[0], np.cumsum(order[:-1] ** 2)]).astype(np.int64)
        self.N = len(self.x)

    def element_boundaries(self):
        """src/fedvr.jl:82 (1-based)."""
        return np.concatenate([[1], self.estart + self.order]).astype(np.int64)

    def elements(self, a, b):
        """src/fedvr.jl:85-93: 1-based element range covering functions a..b."""
        lo = self.estart + 1
        hi = self.estart + self.order
        i = np.nonzero((lo <= a) & (a <= hi))[0][0] + 1
        j = np.nonzero((lo <= b) & (b <= hi))[0][-1] + 1
        return i, j

    def block_orders(self, a=1, b=None):
        """src/fedvr.jl:121-139."""
        b = self.N if b is None else b
        elb = self.element_boundaries()
        e1, e2 = self.elements(a, b)
        els = list(range(e1, e2 + 1))
        if els[0] != self.nel and a == elb[els[0]]:          # elb[first(els)+1], 1-based
            els = els[1:]
        if els[-1] > els[0] and b == elb[els[-1] - 1]:       # elb[last(els)]
            els = els[:-1]
        s = a - elb[els[0] - 1]
        e = elb[els[-1]] - b
        o = [int(self.order[el - 1]) for el in els]
        return o, int(s), int(e)

    def block_structure(self, a=1, b=None):
        """src/fedvr.jl:141-157."""
        o, s, e = self.block_orders(a, b)
        if len(o) > 1:
            bs = []
            for oo in o[1:-1]:
                bs += [oo - 2, 1] if oo > 2 else [1]
            bs = [o[0] - 1 - s, 1] + bs
            bs.append(o[-1] - 1 - e)
            return bs
        return [o[0] - (s + e)]

    def block_bandwidths(self, a=1, b=None):
        """src/fedvr.jl:159-176."""
        rows = self.block_structure(a, b)
        nrows = len(rows)
        if nrows > 1:
            o, _, _ = self.block_orders(a, b)
            bws = [[2, 1] if oo > 2 else [1] for oo in o]
            l = [1] + [v for bw in bws[1:] for v in bw]
            u = [v for bw in bws[:-1] for v in reversed(bw)] + [1]
            if len(l) < nrows:
                l = l + [0]
            if len(u) < nrows:
                u = [0] + u
            return l, u
        return [0], [0]

    def blocks(self, nder):
        """All element matrices diff(B, nder, i) (src/fedvr_derivatives.jl:43-57),
        concatenated column-major."""
        out = np.zeros(int(np.sum(self.order ** 2)))
        o32 = self.order.astype(np.int32)
        lib().cbo_fedvr_blocks(_p(self.x), _p(self.wcat), _p(self.n), _p(self.estart, ip),
                               _p(o32, i32p), i64(self.nel), C.c_int(nder), _p(self.boff, ip), _p(out))
        return out

    def dense(self, blocks, a=1, b=None):
        """The matrix set_elements! builds (src/fedvr_operators.jl:40-154),
        restricted to functions a..b, as a dense array."""
        b = self.N if b is None else b
        n = b - a + 1
        A = np.zeros((n, n))
        o32 = self.order.astype(np.int32)
        blocks = np.ascontiguousarray(blocks)
        lib().cbo_fedvr_dense(_p(blocks), _p(self.boff, ip), _p(self.estart, ip), _p(o32, i32p),
                              i64(self.nel), i64(a), i64(b), _p(A))
        return A.T.copy()  # column-major -> numpy [i, j]

    def apply(self, blocks, X, alpha=1.0, beta=0.0, Y=None, threads=False):
        X = np.asfortranarray(X, dtype=np.float64)
        n, nvec = X.shape
        Y = np.zeros((n, nvec), order="F") if Y is None else np.asfortranarray(Y)
        o32 = self.order.astype(np.int32)
        blocks = np.ascontiguousarray(blocks)
        f = lib().cbo_fedvr_apply_mt if threads else lib().cbo_fedvr_apply
        f(_p(blocks), _p(self.boff, ip), _p(self.estart, ip), _p(o32, i32p), i64(self.nel), i64(n),
          _p(X), i64(n), i64(nvec), C.c_double(alpha), C.c_double(beta), _p(Y), i64(n))
        return Y


def _fma(a, b, c):
    """Correctly rounded fma via exact rationals (test sizes only)."""
    return float(Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c)))


# ---------------------------------------------------------------------------
# Finite differences (src/finite_differences.jl, src/fd_derivatives.jl)
# ---------------------------------------------------------------------------

def fd_cartesian(N):
    """alpha, beta, gamma, delta of FiniteDifferences: src/fd_derivatives.jl:12-15,154-158."""
    return np.ones(N - 1), np.ones(N), np.zeros(N), np.ones(N - 1)


def fd_staggered_uniform(N):
    """src/fd_derivatives.jl:20-25."""
    j = np.arange(1, N, dtype=np.float64)
    alpha = j ** 2 / (j ** 2 - 1.0 / 4)
    j = np.arange(1, N + 1, dtype=np.float64)
    beta = (j ** 2 - j + 1.0 / 2) / (j ** 2 - j + 1.0 / 4)
    return alpha, beta, np.zeros(N), alpha.copy()


def _staggered_common(fun, r, dN=0):
    """src/fd_derivatives.jl:27-45."""
    r = np.asarray(r, dtype=np.float64)
    N = len(r)
    v = np.zeros(N + dN)
    rt = np.concatenate([r, [2 * r[-1] - r[-2]]])
    a = 2 * r[0] - r[1]
    b, c, d = r[0], r[1], r[2]
    for j in range(1, N + dN + 1):
        if j < N:
            d = rt[j + 1]
        v[j - 1] = fun(a, b, c, d)
        a, b, c = b, c, d
    return v


def fd_staggered_nonuniform(r, f=None):
    """alpha, beta, gamma, delta for a non-uniform staggered grid, Krause &
    Schafer (A13)/(A14): src/fd_derivatives.jl:47-92.  `f` (optional callable)
    gives the weighted forms :54-60,71-81 (f is called on max(0, r))."""
    sq = np.sqrt
    if f is None:
        alpha = _staggered_common(lambda a, b, c, d: 2 / ((c - b) * sq((d - b) * (c - a))) * ((b + c) / 2) ** 2 / (b * c), r, -1)

        def bet(a, b, c, d):
            fp = ((c + b) / (2 * b)) ** 2
            fm = ((b + a) / (2 * b)) ** 2
            return 1 / (c - a) * (1 / (c - b) * fp + 1 / (b - a) * fm)
        beta = _staggered_common(bet, r)
    else:
        ff = lambda rr: f(max(0.0, rr))

        def alp(a, b, c, d):
            rh = (b + c) / 2
            return 2 * ff(rh) / ((c - b) * sq((d - b) * (c - a))) * rh ** 2 / (b * c)
        alpha = _staggered_common(alp, r, -1)

        def bet(a, b, c, d):
            bm2 = 1.0 / (b ** 2)
            rp = (b + c) / 2
            rm = (a + b) / 2
            fp = rp ** 2 * bm2
            fm = rm ** 2 * bm2
            return 1 / (c - a) * (ff(rp) / (c - b) * fp + ff(rm) / (b - a) * fm)
        beta = _staggered_common(bet, r)
    delta = _staggered_common(lambda a, b, c, d: 2 / (sq((d - b) * (c - a))) * ((b + c) / 2) ** 2 / (b * c), r, -1)
    return alpha, beta, np.zeros(len(r)), delta


def fd_gradient(delta, gamma, step):
    """Tridiagonal dl, d, du of R'DR: src/fd_derivatives.jl:233-237."""
    mu = 2 * step
    return -delta / mu, gamma / mu, delta / mu


def fd_laplacian(alpha, beta, step):
    """SymTridiagonal dv, ev of R'D'DR: src/fd_derivatives.jl:265-267."""
    d2 = step ** 2
    return -2 * beta / d2, alpha / d2


def staggered_local_step(r, j, uniform_step=None):
    """src/finite_differences.jl:278-288 (j 1-based)."""
    if uniform_step is not None:
        return uniform_step
    if j == 1:
        return r[1] - r[0]
    if j == len(r):
        return r[j - 1] - r[j - 2]
    return (r[j] - r[j - 2]) / 2


def laplacian_correction(dv, rho, Z, ell):
    """src/fd_derivatives.jl:542-551 (staggered): d2[1,1] -= 2*dbeta1."""
    dv = dv.copy()
    if ell == 0 and Z != 0:
        db = 1 / rho ** 2 * Z * rho / 8 * (1 + Z * rho)
        dv[0] -= 2 * db
    return dv


def tridiag_apply(dl, d, du, X, alpha=1.0, beta=0.0, Y=None, threads=False):
    X = np.asfortranarray(X, dtype=np.float64)
    n, nvec = X.shape
    Y = np.zeros((n, nvec), order="F") if Y is None else np.asfortranarray(Y)
    dl, d, du = (np.ascontiguousarray(v, dtype=np.float64) for v in (dl, d, du))
    f = lib().cbo_tridiag_apply_mt if threads else lib().cbo_tridiag_apply
    f(_p(dl), _p(d), _p(du), i64(n), _p(X), i64(n), i64(nvec), C.c_double(alpha), C.c_double(beta), _p(Y), i64(n))
    return Y


def banded_apply(band, m, n, l, u, X, alpha=1.0, beta=0.0, Y=None, threads=False):
    X = np.asfortranarray(X, dtype=np.float64)
    nvec = X.shape[1]
    Y = np.zeros((m, nvec), order="F") if Y is None else np.asfortranarray(Y)
    band = np.ascontiguousarray(band, dtype=np.float64)
    f = lib().cbo_banded_apply_mt if threads else lib().cbo_banded_apply
    f(_p(band), i64(m), i64(n), C.c_int(l), C.c_int(u), _p(X), i64(X.shape[0]), i64(nvec),
      C.c_double(alpha), C.c_double(beta), _p(Y), i64(m))
    return Y


# ---------------------------------------------------------------------------
# Densities (src/densities.jl)
# ---------------------------------------------------------------------------

def density_bspline(V, f, g, wvals=None, conj=True):
    """General (non-orthogonal) branch, src/densities.jl:92-99,156-164:
    rho = pinv(Matrix(RV)) * w * (conj(RV f) .* (LV g)); V = B[locs(B), sel]
    dense.  f, g are [nf] or [nf, P]; 1 x P broadcasting as :57-72."""
    Cm = np.linalg.pinv(V)
    if wvals is not None:
        Cm = Cm * np.asarray(wvals)[None, :]
    ft = V @ f
    if conj:
        ft = np.conj(ft)
    gt = V @ g
    if ft.ndim == 1 and gt.ndim == 2:
        ft = ft[:, None]
    if gt.ndim == 1 and ft.ndim == 2:
        gt = gt[:, None]
    return Cm @ (ft * gt)


def density_bspline_qr(t, k, x, q, f, g, wvals=None, col0=0, ncols=None):
    """The same projection at linear cost (full BASELINE sizes): the nodal product w .* (V f) .* (V g) followed by
    the least-squares solve V rho = b through a banded Givens QR (cbo_banded_lstsq) instead of the dense pinv of
    src/densities.jl:96.  t: expanded knots; f, g: [ncols, P] (or [ncols]); returns [ncols, P]."""
    tab = basis_table(t, k, x, q, 0)
    return density_from_table(tab, nonempty_intervals(t), k, q, f, g, wvals, col0, len(t) - k - col0 if ncols is None else ncols)


def density_from_table(tab, iv, k, q, f, g, wvals=None, col0=0, ncols=None):
    f = np.asfortranarray(np.atleast_2d(np.asarray(f, dtype=np.float64).T).T)
    g = np.asfortranarray(np.atleast_2d(np.asarray(g, dtype=np.float64).T).T)
    ncols = f.shape[0] if ncols is None else ncols
    iv = np.ascontiguousarray(iv, dtype=np.int64)
    ni = len(iv)
    Pf, Pg = f.shape[1], g.shape[1]
    P = max(Pf, Pg)
    nq = ni * q
    Bm = np.zeros((nq, P), order="F")
    wp = None
    if wvals is not None:
        wvals, wp = _d(wvals)
    tab = np.ascontiguousarray(tab, dtype=np.float64)
    lib().cbo_density_nodal(_p(tab), _p(iv, ip), i64(ni), C.c_int(q), C.c_int(k), i64(col0), i64(ncols), wp,
                            _p(f), i64(f.shape[0]), i64(Pf), _p(g), i64(g.shape[0]), i64(Pg), _p(Bm), i64(nq))
    return banded_lstsq(tab, iv, k, q, Bm, col0, ncols)


def banded_lstsq(tab, iv, k, q, Bm, col0, ncols):
    """argmin_c |V c - b| for every column b of Bm, V the (restricted) Vandermonde matrix at the nodes given by its
    [ni, q, k] table: ``B \\ f`` (src/bsplines.jl:329-343) and pinv(V) b (src/densities.jl:96) at linear cost."""
    tab = np.ascontiguousarray(tab, dtype=np.float64)
    iv = np.ascontiguousarray(iv, dtype=np.int64)
    Bm = np.asfortranarray(np.atleast_2d(np.asarray(Bm, dtype=np.float64).T).T).copy(order="F")
    nq, P = Bm.shape
    out = np.zeros((ncols, P), order="F")
    fn = lib().cbo_banded_lstsq_mt if int(lib().cbo_get_threads()) > 1 else lib().cbo_banded_lstsq
    fn(_p(tab), _p(iv, ip), i64(len(iv)), C.c_int(q), C.c_int(k), i64(col0), i64(ncols),
       _p(Bm), i64(nq), i64(P), _p(out), i64(ncols))
    return out


def set_threads(n: int):
    """Threads for the table / assembly / *_mt loops (1 = single-threaded like the reference)."""
    lib().cbo_set_threads(C.c_int(int(n)))


def max_threads() -> int:
    return int(lib().cbo_max_threads())


def density_orthogonal(f, g, c=None, conj=True):
    """Orthogonal branches, src/densities.jl:166-185: rho = C .* (conj(f) .* g)
    with C = RV*w (diagonal) or I."""
    fg = (np.conj(f) if conj else f)
    if fg.ndim == 1 and np.ndim(g) == 2:
        fg = fg[:, None]
    if np.ndim(g) == 1 and fg.ndim == 2:
        g = g[:, None]
    out = fg * g
    if c is not None:
        out = (c[:, None] if out.ndim == 2 else c) * out
    return out


# ---------------------------------------------------------------------------
# Function interpolation and the metric inverse ("next" rows)
# ---------------------------------------------------------------------------

def bspline_ldiv(V, fvals):
    """``B \\ f`` for B-splines, src/bsplines.jl:329-343: ``B.B \\ getindex.(Ref(f), B.x)`` — the sparse
    least-squares solve of V c = f(x) at the quadrature nodes (restricted bases: V = B[x, :] over all
    nodes of the parent).  V dense [nquad, n]; fvals [nquad] or [nquad, P]."""
    return np.linalg.lstsq(V, fvals, rcond=None)[0]


def linear_operator_apply(A, S, X, alpha=1.0, beta=0.0, Y=None):
    """``mul!(y, LinearOperator(A, factorize(S)), x, alpha, beta)``, src/linear_operators.jl:38-48:
    temp = alpha*A*x, then y = S \\ temp (beta == 0) or y = beta*y + S \\ temp.  A, S dense."""
    t = np.linalg.solve(S, alpha * (A @ X))
    return t if beta == 0.0 or Y is None else beta * Y + t


# ---------------------------------------------------------------------------
# "Next" rows: DiagonalOperator and ShiftAndInvert (src/linear_operators.jl:85-236)
# ---------------------------------------------------------------------------

def diagonal_operator_apply(density, diag, X, alpha=1.0, beta=0.0, Y=None):
    """mul!(y, L::DiagonalOperator, x, alpha, beta), src/linear_operators.jl:129-138:
    copyto!(L.rho, L.diag, x) with the FunctionProduct{false} of the basis (``density(diag, x)`` is one of
    density_bspline / density_orthogonal with conj=False), then y = beta*y + alpha*rho."""
    rho = density(diag, X)
    if beta == 0.0:
        return alpha * rho
    return beta * np.asarray(Y) + alpha * rho


def shift_invert_apply(A, B, sigma, X):
    """mul!(y, M::ShiftAndInvert, x), src/linear_operators.jl:229-236: temp = B x; y = factorize(A - sigma*B) \\ temp,
    on dense copies of the operators (B = None: identity metric)."""
    A = np.asarray(A, dtype=np.float64)
    n = A.shape[0]
    Bm = np.eye(n) if B is None else np.asarray(B, dtype=np.float64)
    return np.linalg.solve(A - sigma * Bm, Bm @ np.asarray(X, dtype=np.float64))


def implicit_lhs_uniform(n, difforder):
    """implicit_lhs(::Uniform, B, difforder), src/fd_derivatives.jl:402-424: SymTridiagonal (dv, ev) of the compact
    scheme, M1 = (4, 1)/6 (Eq. 19), M2 = (10, 1)/12 (Eq. 13)."""
    dv, ev = (4.0 / 6.0, 1.0 / 6.0) if difforder == 1 else (10.0 / 12.0, 1.0 / 12.0)
    return np.full(n, dv), np.full(max(n - 1, 0), ev)


def implicit_derivative_apply(dl, d, du, mdv, mev, c, X, alpha=1.0, beta=0.0, Y=None):
    """mul!(y, d::ImplicitDerivative, x, alpha, beta), src/finite_differences.jl:388-396:
    mul!(y, Delta, x, alpha, beta); ldiv!(M^-1, y); lmul!(c, y) — on dense copies."""
    Delta = np.diag(d) + np.diag(dl, -1) + np.diag(du, 1)
    M = np.diag(mdv) + np.diag(mev, -1) + np.diag(mev, 1)
    rhs = alpha * (Delta @ np.asarray(X, dtype=np.float64))
    if beta != 0.0:
        rhs = rhs + beta * np.asarray(Y, dtype=np.float64)
    return c * np.linalg.solve(M, rhs)


def bandlu_factor(band, n, l, u):
    """factorize(A) of a general BandedMatrix = LAPACK dgbtrf (BandedMatrices 0.16 -> LAPACK, un-vendored; this is the
    routine itself, through SciPy).  band: BandedMatrices data as (n, l+u+1) rows [column j -> band[j, u + i - j]].
    Returns (ab, ipiv, info) with ab the (2l+u+1) x n LAPACK band array and 0-based pivots."""
    from scipy.linalg import lapack
    band = np.asarray(band, dtype=np.float64)
    ab = np.zeros((2 * l + u + 1, n), order="F")
    ab[l:, :] = band.T
    lu, piv, info = lapack.dgbtrf(ab, l, u)
    return lu, piv, info


def bandlu_solve(band, n, l, u, X):
    """ldiv!(factorize(A), X): LAPACK dgbtrf + dgbtrs."""
    from scipy.linalg import lapack
    lu, piv, info = bandlu_factor(band, n, l, u)
    if info != 0:
        raise np.linalg.LinAlgError(f"singular band (info = {info})")
    x, info = lapack.dgbtrs(lu, l, u, np.asfortranarray(X, dtype=np.float64), piv)
    return x


def inner_products(U, V, S=None, scale=1.0):
    """Pairwise u_p' S v_p (src/inner_products.jl:22-46; S = None: orthogonal bases, scale = step(B) for finite
    differences, :3-4) on dense copies."""
    U, V = np.asarray(U, dtype=np.float64), np.asarray(V, dtype=np.float64)
    W = V if S is None else np.asarray(S) @ V
    return scale * np.einsum("ip,ip->p", U, W)


def implicit_stencil_common(r, fun, s=1, dN=0):
    """``implicit_stencil_common(fun, B, s, dN)`` (src/fd_derivatives.jl:105-125), the reference's own loop: ``fun(a, b, c)``
    of the three consecutive local steps around every node, the grid continued by one mirrored point at the end and
    starting from the origin."""
    r = np.asarray(r, dtype=np.float64)
    N = r.size
    v = np.zeros(N + dN)
    rt = np.concatenate([r[s - 1:], [2 * r[-1] - r[-2]]])
    a = rt[0] - (r[s - 2] if s > 1 else 0.0)
    b = rt[1] - rt[0]
    c = 0.0
    for j in range(1, N + dN + 1):
        if j < N - (s - 1):
            c = rt[j + 1] - rt[j]
        v[j - 1] = fun(a, b, c)
        a, b = b, c
    return v


IMPLICIT_STENCIL_FUNS = {   # name -> (fun, s, dN): src/fd_derivatives.jl:127-143 (Eq. 33) and implicit_lhs :410-424
    "alpha_t": (lambda a, b, c: 2 / (a * (a + b)), 2, -1), "alpha": (lambda a, b, c: 2 / (b * (a + b)), 1, -1),
    "beta": (lambda a, b, c: 1 / (a * b), 1, 0),
    "m_dl": (lambda a, b, c: (a * a + a * b - b * b) / (6 * a * (a + b)), 2, -1),
    "m_d": (lambda a, b, c: (a * a + 3 * a * b + b * b) / (6 * a * b), 1, 0),
    "m_du": (lambda a, b, c: (-a * a + a * b + b * b) / (6 * b * (a + b)), 1, -1)}


def implicit_stencils_loop(r):
    """All six stencil vectors by the reference's loop."""
    return {k: implicit_stencil_common(r, f, s, dn) for k, (f, s, dn) in IMPLICIT_STENCIL_FUNS.items()}


def implicit_stencils_nonuniform(r):
    """Patchkovskii's non-uniform compact stencils (src/fd_derivatives.jl:105-143 alpha~, alpha, beta of Eq. 33 and
    implicit_lhs :410-424) in closed form on the local steps h_j = r_j - r_{j-1}, r_0 = 0, r_{N+1} = 2 r_N - r_{N-1}."""
    r = np.asarray(r, dtype=np.float64)
    h = np.diff(np.concatenate([[0.0], r, [2 * r[-1] - r[-2]]]))      # h[j-1] = h_j, j = 1..N+1
    a, b = h[:-1], h[1:]                                              # node j: a = h_j, b = h_{j+1}
    out = {
        "beta": 1 / (a * b),
        "alpha": (2 / (b * (a + b)))[:-1],
        "m_d": (a * a + 3 * a * b + b * b) / (6 * a * b),
        "m_du": ((-a * a + a * b + b * b) / (6 * b * (a + b)))[:-1],
    }
    a2, b2 = h[1:-1], np.concatenate([h[2:-1], [h[-1]]])              # entry j: a = h_{j+1}, b = h_{j+2} (mirrored end)
    out["alpha_t"] = 2 / (a2 * (a2 + b2))
    out["m_dl"] = (a2 * a2 + a2 * b2 - b2 * b2) / (6 * a2 * (a2 + b2))
    return out


def tridiag_dense(dl, d, du):
    return np.diag(d) + np.diag(dl, -1) + np.diag(du, 1)


def implicit_derivative_apply_general(Delta, M, c, X, alpha=1.0, beta=0.0, Y=None):
    """mul!(y, d::ImplicitDerivative, x, alpha, beta) with dense Delta and M (src/finite_differences.jl:388-396)."""
    rhs = alpha * (Delta @ np.asarray(X, dtype=np.float64))
    if beta != 0.0:
        rhs = rhs + beta * np.asarray(Y, dtype=np.float64)
    return c * np.linalg.solve(M, rhs)


# ---- complex-scaled FE-DVR (src/fedvr.jl:9-53, src/fedvr_derivatives.jl:15-57), numpy restatement ------------------------
# PARITY UNPINNED: the reference holds no golden values for the complex-scaled operators (test/fedvr/complex_scaling.jl
# checks complex_rotate and real_locs only, src/fedvr_derivatives.jl:55 carries a "TODO Check if correct"); this is a
# literal transcription of the Julia loops, including the dropped rotation of the nodes (src/quadrature.jl:25-27 calls
# change_interval! without its gamma argument) and the conjugation inside LinearAlgebra.dot.
def fd_eval_dense(l, w, xq):
    """B[x, :] of a finite-difference basis as a dense matrix (getindex, src/finite_differences.jl:63-93): column j is
    weight_j * tent(x, l_{j-1}, l_j, l_{j+1}) on the points of [a, c], the neighbours mirrored beyond the ends."""
    l = np.asarray(l, dtype=np.float64)
    xq = np.asarray(xq, dtype=np.float64)
    n = l.size
    out = np.zeros((xq.size, n))
    clamp = lambda v: min(max(v, 0.0), 1.0)
    for j in range(n):
        b = l[j]
        a = l[j - 1] if j > 0 else 2 * b - l[j + 1]
        c = l[j + 1] if j < n - 1 else 2 * b - l[j - 1]
        wj = 1.0 if w is None else w[j]
        for p in np.nonzero((xq >= a) & (xq <= c))[0]:
            out[p, j] = wj * (1 - clamp((b - xq[p]) / (b - a)) - clamp((b - xq[p]) / (b - c)))
    return out


def fedvr_eval_dense(O, xq):
    """B[x, :] of an FE-DVR basis as a dense matrix, following getindex (src/fedvr.jl:211-222, 272-295): column by
    column over (element i, function m), chi = n_m prod_{j != m} (x - x_j) / (x_m - x_j) for the points in the closed
    element; a bridge function is written by both its elements (the right one last).  O: FEDVROracle."""
    xq = np.asarray(xq, dtype=np.float64)
    out = np.zeros((xq.size, O.N))
    for i in range(O.nel):
        o = int(O.order[i])
        xi = O.x[O.estart[i]:O.estart[i] + o]
        ni = O.n[O.estart[i]:O.estart[i] + o]
        inside = np.nonzero((xq >= O.t[i]) & (xq <= O.t[i + 1]))[0]
        for m in range(o):
            for p in inside:
                chi = ni[m]
                for j in range(o):
                    if j == m:
                        continue
                    chi = chi * ((xq[p] - xi[j]) / (xi[m] - xi[j]))
                out[p, O.estart[i] + m] = chi
    return out


def fedvr_complex_fields(t, order, t0, phi):
    t = np.asarray(t, dtype=np.float64)
    nel = len(t) - 1
    order = np.full(nel, order, dtype=np.int64) if np.isscalar(order) else np.asarray(order, dtype=np.int64)
    idx = np.nonzero(t >= t0)[0]
    if idx.size == 0:
        raise ValueError("Complex scaling starting point outside grid")
    i0 = int(idx[0]) + 1
    t0 = t[i0 - 1]
    eiphi = np.exp(1j * phi)
    xs, ws = [], []
    for i in range(nel):
        xi, wr = gausslobatto(int(order[i]))
        a, b = t[i] - t0, t[i + 1] - t0
        tt_ = (xi + 1) / 2
        xloc = np.array([np.float64(_fma(tv, b, _fma(-tv, a, a))) for tv in tt_])     # gamma = one(T): no rotation
        xs.append((t0 + xloc).astype(np.complex128))
        ws.append((b - a) * wr / 2)
    rot = lambda i, v: eiphi * v if i + 1 >= i0 else v + 0j
    ns = [1.0 / np.sqrt(rot(i, w_)) for i, w_ in enumerate(ws)]
    for i in range(nel - 1):
        v = 1.0 / np.sqrt(rot(i, ws[i][-1]) + rot(i + 1, ws[i + 1][0]))
        ns[i][-1] = v
        ns[i + 1][0] = v
    return dict(order=order, i0=i0, t0=t0, eiphi=eiphi, x=xs, w=ws, n=ns)


def fedvr_complex_element(xi, wi, ni, nder, rotated, eiphi):
    """diff(B, n, i) (src/fedvr_derivatives.jl:43-57) for one element with complex nodes xi, real weights wi, complex
    normalisation ni."""
    o = len(xi)
    Lp = np.zeros((o, o), dtype=np.complex128)
    for m in range(o):                                   # lagrangeder!, :15-31
        Lp[m, m] = ((1 if m == o - 1 else 0) - (1 if m == 0 else 0)) / (2 * wi[m])
        for mp in range(o):
            if mp == m:
                continue
            f = 1.0 + 0j
            for k in range(o):
                if k == m or k == mp:
                    continue
                f *= (xi[mp] - xi[k]) / (xi[m] - xi[k])
            Lp[m, mp] = f / (xi[m] - xi[mp])
    Lt = np.eye(o, dtype=np.complex128) if nder == 1 else -Lp
    D = np.zeros((o, o), dtype=np.complex128)
    for m in range(o):                                   # diff!, :33-41 (dot conjugates its first argument)
        for mp in range(o):
            D[m, mp] = ni[m] * ni[mp] * np.vdot(Lt[m, :], wi * Lp[mp, :])
    if rotated:
        D = D / np.sqrt(eiphi)
    return D


def fedvr_complex_dense(fields, nder):
    """The matrix set_elements! builds from the complex element blocks (bridge entries summed)."""
    order = fields["order"]
    N = int(np.sum(order - 1)) + 1
    A = np.zeros((N, N), dtype=np.complex128)
    start = 0
    blocks = []
    for i in range(len(order)):
        o = int(order[i])
        D = fedvr_complex_element(fields["x"][i], fields["w"][i], fields["n"][i], nder, i + 1 >= fields["i0"], fields["eiphi"])
        A[start:start + o, start:start + o] += D
        blocks.append(D)
        start += o - 1
    return A, blocks
