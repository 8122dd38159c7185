"""B-spline bases — host mirror of src/bsplines.jl / src/bspline_derivatives.jl /
src/restricted_bases.jl (CompactBases.jl v0.3.1) over the CUDA C ABI.

    B = BSpline(LinearKnotSet(7, 0, 100, 1000))      # k' = 3 by default
    chi = B[x, :]                                    # SparseMatrixCSC on the device
    S = B.T @ B                                      # BandedMatrix
    T = B.T @ D.T @ D @ B                            # weak Laplacian
    V = B.T @ QuasiDiagonal(lambda r: -1/r) @ B
    Br = B[:, 2:B.nf-1]                              # restriction (Julia 1-based, inclusive)
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import DimensionMismatch
from .knot_sets import AbstractKnotSet
from .lazy import AdjointBasis, Derivative, QuasiDiagonal
from .operators import BandedMatrix, ColMajor, Tridiagonal, _ptr, _stream, device_of, to_device
from .quadrature import gausslegendre, num_quadrature_points


class SparseMatrixCSC:
    """Julia's ``SparseMatrixCSC{Float64,Int64}``: 1-based ``colptr``/``rowval`` (device tensors)."""

    def __init__(self, m, n, colptr, rowval, nzval):
        self.m, self.n, self.colptr, self.rowval, self.nzval = m, n, colptr, rowval, nzval

    @property
    def shape(self):
        return (self.m, self.n)

    def nnz(self):
        return int(self.nzval.numel())

    def toarray(self) -> np.ndarray:
        cp, rv, nz = self.colptr.cpu().numpy(), self.rowval.cpu().numpy(), self.nzval.cpu().numpy()
        A = np.zeros((self.m, self.n))
        for j in range(self.n):
            sl = slice(cp[j] - 1, cp[j + 1] - 1)
            A[rv[sl] - 1, j] = nz[sl]
        return A


class CoulombPotential(QuasiDiagonal):
    """``QuasiDiagonal(-Z ./ r)`` evaluated on chip (no array of operator values is read)."""

    def __init__(self, Z=1.0):
        super().__init__(lambda r: -float(Z) / r)
        self.Z = float(Z)


class _BSplineOrRestricted:
    is_basis = True

    @property
    def T(self):
        return AdjointBasis(self)

    # -- src/restricted_bases.jl:22-27 ------------------------------------
    def indices(self):
        return (self.col0 + 1, self.col0 + self.ncols)

    def size(self):
        return self.ncols

    def order(self):
        return self.parent_basis.k

    def locs(self):
        return self.parent_basis.x

    def weights(self):
        return self.parent_basis.w

    # -- evaluation: getindex (src/bsplines.jl:204-236) --------------------
    def __getitem__(self, key):
        x, sel = key
        if isinstance(x, slice) and x == slice(None):
            # B[:, a:b] (Julia indices, inclusive) -> restriction
            if isinstance(sel, slice):
                a = 1 if sel.start is None else sel.start
                b = self.ncols if sel.stop is None else sel.stop
            else:
                a, b = sel
            return self.restrict(a, b)
        if isinstance(sel, slice) and sel == slice(None):
            col0, ncols = self.col0, self.ncols
        elif isinstance(sel, slice):
            a = 1 if sel.start is None else sel.start
            b = self.ncols if sel.stop is None else sel.stop
            col0, ncols = self.col0 + a - 1, b - a + 1
        else:                                  # B[x, j]
            col0, ncols = self.col0 + int(sel) - 1, 1
        if np.isscalar(x):
            row = self.parent_basis.eval_csc(np.array([float(x)]), col0, ncols).toarray()[0]
            return float(row[0]) if not isinstance(sel, slice) else row
        return self.parent_basis.eval_csc(x, col0, ncols)

    def restrict(self, a: int, b: int):
        """``B[:, a:b]`` relative to this basis' own function numbering."""
        if not (1 <= a <= b <= self.ncols):
            raise IndexError(f"BoundsError: restriction {a}:{b} outside 1:{self.ncols}")
        return RestrictedBSpline(self.parent_basis, self.col0 + a - 1, b - a + 1)

    # -- operator materialisation (the @materialize bodies) ----------------
    def _materialize(self, L, middle):
        """``copyto!(similar(applied), applied)`` for ``L' * middle... * self``."""
        if not isinstance(L, _BSplineOrRestricted):
            raise ValueError("Can only multiply B-spline bases with B-spline bases")
        a = b = 0
        ops = []
        mid = list(middle)
        if mid and isinstance(mid[0], Derivative) and mid[0].adjoint:
            a = 1                               # D' on the left basis
            mid = mid[1:]
        if mid and isinstance(mid[-1], Derivative) and not mid[-1].adjoint:
            b = 1
            mid = mid[:-1]
        if a == 0 and b == 0 and middle and not all(isinstance(f, QuasiDiagonal) for f in middle):
            raise NotImplementedError(f"no materialiser for {middle}")
        for f in mid:
            if not isinstance(f, QuasiDiagonal):
                raise NotImplementedError(f"no materialiser for {middle}")
            ops.append(f)
        if a == 1 and b == 0:
            raise NotImplementedError("B'*D'*B is not defined by the reference; use B'*D*B")
        return diffop(L, self, a, b, ops)

    # -- fused spline evaluation (B*c)[x] -----------------------------------
    def evaluate(self, x, coeffs, m=0):
        """``(B*C)[x]`` for a coefficient batch C (ColMajor, ncols x nvec) -> ColMajor (nx x nvec).
        Only for an unrestricted basis (restricted: pad the coefficients)."""
        P = self.parent_basis
        if self.col0 != 0 or self.ncols != P.nf:
            raise NotImplementedError("evaluate() needs the unrestricted basis")
        xd = to_device(x, P.device)
        Cm = coeffs if isinstance(coeffs, ColMajor) else ColMajor(coeffs)
        if Cm.n != P.nf:
            raise DimensionMismatch(f"DimensionMismatch: basis has {P.nf} functions, coefficients have {Cm.n} rows")
        out = ColMajor.zeros(xd.numel(), Cm.nvec, P.device)
        _lib.call("cb_bspline_eval_spline", P.handle, _ptr(xd), xd.numel(), m, _ptr(Cm.data), Cm.ld, Cm.nvec,
                  _ptr(out.data), out.ld, _stream())
        return out


    # -- function interpolation B \ f (src/bsplines.jl:329-343) -------------
    def ldiv(self, f) -> ColMajor:
        r"""``B \ f``: least-squares fit of f at the quadrature nodes of the parent basis (the reference
        solves ``B.B \ f.(B.x)``, restricted: ``B[x,:] \ f.(x)``).  f: a callable, or its nodal values as a
        tensor of shape (P, nquad) / (nquad,).  Returns the ncols x P coefficients."""
        P = self.parent_basis
        if callable(f):
            vals = P.nodal_values(f).reshape(1, -1)
        else:
            vals = f if isinstance(f, torch.Tensor) else to_device(np.atleast_2d(np.asarray(f, dtype=np.float64)), P.device)
            vals = vals.reshape(1, -1) if vals.dim() == 1 else vals
        nq = P.ni * P.q
        if vals.shape[1] != nq:
            raise DimensionMismatch(f"DimensionMismatch: the function must be given at the {nq} quadrature nodes")
        vals = vals.to(P.device, torch.float64).contiguous()
        plan = getattr(self, "_ldiv_plan", None)
        if plan is None:
            from .densities import FunctionProduct
            plan = self._ldiv_plan = FunctionProduct(self)
        out = ColMajor.zeros(self.ncols, vals.shape[0], P.device)
        _lib.call("cb_density_project", plan.handle, _ptr(vals), vals.shape[1], vals.shape[0], _ptr(out.data), out.ld, _stream())
        return out


def band_shape(L, R):
    """(l, u) of ``Matrix(undef, A, B)`` (src/bsplines.jl:264-269)."""
    k = max(L.order(), R.order())
    ij = (R.col0 + 1) - (L.col0 + 1)
    return k - 1 + ij, k - 1 - ij


def diffop(L, R, a: int, b: int, ops=()):
    """``diffop!(dest, L, R, a, b, op)`` (src/bspline_derivatives.jl:3-20) into a freshly
    allocated destination; a = b = 0 is ``overlap_matrix!`` (src/bsplines.jl:84-93).
    ``ops``: QuasiDiagonal factors multiplied together at the quadrature nodes."""
    PL, PR = L.parent_basis, R.parent_basis
    l, u = band_shape(L, R)
    m, n = L.ncols, R.ncols
    dest = BandedMatrix.undef(m, n, l, u, PR.device)
    opvals = None
    coulomb = None
    for f in ops:
        if isinstance(f, CoulombPotential) and len(ops) == 1:
            coulomb = f.Z
            break
        v = PR.nodal_values(f.diag)
        opvals = v if opvals is None else opvals * v
    if coulomb is not None:
        if a or b:
            raise NotImplementedError("CoulombPotential is a plain potential term")
        _lib.call("cb_bspline_assemble_coulomb", PL.handle, PR.handle, coulomb, L.col0, m, R.col0, n, l, u,
                  _ptr(dest.data), _stream())
    else:
        _lib.call("cb_bspline_assemble", PL.handle, PR.handle, a, b, _ptr(opvals), L.col0, m, R.col0, n, l, u,
                  _ptr(dest.data), _stream())
    if max(L.order(), R.order()) == 2 and m == n:   # src/bsplines.jl:258-262
        dl = torch.empty(max(n - 1, 0), dtype=torch.float64, device=PR.device)
        d = torch.empty(n, dtype=torch.float64, device=PR.device)
        du = torch.empty(max(n - 1, 0), dtype=torch.float64, device=PR.device)
        _lib.call("cb_band_to_tridiag", _ptr(dest.data), n, _ptr(dl), _ptr(d), _ptr(du), _stream())
        return Tridiagonal(dl, d, du)
    return dest


def hamiltonian_bands(R, op=None, out=None):
    """``(R'*QuasiDiagonal(op)*R, R'*D'*D*R)`` from one launch (``cb_bspline_assemble_pair``): what a radial
    Hamiltonian needs (test/bsplines/operators.jl:15-22, docs/src/bsplines/eigenproblems.md).  ``op``: None (mass
    matrix), a QuasiDiagonal / callable / nodal-value tensor, or a CoulombPotential (evaluated on chip).  The entries
    are bit-identical to ``R.T @ op @ R`` and ``R.T @ D.T @ D @ R``."""
    P = R.parent_basis
    k = R.order()
    n = R.ncols
    V, T = out if out is not None else (BandedMatrix.undef(n, n, k - 1, k - 1, P.device), BandedMatrix.undef(n, n, k - 1, k - 1, P.device))
    opvals, coulomb, Z = None, 0, 0.0
    if isinstance(op, CoulombPotential):
        coulomb, Z = 1, float(op.Z)
    elif op is not None:
        opvals = P.nodal_values(op.diag if isinstance(op, QuasiDiagonal) else op)
    _lib.call("cb_bspline_assemble_pair", P.handle, _ptr(opvals), coulomb, Z, R.col0, n, _ptr(V.data), _ptr(T.data), _stream())
    return V, T


def diffop_host(L, R, a: int, b: int, ops=()) -> np.ndarray:
    """``diffop!`` into a HOST BandedMatrix (``cb_bspline_assemble_host``): returns the band data as an (n, l+u+1)
    numpy array (column-major (l+u+1) x n).  ops: QuasiDiagonal factors (callables are evaluated on the host at
    ``locs(B)``) or one CoulombPotential."""
    PL, PR = L.parent_basis, R.parent_basis
    l, u = band_shape(L, R)
    m, n = L.ncols, R.ncols
    out = np.empty((n, l + u + 1), dtype=np.float64)
    opvals, coulomb, Z = None, 0, 0.0
    for f in ops:
        if isinstance(f, CoulombPotential) and len(ops) == 1:
            coulomb, Z = 1, f.Z
            break
        v = PR.nodal_values(f.diag).cpu().numpy()
        opvals = v if opvals is None else opvals * v
    if opvals is not None:
        opvals = np.ascontiguousarray(opvals, dtype=np.float64)
    _lib.call("cb_bspline_assemble_host", PL.handle, PR.handle, a, b,
              None if opvals is None else C.c_void_p(opvals.ctypes.data), coulomb, Z, L.col0, m, R.col0, n, l, u,
              C.c_void_p(out.ctypes.data), _stream())
    return out


class BSpline(_BSplineOrRestricted):
    """``BSpline(t[, N]; k'=3)`` (src/bsplines.jl:95-119): the basis on knot set ``t`` with N
    Gauss-Legendre points per interval.  Construction uploads the knots and runs ``lgwt`` on
    the device; the mass matrix ``B.S`` is assembled on first use."""

    def __init__(self, t: AbstractKnotSet, N: int | None = None, kp: int = 3, device=None):
        self.t = t
        self.k = t.k
        self.q = num_quadrature_points(t.k, kp) if N is None else int(N)
        self.device = device_of(device)
        xi, wr = gausslegendre(self.q)
        self._keep = (np.ascontiguousarray(t.t), xi, wr)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.call("cb_bspline_create", C.c_void_p(self._keep[0].ctypes.data), t.t.size, t.k, t.ml, t.mr,
                      C.c_void_p(xi.ctypes.data), C.c_void_p(wr.ctypes.data), self.q, C.byref(h))
        self.handle = h
        lib = _lib.load()
        self.nf = int(lib.cb_bspline_numfunctions(h))
        self.nt = int(lib.cb_bspline_numknots(h))
        self.ni = int(lib.cb_bspline_num_nonempty(h))
        self.parent_basis, self.col0, self.ncols = self, 0, self.nf
        self._x = self._w = self._S = None

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.load().cb_bspline_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def __eq__(self, other):
        return isinstance(other, BSpline) and self.t == other.t   # src/bsplines.jl:127

    def __hash__(self):
        return hash(self.t)

    def __repr__(self):
        return f"BSpline{{Float64}} basis with {self.t!r}"

    def _node_views(self):
        if self._x is None:
            n = self.ni * self.q
            self._x = torch.empty(n, dtype=torch.float64, device=self.device)
            self._w = torch.empty(n, dtype=torch.float64, device=self.device)
            _lib.call("cb_bspline_nodes_copy", self.handle, _ptr(self._x), _ptr(self._w), _stream())
        return self._x, self._w

    @property
    def x(self):
        return self._node_views()[0]

    @property
    def w(self):
        return self._node_views()[1]

    @property
    def S(self):
        """``metric(B)`` = B.S (src/bsplines.jl:292)."""
        if self._S is None:
            self._S = self.T @ self
        return self._S

    def nodal_values(self, f):
        """``getindex.(Ref(D.diag), locs(B))`` (src/bsplines.jl:305)."""
        if isinstance(f, torch.Tensor):
            if f.numel() != self.ni * self.q:
                raise DimensionMismatch("DimensionMismatch: operator values must be given at the quadrature nodes")
            return f.to(self.device, torch.float64)
        return to_device(np.asarray(f(self.x.cpu().numpy()), dtype=np.float64), self.device)

    # -- device evaluation entry points ------------------------------------
    def eval_ell(self, x, m=0):
        """Row form of ``B[x,:]``: (interval int32[nx], vals float64[nx, k])."""
        xd = to_device(x, self.device)
        nx = xd.numel()
        iv = torch.empty(nx, dtype=torch.int32, device=self.device)
        vals = torch.empty((nx, self.k), dtype=torch.float64, device=self.device)
        _lib.call("cb_bspline_eval", self.handle, _ptr(xd), nx, m, _ptr(iv), _ptr(vals), _stream())
        return iv, vals

    def eval_ell_host(self, x: np.ndarray, m=0):
        """Host-pointer form (``cb_bspline_eval_host``): numpy in, numpy out."""
        x = np.ascontiguousarray(x, dtype=np.float64)
        iv = np.empty(x.size, dtype=np.int32)
        vals = np.empty((x.size, self.k), dtype=np.float64)
        with torch.cuda.device(self.device):
            _lib.call("cb_bspline_eval_host", self.handle, C.c_void_p(x.ctypes.data), x.size, m, C.c_void_p(iv.ctypes.data),
                      C.c_void_p(vals.ctypes.data))
        return iv, vals

    def eval_csc(self, x, col0=0, ncols=None) -> SparseMatrixCSC:
        ncols = self.nf - col0 if ncols is None else ncols
        xd = to_device(x, self.device)
        nx = xd.numel()
        colptr = torch.empty(ncols + 1, dtype=torch.int64, device=self.device)
        cap = max(1, nx * self.k)
        rowval = torch.empty(cap, dtype=torch.int64, device=self.device)
        nzval = torch.empty(cap, dtype=torch.float64, device=self.device)
        nnz = C.c_int64(0)
        _lib.call("cb_bspline_eval_csc", self.handle, _ptr(xd), nx, col0, ncols, _ptr(colptr), _ptr(rowval),
                  _ptr(nzval), C.byref(nnz), _stream())
        return SparseMatrixCSC(nx, ncols, colptr, rowval[:nnz.value], nzval[:nnz.value])

    def basis_functions(self, m=0) -> torch.Tensor:
        """``basis_functions(t, x, m)`` (src/bsplines.jl:28-52) as a dense [ni, q, k] table."""
        tab = torch.empty((self.ni, self.q, self.k), dtype=torch.float64, device=self.device)
        _lib.call("cb_bspline_quad_table", self.handle, m, _ptr(tab), _stream())
        return tab


class RestrictedBSpline(_BSplineOrRestricted):
    """``B[:, a:b]`` (src/restricted_bases.jl): functions col0+1 .. col0+ncols of the parent."""

    def __init__(self, parent: BSpline, col0: int, ncols: int):
        self.parent_basis, self.col0, self.ncols = parent, col0, ncols

    def __eq__(self, other):
        return (isinstance(other, RestrictedBSpline) and self.parent_basis == other.parent_basis
                and (self.col0, self.ncols) == (other.col0, other.ncols))

    def __hash__(self):
        return hash((self.parent_basis, self.col0, self.ncols))
