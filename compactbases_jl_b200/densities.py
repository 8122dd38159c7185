"""Mutual densities — host mirror of src/densities.jl over the CUDA C ABI.

    rho = Density(B, nlhs=P, nrhs=P)        # FunctionProduct{true}(B, B, Float64, P, P)
    rho.copyto(f, g)                        # copyto!(rho, f, g): rho.rho = coefficients of conj(f) g
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import DimensionMismatch
from .bsplines import _BSplineOrRestricted, diffop
from .lazy import QuasiDiagonal
from .operators import ColMajor, _overlaps, _ptr, _stream, axpby, to_device


class FunctionProduct:
    """``FunctionProduct{Conjugated}(L, R, T, nlhs, nrhs; w=one)`` (src/densities.jl:48-101).

    B-spline bases take the general branch (:92-99); the dense ``pinv(Matrix(RV))`` of the
    reference is replaced by a banded Cholesky factorisation of RV'RV held in the plan.
    FE-DVR and finite-difference bases are orthogonal (:73-91): C is ``I`` or diagonal."""

    def __init__(self, L, R=None, nlhs=None, nrhs=None, w=None, conjugated=True):
        R = L if R is None else R
        self.L, self.R, self.conjugated = L, R, conjugated
        rl, rr = nlhs or 1, nrhs or 1
        if not (rl == rr or rl == 1 or rr == 1):                         # :57-60
            raise DimensionMismatch(f"DimensionMismatch: Cannot broadcast {rl} LHS and {rr} RHS to common size")
        self.nrho = None if (rl == 1 and rr == 1) else max(rl, rr)
        self.handle = None
        self.c = None
        P = R.parent_basis
        self.device = P.device
        if isinstance(R, _BSplineOrRestricted):
            if not (isinstance(L, _BSplineOrRestricted) and L.parent_basis == P and (L.col0, L.ncols) == (R.col0, R.ncols)):
                raise ValueError("Can only multiply B-spline bases with identical knot sets and restrictions")
            wv = None
            if w is not None:
                wv = np.ascontiguousarray(w(P.x.cpu().numpy()), dtype=np.float64)
            h = C.c_void_p()
            with torch.cuda.device(self.device):
                _lib.call("cb_density_bspline_create", P.handle, R.col0, R.ncols,
                          None if wv is None else C.c_void_p(wv.ctypes.data), C.byref(h), _stream())
            self.handle = h
            self.n = R.ncols
        else:
            if L is not R and not (L.parent_basis == P and L.indices() == R.indices()):
                raise ValueError("Cannot multiply functions on different grids")
            self.n = R.size()
            i0 = R.indices()[0] - 1
            # C = w if the Vandermonde diagonal is all ones, else RV*w (:76-84)
            vd = P.vandermonde_diag()
            c = None if vd is None else np.asarray(vd, dtype=np.float64)[i0:i0 + self.n]
            if w is not None:
                wv = np.asarray(w(np.asarray(P.locs())[i0:i0 + self.n]), dtype=np.float64)
                c = wv if c is None else c * wv
            if c is not None:
                self.c = to_device(c, self.device)
        self.rho = None
        self.tmp = None

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.load().cb_density_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def set_refine(self, steps: int):
        """Corrected-semi-normal-equation refinement steps (0 = plain normal equations)."""
        if self.handle:
            _lib.call("cb_density_set_refine", self.handle, int(steps))

    def copyto(self, f, g, out: ColMajor | None = None) -> ColMajor:
        """``copyto!(rho, f, g)`` (src/densities.jl:156-185).  f, g: ColMajor coefficient batches
        (n x Pf, n x Pg; one of them may have a single column)."""
        f = f if isinstance(f, ColMajor) else ColMajor(f)
        g = g if isinstance(g, ColMajor) else ColMajor(g)
        if f.n != self.n or g.n != self.n:
            raise DimensionMismatch(f"DimensionMismatch: basis has {self.n} functions, got {f.n} and {g.n} rows")
        P = max(f.nvec, g.nvec)
        rho = ColMajor.zeros(self.n, P, self.device) if out is None else out
        if self.handle:
            _lib.call("cb_density_apply", self.handle, _ptr(f.data), f.ld, f.nvec, _ptr(g.data), g.ld, g.nvec,
                      int(self.conjugated), _ptr(rho.data), rho.ld, _stream())
        else:
            _lib.call("cb_density_orthogonal", _ptr(self.c), self.n, _ptr(f.data), f.ld, f.nvec,
                      _ptr(g.data), g.ld, g.nvec, int(self.conjugated), _ptr(rho.data), rho.ld, _stream())
        self.rho = rho
        self._last = (f, g)
        return rho

    def copyto_host(self, f: np.ndarray, g: np.ndarray, out: np.ndarray | None = None) -> np.ndarray:
        """Host-buffer form (``cb_density_apply_host``, B-spline branch): Fortran-ordered (n, P) numpy arrays (pinned
        for PCIe speed); the pair batch is streamed through the device in tiles."""
        if not self.handle:
            raise NotImplementedError("host-buffer form exists for the B-spline branch")
        f = np.asfortranarray(f.reshape(self.n, -1), dtype=np.float64) if not f.flags.f_contiguous else f.reshape(self.n, -1, order="F")
        g = np.asfortranarray(g.reshape(self.n, -1), dtype=np.float64) if not g.flags.f_contiguous else g.reshape(self.n, -1, order="F")
        P = max(f.shape[1], g.shape[1])
        out = np.empty((self.n, P), order="F") if out is None else out
        assert out.flags.f_contiguous and out.shape == (self.n, P)
        with torch.cuda.device(self.device):
            _lib.call("cb_density_apply_host", self.handle, C.c_void_p(f.ctypes.data), f.shape[0], f.shape[1],
                      C.c_void_p(g.ctypes.data), g.shape[0], g.shape[1], int(self.conjugated), C.c_void_p(out.ctypes.data),
                      out.shape[0], _stream())
        return out

    def matrix(self):
        """``copyto!(M, rho)`` (src/densities.jl:202-208): the operator of multiplication by the
        density, ``overlap_matrix!(M, L', R, Diagonal(rho.tmp))`` for B-splines."""
        if not self.handle:
            # copyto!(M::Diagonal, rho) = mul!(M.diag, rho.C, rho.rho) (:202-204): the nodal values of the product,
            # i.e. C .* rho.rho once more (C = Diagonal(B.n) for FE-DVR, 1/sqrt(h) on non-uniform grids, I otherwise)
            from .operators import Diagonal
            d = self.rho.data[0, :self.n]
            return Diagonal((self.c * d if self.c is not None else d).contiguous())
        f, g = self._last
        P = self.R.parent_basis
        nn = P.ni * P.q
        tmp = torch.empty(nn, dtype=torch.float64, device=self.device)
        _lib.call("cb_density_nodal_product", self.handle, _ptr(f.data), f.ld, 1, _ptr(g.data), g.ld, 1,
                  int(self.conjugated), _ptr(tmp), nn, _stream())
        self.tmp = tmp
        return diffop(self.L, self.R, 0, 0, [QuasiDiagonal(tmp)])


def Density(L, R=None, nlhs=None, nrhs=None, w=None):
    """``Density = FunctionProduct{true}`` (src/densities.jl:30)."""
    return FunctionProduct(L, R, nlhs, nrhs, w, conjugated=True)


class DiagonalOperator:
    """``DiagonalOperator(f)`` (src/linear_operators.jl:85-173): multiplication by the function ``f = R*c`` in
    coefficient space, ``y = alpha * (coefficients of f*x) + beta * y``, through the ``FunctionProduct{false}`` of the
    two expansions instead of an assembled matrix.  diag: the coefficients c (length n); nrhs: the number of vectors
    ``mul`` is applied to at once (the reference applies one)."""

    def __init__(self, R, diag, nrhs=None, w=None):
        self.R = R
        self.rho = FunctionProduct(R, R, 1, nrhs, w, conjugated=False)
        self.n = self.rho.n
        self.diag = None
        self.copyto(diag)

    @property
    def shape(self):
        return (self.n, self.n)

    def copyto(self, diag):
        """``copyto!(o, diag)`` (:121-124): represent multiplication by the function with coefficients diag."""
        d = diag if isinstance(diag, ColMajor) else ColMajor(to_device(np.asarray(diag, dtype=np.float64).reshape(1, -1), self.rho.device)
                                                             if not isinstance(diag, torch.Tensor) else diag.reshape(1, -1))
        if d.n != self.n or d.nvec != 1:
            raise DimensionMismatch(f"DimensionMismatch: the basis has {self.n} functions, diag has {d.n} x {d.nvec}")
        self.diag = d
        return self

    def mul(self, Y, X, alpha: float = 1.0, beta: float = 0.0):
        """``mul!(y, L, x, alpha, beta)`` (:129-138)."""
        Xc = X if isinstance(X, ColMajor) else ColMajor(X)
        Yc = Y if isinstance(Y, ColMajor) else ColMajor(Y)
        if Xc.n != self.n or Yc.n != self.n or Xc.nvec != Yc.nvec:
            raise DimensionMismatch(f"DimensionMismatch: operator of order {self.n}, X {Xc.n}x{Xc.nvec}, Y {Yc.n}x{Yc.nvec}")
        if alpha == 1.0 and beta == 0.0 and not _overlaps(Yc.data, Xc.data):
            self.rho.copyto(self.diag, Xc, out=Yc)                   # the density lands in y directly
        else:
            rho = self.rho.copyto(self.diag, Xc)
            axpby(Yc, rho, alpha, beta)
        return Y

    def __matmul__(self, X):
        Xc = X if isinstance(X, ColMajor) else ColMajor(X)
        Y = ColMajor.zeros(self.n, Xc.nvec, self.rho.device)
        self.mul(Y, Xc)
        return Y

    def matrix(self, x=None):
        r"""``Matrix(L[, x])`` (:143-173): the matrix of the operator, optionally weighted by the function with
        coefficients x (default ``R \ 1``): ``copyto!(L.rho, L.diag, x); copyto!(M, L.rho)``."""
        if x is None:
            x = self.R.ldiv(lambda r: np.ones_like(r))
        xc = x if isinstance(x, ColMajor) else ColMajor(to_device(np.asarray(x, dtype=np.float64).reshape(1, -1), self.rho.device)
                                                        if not isinstance(x, torch.Tensor) else x.reshape(1, -1))
        self.rho.copyto(self.diag, xc)
        return self.rho.matrix()
