"""Operator containers and ``mul`` (= Julia's ``mul!(Y, A, X, alpha, beta)``).

The reference materialises its operators into ``BandedMatrix``,
``BlockSkylineMatrix``, ``Tridiagonal``/``SymTridiagonal`` and ``Diagonal``
 and applies them with ``mul!`` (a17; call sites
src/linear_operators.jl:40,50-52, src/inner_products.jl:30,46).  These classes
hold the same storage in device memory (torch is only the allocator) and
``mul`` routes to the CUDA kernels behind the C ABI.

Coefficient batches follow Julia's layout: an ``n x nvec`` column-major matrix,
i.e. a torch tensor of shape ``(nvec, n)`` whose rows are the coefficient
vectors (``colmajor``/``from_colmajor`` convert from/to the (n, nvec) view).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import DimensionMismatch


def _ptr(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def device_of(dev=None) -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("compactbases_jl_b200 needs a CUDA device (no CPU fallback exists)")
    if dev is None:
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device(dev)


def to_device(a, dev=None, dtype=torch.float64) -> torch.Tensor:
    if isinstance(a, torch.Tensor):
        return a.to(device=device_of(dev), dtype=dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(device_of(dev))


class ColMajor:
    """An n x nvec column-major Float64 matrix on the device (Julia ``Matrix``).
    ``data`` has shape (nvec, ld) with ld >= n; column v is ``data[v, :n]``."""

    def __init__(self, data: torch.Tensor, n: int | None = None):
        if data.dim() == 1:
            data = data.unsqueeze(0)
        assert data.dim() == 2 and data.dtype == torch.float64 and data.is_cuda
        if data.shape[1] > 1 and data.stride(1) != 1:
            raise ValueError("coefficient vectors must be contiguous (column-major n x nvec)")
        if data.shape[1] <= 1:
            data = data.contiguous()
        self.data = data
        self.n = data.shape[1] if n is None else n
        self.nvec = data.shape[0]
        self.ld = data.stride(0) if self.nvec > 1 else max(data.shape[1], 1)

    @classmethod
    def zeros(cls, n, nvec, dev=None):
        return cls(torch.zeros((nvec, n), dtype=torch.float64, device=device_of(dev)))

    @classmethod
    def from_numpy(cls, a, dev=None):
        """a: (n,) or (n, nvec) numpy array (any order)."""
        a = np.asarray(a, dtype=np.float64)
        if a.ndim == 1:
            a = a[:, None]
        return cls(to_device(np.ascontiguousarray(a.T), dev))

    def numpy(self) -> np.ndarray:
        """(n, nvec) Fortran-ordered numpy array."""
        return np.asfortranarray(self.data[:, :self.n].cpu().numpy().T)


def _as_colmajor(X) -> ColMajor:
    return X if isinstance(X, ColMajor) else ColMajor(X)


def _overlaps(a: torch.Tensor, b: torch.Tensor) -> bool:
    """True if the storage ranges of two device tensors intersect (an output aliasing an input: the apply and density
    kernels read X while they write Y, whereas the reference goes through ``M.temp`` / ``rho.rho`` and tolerates
    ``y === x``)."""
    if a.numel() == 0 or b.numel() == 0 or a.device != b.device:
        return False
    def span(t):
        lo = t.data_ptr()
        ext = sum((sz - 1) * st for sz, st in zip(t.shape, t.stride())) + 1
        return lo, lo + ext * t.element_size()
    (a0, a1), (b0, b1) = span(a), span(b)
    return a0 < b1 and b0 < a1


class BandedMatrix:
    """BandedMatrices.jl storage: ``data[u + i - j, j]`` (0-based), data is (l+u+1) x n
    column-major == torch shape (n, l+u+1)."""

    def __init__(self, data: torch.Tensor, m: int, n: int, l: int, u: int):
        assert data.shape == (n, l + u + 1) and data.is_contiguous()
        self.data, self.m, self.n, self.l, self.u = data, m, n, l, u

    @classmethod
    def undef(cls, m, n, l, u, dev=None):
        return cls(torch.empty((n, l + u + 1), dtype=torch.float64, device=device_of(dev)), m, n, l, u)

    @property
    def shape(self):
        return (self.m, self.n)

    def bandwidths(self):
        """``bandwidths(A)`` as BandedMatrices reports them: clamped to the matrix size (a band that lies entirely
        outside an m x n matrix gives an empty ``bandrange``; test/fd/derivatives.jl:36-59 counts on it)."""
        return (min(self.l, self.m - 1), min(self.u, self.n - 1))

    def bandrange(self):
        l, u = self.bandwidths()
        return range(-l, u + 1)

    def dense(self) -> np.ndarray:
        band = self.data.cpu().numpy()
        A = np.zeros((self.m, self.n))
        for j in range(self.n):
            for i in range(max(0, j - self.u), min(self.m, j + self.l + 1)):
                A[i, j] = band[j, self.u + i - j]
        return A

    def _mul(self, Y, X, alpha, beta):
        _lib.call("cb_banded_apply", _ptr(self.data), self.m, self.n, self.l, self.u, _ptr(X.data), X.ld, X.nvec,
                  alpha, beta, _ptr(Y.data), Y.ld, _stream())

    def mul_host(self, Yh: np.ndarray, Xh: np.ndarray, alpha=1.0, beta=0.0):
        """Host-buffer entry (cb_banded_apply_host): Xh (n, nvec), Yh (m, nvec) Fortran-ordered numpy arrays; the
        operator arrays may still be in flight on the current stream (the entry orders itself after it)."""
        assert Xh.flags.f_contiguous and Yh.flags.f_contiguous
        _lib.call("cb_banded_apply_host", _ptr(self.data), self.m, self.n, self.l, self.u, C.c_void_p(Xh.ctypes.data),
                  Xh.shape[0], Xh.shape[1] if Xh.ndim > 1 else 1, alpha, beta, C.c_void_p(Yh.ctypes.data), Yh.shape[0],
                  _stream())
        return Yh

    def cholesky(self) -> "BandCholesky":
        """``factorize(S)`` / ``cholesky(S)`` of a symmetric positive definite band (the metric B'B)."""
        return BandCholesky(self)

    # -- the little band algebra the reference's drivers use (``T / -2``, ``A - sigma*B``); host glue on the
    #    band storage, the Julia side keeps BandedMatrices' own broadcasting for it ------------------------
    def scaled(self, a: float) -> "BandedMatrix":
        return BandedMatrix(self.data * float(a), self.m, self.n, self.l, self.u)

    def __mul__(self, a):
        return self.scaled(a)

    __rmul__ = __mul__

    def __truediv__(self, a):
        return self.scaled(1.0 / float(a))

    def __neg__(self):
        return self.scaled(-1.0)

    def widened(self, l: int, u: int) -> "BandedMatrix":
        """The same matrix stored with bandwidths (l, u) >= (self.l, self.u)."""
        if l < self.l or u < self.u:
            raise DimensionMismatch("DimensionMismatch: cannot narrow a band")
        if (l, u) == (self.l, self.u):
            return self
        data = torch.zeros((self.n, l + u + 1), dtype=torch.float64, device=self.data.device)
        data[:, u - self.u:u - self.u + self.l + self.u + 1] = self.data
        return BandedMatrix(data, self.m, self.n, l, u)

    def __sub__(self, other):
        return _band_axpy(self, other, -1.0)

    def __add__(self, other):
        return _band_axpy(self, other, 1.0)


class BandCholesky:
    """Cholesky factor of an SPD ``BandedMatrix`` held on the device (``cb_bandchol_*``); ``ldiv``
    is the batched ``ldiv!`` of the reference's ``LinearOperator`` (src/linear_operators.jl:38-48)."""

    def __init__(self, A: BandedMatrix, indefinite: bool = False):
        """indefinite=True: ``L D L'`` without pivoting of a symmetric indefinite band (``cb_bandldl_create``), same
        partitioned solve; raises ValueError on a pivot breakdown."""
        if A.m != A.n or A.l != A.u:
            raise DimensionMismatch("DimensionMismatch: cholesky needs a square band with equal bandwidths")
        h = C.c_void_p()
        with torch.cuda.device(A.data.device):
            _lib.call("cb_bandldl_create" if indefinite else "cb_bandchol_create", _ptr(A.data), A.n, A.l, C.byref(h), _stream())
        self.handle, self.n, self.device, self.indefinite = h, A.n, A.data.device, bool(indefinite)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.load().cb_bandchol_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def ldiv(self, X):
        """In place ``ldiv!(F, X)``: X <- A^-1 X for a ColMajor batch (or a (nvec, n) tensor)."""
        Xc = _as_colmajor(X)
        if Xc.n != self.n:
            raise DimensionMismatch(f"DimensionMismatch: factor has order {self.n}, X has {Xc.n} rows")
        _lib.call("cb_bandchol_solve", self.handle, _ptr(Xc.data), Xc.ld, Xc.nvec, _stream())
        return X


def as_banded(A) -> BandedMatrix:
    """BandedMatrix view of a Tridiagonal / SymTridiagonal / Diagonal / BandedMatrix operator."""
    if isinstance(A, BandedMatrix):
        return A
    if isinstance(A, Tridiagonal):
        n = A.d.shape[0]
        band = torch.zeros((n, 3), dtype=torch.float64, device=A.d.device)
        band[:, 1] = A.d
        band[1:, 0] = A.du                           # column j holds A[j-1, j] = du[j-1] in band row u + i - j = 0
        band[:-1, 2] = A.dl
        return BandedMatrix(band.contiguous(), n, n, 1, 1)
    if isinstance(A, Diagonal):
        n = A.diag.shape[0]
        return BandedMatrix(A.diag.reshape(n, 1).contiguous(), n, n, 0, 0)
    if isinstance(A, FEDVRMatrix):
        return A.banded()
    raise NotImplementedError(f"no band storage for {type(A).__name__}")


def _band_axpy(A: BandedMatrix, B, s: float) -> BandedMatrix:
    """A + s*B for band operators of equal shape (bandwidths are merged)."""
    B = as_banded(B)
    if (A.m, A.n) != (B.m, B.n):
        raise DimensionMismatch(f"DimensionMismatch: {A.shape} and {B.shape}")
    l, u = max(A.l, B.l), max(A.u, B.u)
    Aw, Bw = A.widened(l, u), B.widened(l, u)
    return BandedMatrix(Aw.data + float(s) * Bw.data, A.m, A.n, l, u)


class ShiftAndInvert:
    r"""``ShiftAndInvert(A, B_or_R, sigma)`` (src/linear_operators.jl:185-236): the operator
    ``x -> (A - sigma*B)^-1 B x`` of the generalised eigenproblem ``A x = lambda B x``, with B the operator metric of
    the basis R (``B'B`` for B-splines, ``I`` for FE-DVR / finite differences), a band matrix, or None (``I``).

    ``factorize(A - sigma*B)``: the banded Cholesky factorisation of the C ABI (``cb_bandchol_create``, partitioned
    batched solve) when the shifted operator is positive definite — the reference's own tests (kinetic operator, sigma
    below the spectrum; test/linear_operators.jl:72-88) — and otherwise the pivoted band LU ``cb_bandlu_create`` (what
    BandedMatrices' ``factorize`` calls: LAPACK gbtrf), e.g. for the interior shift of
    docs/src/bsplines/eigenproblems.md:96-110.  ``method`` = "auto" | "cholesky" | "lu".  ``mul`` is batched over the
    columns of X."""

    def __init__(self, A, B=None, sigma: float = 0.0, method: str = "auto"):
        from .bsplines import _BSplineOrRestricted
        metric = B
        if isinstance(B, _BSplineOrRestricted):                      # operator_metric(R) = R'R
            metric = B.S if hasattr(B, "S") else B.T @ B
        elif B is not None and not isinstance(B, (BandedMatrix, Tridiagonal, Diagonal)):
            metric = None                                            # FE-DVR, finite differences: I
        if isinstance(A, ImplicitDerivative):
            # factorize(d - sigma*I) = ImplicitFactorization(factorize(c*Delta - sigma*M), M), ldiv!: y = M x, then the
            # tridiagonal solve (src/finite_differences.jl:398-444)
            if metric is not None:
                raise NotImplementedError("ShiftAndInvert of an ImplicitDerivative with a non-trivial metric")
            shifted = _band_axpy(as_banded(A.Delta.scaled(A.c)), A.M, -sigma)
            self.Ainv = factorize(shifted, method)
            self.B = A.M
            self.n = A.n
            return
        Ab = as_banded(A)
        if Ab.m != Ab.n:
            raise DimensionMismatch("DimensionMismatch: ShiftAndInvert needs a square operator")
        if metric is None:
            shifted = Ab if sigma == 0.0 else _band_axpy(Ab, Diagonal(torch.ones(Ab.n, dtype=torch.float64, device=Ab.data.device)), -sigma)
        else:
            shifted = Ab if sigma == 0.0 else _band_axpy(Ab, metric, -sigma)
        self.Ainv = factorize(shifted, method)
        self.B = metric
        self.n = Ab.n

    @property
    def shape(self):
        return (self.n, self.n)

    def mul(self, Y, X):
        """``mul!(y, M, x)`` (src/linear_operators.jl:229-236): ``temp = B x``, ``y = A^-1 temp``."""
        Xc, Yc = _as_colmajor(X), _as_colmajor(Y)
        if self.B is None:
            if Yc.data.data_ptr() != Xc.data.data_ptr():
                Yc.data[:, :self.n].copy_(Xc.data[:, :self.n])
        elif _overlaps(Yc.data, Xc.data):                            # mul!(x, M, x): the reference's M.temp
            tmp = ColMajor(torch.empty((Xc.nvec, self.n), dtype=torch.float64, device=Xc.data.device))
            mul(tmp, self.B, Xc)
            Yc.data[:, :self.n].copy_(tmp.data)
        else:
            mul(Yc, self.B, Xc)
        self.Ainv.ldiv(Yc)
        return Y

    def __matmul__(self, X):
        Xc = _as_colmajor(X)
        Y = ColMajor(torch.empty((Xc.nvec, self.n), dtype=torch.float64, device=Xc.data.device))
        self.mul(Y, Xc)
        return Y


class BandLU:
    """``lu(A)`` / ``factorize(A)`` of a general ``BandedMatrix`` held on the device (``cb_bandlu_*``: LAPACK gbtrf /
    gbtrs semantics, partial pivoting); ``ldiv`` is the batched in-place solve."""

    def __init__(self, A: BandedMatrix):
        if A.m != A.n:
            raise DimensionMismatch("DimensionMismatch: lu needs a square band")
        h = C.c_void_p()
        with torch.cuda.device(A.data.device):
            _lib.call("cb_bandlu_create", _ptr(A.data), A.n, A.l, A.u, C.byref(h), _stream())
        self.handle, self.n, self.device = h, A.n, A.data.device

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                _lib.load().cb_bandlu_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def ldiv(self, X):
        Xc = _as_colmajor(X)
        if Xc.n != self.n:
            raise DimensionMismatch(f"DimensionMismatch: factor has order {self.n}, X has {Xc.n} rows")
        _lib.call("cb_bandlu_solve", self.handle, _ptr(Xc.data), Xc.ld, Xc.nvec, _stream())
        return X


def _band_is_symmetric(A: BandedMatrix) -> bool:
    """True if the band is symmetric to rounding (square, l == u, |A[j+d, j] - A[j, j+d]| <= 1e-13 max|A|: assembled
    operators sum their quadrature terms per entry, so the two triangles may differ in the last bits)."""
    if A.m != A.n or A.l != A.u:
        return False
    tol = 1e-13 * float(A.data.abs().max()) if A.data.numel() else 0.0
    for d in range(1, A.l + 1):
        if float((A.data[:A.n - d, A.u + d] - A.data[d:, A.u - d]).abs().max()) > tol:
            return False
    return True


def factorize(A: BandedMatrix, method: str = "auto"):
    """``factorize(A)``: "cholesky" (SPD bands, partitioned batched solve), "ldl" (symmetric indefinite, unpivoted
    L D L', same solve), "lu" (pivoted band LU, any non-singular band) or "auto" = for a symmetric band Cholesky, then
    L D L' if it passes a residual probe, else LU."""
    if method == "lu":
        return BandLU(A)
    if method == "cholesky":
        lu = max(A.l, A.u)
        return A.widened(lu, lu).cholesky()
    if method == "ldl":
        return BandCholesky(A, indefinite=True)
    if _band_is_symmetric(A):
        try:
            return A.cholesky()
        except ValueError:                                           # not positive definite
            pass
        try:
            # symmetric indefinite: L D L' without pivoting keeps the partitioned solve; there is no growth bound
            # without pivoting, so the factor is accepted only if it solves a probe system to working accuracy
            F = BandCholesky(A, indefinite=True)
            g = torch.Generator(device=A.data.device).manual_seed(1234)
            b = torch.randn((1, A.n), dtype=torch.float64, device=A.data.device, generator=g)
            x = ColMajor(b.clone())
            F.ldiv(x)
            r = ColMajor(torch.empty_like(b))
            mul(r, A, x)
            if float(torch.linalg.norm(r.data - b) / torch.linalg.norm(b)) <= 1e-9:
                return F
        except ValueError:                                           # pivot breakdown
            pass
    return BandLU(A)


class LinearOperator:
    """``LinearOperator(A, R)`` (src/linear_operators.jl:26-75): the action of the operator with matrix A in the
    basis R, i.e. ``y = S^-1 (A x)`` with the metric S = R'R for non-orthogonal bases (B-splines) and ``y = A x``
    for FE-DVR and finite differences (``operator_metric`` is ``I``, :4-17)."""

    def __init__(self, A, R=None):
        self.A = A
        self.Binv = None
        self.temp = None
        metric = None
        if R is not None:
            from .bsplines import _BSplineOrRestricted
            if isinstance(R, _BSplineOrRestricted):                  # operator_metric(B) = B'B
                metric = R.S if hasattr(R, "S") else R.T @ R
        if isinstance(metric, BandedMatrix):
            self.Binv = metric.cholesky()
        elif isinstance(metric, Tridiagonal):                        # k = 2 B-splines: the metric is SymTridiagonal
            n = metric.d.shape[0]
            band = torch.zeros((n, 3), dtype=torch.float64, device=metric.d.device)
            band[:, 1] = metric.d
            band[1:, 0] = metric.du
            band[:-1, 2] = metric.dl
            self.Binv = BandedMatrix(band.contiguous(), n, n, 1, 1).cholesky()

    @property
    def shape(self):
        return self.A.shape

    def mul(self, Y, X, alpha: float = 1.0, beta: float = 0.0):
        """``mul!(y, M, x, alpha, beta)`` (src/linear_operators.jl:38-52)."""
        if self.Binv is None:
            return mul(Y, self.A, X, alpha, beta)
        Xc, Yc = _as_colmajor(X), _as_colmajor(Y)
        if self.temp is None or self.temp.data.shape != Yc.data.shape or self.temp.data.device != Yc.data.device:
            self.temp = ColMajor(torch.empty_like(Yc.data))
        mul(self.temp, self.A, Xc, alpha, 0.0)
        self.Binv.ldiv(self.temp)
        if beta == 0.0:
            Yc.data.copy_(self.temp.data)
        else:
            Yc.data.mul_(beta).add_(self.temp.data)
        return Y

    def __matmul__(self, X):
        Xc = _as_colmajor(X)
        Y = ColMajor(torch.empty((Xc.nvec, self.shape[0]), dtype=torch.float64, device=Xc.data.device))
        self.mul(Y, Xc)
        return Y


class Tridiagonal:
    """LinearAlgebra.Tridiagonal(dl, d, du)."""

    def __init__(self, dl, d, du):
        self.dl, self.d, self.du = dl, d, du
        self.n = d.numel()

    @property
    def shape(self):
        return (self.n, self.n)

    def scaled(self, a: float) -> "Tridiagonal":
        a = float(a)
        return Tridiagonal(self.dl * a, self.d * a, self.du * a)

    def __mul__(self, a):
        return self.scaled(a)

    __rmul__ = __mul__

    def __truediv__(self, a):
        return self.scaled(1.0 / float(a))

    def __neg__(self):
        return self.scaled(-1.0)

    def dense(self):
        dl, d, du = (v.cpu().numpy() for v in (self.dl, self.d, self.du))
        return np.diag(d) + np.diag(dl, -1) + np.diag(du, 1)

    def _mul(self, Y, X, alpha, beta):
        _lib.call("cb_tridiag_apply", _ptr(self.dl), _ptr(self.d), _ptr(self.du), self.n, _ptr(X.data), X.ld, X.nvec,
                  alpha, beta, _ptr(Y.data), Y.ld, _stream())

    def mul_host(self, Yh: np.ndarray, Xh: np.ndarray, alpha=1.0, beta=0.0):
        """Host-buffer entry (cb_tridiag_apply_host): Xh, Yh are Fortran-ordered (n, nvec) numpy arrays."""
        assert Xh.flags.f_contiguous and Yh.flags.f_contiguous
        _lib.call("cb_tridiag_apply_host", _ptr(self.dl), _ptr(self.d), _ptr(self.du), self.n,
                  C.c_void_p(Xh.ctypes.data), Xh.shape[0], Xh.shape[1] if Xh.ndim > 1 else 1, alpha, beta,
                  C.c_void_p(Yh.ctypes.data), Yh.shape[0], _stream())
        return Yh


def axpby(Y, X, alpha: float = 1.0, beta: float = 0.0):
    """``Y .= alpha .* X .+ beta .* Y`` for ColMajor batches (``cb_axpby``)."""
    Xc, Yc = _as_colmajor(X), _as_colmajor(Y)
    if (Xc.n, Xc.nvec) != (Yc.n, Yc.nvec):
        raise DimensionMismatch(f"DimensionMismatch: X is {Xc.n}x{Xc.nvec}, Y is {Yc.n}x{Yc.nvec}")
    _lib.call("cb_axpby", Xc.n, Xc.nvec, float(alpha), _ptr(Xc.data), Xc.ld, float(beta), _ptr(Yc.data), Yc.ld, _stream())
    return Y


class SymTridiagonal(Tridiagonal):
    """LinearAlgebra.SymTridiagonal(dv, ev)."""

    def __init__(self, dv, ev):
        super().__init__(ev, dv, ev)
        self.dv, self.ev = dv, ev


class ImplicitDerivative:
    """``ImplicitDerivative(Delta, M, c)`` (src/finite_differences.jl:356-413): the compact finite-difference
    derivative ``c * M^-1 * Delta`` with tridiagonal Delta and M (uniform grids: M1 = SymTridiagonal(4, 1)/6,
    M2 = SymTridiagonal(10, 1)/12; non-uniform grids: Patchkovskii's non-symmetric Tridiagonal,
    src/fd_derivatives.jl:395-424).  ``factorize(M)`` is the banded Cholesky of the C ABI for the symmetric case and the
    pivoted band LU otherwise; ``mul!`` = tridiagonal apply + batched solve + scaling."""

    def __init__(self, Delta, M, c: float = 1.0, Minv=None):
        self.Delta, self.M, self.c = Delta, M, float(c)
        self.Minv = factorize(as_banded(M)) if Minv is None else Minv       # Cholesky (uniform grids) or band LU
        self.n = Delta.n

    @property
    def shape(self):
        return (self.n, self.n)

    def scaled(self, a: float) -> "ImplicitDerivative":           # Base.:(*)(d, a), :376-383
        return ImplicitDerivative(self.Delta, self.M, self.c * float(a), self.Minv)

    def __mul__(self, a):
        return self.scaled(a)

    __rmul__ = __mul__

    def __truediv__(self, a):
        return self.scaled(1.0 / float(a))

    def __neg__(self):
        return self.scaled(-1.0)

    def dense(self):
        return self.c * np.linalg.solve(self.M.dense(), self.Delta.dense())

    def _mul(self, Y, X, alpha, beta):
        """``mul!(y, d, x, alpha, beta)`` (:388-396): ``mul!(y, Delta, x, alpha, beta); ldiv!(M^-1, y); lmul!(c, y)``
        (beta*y passes through M^-1 as well, as in the reference)."""
        self.Delta._mul(Y, X, alpha, beta)
        self.Minv.ldiv(Y)
        if self.c != 1.0:
            axpby(Y, Y, self.c, 0.0)


class Diagonal:
    def __init__(self, diag):
        self.diag = diag
        self.n = diag.numel()

    @property
    def shape(self):
        return (self.n, self.n)

    def dense(self):
        return np.diag(self.diag.cpu().numpy())

    def _mul(self, Y, X, alpha, beta):
        _lib.call("cb_diag_apply", _ptr(self.diag), self.n, _ptr(X.data), X.ld, X.nvec, alpha, beta,
                  _ptr(Y.data), Y.ld, _stream())


class BlockSkylineMatrix:
    """BlockBandedMatrices' ``BlockSkylineMatrix(undef, rows, cols, (l, u))`` (the destination of
    ``Matrix(undef, B)`` for FE-DVR, src/fedvr_operators.jl:24-28): one flat ``data`` vector; block column J keeps
    the block rows ``max(1, J-u[J]) : min(J+l[J], N)`` stacked into one dense column-major panel.  ``block_starts``
    / ``block_strides`` follow BlockBandedMatrices' ``BlockSkylineSizes`` (0-based offsets here).  Host integer
    logic only; the data lives on the device."""

    def __init__(self, rows, l, u, dev=None):
        self.rows = [int(r) for r in rows]
        self.l, self.u = [int(v) for v in l], [int(v) for v in u]
        nb = len(self.rows)
        if len(self.l) != nb or len(self.u) != nb:
            raise DimensionMismatch("DimensionMismatch: one (l, u) pair per block column")
        self.cum = np.concatenate([[0], np.cumsum(self.rows)]).astype(np.int64)
        self.n = int(self.cum[-1])
        self.kfirst = np.array([max(0, J - self.u[J]) for J in range(nb)], dtype=np.int64)
        self.klast = np.array([min(nb - 1, J + self.l[J]) for J in range(nb)], dtype=np.int64)
        self.stride = (self.cum[self.klast + 1] - self.cum[self.kfirst]).astype(np.int64)     # bb_blockstrides
        sizes = self.stride * np.asarray(self.rows, dtype=np.int64)
        self.start = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)            # bb_blockstarts (first K)
        self.ndata = int(np.sum(sizes))
        self.data = torch.empty(self.ndata, dtype=torch.float64, device="cpu" if dev == "cpu" else device_of(dev))

    @property
    def shape(self):
        return (self.n, self.n)

    def block_start(self, K, J):
        """0-based offset in ``data`` of block (K, J) (0-based block indices) or None outside the skyline."""
        if K < self.kfirst[J] or K > self.klast[J]:
            return None
        return int(self.start[J] + self.cum[K] - self.cum[self.kfirst[J]])

    def block(self, K, J) -> np.ndarray:
        """``A[Block(K+1, J+1)]`` as a dense numpy array (zeros outside the skyline)."""
        r, c = self.rows[K], self.rows[J]
        st = self.block_start(K, J)
        if st is None:
            return np.zeros((r, c))
        d = self.data.cpu().numpy()
        sd = int(self.stride[J])
        return np.array([[d[st + i + j * sd] for j in range(c)] for i in range(r)])

    def dense(self) -> np.ndarray:
        A = np.zeros((self.n, self.n))
        d = self.data.cpu().numpy()
        for J in range(len(self.rows)):
            r0, r1 = int(self.cum[self.kfirst[J]]), int(self.cum[self.klast[J] + 1])
            sd = int(self.stride[J])
            for j in range(self.rows[J]):
                A[r0:r1, int(self.cum[J]) + j] = d[int(self.start[J]) + j * sd:int(self.start[J]) + j * sd + (r1 - r0)]
        return A


class FEDVRMatrix:
    """The operator that ``Matrix(undef, B)`` + ``set_elements!`` build for an FE-DVR basis
    (src/fedvr_operators.jl:17-30,148-154), kept as one dense o x o block per element;
    bridge corners are summed when the operator is applied or exported.  ``first..last``
    (1-based) is the restriction ``B[:, first:last]``."""

    def __init__(self, basis, blocks: torch.Tensor, first: int, last: int):
        self.basis, self.blocks, self.first, self.last = basis, blocks, first, last
        self.n = last - first + 1

    @property
    def shape(self):
        return (self.n, self.n)

    def block_structure(self):
        return self.basis._block_structure(self.first, self.last)

    def dense(self) -> np.ndarray:
        b = self.basis
        A = torch.empty((self.n, self.n), dtype=torch.float64, device=self.blocks.device)
        _lib.call("cb_fedvr_to_dense", _ptr(self.blocks), _ptr(b.d_estart), _ptr(b.d_order), _ptr(b.d_boff), b.nel,
                  self.first, self.last, _ptr(A), _stream())
        return A.cpu().numpy().T.copy()

    def scaled(self, a: float) -> "FEDVRMatrix":
        return FEDVRMatrix(self.basis, self.blocks * float(a), self.first, self.last)

    def __mul__(self, a):
        return self.scaled(a)

    __rmul__ = __mul__

    def __truediv__(self, a):
        return self.scaled(1.0 / float(a))

    def __neg__(self):
        return self.scaled(-1.0)

    def banded(self) -> "BandedMatrix":
        """The operator as a ``BandedMatrix`` with l = u = max order - 1 (``cb_fedvr_to_band``): an element couples only
        its own functions.  This is what ``factorize(A - sigma*I)`` of ``ShiftAndInvert`` works on for FE-DVR
        (src/linear_operators.jl:185-236; the reference factorises the BlockSkylineMatrix itself)."""
        b = self.basis
        w = max(int(b.order.max()) - 1, 1)
        A = BandedMatrix.undef(self.n, self.n, w, w, self.blocks.device)
        _lib.call("cb_fedvr_to_band", _ptr(self.blocks), _ptr(b.d_estart), _ptr(b.d_order), _ptr(b.d_boff), b.nel,
                  self.first, self.last, w, _ptr(A.data), _stream())
        return A

    def skyline(self):
        """The operator as the reference's own destination: ``BlockSkylineMatrix`` of ``Matrix(undef, B)``
        (src/fedvr_operators.jl:17-30) filled by ``set_elements!`` (:148-154), or ``Tridiagonal`` when every element
        has order 2."""
        b = self.basis
        R = b if (self.first, self.last) == (1, b.N) else b.restrict(self.first, self.last)
        dev = self.blocks.device
        if np.all(b.order == 2):
            n = self.n
            dl = torch.zeros(max(n - 1, 0), dtype=torch.float64, device=dev)
            du = torch.zeros(max(n - 1, 0), dtype=torch.float64, device=dev)
            d = torch.empty(n, dtype=torch.float64, device=dev)
            _lib.call("cb_fedvr_to_tridiag", _ptr(self.blocks), _ptr(b.d_estart), _ptr(b.d_order), _ptr(b.d_boff), b.nel,
                      self.first, self.last, _ptr(dl), _ptr(d), _ptr(du), _stream())
            return Tridiagonal(dl, d, du)
        rows = R.block_structure()
        l, u = R.block_bandwidths(rows)
        A = BlockSkylineMatrix(rows, l, u, dev)
        assert A.n == self.n
        tod = lambda a: to_device(a, dev, torch.int64)
        cum, kf, kl, st, sd = (tod(a) for a in (A.cum, A.kfirst, A.klast, A.start, A.stride))
        _lib.call("cb_fedvr_to_skyline", _ptr(self.blocks), _ptr(b.d_estart), _ptr(b.d_order), _ptr(b.d_boff), b.nel,
                  self.first, self.last, _ptr(cum), len(rows), _ptr(kf), _ptr(kl), _ptr(st), _ptr(sd), _ptr(A.data),
                  A.ndata, _stream())
        return A

    def _mul(self, Y, X, alpha, beta):
        b = self.basis
        _lib.call("cb_fedvr_apply", _ptr(self.blocks), _ptr(b.d_estart), _ptr(b.d_order), _ptr(b.d_boff), b.nel,
                  b.uniform_order, self.first, self.last, _ptr(X.data), X.ld, X.nvec, alpha, beta,
                  _ptr(Y.data), Y.ld, _stream())

    def mul_host(self, Yh: np.ndarray, Xh: np.ndarray, alpha=1.0, beta=0.0):
        """Host-buffer entry (cb_fedvr_apply_host); Fortran-ordered (n, nvec) numpy arrays."""
        assert Xh.flags.f_contiguous and Yh.flags.f_contiguous
        b = self.basis
        _lib.call("cb_fedvr_apply_host", _ptr(self.blocks), _ptr(b.d_estart), _ptr(b.d_order), _ptr(b.d_boff), b.nel,
                  b.uniform_order, self.first, self.last, C.c_void_p(Xh.ctypes.data), Xh.shape[0],
                  Xh.shape[1] if Xh.ndim > 1 else 1, alpha, beta, C.c_void_p(Yh.ctypes.data), Yh.shape[0], _stream())
        return Yh


def mul(Y, A, X, alpha: float = 1.0, beta: float = 0.0):
    """``mul!(Y, A, X, alpha, beta)``: Y = alpha*A*X + beta*Y.  X, Y: ColMajor (or torch
    tensors of shape (nvec, n)).  Throws DimensionMismatch like the reference."""
    X, Yc = _as_colmajor(X), _as_colmajor(Y)
    m, n = A.shape
    if X.n != n or Yc.n != m or X.nvec != Yc.nvec:
        raise DimensionMismatch(f"DimensionMismatch: A has size ({m},{n}), X has size ({X.n},{X.nvec}), "
                                f"Y has size ({Yc.n},{Yc.nvec})")
    A._mul(Yc, X, float(alpha), float(beta))
    return Y


def batched_dot(U, W, scale: float = 1.0) -> torch.Tensor:
    """``out[p] = scale * sum_i U[i,p] W[i,p]`` for ColMajor batches of equal shape (``cb_batched_dot``); returns a
    device tensor of P doubles."""
    Uc, Wc = _as_colmajor(U), _as_colmajor(W)
    if (Uc.n, Uc.nvec) != (Wc.n, Wc.nvec):
        raise DimensionMismatch(f"DimensionMismatch: U is {Uc.n}x{Uc.nvec}, W is {Wc.n}x{Wc.nvec}")
    out = torch.empty(Uc.nvec, dtype=torch.float64, device=Uc.data.device)
    _lib.call("cb_batched_dot", _ptr(Uc.data), Uc.ld, _ptr(Wc.data), Wc.ld, Uc.n, Uc.nvec, float(scale), _ptr(out), _stream())
    return out


def inner_product(u, R, v) -> torch.Tensor:
    """Pairwise ``(R*u_p)' (R*v_p)`` for coefficient batches u, v in the basis R (src/inner_products.jl:3-46):
    ``u_p' S v_p`` with the metric S = R'R for B-splines (banded apply + batched dot), ``step(R) * u_p' v_p`` for finite
    differences, ``u_p' v_p`` for FE-DVR.  Returns P doubles on the device."""
    from .bsplines import _BSplineOrRestricted
    uc, vc = _as_colmajor(u), _as_colmajor(v)
    if isinstance(R, _BSplineOrRestricted):
        S = R.S if hasattr(R, "S") else R.T @ R
        W = ColMajor(torch.empty_like(vc.data))
        mul(W, S, vc)
        return batched_dot(uc, W)
    scale = R.parent_basis.mass_value() if hasattr(R.parent_basis, "mass_value") else 1.0
    return batched_dot(uc, vc, scale)


def norm(v, R) -> torch.Tensor:
    """``norm(R*v_p)`` per column (src/inner_products.jl:60-84)."""
    return torch.sqrt(inner_product(v, R, v))
