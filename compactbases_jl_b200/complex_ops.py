"""ComplexF64 paths of the reference on top of the real device operators.

The reference runs every basis with ``T = ComplexF64`` (test/linear_operators.jl:1-35: a complex potential applied to
complex coefficients through ``LinearOperator``) and has a complex-scaled FE-DVR basis (src/fedvr.jl:17-37,95-96).
Here a complex operator is a pair of real device operators ``A = Ar + i Ai`` and a complex coefficient batch is split
once into the real batch ``[Re | Im]`` of twice as many vectors (``cb_complex_split``); the real apply kernels run on that
batch and ``cb_complex_combine`` forms ``Y = alpha A X + beta Y`` (csrc/complex.cu).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from ._lib import DimensionMismatch
from .lazy import Derivative, QuasiDiagonal
from .operators import ColMajor, FEDVRMatrix, _ptr, _stream, device_of, to_device


class ComplexColMajor:
    """An n x nvec ComplexF64 column-major matrix on the device (Julia ``Matrix{ComplexF64}``): ``data`` has shape
    (nvec, ld) of complex128, column v is ``data[v, :n]``."""

    def __init__(self, data: torch.Tensor, n: int | None = None):
        if data.dim() == 1:
            data = data.unsqueeze(0)
        assert data.dim() == 2 and data.dtype == torch.complex128 and data.is_cuda and data.stride(1) == 1
        self.data = data
        self.n = data.shape[1] if n is None else n
        self.nvec = data.shape[0]
        self.ld = data.stride(0) if self.nvec > 1 else max(data.shape[1], 1)

    @classmethod
    def zeros(cls, n, nvec, dev=None):
        return cls(torch.zeros((nvec, n), dtype=torch.complex128, device=device_of(dev)))

    @classmethod
    def from_numpy(cls, a, dev=None):
        a = np.asarray(a, dtype=np.complex128)
        if a.ndim == 1:
            a = a[:, None]
        return cls(torch.as_tensor(np.ascontiguousarray(a.T), device=device_of(dev)))

    def numpy(self) -> np.ndarray:
        return np.asfortranarray(self.data[:, :self.n].cpu().numpy().T)

    def split(self) -> ColMajor:
        """The real batch [Re | Im] (2 nvec vectors)."""
        R = torch.empty((2 * self.nvec, self.n), dtype=torch.float64, device=self.data.device)
        _lib.call("cb_complex_split", self.data.data_ptr(), self.ld, self.n, self.nvec, _ptr(R), self.n, _stream())
        return ColMajor(R)


class ComplexOperator:
    """``A = Ar + i Ai`` with real device operators of equal shape (``Ai`` None: a real operator acting on complex
    coefficients).  ``metric_inverse``: the factorised metric of a ``LinearOperator`` (applied to both parts)."""

    def __init__(self, Ar, Ai=None, metric_inverse=None):
        if Ai is not None and tuple(Ai.shape) != tuple(Ar.shape):
            raise DimensionMismatch("DimensionMismatch: real and imaginary parts of the operator differ in size")
        self.Ar, self.Ai, self.Binv = Ar, Ai, metric_inverse

    @property
    def shape(self):
        return self.Ar.shape

    def mul(self, Y: ComplexColMajor, X: ComplexColMajor, alpha: complex = 1.0, beta: complex = 0.0):
        """``mul!(Y, A, X, alpha, beta)`` for ComplexF64 batches."""
        m, n = self.shape
        if X.n != n or Y.n != m or X.nvec != Y.nvec:
            raise DimensionMismatch(f"DimensionMismatch: A has size ({m},{n}), X has size ({X.n},{X.nvec}), Y has size ({Y.n},{Y.nvec})")
        R = X.split()
        T1 = ColMajor(torch.empty((2 * X.nvec, m), dtype=torch.float64, device=X.data.device))
        self.Ar._mul(T1, R, 1.0, 0.0)
        T2 = None
        if self.Ai is not None:
            T2 = ColMajor(torch.empty_like(T1.data))
            self.Ai._mul(T2, R, 1.0, 0.0)
        if self.Binv is not None:                          # S^-1 is real: it acts on every part separately
            self.Binv.ldiv(T1)
            if T2 is not None:
                self.Binv.ldiv(T2)
        a, b = complex(alpha), complex(beta)
        _lib.call("cb_complex_combine", _ptr(T1.data), None if T2 is None else _ptr(T2.data), m, m, X.nvec, a.real, a.imag, b.real, b.imag,
                  Y.data.data_ptr(), Y.ld, _stream())
        return Y

    def __matmul__(self, X: ComplexColMajor):
        Y = ComplexColMajor.zeros(self.shape[0], X.nvec, X.data.device)
        return self.mul(Y, X)

    def dense(self) -> np.ndarray:
        d = np.asarray(self.Ar.dense(), dtype=np.complex128)
        if self.Ai is not None:
            d = d + 1j * np.asarray(self.Ai.dense())
        return d


def complex_potential(R, f):
    """``R' * QuasiDiagonal(f.(x)) * R`` for a complex-valued function f (host callable): the two real materialisations
    of its real and imaginary parts (the assembly is linear in the potential)."""
    fr = QuasiDiagonal(lambda x: np.real(np.asarray(f(x), dtype=np.complex128)))
    fi = QuasiDiagonal(lambda x: np.imag(np.asarray(f(x), dtype=np.complex128)))
    return ComplexOperator(R.T @ fr @ R, R.T @ fi @ R)


def complex_linear_operator(A: ComplexOperator, R) -> ComplexOperator:
    """``LinearOperator(A, R)`` (src/linear_operators.jl:26-75) for a complex operator: y = S^-1 (A x) with the real
    metric S = R'R of a non-orthogonal basis."""
    from .operators import LinearOperator
    L = LinearOperator(A.Ar, R)
    return ComplexOperator(A.Ar, A.Ai, metric_inverse=L.Binv)


def complex_ldiv(R, f) -> ComplexColMajor:
    r"""``R \ f.(x)`` for a complex-valued f: the projection is linear, real and imaginary parts are fitted separately."""
    cr = R.ldiv(lambda x: np.real(np.asarray(f(x), dtype=np.complex128)))
    ci = R.ldiv(lambda x: np.imag(np.asarray(f(x), dtype=np.complex128)))
    return ComplexColMajor(torch.complex(cr.data[:, :cr.n].contiguous(), ci.data[:, :ci.n].contiguous()))


def complex_density(rho, f: ComplexColMajor, g: ComplexColMajor, out: ComplexColMajor | None = None,
                    alpha: complex = 1.0, beta: complex = 0.0) -> ComplexColMajor:
    """``copyto!(rho, f, g)`` for ComplexF64 coefficient batches (src/densities.jl:156-185 with ``T = ComplexF64``):
    ``Density`` conjugates its first factor, ``FunctionProduct{false}`` does not.  The projection onto the basis is a
    real linear map, so the complex density is four real ones on the split batches [Re | Im]:
    ``conj(f) g = (fr gr + fi gi) + i (fr gi - fi gr)``,  ``f g = (fr gr - fi gi) + i (fr gi + fi gr)``;
    every real density is the same ``cb_density_apply`` / ``cb_density_orthogonal`` call (the object ``rho`` carries the
    plan), the sums are ``cb_axpby``.  Column broadcasting (one of the factors with a single column) carries over.
    ``out``, ``alpha``, ``beta``: ``out = alpha * rho + beta * out`` (what ``DiagonalOperator``'s ``mul!`` needs)."""
    from .operators import axpby
    if f.n != rho.n or g.n != rho.n:
        raise DimensionMismatch(f"DimensionMismatch: basis has {rho.n} functions, got {f.n} and {g.n} rows")
    if f.nvec != g.nvec and 1 not in (f.nvec, g.nvec):
        raise DimensionMismatch(f"DimensionMismatch: {f.nvec} and {g.nvec} columns")
    fs, gs = f.split(), g.split()
    fr, fi = ColMajor(fs.data[:f.nvec]), ColMajor(fs.data[f.nvec:])
    gr, gi = ColMajor(gs.data[:g.nvec]), ColMajor(gs.data[g.nvec:])
    P = max(f.nvec, g.nvec)
    dev = f.data.device
    re_im = torch.empty((2 * P, rho.n), dtype=torch.float64, device=dev)     # [Re | Im] of the density
    re, im = ColMajor(re_im[:P]), ColMajor(re_im[P:])
    tmp = ColMajor(torch.empty((P, rho.n), dtype=torch.float64, device=dev))
    sgn = 1.0 if rho.conjugated else -1.0
    rho.copyto(fr, gr, re)
    rho.copyto(fi, gi, tmp)
    axpby(re, tmp, sgn, 1.0)                        # Re: fr gr +- fi gi
    rho.copyto(fr, gi, im)
    rho.copyto(fi, gr, tmp)
    axpby(im, tmp, -sgn, 1.0)                       # Im: fr gi -+ fi gr
    Z = ComplexColMajor.zeros(rho.n, P, dev) if out is None else out
    if Z.n != rho.n or Z.nvec != P:
        raise DimensionMismatch(f"DimensionMismatch: the density is {rho.n} x {P}, the destination {Z.n} x {Z.nvec}")
    a, b = complex(alpha), complex(beta)
    _lib.call("cb_complex_combine", _ptr(re_im), None, rho.n, rho.n, P, a.real, a.imag, b.real, b.imag,
              Z.data.data_ptr(), Z.ld, _stream())
    return Z


class ComplexDiagonalOperator:
    """``DiagonalOperator(f)`` with ComplexF64 coefficients (src/linear_operators.jl:85-173; the "Diagonal" operators of
    test/linear_operators.jl:1-35 with T = ComplexF64): ``mul!(y, L, x, alpha, beta)`` is the unconjugated product
    ``FunctionProduct{false}`` of the two expansions, ``y = alpha * (coefficients of f*x) + beta * y``."""

    def __init__(self, R, diag: ComplexColMajor, nrhs=None, w=None):
        from .densities import FunctionProduct
        self.R = R
        self.rho = FunctionProduct(R, R, 1, nrhs, w, conjugated=False)
        self.n = self.rho.n
        if diag.n != self.n or diag.nvec != 1:
            raise DimensionMismatch(f"DimensionMismatch: the basis has {self.n} functions, diag has {diag.n} x {diag.nvec}")
        self.diag = diag

    @property
    def shape(self):
        return (self.n, self.n)

    def mul(self, Y: ComplexColMajor, X: ComplexColMajor, alpha: complex = 1.0, beta: complex = 0.0):
        if X.n != self.n or Y.n != self.n or X.nvec != Y.nvec:
            raise DimensionMismatch(f"DimensionMismatch: operator of order {self.n}, X {X.n}x{X.nvec}, Y {Y.n}x{Y.nvec}")
        return complex_density(self.rho, self.diag, X, out=Y, alpha=alpha, beta=beta)

    def __matmul__(self, X: ComplexColMajor):
        return self.mul(ComplexColMajor.zeros(self.n, X.nvec, X.data.device), X)


def cmul(Y: ComplexColMajor, A, X: ComplexColMajor, alpha: complex = 1.0, beta: complex = 0.0):
    """``mul!`` for complex batches; A may be a ComplexOperator or any real device operator."""
    if not isinstance(A, ComplexOperator):
        A = ComplexOperator(A)
    return A.mul(Y, X, alpha, beta)


# ---- complex-scaled FE-DVR ---------------------------------------------------------------------------------------------
class ComplexScaledFEDVR:
    """``FEDVR(t, order, t0 = ..., phi = ...)`` with phi != 0 (src/fedvr.jl:9-53): exterior complex scaling from the
    first grid point >= t0 (``complex_rotate``: x -> t0 + (x - t0) e^{i phi}).  Fields as the reference computes them (x
    and n of complex type, weights real — see the note on the nodes below); the derivative operators are complex block
    operators (``ComplexOperator`` of two ``FEDVRMatrix``), assembled by ``cb_fedvr_assemble_complex``."""

    is_basis = True

    def __init__(self, t, order, t0: float = 0.0, phi: float = 0.0, device=None):
        from .fedvr import FEDVR
        from .quadrature import gausslobatto, lerp
        if phi == 0.0:
            raise ValueError("phi = 0: use FEDVR")
        t = np.ascontiguousarray(t, dtype=np.float64)
        self.real_basis = FEDVR(t, order, device=device)            # block structure, real_locs, device index arrays
        RB = self.real_basis
        i0 = np.nonzero(t >= t0)[0]
        if i0.size == 0:
            raise ValueError(f"ArgumentError: Complex scaling starting point outside grid {t}")
        self.i0 = int(i0[0]) + 1                                    # 1-based element index of the first rotated element
        self.t0 = float(t[self.i0 - 1])
        self.phi = float(phi)
        self.eiphi = np.exp(1j * phi)
        self.t, self.order, self.nel = t, RB.order, RB.nel
        # element_grid(order, t[i], t[i+1], t0, e^{i phi}) (src/quadrature.jl:81-85).  NOTE: in the reference (v0.3.1) the
        # non-mutating change_interval drops its rotation argument (src/quadrature.jl:25-27 calls change_interval! without
        # gamma), so the nodes come out complex-TYPED but unrotated; the rotation enters through the normalisation n
        # (rot(i, w), src/fedvr.jl:33-37) and the 1/sqrt(e^{i phi}) of the element matrices (src/fedvr_derivatives.jl:55).
        # This mirror reproduces what the reference computes, not what its docstring intends.
        xs, ws, ns = [], [], []
        for i in range(self.nel):
            xi, wr = gausslobatto(int(self.order[i]))
            a, b = t[i] - self.t0, t[i + 1] - self.t0
            rot = self.eiphi if i + 1 >= self.i0 else 1.0
            xs.append(self.t0 + lerp(a, b, (xi + 1) / 2))
            ws.append((b - a) * wr / 2)
            ns.append(1.0 / np.sqrt(rot * ws[-1] + 0j))
        for i in range(self.nel - 1):                               # bridge normalisation, :35-37
            ra = self.eiphi if i + 1 >= self.i0 else 1.0
            rb = self.eiphi if i + 2 >= self.i0 else 1.0
            v = 1.0 / np.sqrt(ra * ws[i][-1] + rb * ws[i + 1][0] + 0j)
            ns[i][-1] = v
            ns[i + 1][0] = v
        self.x = np.concatenate([xs[0]] + [v[1:] for v in xs[1:]]).astype(np.complex128)
        self.n = np.concatenate([ns[0]] + [v[1:] for v in ns[1:]]).astype(np.complex128)
        self.wi = ws
        self.wcat = np.concatenate(ws)
        self.d_wcat = to_device(self.wcat, RB.device)
        self.N = self.x.size
        self.device = RB.device
        self.parent_basis, self.first, self.last = self, 1, self.N
        self.d_x = torch.as_tensor(self.x, device=self.device)
        self.d_n = torch.as_tensor(self.n, device=self.device)
        self._ops = {}

    @property
    def T(self):
        from .lazy import AdjointBasis
        return AdjointBasis(self)

    def size(self):
        return self.N

    def locs(self):
        return self.x

    def real_locs(self):
        """``real_locs(B)`` (src/fedvr.jl:182-193): the grid of the unscaled basis (element_grid relative to t[1])."""
        return self.real_basis.x

    def complex_rotate(self, x):
        """``complex_rotate(x, B)`` (src/fedvr.jl:96)."""
        return x if x < self.t0 else self.t0 + (x - self.t0) * self.eiphi

    def derop(self, nder: int):
        """``derop!(A, B, nder)``: (blocks_re, blocks_im)."""
        if nder not in self._ops:
            RB = self.real_basis
            nb = int(np.sum(RB.order.astype(np.int64) ** 2))
            bre = torch.empty(nb, dtype=torch.float64, device=self.device)
            bim = torch.empty(nb, dtype=torch.float64, device=self.device)
            _lib.call("cb_fedvr_assemble_complex", self.d_x.data_ptr(), _ptr(self.d_wcat), self.d_n.data_ptr(), _ptr(RB.d_estart),
                      _ptr(RB.d_order), _ptr(RB.d_boff), _ptr(RB.d_woff), RB.nel, RB.uniform_order, int(RB.order.max()), nder,
                      self.i0 - 1, self.phi, _ptr(bre), _ptr(bim), _stream())
            self._ops[nder] = (bre, bim)
        return self._ops[nder]

    def _materialize(self, L, middle):
        if L is not self:
            raise ValueError("Cannot multiply functions on different grids")
        mid = list(middle)
        if len(mid) == 1 and isinstance(mid[0], Derivative) and not mid[0].adjoint:
            nder = 1
        elif (len(mid) == 2 and isinstance(mid[0], Derivative) and mid[0].adjoint and isinstance(mid[1], Derivative)
              and not mid[1].adjoint):
            nder = 2
        else:
            raise NotImplementedError(f"no materialiser for {middle} on the complex-scaled FE-DVR basis")
        bre, bim = self.derop(nder)
        RB = self.real_basis
        return ComplexOperator(FEDVRMatrix(RB, bre, 1, self.N), FEDVRMatrix(RB, bim, 1, self.N))

    def __repr__(self):
        return (f"FEDVR{{ComplexF64}} basis with {self.nel} elements on {self.t[0]}..{self.t[-1]} with ECS @ "
                f"{np.degrees(self.phi):0.2f}deg starting at {self.t0}")
