"""Finite-difference bases — host mirror of src/finite_differences.jl, src/fd_derivatives.jl
and src/fd_operators.jl (explicit three-point schemes) over the CUDA C ABI.

    R = StaggeredFiniteDifferences(10_000_000, 1e-3)
    lap = R.T @ D.T @ D @ R                  # SymTridiagonal
    lapZ = R.T @ CoulombDerivative(D, 1, 0).T @ CoulombDerivative(D, 1, 0) @ R
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .lazy import AdjointBasis, CoulombDerivative, Derivative, QuasiDiagonal
from .operators import Diagonal, ImplicitDerivative, SymTridiagonal, Tridiagonal, _ptr, _stream, device_of, to_device

CARTESIAN, STAGGERED_UNIFORM, STAGGERED_NONUNIFORM = 0, 1, 2


class _FDOrRestricted:
    is_basis = True

    @property
    def T(self):
        return AdjointBasis(self)

    def size(self):
        return self.last - self.first + 1

    def indices(self):
        return (self.first, self.last)

    def __getitem__(self, key):
        x, sel = key
        if not (isinstance(x, slice) and x == slice(None)):
            raise NotImplementedError("finite-difference tent-function evaluation is outside the accelerated path")
        a = 1 if sel.start is None else sel.start
        b = self.size() if sel.stop is None else sel.stop
        if not (1 <= a <= b <= self.size()):
            raise IndexError(f"BoundsError: restriction {a}:{b} outside 1:{self.size()}")
        return RestrictedFiniteDifferences(self.parent_basis, self.first + a - 1, self.first + b - 1)

    def ldiv(self, f):
        r"""``B \ f`` (src/finite_differences.jl:448-458): the function on the grid points over ``weights(B)``.
        Host glue (pointwise); returns the size() x 1 coefficients on the device."""
        from .operators import ColMajor, to_device
        P = self.parent_basis
        a, b = self.indices()
        x = np.asarray(P.locs())[a - 1:b]
        vd = P.vandermonde_diag()
        v = np.asarray(f(x), dtype=np.float64)
        if vd is not None:
            v = v / np.asarray(vd)[a - 1:b]
        return ColMajor(to_device(v.reshape(1, -1), P.device))

    def _materialize(self, L, middle):
        P = self.parent_basis
        # derivative_matrix, src/fd_derivatives.jl:200-203
        if not isinstance(L, _FDOrRestricted) or L.parent_basis != P:
            raise ValueError("Cannot multiply functions on different grids")
        if (L.first, L.last) != (self.first, self.last):
            raise NotImplementedError("FD operators between different restrictions (3-band BandedMatrix)")
        mid = list(middle)
        n, i0 = self.size(), self.first - 1
        if not mid:                                   # src/finite_differences.jl:141-143
            return Diagonal(torch.full((n,), P.mass_value(), dtype=torch.float64, device=P.device))
        if all(isinstance(f, QuasiDiagonal) for f in mid):   # src/fd_operators.jl:3-35
            d = None
            for f in mid:
                v = P.nodal_values(f.diag)[i0:i0 + n]
                d = v if d is None else d * v
            return Diagonal(d.contiguous())
        Z, ell = 0.0, 0
        ders = [f for f in mid if isinstance(f, (Derivative, CoulombDerivative))]
        for f in ders:
            if isinstance(f, CoulombDerivative):
                Z, ell = f.Z, f.ell
        weights = [f for f in mid if isinstance(f, QuasiDiagonal)]
        if len(ders) == 1 and not ders[0].adjoint and not weights:
            nder = 1
            if Z != 0.0:
                raise NotImplementedError("gradient_correction! (CoulombDerivative of first order)")
        elif len(ders) == 2 and ders[0].adjoint and not ders[1].adjoint and mid[0] is ders[0] and mid[-1] is ders[1]:
            nder = 2
        else:
            raise NotImplementedError(f"no materialiser for {middle}")
        if getattr(P, "implicit", False) and not P.uniform:
            # non-uniform compact scheme (Patchkovskii 2016, Eq. 33): both sides come from the three local steps
            # a, b, c around a node (implicit_stencil_common, src/fd_derivatives.jl:105-150, :410-424)
            if nder != 2:
                raise NotImplementedError("the reference has no first-derivative left-hand side on non-uniform grids (src/fd_derivatives.jl:411)")
            if Z != 0.0:
                raise NotImplementedError("CoulombDerivative corrections of the implicit scheme (src/fd_derivatives.jl:553-667)")
            st = P.nonuniform_stencils()
            sl = slice(i0, i0 + n)
            so = slice(i0, i0 + n - 1)
            dev = lambda a: to_device(np.ascontiguousarray(a), P.device)
            Delta = Tridiagonal(dev(st["alpha_t"][so]), dev(-2.0 * st["beta"][sl]), dev(st["alpha"][so]))   # step = 1
            M = Tridiagonal(dev(st["m_dl"][so]), dev(st["m_d"][sl]), dev(st["m_du"][so]))
            return ImplicitDerivative(Delta, M)
        fhalf = None
        if weights:                                   # src/fd_derivatives.jl:292-319
            if P.scheme != STAGGERED_NONUNIFORM or len(weights) != 1:
                raise NotImplementedError("weighted Laplacian exists for non-uniform staggered grids only")
            fhalf = P.half_node_values(weights[0].diag)
        dl = torch.empty(max(n - 1, 0), dtype=torch.float64, device=P.device)
        d = torch.empty(n, dtype=torch.float64, device=P.device)
        du = dl if nder == 2 else torch.empty_like(dl)
        _lib.call("cb_fd_derivative", P.scheme, nder, _ptr(P.d_r), _ptr(fhalf), P.N, P.step, i0, n, Z, ell,
                  _ptr(dl), _ptr(d), None if nder == 2 else _ptr(du), _stream())
        Delta = SymTridiagonal(d, dl) if nder == 2 else Tridiagonal(dl, d, du)
        if getattr(P, "implicit", False):
            # implicit_derivative, src/fd_derivatives.jl:426-432: the explicit stencil becomes the right-hand side
            # Delta, implicit_lhs (:402-424, uniform grids) the left-hand side M
            if Z != 0.0:
                raise NotImplementedError("CoulombDerivative corrections of the implicit scheme (src/fd_derivatives.jl:553-667)")
            dv, ev = (4.0 / 6.0, 1.0 / 6.0) if nder == 1 else (10.0 / 12.0, 1.0 / 12.0)
            M = SymTridiagonal(torch.full((n,), dv, dtype=torch.float64, device=P.device),
                               torch.full((max(n - 1, 0),), ev, dtype=torch.float64, device=P.device))
            return ImplicitDerivative(Delta, M)
        return Delta


class AbstractFiniteDifferences(_FDOrRestricted):
    scheme = CARTESIAN

    def _setup(self, r, step, device):
        self.r = np.ascontiguousarray(r, dtype=np.float64)
        self.N = self.r.size
        self.step = float(step)
        self.device = device_of(device)
        self.parent_basis, self.first, self.last = self, 1, self.N
        self._d_r = None

    @property
    def d_r(self):
        """Node vector on the device (only the non-uniform scheme reads it)."""
        if self.scheme != STAGGERED_NONUNIFORM:
            return None
        if self._d_r is None:
            self._d_r = to_device(self.r, self.device)
        return self._d_r

    def locs(self):
        return self.r

    def vandermonde_diag(self):
        """Diagonal of ``vandermonde(B)`` (src/finite_differences.jl:101-107): all ones (None) on
        uniform grids, 1/sqrt(local_step) (:30, :278-288) on non-uniform ones."""
        if self.scheme != STAGGERED_NONUNIFORM:
            return None
        return self._nonuniform_weights()

    def _nonuniform_weights(self):
        r = self.r
        h = np.empty_like(r)
        h[0], h[-1] = r[1] - r[0], r[-1] - r[-2]
        h[1:-1] = (r[2:] - r[:-2]) / 2
        return 1 / np.sqrt(h)

    def mass_value(self):
        return self.step                               # src/finite_differences.jl:141-143

    def nodal_values(self, f):
        if isinstance(f, torch.Tensor):
            return f.to(self.device, torch.float64)
        return to_device(np.asarray(f(self.r), dtype=np.float64), self.device)

    def __eq__(self, other):
        return type(self) is type(other) and np.array_equal(self.r, other.r)

    def __hash__(self):
        return hash((type(self).__name__, self.r.tobytes()))


class FiniteDifferences(AbstractFiniteDifferences):
    """``FiniteDifferences(n, dx)``: Cartesian grid (1:n)*dx (src/finite_differences.jl:175-187)."""
    scheme = CARTESIAN

    def __init__(self, n, dx, device=None):
        self._setup(np.arange(1, n + 1, dtype=np.float64) * dx, dx, device)


def implicit_stencil_common(r, fun, s=1, dN=0):
    """``implicit_stencil_common(fun, B, s, dN)`` (src/fd_derivatives.jl:105-125): ``fun(a, b, c)`` of the three
    consecutive local steps around every node, the grid continued by one mirrored point at the end and starting from
    the origin."""
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


class ImplicitFiniteDifferences(AbstractFiniteDifferences):
    """``ImplicitFiniteDifferences(n, rho)``: compact (implicit) finite differences on the uniform grid rho*(1:n), or
    ``ImplicitFiniteDifferences(r)`` on arbitrary non-decreasing nodes (src/finite_differences.jl:303-342).  Derivatives
    materialise to ``ImplicitDerivative``; on non-uniform grids (Patchkovskii 2016) only the second derivative exists,
    as in the reference, and its left-hand side is a non-symmetric tridiagonal matrix (band LU)."""
    scheme = CARTESIAN
    implicit = True

    def __init__(self, n_or_r, rho=None, device=None):
        if rho is not None:
            self.uniform = True
            self._setup(np.arange(1, int(n_or_r) + 1, dtype=np.float64) * rho, rho, device)
        else:
            r = np.asarray(n_or_r, dtype=np.float64)
            if np.any(np.diff(r) < 0):
                raise ValueError("Node locations must be non-decreasing")
            self.uniform = False
            self._setup(r, 1.0, device)                # step(::NonUniform) = 1
        self._stencils = None

    def nonuniform_stencils(self):
        """alpha~, alpha, beta (Eq. 33; src/fd_derivatives.jl:127-143) and the left-hand side of the second derivative
        (implicit_lhs, :410-424)."""
        if self._stencils is None:
            r = self.r
            self._stencils = {
                "alpha_t": implicit_stencil_common(r, lambda a, b, c: 2 / (a * (a + b)), 2, -1),
                "alpha": implicit_stencil_common(r, lambda a, b, c: 2 / (b * (a + b)), 1, -1),
                "beta": implicit_stencil_common(r, lambda a, b, c: 1 / (a * b)),
                "m_dl": implicit_stencil_common(r, lambda a, b, c: (a * a + a * b - b * b) / (6 * a * (a + b)), 2, -1),
                "m_d": implicit_stencil_common(r, lambda a, b, c: (a * a + 3 * a * b + b * b) / (6 * a * b)),
                "m_du": implicit_stencil_common(r, lambda a, b, c: (-a * a + a * b + b * b) / (6 * b * (a + b)), 1, -1),
            }
        return self._stencils

    def vandermonde_diag(self):
        return None if self.uniform else self._nonuniform_weights()


class StaggeredFiniteDifferences(AbstractFiniteDifferences):
    """``StaggeredFiniteDifferences(n, rho)`` (uniform, r_j = (j - 1/2) rho) or
    ``StaggeredFiniteDifferences(r)`` (arbitrary nodes): src/finite_differences.jl:236-275."""

    def __init__(self, n_or_r, rho=None, device=None):
        if rho is not None:
            n = int(n_or_r)
            self.scheme = STAGGERED_UNIFORM
            self._setup(rho * np.arange(1, n + 1, dtype=np.float64) - rho / 2, rho, device)
        else:
            r = np.asarray(n_or_r, dtype=np.float64)
            if np.any(np.diff(r) < 0):
                raise ValueError("Node locations must be non-decreasing")
            self.scheme = STAGGERED_NONUNIFORM
            self._setup(r, 1.0, device)                # step(::NonUniform) = 1, src/finite_differences.jl:18

    def half_node_values(self, f):
        """f(max(0, r_{j-1/2})), j = 1..N+1, with r_0 = 2r_1 - r_2 and r_{N+1} = 2r_N - r_{N-1}
        (the closure ``r -> O.diag[max(zero(T), r)]`` of src/fd_derivatives.jl:306)."""
        r = self.r
        rt = np.concatenate([[2 * r[0] - r[1]], r, [2 * r[-1] - r[-2]]])
        h = np.maximum(0.0, (rt[:-1] + rt[1:]) / 2)
        if isinstance(f, torch.Tensor):
            raise NotImplementedError("weights must be given as a callable for the half-node stencil")
        return to_device(np.asarray(f(h), dtype=np.float64), self.device)


class RestrictedFiniteDifferences(_FDOrRestricted):
    def __init__(self, parent, first, last):
        self.parent_basis, self.first, self.last = parent, first, last
