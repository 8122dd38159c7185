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
from ._lib import DimensionMismatch
from .operators import (BandedMatrix, Diagonal, ImplicitDerivative, SymTridiagonal, Tridiagonal, _ptr, _stream, device_of,
                        to_device)

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
        same = (L.first, L.last) == (self.first, self.last)
        mid = list(middle)
        n, i0 = self.size(), self.first - 1
        m, r0 = L.size(), L.first - 1
        implicit = getattr(P, "implicit", False)
        if not mid:                                   # src/finite_differences.jl:141-143
            if not same:
                raise NotImplementedError("mass matrix between different restrictions")
            return Diagonal(torch.full((n,), P.mass_value(), dtype=torch.float64, device=P.device))
        if all(isinstance(f, QuasiDiagonal) for f in mid):   # src/fd_operators.jl:3-35
            d = None
            for f in mid:
                v = P.nodal_values(f.diag)
                d = v if d is None else d * v
            if same:
                return Diagonal(d[i0:i0 + n].contiguous())
            # different restrictions (:20-35): a single-band BandedMatrix with (l, u) = (-offset, offset),
            # offset = first(A) - first(B), holding the function at the common nodes
            off = r0 - i0
            lo, hi = max(r0, i0), min(r0 + m, i0 + n)
            band = torch.zeros((n, 1), dtype=torch.float64, device=P.device)
            if hi > lo:
                band[lo - i0:hi - i0, 0] = d[lo:hi]
            return BandedMatrix(band.contiguous(), m, n, -off, off)
        Z, ell = 0.0, 0
        ders = [f for f in mid if isinstance(f, (Derivative, CoulombDerivative))]
        coulomb = any(isinstance(f, CoulombDerivative) for f in ders)
        for f in ders:
            if isinstance(f, CoulombDerivative):
                Z, ell = f.Z, f.ell
        weights = [f for f in mid if isinstance(f, QuasiDiagonal)]
        if len(ders) == 1 and not ders[0].adjoint and not weights:
            nder = 1
        elif len(ders) == 2 and ders[0].adjoint and not ders[1].adjoint and mid[0] is ders[0] and mid[-1] is ders[1]:
            nder = 2
        else:
            raise NotImplementedError(f"no materialiser for {middle}")
        at_origin = L.first == 1 and self.first == 1   # corrections only touch the [1,1] corner (:521-534, :596-643)
        # explicit schemes: gradient_correction! is the identity (:485), laplacian_correction! exists for the staggered
        # grids (:542-551) and is applied on the device
        Zdev = Z if (nder == 2 and at_origin) else 0.0
        if implicit and not same:                     # implicit_derivative, :426-430
            raise DimensionMismatch("DimensionMismatch: Implicit finite-differences does not support different restrictions")
        if implicit and not P.uniform:
            # non-uniform compact scheme (Patchkovskii 2016, Eq. 33): both sides come from the local steps around a
            # node (implicit_stencil_common, src/fd_derivatives.jl:105-150, :410-424), computed by cb_fd_implicit_stencils
            if nder != 2:
                raise NotImplementedError("the reference has no first-derivative left-hand side on non-uniform grids (src/fd_derivatives.jl:411)")
            new = lambda k: torch.empty(max(k, 0), dtype=torch.float64, device=P.device)
            at, al, be, mdl, md, mdu = new(n - 1), new(n - 1), new(n), new(n - 1), new(n), new(n - 1)
            _lib.call("cb_fd_implicit_stencils", _ptr(P.d_nodes), P.N, i0, n, _ptr(at), _ptr(al), _ptr(be), _ptr(mdl), _ptr(md),
                      _ptr(mdu), _stream())
            Delta = Tridiagonal(at, -2.0 * be, al)                     # step(::NonUniform) = 1
            M = Tridiagonal(mdl, md, mdu)
            if coulomb and at_origin and Z != 0.0:
                # laplacian_correction!(::ImplicitDerivative, ::NonUniform, ...), :573-593, Eq. (34)
                if ell != 0:
                    raise NotImplementedError("the reference's Eq. (35) branch (l != 0) throws UndefVarError (src/fd_derivatives.jl:587)")
                a = float(P.r[0])
                b = float(P.r[1]) - a
                num = b * (a * (3 * a + 4 * b) * Z - 6 * (a + b))
                Delta.d[0] = 2 * (a + b) * (6 * a - (3 * a - b) * (a + b) * Z) / (a ** 2 * num)
                M.d[0] = (a + b) * (a * (a + b) * (a + 3 * b) * Z - 3 * (a ** 2 + 3 * b * a + b ** 2)) / (3 * a * num)
                if n > 1:
                    M.du[0] = (a ** 3 * (-Z) + a ** 2 * (b * Z + 3) + b * a * (2 * b * Z - 3) - 3 * b ** 2) / (3 * num)
            return ImplicitDerivative(Delta, M)
        fhalf = None
        if weights:                                   # src/fd_derivatives.jl:292-319
            if P.scheme != STAGGERED_NONUNIFORM or len(weights) != 1:
                raise NotImplementedError("weighted Laplacian exists for non-uniform staggered grids only")
            if not same:
                raise NotImplementedError("the reference's weighted Laplacian between different restrictions is commented out (src/fd_derivatives.jl:320-331)")
            fhalf = P.half_node_values(weights[0].diag)
        if not same:
            # derivative_matrix (:214-219): BandedMatrix{T}(undef, (m, n), (1 - offset, 1 + offset)), filled band by
            # band from the parent's stencils (:244-249, :278-289)
            off = r0 - i0
            dest = BandedMatrix.undef(m, n, 1 - off, 1 + off, P.device)
            _lib.call("cb_fd_derivative_banded", P.scheme, nder, _ptr(P.d_r), None, P.N, P.step, r0, m, i0, n, Zdev, ell,
                      _ptr(dest.data), _stream())
            return dest
        dl = torch.empty(max(n - 1, 0), dtype=torch.float64, device=P.device)
        d = torch.empty(n, dtype=torch.float64, device=P.device)
        du = dl if nder == 2 else torch.empty_like(dl)
        _lib.call("cb_fd_derivative", P.scheme, nder, _ptr(P.d_r), _ptr(fhalf), P.N, P.step, i0, n, Zdev, ell,
                  _ptr(dl), _ptr(d), None if nder == 2 else _ptr(du), _stream())
        Delta = SymTridiagonal(d, dl) if nder == 2 else Tridiagonal(dl, d, du)
        if implicit:
            # implicit_derivative, src/fd_derivatives.jl:426-432: the explicit stencil becomes the right-hand side
            # Delta, implicit_lhs (:402-424, uniform grids) the left-hand side M
            dv, ev = (4.0 / 6.0, 1.0 / 6.0) if nder == 1 else (10.0 / 12.0, 1.0 / 12.0)
            Md = torch.full((n,), dv, dtype=torch.float64, device=P.device)
            Me = torch.full((max(n - 1, 0),), ev, dtype=torch.float64, device=P.device)
            if coulomb and at_origin:
                rho = P.step
                if nder == 1:
                    # gradient_correction!(::ImplicitDerivative, ::Uniform, ...), :487-499 (Muller 1999, Eq. 20ff):
                    # applied whatever Z and l are
                    Delta.d[0] += (np.sqrt(3.0) - 2) / (2 * rho)
                    Md[0] += (np.sqrt(3.0) - 2) / 6
                elif ell == 0 and Z != 0.0:
                    # laplacian_correction!(::ImplicitDerivative, ::Uniform, ...), :553-565 (Muller 1999, Eq. 17)
                    db1 = -Z * rho / (12 - 10 * Z * rho)
                    Delta.d[0] -= 2 * db1 / rho ** 2
                    Md[0] -= 2 * db1 / ((-6) * (-2))
            return ImplicitDerivative(Delta, SymTridiagonal(Md, Me))
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

    @property
    def d_nodes(self):
        """Node vector on the device (any scheme)."""
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

    def eval(self, x):
        """``B[x, :]`` in row form (``cb_fd_eval``; getindex, src/finite_differences.jl:63-93): the 1-based first function
        that can be non-zero at every point (0: none) and the values of it and of the next one."""
        xq = x if isinstance(x, torch.Tensor) else to_device(np.asarray(x, dtype=np.float64), self.device)
        xq = xq.to(self.device, torch.float64).contiguous()
        nx = xq.numel()
        wv = self.vandermonde_diag()
        d_w = None if wv is None else to_device(np.asarray(wv, dtype=np.float64), self.device)
        first = torch.empty(nx, dtype=torch.int32, device=self.device)
        vals = torch.empty((nx, 2), dtype=torch.float64, device=self.device)
        _lib.call("cb_fd_eval", _ptr(self.d_nodes), None if d_w is None else _ptr(d_w), self.N, _ptr(xq), nx,
                  _ptr(first), _ptr(vals), _stream())
        return first, vals

    def eval_dense(self, x, first: int = 1, last: int | None = None) -> np.ndarray:
        """``Matrix(B[x, first:last])`` on the host (tests, small grids)."""
        last = self.N if last is None else last
        f, v = self.eval(x)
        f, v = f.cpu().numpy(), v.cpu().numpy()
        out = np.zeros((f.size, self.N + 1))
        idx = np.nonzero(f)[0]
        out[idx, f[idx] - 1] = v[idx, 0]
        out[idx, f[idx]] = v[idx, 1]
        return out[:, first - 1:last]

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
