"""FE-DVR bases — host mirror of src/fedvr.jl, src/fedvr_operators.jl and
src/fedvr_derivatives.jl (real, non-complex-scaled case) over the CUDA C ABI.

    B = FEDVR(np.linspace(0, 1000, 10001), 10)
    lap = B.T @ D.T @ D @ B          # FEDVRMatrix (element blocks on the device)
    mul(Y, lap, X)
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .lazy import AdjointBasis, Derivative, QuasiDiagonal
from .operators import Diagonal, FEDVRMatrix, _ptr, _stream, device_of, to_device
from .quadrature import gausslobatto, lerp


class _FEDVROrRestricted:
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
            raise NotImplementedError("FE-DVR basis-function evaluation B[x,:] is outside the accelerated path")
        a = 1 if sel.start is None else sel.start
        b = self.size() if sel.stop is None else sel.stop
        return self.restrict(a, b)

    def restrict(self, a, b):
        if not (1 <= a <= b <= self.size()):
            raise IndexError(f"BoundsError: restriction {a}:{b} outside 1:{self.size()}")
        return RestrictedFEDVR(self.parent_basis, self.first + a - 1, self.first + b - 1)

    def block_structure(self):
        return self.parent_basis._block_structure(self.first, self.last)

    def block_bandwidths(self, rows=None):
        """``block_bandwidths(B, rows)`` (src/fedvr.jl:159-176): per-block-column (l, u) of the skyline."""
        P = self.parent_basis
        rows = self.block_structure() if rows is None else rows
        nrows = len(rows)
        if nrows <= 1:
            return [0], [0]
        o, _, _ = P.block_orders(self.first, self.last)
        bws = [[2, 1] if oo > 2 else [1] for oo in o]
        l = [1] + [v for bw in bws[1:] for v in bw]
        u = [v for bw in bws[:-1] for v in reversed(bw)] + [1]
        if len(l) < nrows:
            l = l + [0]
        if len(u) < nrows:
            u = [0] + u
        return l[:nrows], u[:nrows]

    def ldiv(self, f):
        r"""``B \ f`` (src/fedvr.jl:369-412): quadrature projection, ``v_j = n_j * (sum of the element weights at node
        j) * f(x_j) = f(x_j) / n_j``.  Host glue (pointwise); returns the size() x 1 coefficients on the device."""
        from .operators import ColMajor, to_device
        P = self.parent_basis
        x = np.asarray(P.x)[self.first - 1:self.last]
        n = np.asarray(P.n)[self.first - 1:self.last]
        return ColMajor(to_device((np.asarray(f(x), dtype=np.float64) / n).reshape(1, -1), P.device))

    def _materialize(self, L, middle):
        if not isinstance(L, _FEDVROrRestricted) or L.parent_basis is not self.parent_basis:
            raise ValueError("Cannot multiply functions on different grids")
        # The reference's materialisers (src/fedvr_derivatives.jl:66-97) check the parents only and build the destination
        # from B alone (Matrix(undef, B), derop!(dest, B, n)): the restriction of the adjoint factor is ignored.  Same here.
        P = self.parent_basis
        mid = list(middle)
        same = (L.first, L.last) == (self.first, self.last)
        if not same and (not mid or all(isinstance(f, QuasiDiagonal) for f in mid)):
            raise NotImplementedError("FE-DVR mass / potential matrices between different restrictions")
        if not mid:                                  # mass matrix = identity, src/fedvr.jl:331-333
            return Diagonal(torch.ones(self.size(), dtype=torch.float64, device=P.device))
        if all(isinstance(f, QuasiDiagonal) for f in mid):   # src/fedvr_operators.jl:158-185
            d = None
            for f in mid:
                v = P.nodal_values(f.diag)[self.first - 1:self.last]
                d = v if d is None else d * v
            return Diagonal(d.contiguous())
        if len(mid) == 1 and isinstance(mid[0], Derivative) and not mid[0].adjoint:
            nder = 1                                 # src/fedvr_derivatives.jl:68-78
        elif (len(mid) == 2 and isinstance(mid[0], Derivative) and mid[0].adjoint
              and isinstance(mid[1], Derivative) and not mid[1].adjoint):
            nder = 2                                 # :80-97
        else:
            raise NotImplementedError(f"no materialiser for {middle}")
        return FEDVRMatrix(P, P.derop(nder), self.first, self.last)


class FEDVR(_FEDVROrRestricted):
    """``FEDVR(t, order)`` (src/fedvr.jl:3-56): finite elements t[i]..t[i+1], each with a
    Gauss-Lobatto grid of ``order[i]`` points (element_grid, src/quadrature.jl:81-85);
    neighbouring elements share their boundary node (bridge function)."""

    def __init__(self, t, order, device=None):
        t = np.ascontiguousarray(t, dtype=np.float64)
        nel = t.size - 1
        order = np.full(nel, int(order), dtype=np.int64) if np.isscalar(order) else np.asarray(order, dtype=np.int64)
        if order.size != nel or np.any(order <= 1):
            raise ValueError("need one order > 1 per finite element")
        self.t, self.order, self.nel = t, order, nel
        host_only = device == "host"                   # grid + block structure only (no uploads)
        self.device = None if host_only else device_of(device)
        t0 = t[0]
        a, b = t[:-1] - t0, t[1:] - t0
        xs, ws = [], []
        for o in np.unique(order):
            xi, wr = gausslobatto(int(o))
            sel = np.nonzero(order == o)[0]
            tt = (xi + 1) / 2
            xe = t0 + lerp(a[sel, None], b[sel, None], tt[None, :])        # change_interval!, src/quadrature.jl:18-23
            we = (b[sel, None] - a[sel, None]) * wr[None, :] / 2
            xs.append((sel, xe))
            ws.append((sel, we))
        xel = [None] * nel
        wel = [None] * nel
        for (sel, xe), (_, we) in zip(xs, ws):
            for r, e in enumerate(sel):
                xel[e], wel[e] = xe[r], we[r]
        nrm = [1.0 / np.sqrt(w) for w in wel]                        # src/fedvr.jl:34
        for i in range(nel - 1):                                     # :35-37
            v = 1.0 / np.sqrt(wel[i][-1] + wel[i + 1][0])
            nrm[i][-1] = v
            nrm[i + 1][0] = v
        self.x = np.concatenate([xel[0]] + [v[1:] for v in xel[1:]])
        self.n = np.concatenate([nrm[0]] + [v[1:] for v in nrm[1:]])
        self.wi = wel
        self.wcat = np.concatenate(wel)
        self.estart = np.concatenate([[0], np.cumsum(order[:-1] - 1)]).astype(np.int64)
        self.boff = np.concatenate([[0], np.cumsum(order[:-1] ** 2)]).astype(np.int64)
        self.woff = np.concatenate([[0], np.cumsum(order[:-1])]).astype(np.int64)
        self.N = self.x.size
        self.uniform_order = int(order[0]) if np.all(order == order[0]) else 0
        self.parent_basis, self.first, self.last = self, 1, self.N
        self._ops = {}
        if host_only:
            return
        dev = self.device
        self.d_x, self.d_wcat, self.d_n = (to_device(v, dev) for v in (self.x, self.wcat, self.n))
        self.d_estart = to_device(self.estart, dev, torch.int64)
        self.d_boff = to_device(self.boff, dev, torch.int64)
        self.d_woff = to_device(self.woff, dev, torch.int64)
        self.d_order = to_device(order.astype(np.int32), dev, torch.int32)
        self._ops = {}

    def __repr__(self):
        return f"FEDVR{{Float64}} basis with {self.nel} elements on {self.t[0]}..{self.t[-1]}"

    def locs(self):
        return self.x

    def vandermonde_diag(self):
        """``vandermonde(B) = BandedMatrix(0 => B.n)`` (src/fedvr.jl:204)."""
        return self.n

    def nodal_values(self, f):
        if isinstance(f, torch.Tensor):
            return f.to(self.device, torch.float64)
        return to_device(np.asarray(f(self.x), dtype=np.float64), self.device)

    def eval(self, x):
        """``B[x, :]`` in row form (``cb_fedvr_eval``; getindex, src/fedvr.jl:211-222, 272-295): for every point the
        1-based element that contains it (0 outside the basis) and the values of that element's functions (function
        ``estart[e] + m + 1``, zero-padded to the largest order).  x: device tensor or array."""
        xq = x if isinstance(x, torch.Tensor) else to_device(np.asarray(x, dtype=np.float64), self.device)
        xq = xq.to(self.device, torch.float64).contiguous()
        nx, omax = xq.numel(), int(self.order.max())
        if not hasattr(self, "d_t"):
            self.d_t = to_device(np.asarray(self.t, dtype=np.float64), self.device)
        elem = torch.empty(nx, dtype=torch.int32, device=self.device)
        vals = torch.empty((nx, omax), dtype=torch.float64, device=self.device)
        _lib.call("cb_fedvr_eval", _ptr(self.d_x), _ptr(self.d_n), _ptr(self.d_t), _ptr(self.d_estart), _ptr(self.d_order),
                  self.nel, omax, _ptr(xq), nx, _ptr(elem), _ptr(vals), _stream())
        return elem, vals

    def eval_dense(self, x, first: int = 1, last: int | None = None) -> np.ndarray:
        """``Matrix(B[x, first:last])`` on the host (tests, small bases): scatters the row form."""
        last = self.N if last is None else last
        elem, vals = self.eval(x)
        e, v = elem.cpu().numpy(), vals.cpu().numpy()
        out = np.zeros((e.size, self.N))
        for p in np.nonzero(e)[0]:
            el = e[p] - 1
            o = int(self.order[el])
            out[p, self.estart[el]:self.estart[el] + o] = v[p, :o]
        return out[:, first - 1:last]

    def derop(self, nder: int) -> torch.Tensor:
        """``derop!(A, B, n)`` (src/fedvr_derivatives.jl:62-63): all element blocks."""
        if nder not in self._ops:
            blocks = torch.empty(int(np.sum(self.order ** 2)), dtype=torch.float64, device=self.device)
            _lib.call("cb_fedvr_assemble", _ptr(self.d_x), _ptr(self.d_wcat), _ptr(self.d_n), _ptr(self.d_estart),
                      _ptr(self.d_order), _ptr(self.d_boff), _ptr(self.d_woff), self.nel, self.uniform_order,
                      int(self.order.max()), nder, _ptr(blocks), _stream())
            self._ops[nder] = blocks
        return self._ops[nder]

    def derop_mul(self, nder: int, Y, X, alpha: float = 1.0, beta: float = 0.0):
        """``mul!(Y, derop!(A, B, nder), X, alpha, beta)`` in one foreign call (``cb_fedvr_assemble_apply``): the
        operator is re-assembled into its cached block array and applied; returns the FEDVRMatrix."""
        from .operators import ColMajor
        Xc = X if isinstance(X, ColMajor) else ColMajor(X)
        Yc = Y if isinstance(Y, ColMajor) else ColMajor(Y)
        if nder not in self._ops:
            self._ops[nder] = torch.empty(int(np.sum(self.order ** 2)), dtype=torch.float64, device=self.device)
        blocks = self._ops[nder]
        _lib.call("cb_fedvr_assemble_apply", _ptr(self.d_x), _ptr(self.d_wcat), _ptr(self.d_n), _ptr(self.d_estart),
                  _ptr(self.d_order), _ptr(self.d_boff), _ptr(self.d_woff), self.nel, self.uniform_order,
                  int(self.order.max()), nder, _ptr(blocks), 1, self.N, _ptr(Xc.data), Xc.ld, Xc.nvec, float(alpha), float(beta),
                  _ptr(Yc.data), Yc.ld, _stream())
        return FEDVRMatrix(self, blocks, 1, self.N)

    # -- block structure (src/fedvr.jl:82-176), host integer logic ----------
    def element_boundaries(self):
        return np.concatenate([[1], self.estart + self.order]).astype(np.int64)

    def elements(self, a, b):
        lo, hi = self.estart + 1, self.estart + self.order
        ia = np.nonzero((lo <= a) & (a <= hi))[0]
        ib = np.nonzero((lo <= b) & (b <= hi))[0]
        if ia.size == 0 or ib.size == 0:
            raise ValueError(f"Invalid restriction {a}..{b}")
        return int(ia[0]) + 1, int(ib[-1]) + 1

    def block_orders(self, a=1, b=None):
        b = self.N if b is None else b
        elb = self.element_boundaries()
        e1, e2 = self.elements(a, b)
        els = list(range(e1, e2 + 1))
        if els[0] != self.nel and a == elb[els[0]]:
            els = els[1:]
        if els[-1] > els[0] and b == elb[els[-1] - 1]:
            els = els[:-1]
        s = int(a - elb[els[0] - 1])
        e = int(elb[els[-1]] - b)
        return [int(self.order[el - 1]) for el in els], s, e

    def _block_structure(self, a=1, b=None):
        o, s, e = self.block_orders(a, b)
        if len(o) > 1:
            mid = []
            for oo in o[1:-1]:
                mid += [oo - 2, 1] if oo > 2 else [1]
            return [o[0] - 1 - s, 1] + mid + [o[-1] - 1 - e]
        return [o[0] - (s + e)]


class RestrictedFEDVR(_FEDVROrRestricted):
    def __init__(self, parent: FEDVR, first: int, last: int):
        self.parent_basis, self.first, self.last = parent, first, last
