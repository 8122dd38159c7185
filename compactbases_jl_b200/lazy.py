"""Lazy operator algebra: the slice of ContinuumArrays/LazyArrays the hot path needs.

In the reference ``B'*D'*D*B`` builds an ``ApplyQuasiArray(*, ...)`` and the
``@materialize`` DSL (src/materialize_dsl.jl:76-123) registers, per factor-type
tuple, ``similar`` (allocate the destination) + ``copyto!`` (fill it) +
``LazyArrays._simplify``.  Here ``@`` chains factors the same way and the
product is materialised as soon as it reads ``adjoint-basis ... basis``; the
basis on the right chooses the materialiser (its ``_materialize``), which is
where the CUDA assembly kernels sit.

    B.T @ B                      mass matrix
    B.T @ QuasiDiagonal(f) @ B   potential
    B.T @ D @ B                  first derivative
    B.T @ D.T @ D @ B            weak Laplacian
    B.T @ D.T @ QuasiDiagonal(f) @ D @ B
"""
from __future__ import annotations


class Derivative:
    """``Derivative(axes(B,1))``."""

    adjoint = False

    @property
    def T(self):
        return _AdjointDerivative()

    def __matmul__(self, other):
        return Applied([self]) @ other

    def __repr__(self):
        return "Derivative"


class _AdjointDerivative(Derivative):
    adjoint = True

    @property
    def T(self):
        return Derivative()

    def __repr__(self):
        return "Derivative'"


class QuasiDiagonal:
    """``QuasiDiagonal(f.(x))``: ``diag`` is a callable evaluated at points on the host
    (numpy array in, numpy array out) — the reference evaluates the Julia closure at
    ``locs(B)`` on the host too (src/bsplines.jl:305) — or a device tensor of values at
    the quadrature nodes / grid points."""

    def __init__(self, diag):
        self.diag = diag

    def __matmul__(self, other):
        return Applied([self]) @ other

    def __repr__(self):
        return "QuasiDiagonal"


class CoulombDerivative:
    """``CoulombDerivative(D, Z, ell)`` (src/coulomb_derivative.jl:9-13): a derivative that
    carries the nuclear charge so finite-difference materialisers can apply the
    Schafer / Muller corrections to the first row."""

    adjoint = False

    def __init__(self, D, Z, ell):
        self.D, self.Z, self.ell = D, float(Z), int(ell)

    @property
    def T(self):
        c = CoulombDerivative(self.D, self.Z, self.ell)
        c.adjoint = not self.adjoint
        return c

    def __matmul__(self, other):
        return Applied([self]) @ other


class AdjointBasis:
    """``B'``."""

    def __init__(self, parent):
        self.parent = parent

    @property
    def T(self):
        return self.parent

    def __matmul__(self, other):
        return Applied([self]) @ other


class Applied:
    """A pending product of factors (``ApplyQuasiArray(*, factors...)``)."""

    def __init__(self, factors):
        self.factors = list(factors)

    def __matmul__(self, other):
        factors = self.factors + (other.factors if isinstance(other, Applied) else [other])
        if isinstance(factors[0], AdjointBasis) and getattr(factors[-1], "is_basis", False):
            return factors[-1]._materialize(factors[0].parent, factors[1:-1])
        return Applied(factors)


def materialize(applied: Applied):
    f = applied.factors
    return f[-1]._materialize(f[0].parent, f[1:-1])
