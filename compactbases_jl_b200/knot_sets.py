"""Knot sets (host side) — mirrors src/knot_sets.jl of CompactBases.jl v0.3.1.

``AbstractKnotSet{k,ml,mr,T}`` (src/knot_sets.jl:14) stores the interior knot
vector ``t.t`` once and presents the end knots with implicit multiplicities
``ml``/``mr`` (getindex, :138-151).  Only construction and indexing live here;
interval search on point batches is part of the evaluation kernel.
"""
from __future__ import annotations

from fractions import Fraction

import numpy as np


def _julia_range(a: float, b: float, n: int) -> np.ndarray:
    """``range(a, stop=b, length=n)`` as Julia materialises it: TwicePrecision
    arithmetic makes every element the correctly rounded a + i*(b-a)/(n-1)."""
    if n == 1:
        return np.array([float(a)])
    fa, fb = Fraction(float(a)), Fraction(float(b))
    step = (fb - fa) / (n - 1)
    return np.array([float(fa + step * i) for i in range(n)], dtype=np.float64)


class AbstractKnotSet:
    """k = order, ml/mr = left/right multiplicities (1..k), t = t.t."""

    def __init__(self, k: int, t, ml: int | None = None, mr: int | None = None):
        ml = k if ml is None else ml
        mr = k if mr is None else mr
        # assert_multiplicities, src/knot_sets.jl:26-33
        if k <= 0:
            raise ValueError(f"Polynomial order must be a positive integer, got k = {k}")
        if not 1 <= ml <= k:
            raise ValueError(f"Left multiplicity has to be in range 1..{k}, got {ml}")
        if not 1 <= mr <= k:
            raise ValueError(f"Right multiplicity has to be in range 1..{k}, got {mr}")
        t = np.ascontiguousarray(t, dtype=np.float64)
        nreq = max(2, k + 3 - ml - mr)
        if t.size < nreq:
            raise ValueError(f"Polynomial order k = {k} and left,right multiplicities {ml},{mr} "
                             f"require a knot sequence of at least {nreq} points")
        if np.any(np.diff(t) < 0):
            raise ValueError("Knot sequence must be non-decreasing")
        self.k, self.ml, self.mr, self.t = k, ml, mr, t

    # src/knot_sets.jl:60-71,116
    def order(self) -> int:
        return self.k

    def numintervals(self) -> int:
        return self.t.size - 1

    def __len__(self) -> int:
        return self.t.size + self.ml + self.mr - 2

    def numfunctions(self) -> int:
        return len(self) - self.k

    def first(self) -> float:
        return float(self.t[0])

    def last(self) -> float:
        return float(self.t[-1])

    def __getitem__(self, i: int) -> float:
        """1-based, as in Julia (src/knot_sets.jl:138-151)."""
        if not 1 <= i <= len(self):
            raise IndexError(f"BoundsError: attempt to access {len(self)}-element knot set at index [{i}]")
        j = min(max(i - (self.ml - 1), 1), self.t.size)
        return float(self.t[j - 1])

    def collect(self) -> np.ndarray:
        return np.concatenate([np.full(self.ml - 1, self.t[0]), self.t, np.full(self.mr - 1, self.t[-1])])

    def nonempty_intervals(self) -> np.ndarray:
        """src/knot_sets.jl:196-198 (1-based indices into the expanded vector)."""
        t = self.collect()
        return np.nonzero(t[:-1] != t[1:])[0] + 1

    def __eq__(self, other):
        return (isinstance(other, AbstractKnotSet) and self.k == other.k and self.ml == other.ml
                and self.mr == other.mr and np.array_equal(self.t, other.t))

    def __hash__(self):
        return hash((self.k, self.ml, self.mr, self.t.tobytes()))

    def __repr__(self):
        return (f"{type(self).__name__}(order k = {self.k} (maximum polynomial power = {self.k - 1}) "
                f"on {self.first()}..{self.last()} ({self.numintervals()} intervals))")


class ArbitraryKnotSet(AbstractKnotSet):
    """src/knot_sets.jl:250-266."""

    def __init__(self, k, t, ml=None, mr=None):
        super().__init__(k, t, ml, mr)


class LinearKnotSet(AbstractKnotSet):
    """``LinearKnotSet(k, a, b, N[, ml, mr])``: N intervals on a..b, src/knot_sets.jl:276-306."""

    def __init__(self, k, a, b, N, ml=None, mr=None):
        super().__init__(k, _julia_range(a, b, N + 1), ml, mr)


class ExpKnotSet(AbstractKnotSet):
    """``ExpKnotSet(k, a, b, N; base=10, include0=true)``: knots base^range(a, b), with an
    optional leading zero (src/knot_sets.jl:318-357)."""

    def __init__(self, k, a, b, N, ml=None, mr=None, base=10.0, include0=True):
        e = _julia_range(a, b, N if include0 else N + 1)
        t = np.power(float(base), e)
        if include0:
            t = np.concatenate([[0.0], t])
        super().__init__(k, t, ml, mr)
