"""Reference quadrature rules on [-1, 1] (host side).

The reference takes them from FastGaussQuadrature 0.4 (``gausslegendre``,
``gausslobatto``; src/quadrature.jl:62,82).  Here they are computed by Newton
iteration on the Legendre recurrence in extended precision, then rounded to
Float64; the mapping to the knot intervals (``lgwt``, ``element_grid``) happens
on the device / in fedvr.py.
"""
from __future__ import annotations

import math
from functools import lru_cache

import numpy as np

_LD = np.longdouble


def num_quadrature_points(k: int, kp: int = 3) -> int:
    """src/quadrature.jl:38-41: ceil((2(k-1) + k')/2)."""
    n2 = 2 * (k - 1) + kp
    return (n2 >> 1) + (n2 & 1)


def _legendre(n: int, x):
    """(P_n, P_{n-1}) at x."""
    p_prev, p = np.ones_like(x), x.copy()
    if n == 0:
        return p_prev, np.zeros_like(x)
    for m in range(2, n + 1):
        p_prev, p = p, ((2 * m - 1) * x * p - (m - 1) * p_prev) / m
    return p, p_prev


@lru_cache(maxsize=None)
def _gausslegendre(n: int):
    if n == 1:
        return (0.0,), (2.0,)
    i = np.arange(1, n + 1, dtype=_LD)
    x = -np.cos(_LD(math.pi) * (i - _LD(0.25)) / (n + _LD(0.5)))     # Tricomi start
    for _ in range(8):
        p, pm = _legendre(n, x)
        dp = n * (x * p - pm) / (x * x - 1)
        x = x - p / dp
    p, pm = _legendre(n, x)
    dp = n * (x * p - pm) / (x * x - 1)
    w = 2 / ((1 - x * x) * dp * dp)
    x = (x - x[::-1]) / 2                                            # exact antisymmetry
    w = (w + w[::-1]) / 2
    return tuple(float(v) for v in x), tuple(float(v) for v in w)


@lru_cache(maxsize=None)
def _gausslobatto(n: int):
    if n < 2:
        raise ValueError("Gauss-Lobatto rules need at least 2 points")
    if n == 2:
        return (-1.0, 1.0), (1.0, 1.0)
    N = n - 1
    x = -np.cos(_LD(math.pi) * np.arange(n, dtype=_LD) / N)
    for _ in range(60):
        p, pm = _legendre(N, x)
        q = N * (pm - x * p)               # (1 - x^2) P'_N
        dq = -N * (N + 1) * p
        dx = q / dq
        dx[0] = dx[-1] = 0
        x = x - dx
        if float(np.max(np.abs(dx))) < 1e-19:
            break
    x[0], x[-1] = -1, 1
    p, _ = _legendre(N, x)
    w = 2 / (N * (N + 1) * p * p)
    x = (x - x[::-1]) / 2
    w = (w + w[::-1]) / 2
    return tuple(float(v) for v in x), tuple(float(v) for v in w)


def gausslegendre(n: int):
    x, w = _gausslegendre(int(n))
    return np.array(x), np.array(w)


def gausslobatto(n: int):
    x, w = _gausslobatto(int(n))
    return np.array(x), np.array(w)


def _two_prod(a, b):
    """Error-free product (Dekker/Veltkamp): a*b = p + e exactly."""
    p = a * b
    c = 134217729.0                       # 2^27 + 1
    ah = a * c
    ah = ah - (ah - a)
    al = a - ah
    bh = b * c
    bh = bh - (bh - b)
    bl = b - bh
    e = ((ah * bh - p) + ah * bl + al * bh) + al * bl
    return p, e


def fma(a, b, c):
    """Fused multiply-add on numpy arrays through error-free transformations (numpy has
    no fma; Julia's ``lerp`` uses one, src/quadrature.jl:4).  Equal to the correctly
    rounded a*b + c unless the exact value lies within 2^-106 (relative) of a rounding
    boundary."""
    a, b, c = np.broadcast_arrays(np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64),
                                  np.asarray(c, dtype=np.float64))
    p, e = _two_prod(a, b)
    s = p + c
    bb = s - p
    err = (p - (s - bb)) + (c - bb)       # TwoSum
    return s + (err + e)


def lerp(a, b, t):
    """``lerp(a, b, t) = fma(t, b, fma(-t, a, a))`` (src/quadrature.jl:4)."""
    return fma(t, b, fma(-t, a, a))
