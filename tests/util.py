"""Shared helpers for the parity tests (the oracle is the checker, never the product)."""
import numpy as np

from oracle import cbo

RTOL = 1e-12   # north_star: <= 1e-12 relative in Float64


def relerr(a, b):
    """Norm-wise relative error max|a-b| / max|b| (SURVEY §7 'summation order')."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if a.size == 0:
        return 0.0
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))


def ell_to_dense(iv, vals, k, nf):
    nx = len(iv)
    D = np.zeros((nx, nf))
    for p in range(nx):
        I = int(iv[p])
        if I > 0:
            for a in range(k):
                j = I - k + 1 + a
                if 1 <= j <= nf:
                    D[p, j - 1] = vals[p, a]
    return D


def csc_to_dense(m, n, colptr, rowval, nzval):
    A = np.zeros((m, n))
    for c in range(n):
        sl = slice(colptr[c] - 1, colptr[c + 1] - 1)
        A[rowval[sl] - 1, c] = nzval[sl]
    return A


def oracle_basis(t_knotset, q=None, kp=3):
    """(expanded knots, q, x, w) of the oracle for a product-side knot set."""
    t = cbo.expand_knots(t_knotset.t, t_knotset.k, t_knotset.ml, t_knotset.mr)
    q = cbo.num_quadrature_points(t_knotset.k, kp) if q is None else q
    x, w = cbo.lgwt(t, q)
    return t, q, x, w
