"""GPU parity for the "next" rows of the scope table: function interpolation
``B \\ f`` (src/bsplines.jl:329-343) and ``LinearOperator`` with the metric inverse
(src/linear_operators.jl:26-75), both on the partitioned banded Cholesky solver of the density plan."""
import numpy as np
import pytest

from oracle import cbo
from tests.util import RTOL, oracle_basis, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cb():
    import torch
    assert torch.cuda.is_available()
    import compactbases_jl_b200 as cb
    return cb


LDIV_CASES = [
    ("k7_lin71", lambda cb: cb.LinearKnotSet(7, 0, 1.5, 71)),          # docs/src/densities.md setup
    ("k4_exp", lambda cb: cb.ExpKnotSet(4, -2.0, 1.0, 40)),
    ("k3_repeated", lambda cb: cb.ArbitraryKnotSet(3, [0.0, 1, 1, 3, 4, 6], 1, 3)),
    ("k5_lin700", lambda cb: cb.LinearKnotSet(5, 0, 10, 700)),         # > 2*256 functions: partitioned substitution
]


@pytest.mark.parametrize("name,make", LDIV_CASES, ids=[c[0] for c in LDIV_CASES])
def test_bspline_ldiv_matches_least_squares(cb, name, make):
    t = make(cb)
    B = cb.BSpline(t)
    te, q, x, w = oracle_basis(t)
    V = cbo.basis_matrix(te, t.k, x, q)
    f = lambda r: np.sin(2 * np.pi * r) * np.exp(-0.3 * r)
    got = B.ldiv(f).numpy()[:, 0]
    assert relerr(got, cbo.bspline_ldiv(V, f(x))) <= 1e-11, name
    # a batch of nodal-value vectors
    rng = np.random.default_rng(3)
    F = rng.standard_normal((x.size, 9))
    got = B.ldiv(np.ascontiguousarray(F.T)).numpy()
    assert relerr(got, cbo.bspline_ldiv(V, F)) <= 1e-10, name
    # restricted basis: the fit uses all nodes of the parent (src/bsplines.jl:335-343)
    if B.nf > 4:
        Br = B[:, 2:B.nf - 1]
        got = Br.ldiv(f).numpy()[:, 0]
        assert relerr(got, cbo.bspline_ldiv(V[:, 1:-1], f(x))) <= 1e-11, name
    with pytest.raises(cb.DimensionMismatch):
        B.ldiv(np.zeros(x.size + 1))


def test_bspline_ldiv_reproduces_polynomials_at_scale(cb):
    """Splines of order k contain the polynomials of degree < k: the fit of x^3 evaluates back to x^3
    (5e4 intervals: the partitioned solver with ~200 chunks)."""
    B = cb.BSpline(cb.LinearKnotSet(5, 0, 2, 50_000))
    c = B.ldiv(lambda r: r ** 3 - 2 * r)
    xs = np.random.default_rng(0).uniform(0, 2, 20_000)
    vals = B.evaluate(xs, c).numpy()[:, 0]
    np.testing.assert_allclose(vals, xs ** 3 - 2 * xs, atol=1e-10, rtol=0)


@pytest.mark.parametrize("k,nint", [(4, 60), (7, 40), (2, 30), (6, 800)])
def test_linear_operator_applies_metric_inverse(cb, k, nint):
    import torch
    B = cb.BSpline(cb.LinearKnotSet(k, 0, 3, nint))
    Br = B[:, 2:B.nf - 1] if k > 2 else B
    D = cb.Derivative()
    A = Br.T @ D.T @ D @ Br
    n = A.shape[0]
    M = cb.LinearOperator(A, Br)
    assert M.Binv is not None
    Ad, Sd = A.dense(), (Br.T @ Br).dense()
    rng = np.random.default_rng(k)
    for nvec in (1, 37):
        X = rng.standard_normal((n, nvec))
        Y0 = rng.standard_normal((n, nvec))
        Y = cb.ColMajor.from_numpy(Y0)
        M.mul(Y, cb.ColMajor.from_numpy(X))
        ref = cbo.linear_operator_apply(Ad, Sd, X)
        assert relerr(Y.numpy(), ref) <= 1e-10, (k, nvec)
        Y = cb.ColMajor.from_numpy(Y0)
        M.mul(Y, cb.ColMajor.from_numpy(X), 0.5, -2.0)
        assert relerr(Y.numpy(), cbo.linear_operator_apply(Ad, Sd, X, 0.5, -2.0, Y0)) <= 1e-10, (k, nvec)
    # S^-1 S = I through the factor alone
    F = (Br.T @ Br)
    if isinstance(F, cb.BandedMatrix):
        Xc = cb.ColMajor.from_numpy(rng.standard_normal((n, 5)))
        X0 = Xc.numpy().copy()
        SX = cb.ColMajor.zeros(n, 5)
        cb.mul(SX, F, Xc)
        F.cholesky().ldiv(SX)
        assert relerr(SX.numpy(), X0) <= 1e-11
        # the weak Laplacian is negative semi-definite: no Cholesky factor (PosDefException in the reference)
        with pytest.raises(ValueError):
            A.cholesky()


def test_linear_operator_orthogonal_bases_have_no_metric(cb):
    """operator_metric(FEDVR / FiniteDifferences) = I (src/linear_operators.jl:4-17): mul! is the plain product."""
    F = cb.FEDVR(np.linspace(0, 5, 11), 6)
    D = cb.Derivative()
    A = F.T @ D.T @ D @ F
    M = cb.LinearOperator(A, F)
    assert M.Binv is None
    X = np.random.default_rng(1).standard_normal((F.N, 3))
    Y1, Y2 = cb.ColMajor.zeros(F.N, 3), cb.ColMajor.zeros(F.N, 3)
    M.mul(Y1, cb.ColMajor.from_numpy(X))
    cb.mul(Y2, A, cb.ColMajor.from_numpy(X))
    assert np.array_equal(Y1.numpy(), Y2.numpy())


def test_density_docs_transcript_on_gpu(cb):
    """The reference's own transcript (docs/src/densities.md:119-182), B \\ f and Density both on the GPU."""
    import json, os
    g = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "bspline_rationals.json")))["docs"]["density_k7"]
    R = cb.BSpline(cb.LinearKnotSet(g["k"], g["a"], g["b"], g["N"]))
    cf = R.ldiv(lambda r: np.sin(2 * np.pi * r))
    cg = R.ldiv(lambda r: r * np.exp(-r))
    ch = R.ldiv(lambda r: np.sin(2 * np.pi * r) * r * np.exp(-r)).numpy()[:, 0]
    rho = cb.Density(R).copyto(cf, cg).numpy()[:, 0]
    np.testing.assert_allclose(rho[:16], g["first16"], rtol=0, atol=2e-10)
    np.testing.assert_allclose(rho[-16:], g["last16"], rtol=0, atol=2e-10)
    assert abs(np.linalg.norm(rho - ch) - g["norm_rho_minus_ch"]) <= 1e-9
