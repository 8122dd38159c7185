"""Parity at the full BASELINE sizes (the full-size configs never met the oracle).

C3  BSpline k=10 on ExpKnotSet(10, -3, 2, 200000): Coulomb and Laplacian bands, every entry, vs the oracle
C4  StaggeredFiniteDifferences N = 1e7: tridiagonal Laplacian on 256 vectors, whole vectors vs the oracle
C5  Density on BSpline k=8, 5e4 intervals: 16 pairs vs the linear-cost least-squares oracle (banded QR)
and the reference's FE-DVR derivative-accuracy test (test/fedvr/derivatives.jl:127-174) on the GPU operators.
Tolerance: the north_star's 1e-12 relative (norm-wise per output array), exact for index structure.
"""
import numpy as np
import pytest

from oracle import cbo
from tests.util import RTOL, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cb():
    import torch
    assert torch.cuda.is_available()
    import compactbases_jl_b200 as cb
    return cb


def test_c3_full_size_bands_match_oracle(cb):
    t = cb.ExpKnotSet(10, -3.0, 2.0, 200_000)
    B = cb.BSpline(t)
    assert (B.nf, B.q, B.ni) == (200_009, 11, 200_000)
    D = cb.Derivative()
    V = B.T @ cb.CoulombPotential(1.0) @ B
    T = B.T @ D.T @ D @ B
    assert V.bandwidths() == (9, 9) and T.bandwidths() == (9, 9) and V.data.shape == (200_009, 19)
    te = cbo.expand_knots(t.t, 10)
    x, w = cbo.lgwt(te, 11)
    assert np.array_equal(B.x.cpu().numpy(), x) and relerr(B.w.cpu().numpy(), w) <= 1e-15
    cbo.set_threads(cbo.max_threads())
    try:
        Vo, l, u = cbo.bspline_assemble(te, 10, te, 10, x, w, 11, 0, 0, -1.0 / x)
        To, _, _ = cbo.bspline_assemble(te, 10, te, 10, x, w, 11, 1, 1)
    finally:
        cbo.set_threads(1)
    for name, got, ref in (("V", V.data.cpu().numpy(), Vo), ("T", T.data.cpu().numpy(), To)):
        assert relerr(got, ref) <= RTOL, name
        # the bands span 10 orders of magnitude along the exponential grid: every column on its own scale too
        colmax = np.max(np.abs(ref), axis=1)
        colerr = np.max(np.abs(got - ref), axis=1) / colmax
        assert np.max(colerr) <= 1e-11, (name, float(np.max(colerr)), int(np.argmax(colerr)))
        # band slots outside the matrix are exactly zero (BandedMatrices never reads them, the oracle zeroes them)
        assert np.array_equal(got == 0.0, ref == 0.0) or relerr(got[ref == 0.0], ref[ref == 0.0]) == 0.0


def test_c4_full_size_tridiagonal_apply(cb):
    import torch
    N, nv = 10_000_000, 256
    R = cb.StaggeredFiniteDifferences(N, 1e-3)
    D = cb.Derivative()
    Lp = R.T @ D.T @ D @ R
    al, be, _, _ = cbo.fd_staggered_uniform(N)
    dv, ev = cbo.fd_laplacian(al, be, 1e-3)
    assert np.array_equal(Lp.dv.cpu().numpy(), dv) and np.array_equal(Lp.ev.cpu().numpy(), ev)   # bit-exact stencils
    g = torch.Generator(device="cuda").manual_seed(4)
    X = torch.randn((nv, N), dtype=torch.float64, device="cuda", generator=g)
    Y = torch.full_like(X, float("nan"))
    cb.mul(Y, Lp, X)
    # 2.56e9 elements: the first size where 32-bit index arithmetic would break — whole vectors incl. the last
    for v in (0, 1, 127, 254, 255):
        ref = cbo.tridiag_apply(ev, dv, ev, X[v].cpu().numpy()[:, None])[:, 0]
        assert relerr(Y[v].cpu().numpy(), ref) <= RTOL, v
    assert bool(torch.isfinite(Y).all())
    # linearity over the whole batch: A(2X) - 2AX == 0 exactly (scaling by 2 is exact)
    X.mul_(2.0)
    Y2 = torch.empty_like(Y)
    cb.mul(Y2, Lp, X)
    assert float((Y2 - 2.0 * Y).abs().max()) == 0.0


def test_c5_full_size_density_matches_least_squares_oracle(cb):
    import torch
    t = cb.LinearKnotSet(8, 0, 500, 50_000)
    B = cb.BSpline(t)
    assert (B.nf, B.q) == (50_007, 9)
    P = 16
    rng = np.random.default_rng(5)
    xs = np.linspace(0, 1, B.nf)[:, None]
    env = np.exp(-((xs - rng.uniform(0.2, 0.8, (1, P))) / 0.25) ** 2)           # N(0,1) x smooth envelope
    f = rng.standard_normal((B.nf, P)) * env
    g = np.random.default_rng(6).standard_normal((B.nf, P)) * env
    rho = cb.Density(B, nlhs=P, nrhs=P)
    out = rho.copyto(cb.ColMajor.from_numpy(f), cb.ColMajor.from_numpy(g)).numpy()
    te = cbo.expand_knots(t.t, 8)
    x, w = cbo.lgwt(te, 9)
    cbo.set_threads(cbo.max_threads())
    try:
        ref = cbo.density_bspline_qr(te, 8, x, 9, f, g)
    finally:
        cbo.set_threads(1)
    assert out.shape == ref.shape == (50_007, P)
    for p in range(P):
        assert relerr(out[:, p], ref[:, p]) <= RTOL, p
    # broadcasting 1 x P (src/densities.jl:57-72) at full size
    out1 = rho.copyto(cb.ColMajor.from_numpy(f[:, :1]), cb.ColMajor.from_numpy(g)).numpy()
    ref1 = cbo.density_bspline_qr(te, 8, x, 9, f[:, :1], g)
    assert relerr(out1, ref1) <= RTOL


def _slopes(hs, errs):
    """estimate_convergence_rate (test/derivative_accuracy_utils.jl:45-60): the largest least-squares slope of
    log10(error) over log10(h) among consecutive groups of three points (cut at the first negative slope after
    the fifth)."""
    x, y = np.log10(hs), np.log10(np.abs(errs))
    sl = [np.polyfit(x[i:i + 3], y[i:i + 3], 1)[0] for i in range(len(x) - 2)]
    neg = [i for i, s in enumerate(sl) if s < 0]
    if neg and neg[0] + 1 > 5:
        sl = sl[:max(1, neg[0])]
    return max(sl)


def test_fedvr_derivative_accuracy_slopes(cb):
    """test/fedvr/derivatives.jl:127-174 on the device operators: Dirichlet-restricted FE-DVR of orders 2..10,
    f = exp(-1/(1-x^2)); errors of grad f, grad grad f and lap f in the (identity) metric; the convergence rates
    must exceed order - 2.5 / - 3.5 / - 2.5."""
    f = lambda x: np.exp(-1 / (1 - x ** 2))
    g = lambda x: -2 * np.exp(-1 / (1 - x ** 2)) * x / ((1 - x ** 2) ** 2)
    h = lambda x: -2 * np.exp(-1 / (1 - x ** 2)) * (1 - 3 * x ** 4) / ((1 - x ** 2) ** 4)
    D = cb.Derivative()
    orders = [(2, 1.5, 3), (3, 1.2, 2.5), (4, 1.3, 2.5), (5, 1.3, 2.5), (6, 1.3, 2.5), (7, 1, 2.5), (8, 1, 2.5),
              (9, 1, 2.5), (10, 1, 2.5)]
    for order, oa, ob in orders:
        Ns = np.ceil(10 ** np.linspace(oa, ob, 10)).astype(int)
        errs = []
        for N in Ns:
            R0 = cb.FEDVR(cbo.julia_range(-1.0, 1.0, int(N)), order)
            R = R0[:, 2:R0.N - 1]
            grad, lap = R.T @ D @ R, R.T @ D.T @ D @ R
            fv, gref, href = R.ldiv(f), R.ldiv(g), R.ldiv(h)
            gv = cb.ColMajor.zeros(R.size(), 1)
            hv = cb.ColMajor.zeros(R.size(), 1)
            hv2 = cb.ColMajor.zeros(R.size(), 1)
            cb.mul(gv, grad, fv)
            cb.mul(hv, grad, gv)
            cb.mul(hv2, lap, fv)
            errs.append([np.linalg.norm(gref.numpy() - gv.numpy()), np.linalg.norm(href.numpy() - hv.numpy()),
                         np.linalg.norm(href.numpy() - hv2.numpy())])
        errs = np.array(errs)
        hs = 1.0 / Ns
        pg, ph, ph2 = (_slopes(hs, errs[:, j]) for j in range(3))
        assert pg > order - 2.5 and ph > order - 3.5 and ph2 > order - 2.5, (order, pg, ph, ph2)
