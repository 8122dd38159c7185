"""GPU parity: mutual-density projection (Density / FunctionProduct) vs the dense-pinv oracle."""
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


CASES = [
    ("k4_lin", lambda cb: cb.LinearKnotSet(4, 0, 1, 6)),
    ("k7_lin71", lambda cb: cb.LinearKnotSet(7, 0, 1.5, 71)),
    ("k8_lin", lambda cb: cb.LinearKnotSet(8, 0, 50, 300)),
    ("k3_repeated", lambda cb: cb.ArbitraryKnotSet(3, [0.0, 1, 1, 3, 4, 6], 1, 3)),
    ("k5_exp", lambda cb: cb.ExpKnotSet(5, -2.0, 1.0, 30)),
    ("k2", lambda cb: cb.LinearKnotSet(2, 0, 1, 9)),
    # more than 2*256 functions: the partitioned (chunked, spike-corrected) substitution
    ("k5_lin700", lambda cb: cb.LinearKnotSet(5, 0, 10, 700)),
    ("k8_exp600", lambda cb: cb.ExpKnotSet(8, -2.0, 1.0, 600)),
]


@pytest.mark.parametrize("name,make", CASES, ids=[c[0] for c in CASES])
def test_density_bspline_matches_pinv_oracle(cb, name, make):
    t = make(cb)
    B = cb.BSpline(t)
    te, q, x, w = oracle_basis(t)
    V = cbo.basis_matrix(te, t.k, x, q)
    nf = B.nf
    rng = np.random.default_rng(21)
    for P in (1, 5, 40):
        f, g = rng.standard_normal((nf, P)), rng.standard_normal((nf, P))
        rho = cb.Density(B, nlhs=P, nrhs=P)
        got = rho.copyto(cb.ColMajor.from_numpy(f), cb.ColMajor.from_numpy(g)).numpy()
        ref = cbo.density_bspline(V, f, g)
        assert relerr(got, ref) <= RTOL, (name, P)
        rho.set_refine(1)
        got2 = rho.copyto(cb.ColMajor.from_numpy(f), cb.ColMajor.from_numpy(g)).numpy()
        assert relerr(got2, ref) <= 1e-13, (name, P)
    # 1 x P broadcasting (src/densities.jl:57-72) and a weight function
    f1, g = rng.standard_normal((nf, 1)), rng.standard_normal((nf, 7))
    wfun = lambda r: 1.0 / (1.0 + r)
    rho = cb.Density(B, nlhs=1, nrhs=7, w=wfun)
    got = rho.copyto(cb.ColMajor.from_numpy(f1), cb.ColMajor.from_numpy(g)).numpy()
    assert relerr(got, cbo.density_bspline(V, f1[:, 0], g, wvals=wfun(x))) <= RTOL
    got = rho.copyto(cb.ColMajor.from_numpy(g), cb.ColMajor.from_numpy(f1)).numpy()
    assert relerr(got, cbo.density_bspline(V, g, f1[:, 0], wvals=wfun(x))) <= RTOL
    with pytest.raises(cb.DimensionMismatch):
        cb.Density(B, nlhs=3, nrhs=4)
    if t.ml != t.k or t.mr != t.k:
        return
    # partition of unity (full end multiplicities): density of (1, c) reproduces c
    c = np.sin(np.arange(1, nf + 1, dtype=np.float64))
    got = cb.Density(B).copyto(cb.ColMajor.from_numpy(np.ones(nf)), cb.ColMajor.from_numpy(c)).numpy()[:, 0]
    np.testing.assert_allclose(got, c, atol=1e-12, rtol=0)


def test_density_restricted_and_matrix(cb):
    t = cb.LinearKnotSet(5, 0, 2, 40)
    B = cb.BSpline(t)
    Br = B[:, 2:B.nf - 1]
    te, q, x, w = oracle_basis(t)
    V = cbo.basis_matrix(te, 5, x, q)[:, 1:-1]
    n = B.nf - 2
    rng = np.random.default_rng(22)
    f, g = rng.standard_normal((n, 3)), rng.standard_normal((n, 3))
    rho = cb.Density(Br, nlhs=3, nrhs=3)
    got = rho.copyto(cb.ColMajor.from_numpy(f), cb.ColMajor.from_numpy(g)).numpy()
    assert relerr(got, cbo.density_bspline(V, f, g)) <= RTOL
    # copyto!(M, rho): overlap matrix with the nodal product as diagonal operator
    rho1 = cb.Density(Br)
    rho1.copyto(cb.ColMajor.from_numpy(f[:, 0]), cb.ColMajor.from_numpy(g[:, 0]))
    M = rho1.matrix()
    tmp = (V @ f[:, 0]) * (V @ g[:, 0])
    assert relerr(rho1.tmp.cpu().numpy(), tmp) <= RTOL
    band, l, u = cbo.bspline_assemble(te, 5, te, 5, x, w, q, op=tmp, rowL0=1, m=n, colR0=1, n=n)
    assert relerr(M.dense(), cbo.band_to_dense(band, n, n, l, u)) <= RTOL


def test_density_orthogonal_branches(cb):
    rng = np.random.default_rng(23)
    R = cb.StaggeredFiniteDifferences(500, 0.1)
    f, g = rng.standard_normal((500, 4)), rng.standard_normal((500, 4))
    got = cb.Density(R, nlhs=4, nrhs=4).copyto(cb.ColMajor.from_numpy(f), cb.ColMajor.from_numpy(g)).numpy()
    assert np.array_equal(got, cbo.density_orthogonal(f, g))
    Fb = cb.FEDVR(cbo.julia_range(0, 10, 8), 5)
    n = Fb.N
    f, g = rng.standard_normal((n, 1)), rng.standard_normal((n, 6))
    got = cb.Density(Fb, nlhs=1, nrhs=6).copyto(cb.ColMajor.from_numpy(f), cb.ColMajor.from_numpy(g)).numpy()
    assert relerr(got, cbo.density_orthogonal(f[:, 0], g, c=Fb.n)) <= 1e-15
    wf = lambda r: np.exp(-r)
    got = cb.Density(Fb, nlhs=1, nrhs=6, w=wf).copyto(cb.ColMajor.from_numpy(f), cb.ColMajor.from_numpy(g)).numpy()
    assert relerr(got, cbo.density_orthogonal(f[:, 0], g, c=Fb.n * wf(Fb.x))) <= 1e-15


def test_density_config5_scaled_properties(cb):
    """Config-5 shape (k=8) at a size the dense oracle cannot reach: bilinearity and symmetry,
    plus the oracle on a small sub-problem with the same knot spacing."""
    import torch
    t = cb.LinearKnotSet(8, 0, 50, 5000)
    B = cb.BSpline(t)
    nf = B.nf
    g_ = torch.Generator(device="cuda").manual_seed(5)
    P = 64
    f = torch.randn((P, nf), dtype=torch.float64, device="cuda", generator=g_)
    g = torch.randn((P, nf), dtype=torch.float64, device="cuda", generator=g_)
    rho = cb.Density(B, nlhs=P, nrhs=P)
    r_fg = rho.copyto(cb.ColMajor(f), cb.ColMajor(g)).data.clone()
    r_gf = rho.copyto(cb.ColMajor(g), cb.ColMajor(f)).data.clone()
    scale = torch.max(torch.abs(r_fg)).item()
    assert torch.max(torch.abs(r_fg - r_gf)).item() <= 1e-13 * scale
    r_2 = rho.copyto(cb.ColMajor(2.0 * f), cb.ColMajor(g)).data.clone()
    assert torch.max(torch.abs(r_2 - 2.0 * r_fg)).item() <= 1e-13 * scale
    ones = torch.ones((1, nf), dtype=torch.float64, device="cuda")
    r_1 = cb.Density(B, nlhs=1, nrhs=P).copyto(cb.ColMajor(ones), cb.ColMajor(g)).data
    assert torch.max(torch.abs(r_1 - g)).item() <= 1e-11 * torch.max(torch.abs(g)).item()


def test_density_partitioned_batches_are_independent_of_their_grouping(cb):
    """The partitioned solve groups the pairs by 32 (tiled intermediate) and by 64 (two pairs per lane in the backward
    sweep): a pair's result must not depend on the batch it travels in — sub-batches that end inside a group give the
    same bits as the full batch — and the stand-alone banded solve (LinearOperator's metric inverse) takes the same
    route."""
    t = cb.LinearKnotSet(7, 0, 10, 700)                    # 706 functions: two chunks
    B = cb.BSpline(t)
    nf, P = B.nf, 97
    rng = np.random.default_rng(77)
    f, g = rng.standard_normal((nf, P)), rng.standard_normal((nf, P))
    full = cb.Density(B, nlhs=P, nrhs=P).copyto(cb.ColMajor.from_numpy(f), cb.ColMajor.from_numpy(g)).numpy()
    for Ps in (1, 31, 32, 33, 64, 65, 96):
        sub = cb.Density(B, nlhs=Ps, nrhs=Ps).copyto(cb.ColMajor.from_numpy(f[:, :Ps]), cb.ColMajor.from_numpy(g[:, :Ps])).numpy()
        assert np.array_equal(sub, full[:, :Ps]), Ps
    te, q, x, w = oracle_basis(t)
    V = cbo.basis_matrix(te, t.k, x, q)
    assert relerr(full[:, 90:], cbo.density_bspline(V, f[:, 90:], g[:, 90:])) <= RTOL
    # S \ (S c) = c through cb_bandchol_solve for ragged batches
    S = B.T @ B
    C = rng.standard_normal((nf, 70))
    Y = cb.ColMajor.from_numpy(np.zeros((nf, 70)))
    cb.mul(Y, S, cb.ColMajor.from_numpy(C))
    F = cb.factorize(S, "cholesky")
    F.ldiv(Y)
    assert relerr(Y.numpy(), C) <= 1e-11
