"""GPU parity: FE-DVR / finite-difference assembly and operator application (mul!) vs the
CPU oracle, through the C ABI."""
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


def apply_gpu(cb, A, X, alpha=1.0, beta=0.0, Y0=None):
    Xc = cb.ColMajor.from_numpy(X)
    m = A.shape[0]
    Yc = cb.ColMajor.from_numpy(np.zeros((m, Xc.nvec)) if Y0 is None else Y0)
    if Y0 is None and beta == 0.0:
        Yc.data.fill_(float("nan"))          # beta == 0 must not read Y
    cb.mul(Yc, A, Xc, alpha, beta)
    return Yc.numpy()


FEDVR_CASES = [
    ("o2", lambda: (cbo.julia_range(0, 5, 12), 2)),
    ("o3", lambda: (cbo.julia_range(0, 5, 40), 3)),
    ("o5_readme", lambda: (cbo.julia_range(0, 10, 5), 5)),
    ("o10", lambda: (cbo.julia_range(0, 100, 101), 10)),
    ("o16", lambda: (cbo.julia_range(-3, 3, 9), 16)),
    ("mixed", lambda: (cbo.julia_range(-10.0, 10.0, 10), [4, 3, 4, 2, 5, 4, 2, 4, 8])),
    ("single", lambda: (cbo.julia_range(-10.0, 10.0, 2), 8)),
    ("o20_generic", lambda: (cbo.julia_range(0, 3, 4), 20)),
]


@pytest.mark.parametrize("name,make", FEDVR_CASES, ids=[c[0] for c in FEDVR_CASES])
def test_fedvr_assemble_and_apply(cb, name, make):
    t, order = make()
    B = cb.FEDVR(t, order)
    O = cbo.FEDVROracle(t, order)
    assert np.array_equal(B.x, O.x) and np.array_equal(B.n, O.n) and np.array_equal(B.wcat, O.wcat)
    assert B.block_structure() == O.block_structure()
    D = cb.Derivative()
    rng = np.random.default_rng(7)
    for nder, A in ((1, B.T @ D @ B), (2, B.T @ D.T @ D @ B)):
        blk = O.blocks(nder)
        assert relerr(A.blocks.cpu().numpy(), blk) <= RTOL
        assert relerr(A.dense(), O.dense(blk)) <= RTOL
        for nvec in (1, 3, 70):
            X = rng.standard_normal((B.N, nvec))
            Y0 = rng.standard_normal((B.N, nvec))
            ref = O.apply(blk, X)
            assert relerr(apply_gpu(cb, A, X), ref) <= RTOL
            ref2 = O.apply(blk, X, 0.5, -2.0, Y0.copy(order="F"))
            assert relerr(apply_gpu(cb, A, X, 0.5, -2.0, Y0), ref2) <= RTOL


def test_fedvr_readme_golden_and_restrictions(cb, golden):
    g = golden["docs"]["readme_fedvr"]
    B = cb.FEDVR(cbo.julia_range(g["a"], g["b"], g["nel"] + 1), g["order"])
    assert B.N == 17
    D = cb.Derivative()
    Br = B[:, g["sel"][0]:g["sel"][1]]
    A = (Br.T @ D.T @ D @ Br)
    Ad = A.dense()
    assert Ad.shape == (15, 15) and A.block_structure() == [3, 1, 3, 1, 3, 1, 3]
    np.testing.assert_allclose(Ad[:3, :3], g["block"], rtol=6e-6)
    np.testing.assert_allclose(Ad[:3, 3], g["east"], rtol=6e-6)
    np.testing.assert_allclose(Ad[3, 3], g["bridge_diag"], rtol=6e-6)
    np.testing.assert_allclose(Ad[3, 7], g["bridge_corner"], rtol=6e-6)
    # test/fedvr/derivatives.jl:46-111: restricted == sub-block, and apply of the restriction
    rng = np.random.default_rng(8)
    for order, sel, N in [(4, (2, 12), 5), (4, (1, 13), 5), ([5, 4, 2, 2], (5, 9), 5), ([5, 4, 3, 2], (5, 10), 5),
                          (2, (2, 3), 5), (2, (1, 5), 5), (5, (2, 4), 2), (10, (7, 83), 11)]:
        Bf = cb.FEDVR(cbo.julia_range(0, 20, N), order)
        Bs = Bf[:, sel[0]:sel[1]]
        O = cbo.FEDVROracle(cbo.julia_range(0, 20, N), order)
        assert Bs.block_structure() == O.block_structure(*sel)
        for nder, mid in ((1, [D]), (2, [D.T, D])):
            full = Bf._materialize(Bf, mid)
            sub = Bs._materialize(Bs, mid)
            fd, sd = full.dense(), sub.dense()
            assert np.array_equal(fd[sel[0] - 1:sel[1], sel[0] - 1:sel[1]], sd)
            X = rng.standard_normal((sub.n, 5))
            assert relerr(apply_gpu(cb, sub, X), sd @ X) <= RTOL


def test_fedvr_config2_apply_properties(cb):
    """BASELINE config 2 at full size: linearity, symmetry <Y,AX> = <AY,X>, oracle on a vector slice."""
    import torch
    t = cbo.julia_range(0, 1000, 10001)
    B = cb.FEDVR(t, 10)
    assert B.N == 90001
    D = cb.Derivative()
    A = B.T @ D.T @ D @ B
    O = cbo.FEDVROracle(t, 10)
    blk = O.blocks(2)
    assert relerr(A.blocks.cpu().numpy(), blk) <= RTOL
    g = torch.Generator(device="cuda").manual_seed(2)
    X = torch.randn((1024, B.N), dtype=torch.float64, device="cuda", generator=g)
    Y = torch.empty_like(X)
    cb.mul(Y, A, X)
    ref = O.apply(blk, np.asfortranarray(X[:16].cpu().numpy().T), threads=True)
    assert relerr(Y[:16].cpu().numpy().T, ref) <= RTOL
    ref_last = O.apply(blk, np.asfortranarray(X[-3:].cpu().numpy().T))
    assert relerr(Y[-3:].cpu().numpy().T, ref_last) <= RTOL
    # symmetry of the weak Laplacian: sum(Y .* Z) == sum(X .* (A Z))
    Z = torch.randn((1024, B.N), dtype=torch.float64, device="cuda", generator=g)
    AZ = torch.empty_like(Z)
    cb.mul(AZ, A, Z)
    s1, s2 = torch.sum(Y * Z).item(), torch.sum(X * AZ).item()
    assert abs(s1 - s2) <= 1e-10 * abs(s1)
    # linearity with alpha/beta: Y2 = 2*A*X - Y  == Y
    Y2 = Y.clone()
    cb.mul(Y2, A, X, 2.0, -1.0)
    assert torch.max(torch.abs(Y2 - Y)).item() <= 1e-10 * torch.max(torch.abs(Y)).item()


def test_fedvr_host_pipeline(cb):
    t = cbo.julia_range(0, 50, 501)
    B = cb.FEDVR(t, 6)
    D = cb.Derivative()
    A = B.T @ D.T @ D @ B
    O = cbo.FEDVROracle(t, 6)
    rng = np.random.default_rng(9)
    X = np.asfortranarray(rng.standard_normal((B.N, 37)))
    Y = np.asfortranarray(np.zeros((B.N, 37)))
    A.mul_host(Y, X)
    assert relerr(Y, O.apply(O.blocks(2), X)) <= RTOL
    Y0 = np.asfortranarray(rng.standard_normal((B.N, 37)))
    Y1 = Y0.copy(order="F")
    A.mul_host(Y1, X, 0.25, 3.0)
    assert relerr(Y1, O.apply(O.blocks(2), X, 0.25, 3.0, Y0.copy(order="F"))) <= RTOL


def test_fd_stencils_bit_exact(cb, golden):
    D = cb.Derivative()
    R = cb.FiniteDifferences(5, 1.0)
    G = R.T @ D @ R
    assert np.all(G.d.cpu().numpy() == 0) and np.all(G.du.cpu().numpy() == 0.5) and np.all(G.dl.cpu().numpy() == -0.5)
    Lp = R.T @ D.T @ D @ R
    assert isinstance(Lp, cb.SymTridiagonal)
    assert np.all(Lp.dv.cpu().numpy() == -2) and np.all(Lp.ev.cpu().numpy() == 1)
    for N, rho in ((3, 0.25), (1000, 0.1), (7, 1e-3)):
        R = cb.StaggeredFiniteDifferences(N, rho)
        al, be, ga, de = cbo.fd_staggered_uniform(N)
        dv, ev = cbo.fd_laplacian(al, be, rho)
        Lp = R.T @ D.T @ D @ R
        assert np.array_equal(Lp.dv.cpu().numpy(), dv) and np.array_equal(Lp.ev.cpu().numpy(), ev)
        dl, d, du = cbo.fd_gradient(de, ga, rho)
        G = R.T @ D @ R
        assert np.array_equal(G.dl.cpu().numpy(), dl) and np.array_equal(G.du.cpu().numpy(), du)
        CD = cb.CoulombDerivative(D, 1.0, 0)
        LpZ = R.T @ CD.T @ CD @ R
        assert np.array_equal(LpZ.dv.cpu().numpy(), cbo.laplacian_correction(dv, rho, 1.0, 0))
        CD1 = cb.CoulombDerivative(D, 1.0, 1)
        assert np.array_equal((R.T @ CD1.T @ CD1 @ R).dv.cpu().numpy(), dv)
        if N > 4:                                                  # restriction == sub-block
            Rr = R[:, 2:N - 1]
            Lr = Rr.T @ D.T @ D @ Rr
            assert np.array_equal(Lr.dv.cpu().numpy(), dv[1:-1]) and np.array_equal(Lr.ev.cpu().numpy(), ev[1:-1])
    g = golden["docs"]["fd"]
    R = cb.StaggeredFiniteDifferences(3, 0.25)
    Lp = R.T @ D.T @ D @ R
    np.testing.assert_allclose(Lp.ev.cpu().numpy(), g["staggered_N3_rho025"]["ev"], rtol=6e-6)
    assert Lp.dv.cpu().numpy()[0] == g["staggered_N3_rho025"]["dv_1_uncorrected"]
    # non-uniform staggered grid (log-linear-like), with and without a weight function
    r = np.cumsum(np.geomspace(0.01, 0.3, 60))
    R = cb.StaggeredFiniteDifferences(r)
    al, be, ga, de = cbo.fd_staggered_nonuniform(r)
    dv, ev = cbo.fd_laplacian(al, be, 1.0)
    Lp = R.T @ D.T @ D @ R
    assert np.array_equal(Lp.dv.cpu().numpy(), dv) and np.array_equal(Lp.ev.cpu().numpy(), ev)
    dl, d, du = cbo.fd_gradient(de, ga, 1.0)
    G = R.T @ D @ R
    assert np.array_equal(G.dl.cpu().numpy(), dl) and np.array_equal(G.du.cpu().numpy(), du)
    f = lambda x: 1.0 + x * x
    alf, bef, _, _ = cbo.fd_staggered_nonuniform(r, lambda x: float(f(x)))
    dvf, evf = cbo.fd_laplacian(alf, bef, 1.0)
    Lw = R.T @ D.T @ cb.QuasiDiagonal(f) @ D @ R
    assert np.array_equal(Lw.dv.cpu().numpy(), dvf) and np.array_equal(Lw.ev.cpu().numpy(), evf)


@pytest.mark.parametrize("n,nvec", [(1, 3), (2, 1), (10, 3), (513, 7), (4097, 33), (100000, 5)])
def test_tridiag_apply(cb, n, nvec):
    import torch
    rng = np.random.default_rng(n)
    dl, d, du = rng.standard_normal(max(n - 1, 0)), rng.standard_normal(n), rng.standard_normal(max(n - 1, 0))
    X = rng.standard_normal((n, nvec))
    Y0 = rng.standard_normal((n, nvec))
    T = cb.Tridiagonal(*(torch.as_tensor(v, device="cuda") for v in (dl, d, du)))
    assert relerr(apply_gpu(cb, T, X), cbo.tridiag_apply(dl, d, du, X)) <= RTOL
    assert relerr(apply_gpu(cb, T, X, -1.5, 0.5, Y0), cbo.tridiag_apply(dl, d, du, X, -1.5, 0.5, Y0.copy(order="F"))) <= RTOL
    Dg = cb.Diagonal(torch.as_tensor(d, device="cuda"))
    assert relerr(apply_gpu(cb, Dg, X, 2.0, 0.0), 2.0 * d[:, None] * X) <= RTOL


def test_tridiag_host_pipeline_and_staggered(cb):
    D = cb.Derivative()
    N = 20011
    R = cb.StaggeredFiniteDifferences(N, 1e-3)
    Lp = R.T @ D.T @ D @ R
    rng = np.random.default_rng(4)
    X = np.asfortranarray(rng.standard_normal((N, 9)))
    Y = np.asfortranarray(np.zeros((N, 9)))
    Lp.mul_host(Y, X)
    al, be, _, _ = cbo.fd_staggered_uniform(N)
    dv, ev = cbo.fd_laplacian(al, be, 1e-3)
    assert relerr(Y, cbo.tridiag_apply(ev, dv, ev, X)) <= RTOL
    assert relerr(apply_gpu(cb, Lp, X), Y) <= 1e-15


BANDED = [(10, 10, 2, 2), (50, 53, 3, 6), (53, 50, 6, 3), (300, 300, 0, 0), (129, 129, 12, 12), (1006, 1006, 6, 6),
          (40, 40, -1, 2), (40, 40, 2, -1), (7, 7, 9, 9), (2000, 2000, 9, 9)]


@pytest.mark.parametrize("m,n,l,u", BANDED)
def test_banded_apply(cb, m, n, l, u):
    import torch
    rng = np.random.default_rng(m + n + l)
    band = rng.standard_normal((n, l + u + 1))
    A = cb.BandedMatrix(torch.as_tensor(band, device="cuda").contiguous(), m, n, l, u)
    for nvec in (1, 5, 67):
        X = rng.standard_normal((n, nvec))
        Y0 = rng.standard_normal((m, nvec))
        assert relerr(apply_gpu(cb, A, X), cbo.banded_apply(band, m, n, l, u, X)) <= RTOL
        ref = cbo.banded_apply(band, m, n, l, u, X, 0.3, 1.7, Y0.copy(order="F"))
        assert relerr(apply_gpu(cb, A, X, 0.3, 1.7, Y0), ref) <= RTOL


def test_bspline_operator_apply_end_to_end(cb):
    """B'D'D B assembled on the GPU then applied to a coefficient batch == oracle of both steps."""
    t = cb.LinearKnotSet(7, 0, 100, 1000)
    B = cb.BSpline(t)
    D = cb.Derivative()
    T = B.T @ D.T @ D @ B
    te, q, x, w = oracle_basis(t)
    band, l, u = cbo.bspline_assemble(te, 7, te, 7, x, w, q, 1, 1)
    assert relerr(T.data.cpu().numpy(), band) <= RTOL
    rng = np.random.default_rng(5)
    X = rng.standard_normal((B.nf, 64))
    assert relerr(apply_gpu(cb, T, X), cbo.banded_apply(band, B.nf, B.nf, l, u, X)) <= RTOL


def test_apply_dimension_mismatch(cb):
    import torch
    T = cb.Tridiagonal(*(torch.zeros(s, dtype=torch.float64, device="cuda") for s in (4, 5, 4)))
    X = cb.ColMajor.zeros(6, 2)
    Y = cb.ColMajor.zeros(5, 2)
    with pytest.raises(cb.DimensionMismatch):
        cb.mul(Y, T, X)


# ---------------------------------------------------------------------------
# The warp-specialised TMA pipelines (banded, FE-DVR): enough tiles per CTA for the stage ring to wrap several times,
# odd and even leading dimensions (both 16-byte phases of the columns), partial last vector tile, and X / Y pairs
# whose columns differ in phase (the launcher must then take the non-TMA kernel and still be exact).
# ---------------------------------------------------------------------------
def _shifted_colmajor(cb, a, shift):
    """(n, nvec) numpy -> ColMajor whose storage starts `shift` doubles into a 16-byte aligned allocation."""
    import torch
    n, nvec = a.shape
    buf = torch.zeros(nvec * n + shift + 1, dtype=torch.float64, device="cuda")
    view = buf[shift:shift + nvec * n].view(nvec, n)
    view.copy_(torch.as_tensor(np.ascontiguousarray(a.T), device="cuda"))
    return cb.ColMajor(view)


@pytest.mark.parametrize("n,l,u,nvec,xshift,yshift", [
    (20001, 9, 9, 400, 0, 0),        # odd leading dimension: columns alternate in phase; ring wraps (2 041 tiles)
    (20000, 6, 6, 390, 0, 0),        # even leading dimension, partial last vector tile
    (4099, 3, 5, 130, 1, 1),         # both arrays start on an odd element
    (4099, 3, 5, 130, 0, 1),         # X and Y differ in phase: non-TMA kernel
    (1006, 6, 6, 4096, 0, 0),        # C1-like: few row tiles, first and last clipped
])
def test_banded_apply_pipeline_at_scale(cb, n, l, u, nvec, xshift, yshift):
    import torch
    rng = np.random.default_rng(n + nvec)
    band = rng.standard_normal((n, l + u + 1))
    A = cb.BandedMatrix(torch.as_tensor(band, device="cuda").contiguous(), n, n, l, u)
    X = rng.standard_normal((n, nvec))
    Xc = _shifted_colmajor(cb, X, xshift)
    Yc = _shifted_colmajor(cb, np.full((n, nvec), np.nan), yshift)
    cb.mul(Yc, A, Xc, 1.25, 0.0)
    ref = cbo.banded_apply(band, n, n, l, u, X, 1.25, 0.0, np.zeros((n, nvec), order="F"))
    assert relerr(Yc.numpy(), ref) <= RTOL


@pytest.mark.parametrize("order,nel,nvec,restrict,xshift,yshift", [
    (7, 3000, 330, None, 0, 0),          # 18 001 rows (odd), partial last vector tile, ring wraps
    (10, 1200, 200, (2, 10800), 0, 0),   # restricted basis: first and last functions dropped
    (4, 5001, 96, None, 1, 1),           # 15 004 rows (even), storage starting on an odd element
    (10, 1200, 64, None, 0, 1),          # X and Y differ in phase: lean kernel
])
def test_fedvr_apply_pipeline_at_scale(cb, order, nel, nvec, restrict, xshift, yshift):
    t = cbo.julia_range(0, float(nel), nel + 1)
    B = cb.FEDVR(t, order)
    O = cbo.FEDVROracle(t, order)
    D = cb.Derivative()
    blk = O.blocks(2)
    rng = np.random.default_rng(order + nel)
    if restrict is None:
        A = B.T @ D.T @ D @ B
        nfun, lo = B.N, 0
    else:
        Br = B[:, restrict[0]:restrict[1]]
        A = Br.T @ D.T @ D @ Br
        nfun, lo = restrict[1] - restrict[0] + 1, restrict[0] - 1
    X = rng.standard_normal((nfun, nvec))
    Xfull = np.zeros((B.N, nvec), order="F")
    Xfull[lo:lo + nfun] = X
    ref = O.apply(blk, Xfull, threads=True)[lo:lo + nfun] * 0.75
    Xc = _shifted_colmajor(cb, X, xshift)
    Yc = _shifted_colmajor(cb, np.full((nfun, nvec), np.nan), yshift)
    cb.mul(Yc, A, Xc, 0.75, 0.0)
    assert relerr(Yc.numpy(), ref) <= RTOL


@pytest.mark.parametrize("n,nvec,xshift,yshift", [
    (300001, 200, 0, 0),      # odd leading dimension, partial vector tile, the stage ring wraps many times
    (65536, 96, 0, 0),        # even leading dimension, whole tiles only
    (5000, 70, 1, 1),         # storage starting on an odd element
    (5000, 70, 0, 1),         # X and Y differ in phase: non-TMA kernel
])
def test_tridiag_apply_pipeline_at_scale(cb, n, nvec, xshift, yshift):
    import torch
    rng = np.random.default_rng(n + nvec)
    dl, d, du = rng.standard_normal(n - 1), rng.standard_normal(n), rng.standard_normal(n - 1)
    X = rng.standard_normal((n, nvec))
    T = cb.Tridiagonal(*(torch.as_tensor(v, device="cuda") for v in (dl, d, du)))
    Xc = _shifted_colmajor(cb, X, xshift)
    Yc = _shifted_colmajor(cb, np.full((n, nvec), np.nan), yshift)
    cb.mul(Yc, T, Xc, -0.75, 0.0)
    assert relerr(Yc.numpy(), cbo.tridiag_apply(dl, d, du, X, -0.75, 0.0, np.zeros((n, nvec), order="F"))) <= RTOL


def test_pipelines_random_shapes(cb):
    """Randomised shapes around the tile sizes of the three TMA pipelines (128-row / 384-row / element tiles, 32- and
    16-vector tiles): clipped first / last tiles, partial vector tiles, both column phases, restrictions."""
    import torch
    rng = np.random.default_rng(2024)
    for case in range(24):
        # banded
        n = int(rng.choice([127, 128, 129, 255, 257, 300, 385, 513, 1000, 1025]))
        l, u = int(rng.integers(0, 8)), int(rng.integers(0, 8))
        nvec = int(rng.choice([1, 7, 31, 32, 33, 64, 95]))
        shift = int(rng.integers(0, 2))
        band = rng.standard_normal((n, l + u + 1))
        A = cb.BandedMatrix(torch.as_tensor(band, device="cuda").contiguous(), n, n, l, u)
        X = rng.standard_normal((n, nvec))
        Yc = _shifted_colmajor(cb, np.full((n, nvec), np.nan), shift)
        cb.mul(Yc, A, _shifted_colmajor(cb, X, shift), 1.5, 0.0)
        ref = cbo.banded_apply(band, n, n, l, u, X, 1.5, 0.0, np.zeros((n, nvec), order="F"))
        assert relerr(Yc.numpy(), ref) <= RTOL, ("banded", n, l, u, nvec, shift)
        # tridiagonal
        n = int(rng.choice([256, 257, 383, 384, 385, 767, 769, 1153, 5000]))
        nvec = int(rng.choice([1, 15, 16, 17, 48, 50]))
        dl, d, du = rng.standard_normal(n - 1), rng.standard_normal(n), rng.standard_normal(n - 1)
        T = cb.Tridiagonal(*(torch.as_tensor(v, device="cuda") for v in (dl, d, du)))
        X = rng.standard_normal((n, nvec))
        Yc = _shifted_colmajor(cb, np.full((n, nvec), np.nan), shift)
        cb.mul(Yc, T, _shifted_colmajor(cb, X, shift), 2.0, 0.0)
        assert relerr(Yc.numpy(), cbo.tridiag_apply(dl, d, du, X, 2.0, 0.0, np.zeros((n, nvec), order="F"))) <= RTOL, ("tridiag", n, nvec, shift)
    for case in range(10):
        # FE-DVR, with and without a restriction
        order = int(rng.choice([2, 3, 5, 8, 10, 13, 16]))
        nel = int(rng.choice([1, 2, 14, 15, 16, 17, 31, 46, 100]))
        nvec = int(rng.choice([1, 31, 32, 33, 70]))
        t = cbo.julia_range(0, float(nel), nel + 1)
        B = cb.FEDVR(t, order)
        O = cbo.FEDVROracle(t, order)
        D = cb.Derivative()
        blk = O.blocks(2)
        a, b = 1, B.N
        if B.N > 4 and rng.integers(0, 2):
            a, b = 2, B.N - 1
        R = B if (a, b) == (1, B.N) else B[:, a:b]
        A = R.T @ D.T @ D @ R
        nfun = b - a + 1
        X = rng.standard_normal((nfun, nvec))
        Xfull = np.zeros((B.N, nvec), order="F")
        Xfull[a - 1:b] = X
        ref = O.apply(blk, Xfull)[a - 1:b]
        shift = int(rng.integers(0, 2))
        Yc = _shifted_colmajor(cb, np.full((nfun, nvec), np.nan), shift)
        cb.mul(Yc, A, _shifted_colmajor(cb, X, shift))
        assert relerr(Yc.numpy(), ref) <= RTOL, ("fedvr", order, nel, nvec, a, b, shift)


# ----------------------------------------------------------------------------------------------
# BlockSkylineMatrix / Tridiagonal export (the reference's own destination types)
# ----------------------------------------------------------------------------------------------
def _fedvr_with_int_blocks(cb, t, orders):
    """An FEDVRMatrix whose element i (1-based) is i*ones(o, o): the fills of test/fedvr/block_structure.jl."""
    import torch
    B = cb.FEDVR(t, orders)
    blk = np.concatenate([np.full(int(o) ** 2, float(i + 1)) for i, o in enumerate(B.order)])
    return B, torch.as_tensor(blk, device="cuda")


def test_fedvr_skyline_set_blocks_2(cb):
    """test/fedvr/block_structure.jl:122-164: B[:, 1:round(0.6N)+1] of orders [4,3,4,2,5,4,2,4,8], element i filled
    with i; the value of every listed Block(i, j), and the 4:end-1 restriction == sub-matrix."""
    B, blk = _fedvr_with_int_blocks(cb, cbo.julia_range(-10.0, 10.0, 10), [4, 3, 4, 2, 5, 4, 2, 4, 8])
    N = B.N
    b = int(np.floor(0.6 * N + 0.5)) + 1
    M = cb.FEDVRMatrix(B, blk, 1, b).skyline()
    assert isinstance(M, cb.BlockSkylineMatrix) and M.shape == (18, 18)
    expect = {1: [(1, 1), (1, 2), (2, 1)], 3: [(2, 2)], 2: [(2, 3), (2, 4), (3, 2), (4, 2), (3, 3), (3, 4), (4, 3)],
              5: [(4, 4)], 7: [(6, 6)], 4: [(7, 6), (6, 7)], 9: [(7, 7)],
              11: [(9, 9)], 6: [(10, 9), (11, 9), (9, 10), (9, 11), (10, 10), (11, 10), (10, 11)], 13: [(11, 11)],
              15: [(12, 12)]}
    expect3 = [(5, 4), (6, 4), (4, 5), (4, 6), (5, 5), (6, 5), (5, 6)]
    expect5 = [(8, 7), (9, 7), (7, 8), (7, 9), (8, 8), (9, 8), (8, 9)]
    expect7 = [(12, 11), (11, 12)]
    for v, ijs in list(expect.items()) + [(3, expect3), (5, expect5), (7, expect7)]:
        for i, j in ijs:
            assert np.all(M.block(i - 1, j - 1) == v), (v, i, j)
    M2 = cb.FEDVRMatrix(B, blk, 4, b - 1).skyline()
    assert np.array_equal(M2.dense(), M.dense()[3:-1, 3:-1])
    # and against the oracle's placement (set_elements! restated on a dense matrix)
    O = cbo.FEDVROracle(cbo.julia_range(-10.0, 10.0, 10), [4, 3, 4, 2, 5, 4, 2, 4, 8])
    assert np.array_equal(M.dense(), O.dense(blk.cpu().numpy(), 1, b))


@pytest.mark.parametrize("orders,sel", [([2, 3], None), ([2, 2, 3, 4, 2, 4], None), ([5, 6, 7, 8, 9, 10], None),
                                        (4, (2, 12)), ([5, 4, 2, 2], (5, 9)), ([5, 4, 3, 2], (5, 10)), (10, (2, 45)), (7, (7, 13))])
def test_fedvr_skyline_equals_dense_operator(cb, orders, sel):
    """derop! into the BlockSkylineMatrix destination == the dense operator, for the order mixes of
    test/fedvr/block_structure.jl:96-120 and the restrictions of test/fedvr/derivatives.jl:46-111; block sizes and
    bandwidths are those of the oracle's restatement of src/fedvr.jl:121-176."""
    nel = len(orders) if not np.isscalar(orders) else 5
    t = cbo.julia_range(0.0, 20.0, nel + 1)
    B = cb.FEDVR(t, orders)
    O = cbo.FEDVROracle(t, orders)
    D = cb.Derivative()
    R = B if sel is None else B[:, sel[0]:sel[1]]
    a, b = (1, B.N) if sel is None else sel
    for A in (R.T @ D @ R, R.T @ D.T @ D @ R):
        S = A.skyline()
        assert S.rows == O.block_structure(a, b)
        lo, uo = O.block_bandwidths(a, b)
        assert (S.l, S.u) == (lo[:len(S.rows)], uo[:len(S.rows)])
        assert np.array_equal(S.dense(), A.dense())


def test_fedvr_tridiagonal_destination(cb):
    """All orders 2: Matrix(undef, B) isa Tridiagonal (test/fedvr/block_structure.jl:26, src/fedvr_operators.jl:136-146)."""
    t = cbo.julia_range(1.0, 7.0, 7)
    B = cb.FEDVR(t, 2)
    D = cb.Derivative()
    for R in (B, B[:, 2:6]):
        for A in (R.T @ D @ R, R.T @ D.T @ D @ R):
            T = A.skyline()
            assert isinstance(T, cb.Tridiagonal)
            assert np.array_equal(T.dense(), A.dense())


@pytest.mark.parametrize("world", [2, 3, 8])
def test_fedvr_and_fd_assembly_shards_reassemble_bit_exact(cb, world):
    """K4 / K5 shards: element ranges of the FE-DVR blocks and node ranges of the finite-difference
    stencils, one call per rank; the concatenation is bit-identical to the single assembly (uniform and mixed orders,
    nder = 1, 2; Cartesian, staggered uniform and staggered non-uniform grids with the Coulomb correction on rank 0)."""
    import torch
    from compactbases_jl_b200.sharding import fedvr_assemble_sharded, fd_derivative_sharded
    D = cb.Derivative()
    for order in (7, np.array([4, 6, 3, 5, 7, 4, 6, 2, 5, 9, 3, 4, 8])):
        nel = 13 if not np.isscalar(order) else 37
        F = cb.FEDVR(cbo.julia_range(0.0, 5.0, nel + 1), order)
        for nder in (1, 2):
            full = F.derop(nder).clone()
            parts = [fedvr_assemble_sharded(F, nder, r, world) for r in range(world)]
            assert [p[0] for p in parts[1:]] == [p[1] for p in parts[:-1]] and parts[0][0] == 0 and parts[-1][1] == F.nel
            assert torch.equal(torch.cat([p[2] for p in parts]), full)
    grids = [cb.FiniteDifferences(1000, 0.01), cb.StaggeredFiniteDifferences(1001, 0.02),
             cb.StaggeredFiniteDifferences(0.01 * np.arange(1, 778) ** 1.3)]
    for R in grids:
        for nder, Z in ((1, 0.0), (2, 0.0), (2, 1.0)):
            if Z and isinstance(R, cb.FiniteDifferences):
                continue
            CD = cb.CoulombDerivative(D, Z, 0)
            A = (R.T @ D @ R) if nder == 1 else ((R.T @ CD.T @ CD @ R) if Z else (R.T @ D.T @ D @ R))
            parts = [fd_derivative_sharded(R, nder, r, world, Z=Z) for r in range(world)]
            d = torch.cat([p[3] for p in parts]); du = torch.cat([p[4] for p in parts]); dl = torch.cat([p[2] for p in parts])
            fd, fdu, fdl = A.d, A.du, A.dl
            assert torch.equal(d, fd) and torch.equal(du, fdu) and torch.equal(dl, fdl)


def test_fedvr_derivative_uses_the_right_hand_restriction(cb):
    """src/fedvr_derivatives.jl:66-97: the materialisers of A'*D*B / A'*D'*D*B only compare the parents and size the result
    by B (Matrix(undef, B); derop!(dest, B, n)) — whatever the restriction of A."""
    F = cb.FEDVR(cbo.julia_range(0.0, 3.0, 7), 5)
    D = cb.Derivative()
    A, B = F[:, 1:F.N - 2], F[:, 2:F.N - 1]
    got = (A.T @ D.T @ D @ B).dense()
    ref = (B.T @ D.T @ D @ B).dense()
    assert got.shape == ref.shape and np.array_equal(got, ref)
    with pytest.raises(ValueError):
        cb.FEDVR(cbo.julia_range(0.0, 3.0, 8), 5).T @ D @ B


@pytest.mark.gpu
@pytest.mark.parametrize("order", list(range(2, 17)))
def test_fedvr_rows_kernel_bit_identical_to_generic(cb, order):
    """Uniform orders <= 16 take the several-elements-per-warp kernel (fedvr_assemble_rows_kernel); the generic
    one-warp-per-element kernel (per-element order arrays) must give the same bits, for element counts that leave
    partially filled warps, and both must meet the oracle."""
    import torch
    from compactbases_jl_b200 import _lib
    from compactbases_jl_b200.operators import _ptr, _stream
    rng = np.random.default_rng(order)
    for nel in (1, 2, 3, 5, 16, 17, 333):
        t = np.concatenate([[0.0], np.cumsum(rng.uniform(0.05, 2.0, nel))]) + 900.0
        B = cb.FEDVR(t, order)
        assert B.uniform_order == order
        for nder in (1, 2):
            out = []
            for uni in (order, 0):
                blocks = torch.full((nel * order * order,), float("nan"), dtype=torch.float64, device="cuda")
                _lib.call("cb_fedvr_assemble", _ptr(B.d_x), _ptr(B.d_wcat), _ptr(B.d_n), _ptr(B.d_estart), _ptr(B.d_order),
                          _ptr(B.d_boff), _ptr(B.d_woff), nel, uni, order, nder, _ptr(blocks), _stream())
                out.append(blocks.cpu().numpy())
            assert np.array_equal(out[0], out[1]), (order, nel, nder)
        if nel in (5, 17):
            O = cbo.FEDVROracle(t, order)
            assert relerr(out[0], O.blocks(2)) <= RTOL


@pytest.mark.gpu
@pytest.mark.parametrize("order,nel,nvec", [(10, 400, 64), (10, 37, 5), (4, 300, 33), (16, 50, 32), ([4, 3, 4, 2, 5, 4, 2, 4, 8], 9, 7)])
def test_fedvr_assemble_apply_one_call(cb, order, nel, nvec):
    """cb_fedvr_assemble_apply (derop! followed by mul! in one foreign call; the apply pipeline is launched behind the
    assembly grid with programmatic stream serialisation and waits for it before reading the blocks): same bits as the
    two separate calls, repeatedly (the blocks are rewritten by every call), with alpha / beta, and within tolerance of the
    oracle."""
    import torch
    t = cbo.julia_range(0.0, 25.0, nel + 1)
    B = cb.FEDVR(t, order)
    D = cb.Derivative()
    rng = np.random.default_rng(nel)
    X = rng.standard_normal((B.N, nvec))
    Xc = cb.ColMajor.from_numpy(X)
    for nder, A in ((2, B.T @ D.T @ D @ B), (1, B.T @ D @ B)):
        ref_blocks = A.blocks.clone()
        Ysep = cb.ColMajor.from_numpy(np.zeros((B.N, nvec)))
        cb.mul(Ysep, A, Xc)
        for rep in range(3):
            B._ops[nder].fill_(float("nan"))                     # the call must rebuild the operator
            Y = cb.ColMajor.from_numpy(np.full((B.N, nvec), np.nan))
            A2 = B.derop_mul(nder, Y, Xc)
            torch.cuda.synchronize()
            assert torch.equal(A2.blocks, ref_blocks)
            assert np.array_equal(Y.numpy(), Ysep.numpy()), (nder, rep)
        Y0 = rng.standard_normal((B.N, nvec))
        Yb = cb.ColMajor.from_numpy(Y0)
        B.derop_mul(nder, Yb, Xc, 0.5, -2.0)
        Yc = cb.ColMajor.from_numpy(Y0)
        cb.mul(Yc, A, Xc, 0.5, -2.0)
        assert np.array_equal(Yb.numpy(), Yc.numpy())
        O = cbo.FEDVROracle(t, order)
        assert relerr(Ysep.numpy(), O.apply(O.blocks(nder), X)) <= RTOL


@pytest.mark.gpu
def test_host_register_entries(cb):
    """cb_host_register / cb_host_unregister pin caller-owned host arrays for the *_host entries (what a Julia host does
    with its Arrays): registered and pageable buffers give the same bits; unregistering twice is an error."""
    import ctypes as C
    from compactbases_jl_b200 import _lib
    t = cbo.julia_range(0, 50, 301)
    B = cb.FEDVR(t, 5)
    D = cb.Derivative()
    A = B.T @ D.T @ D @ B
    rng = np.random.default_rng(19)
    X = np.asfortranarray(rng.standard_normal((B.N, 20)))
    Yp = np.asfortranarray(np.zeros((B.N, 20)))
    A.mul_host(Yp, X)                                                  # pageable
    Yr = np.asfortranarray(np.zeros((B.N, 20)))
    lib = _lib.load()
    for arr in (X, Yr):
        assert lib.cb_host_register(C.c_void_p(arr.ctypes.data), C.c_size_t(arr.nbytes)) == 0
    try:
        A.mul_host(Yr, X)
    finally:
        for arr in (X, Yr):
            assert lib.cb_host_unregister(C.c_void_p(arr.ctypes.data)) == 0
    assert np.array_equal(Yr, Yp)
    assert lib.cb_host_unregister(C.c_void_p(X.ctypes.data)) != 0
    assert lib.cb_host_register(None, C.c_size_t(16)) != 0
    # a failed runtime call is reported once: the next entry's launch check must not see it again
    Yd = cb.ColMajor.from_numpy(np.zeros((B.N, 20)))
    cb.mul(Yd, A, cb.ColMajor.from_numpy(X))
    assert np.array_equal(Yd.numpy(), Yp)


@pytest.mark.gpu
@pytest.mark.parametrize("order", [2, 5, 10, [4, 3, 4, 2, 5, 4, 2, 4, 8]])
def test_fedvr_basis_function_evaluation(cb, order):
    """B[x, :] for FE-DVR (getindex, src/fedvr.jl:211-222, 272-295; test/fedvr/basics.jl: B[0.0, 1] != 0, the first
    function of the restriction B[:, 2:end-1] vanishes at the end points): bit-identical to the scalar loop, element
    boundaries and end points included, points outside the basis give empty rows, and the expansion of a smooth function
    reproduces it between the nodes (test/interpolation.jl)."""
    nel = 9
    t = cbo.julia_range(0.0, 20.0, nel + 1)
    B = cb.FEDVR(t, order)
    O = cbo.FEDVROracle(t, order)
    rng = np.random.default_rng(3)
    xq = np.concatenate([rng.uniform(0, 20, 200), np.asarray(t), B.x[::3], [-0.5, 20.5, np.nan]])
    got = B.eval_dense(xq)
    ref = cbo.fedvr_eval_dense(O, xq[:-3])
    assert np.array_equal(got[:-3], ref)
    assert not got[-3:].any()
    assert got[200, 0] != 0.0 and got[200 + nel, -1] != 0.0                  # B[0.0, 1], B[20.0, end]
    Br = B.eval_dense(xq, 2, B.N - 1)
    assert Br[200, 0] == 0.0 and Br[200 + nel, -1] == 0.0                     # restricted basis at the end points
    elem, _ = B.eval(xq)
    e = elem.cpu().numpy()
    assert e[200] == 1 and e[200 + nel] == nel and np.array_equal(e[201:200 + nel], np.arange(2, nel + 1))
    if not isinstance(order, list) and order >= 5:
        f = lambda r: np.exp(-0.1 * r) * np.sin(r)
        c = f(B.x) / B.n                                                       # R \ f for FE-DVR (src/fedvr.jl:369-412)
        assert np.max(np.abs(got[:200] @ c - f(xq[:200]))) <= (2e-3 if order == 5 else 1e-5)   # interpolation error of 2.2-wide elements


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["cartesian", "staggered_uniform", "staggered_nonuniform"])
def test_fd_basis_function_evaluation(cb, kind):
    """B[x, :] for finite-difference bases (getindex, src/finite_differences.jl:63-93; test/fd/basis_functions.jl): the
    weighted tents, bit-identical to the scalar formula, grid points, the mirrored end intervals and points outside
    included; the expansion of a function is its piecewise-linear interpolant."""
    if kind == "cartesian":
        R = cb.FiniteDifferences(40, 0.25)
    elif kind == "staggered_uniform":
        R = cb.StaggeredFiniteDifferences(40, 0.25)
    else:
        R = cb.StaggeredFiniteDifferences(np.cumsum(np.minimum(0.1 + 0.01 * np.arange(60), 0.2))[:40])
    l = np.asarray(R.locs())
    w = R.vandermonde_diag()
    rng = np.random.default_rng(4)
    ext = l[1] - l[0]
    xq = np.concatenate([rng.uniform(l[0] - ext, l[-1] + (l[-1] - l[-2]), 300), l, [l[0] - 2 * ext, l[-1] + 10.0, np.nan]])
    got = R.eval_dense(xq)
    ref = cbo.fd_eval_dense(l, w, xq[:-3])
    assert np.array_equal(got[:-3], ref)
    assert not got[-3:].any()
    wv = np.ones(l.size) if w is None else np.asarray(w)
    assert np.array_equal(np.diag(got[300:300 + l.size]), wv)                  # B[l_j, j] = weight_j, neighbours vanish
    inside = (xq[:300] >= l[0]) & (xq[:300] <= l[-1])
    f = lambda r: np.sin(r) + 0.3 * r
    c = f(l) / wv                                                              # R \\ f (src/finite_differences.jl:452-458)
    assert np.max(np.abs((got[:300] @ c)[inside] - np.interp(xq[:300][inside], l, f(l)))) <= 1e-13
