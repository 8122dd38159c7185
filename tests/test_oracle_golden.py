"""Pins the CPU oracle against the reference's own known-answer fixtures
(SURVEY.md §8c): exact rational matrices from test/bsplines/*.jl, REPL
transcripts from docs/ and README.  CPU only."""
from fractions import Fraction

import numpy as np
import pytest

from oracle import cbo


def frac_matrix(rows):
    return np.array([[float(Fraction(a, b)) for a, b in r] for r in rows])


def bspline(tt, k, ml=None, mr=None, kp=3, q=None):
    t = cbo.expand_knots(tt, k, ml, mr)
    q = cbo.num_quadrature_points(k, kp) if q is None else q
    x, w = cbo.lgwt(t, q)
    return t, q, x, w


def test_knot_vectors():
    # test/bsplines/knot_sets.jl:47-50
    assert list(cbo.expand_knots([0.0, 1, 1, 3, 4, 6], 3, 1, 3)) == [0.0, 1, 1, 3, 4, 6, 6, 6]
    assert list(cbo.expand_knots(cbo.linear_knots(0, 1, 2), 3)) == [0, 0, 0, 0.5, 1, 1, 1]
    assert list(cbo.expand_knots(cbo.linear_knots(0, 1, 2), 2, 1, 1)) == [0, 0.5, 1]
    got = cbo.expand_knots(cbo.exp_knots(-4.0, 2.0, 7), 2)
    np.testing.assert_allclose(got, [0, 0, 0.0001, 0.001, 0.01, 0.1, 1, 10, 100, 100], rtol=2e-16)
    assert len(cbo.nonempty_intervals(cbo.expand_knots([0.0, 1, 1, 3, 4, 6], 3, 1, 3))) == 4
    for k in range(1, 7):  # :53-66
        t = cbo.expand_knots(cbo.linear_knots(0, 1, 1), k)
        assert len(t) == 2 * k and len(cbo.nonempty_intervals(t)) == 1


def test_find_interval():
    # test/bsplines/knot_sets.jl:179-204
    t = cbo.expand_knots([0.0, 1, 1, 3, 4, 6], 3, 1, 3)
    eps = np.finfo(float).eps
    assert cbo.find_interval(t, 3, 0.5) == 1
    cases = [[0, 0.5, 1.0 - eps], [], [1.0, 2.0, 3.0 - np.spacing(3.0)],
             [3, 3.5, 4.0 - np.spacing(4.0)], [4, 5, 6.0 - np.spacing(6.0), 6]]
    for i, xs in enumerate(cases, start=1):
        for j in (1, i):
            for x in xs:
                assert cbo.find_interval(t, 3, x, j) == i
                assert cbo.find_interval_bisect(t, 3, x) == i
                assert cbo.within_support(np.array([x]), [0.0, 1, 1, 3, 4, 6], 3, 1, 3, i)[0][1] == i
    assert cbo.find_interval(t, 3, 0.5, 2) is None
    assert cbo.find_interval_bisect(t, 3, -0.1) is None and cbo.find_interval_bisect(t, 3, 6.1) is None
    # bisect == linear scan on random points incl. the knots themselves
    rng = np.random.default_rng(0)
    tt = np.sort(rng.uniform(0, 1, 40)); tt[5] = tt[6]
    t2 = cbo.expand_knots(tt, 5)
    for x in np.concatenate([rng.uniform(-0.1, 1.1, 500), tt]):
        assert cbo.find_interval(t2, 5, x, 5) == cbo.find_interval_bisect(t2, 5, x)


def test_within_support():
    # test/bsplines/knot_sets.jl:206-250
    x = cbo.julia_range(0, 1, 21)
    tt = cbo.linear_knots(0, 1, 4)
    sup = [cbo.within_support(x, tt, 1, 1, 1, j) for j in range(1, 5)]
    assert [len(s) for s in sup] == [1, 1, 1, 1]
    assert [list(s[0][0]) for s in sup] == [list(range(1, 6)), list(range(6, 11)), list(range(11, 16)), list(range(16, 22))]
    tt = cbo.linear_knots(0, 1, 3)
    sup = [cbo.within_support(x, tt, 2, 2, 2, j) for j in range(1, 5)]
    assert [(list(a), i) for a, i in sup[1]] == [(list(range(1, 8)), 2), (list(range(8, 15)), 3)]
    assert [(list(a), i) for a, i in sup[3]] == [(list(range(15, 22)), 4)]


def test_lgwt():
    # test/bsplines/knot_sets.jl:254-259
    t = cbo.expand_knots(cbo.linear_knots(0, 1, 2), 1)
    x, w = cbo.lgwt(t, 2)
    assert np.all(w == 0.25)
    np.testing.assert_allclose(x, np.array([-1, 1, -1, 1]) / (4 * np.sqrt(3)) + np.array([1, 1, 3, 3]) / 4, rtol=1e-13)
    # docs jldoctest src/quadrature.jl:53-57
    x, w = cbo.lgwt(cbo.expand_knots(cbo.exp_knots(-4.0, 2.0, 7), 2), 2)
    np.testing.assert_allclose(x[:4], [2.11325e-5, 7.88675e-5, 0.000290192, 0.000809808], rtol=6e-6)
    np.testing.assert_allclose(w[-2:], [45.0, 45.0], rtol=1e-14)
    assert cbo.num_quadrature_points(7, 3) == 8 and cbo.num_quadrature_points(10, 3) == 11


def test_gauss_rules():
    for n in (2, 3, 8, 11, 20):
        x, w = cbo.gausslegendre(n)
        xr, wr = np.polynomial.legendre.leggauss(n)
        np.testing.assert_allclose(x, xr, atol=4e-16)
        np.testing.assert_allclose(w, wr, rtol=1e-13)
        for p in range(0, 2 * n, 2):
            assert abs(np.sum(w * x ** p) - 2 / (p + 1)) < 1e-14
    for n in (2, 3, 4, 5, 10, 16):
        x, w = cbo.gausslobatto(n)
        assert x[0] == -1 and x[-1] == 1 and abs(np.sum(w) - 2) < 1e-14
        for p in range(0, 2 * n - 3, 2):
            assert abs(np.sum(w * x ** p) - 2 / (p + 1)) < 1e-14
    x, w = cbo.gausslobatto(3)
    np.testing.assert_allclose(w, [1 / 3, 4 / 3, 1 / 3], rtol=1e-15)


def test_mass_matrix_discontinuous(golden):
    g = golden["mass_k3_discontinuous"]
    t, q, x, w = bspline(g["tt"], g["k"], g["ml"], g["mr"])
    band, l, u = cbo.bspline_assemble(t, 3, t, 3, x, w, q)
    nf = len(t) - 3
    S = cbo.band_to_dense(band, nf, nf, l, u)
    np.testing.assert_allclose(S, frac_matrix(g["matrix"]), atol=4e-16, rtol=0)
    # test/bsplines/splines.jl:6-7,89-96: zero at x=0 (ml=1), finite with repeated knots
    iv, vals = cbo.bspline_eval_ell(t, 3, g["mr"], [0.0, 0.5, 1.0])
    assert np.all(vals[0] == 0)
    iv, vals = cbo.bspline_eval_ell(t, 3, g["mr"], x)
    assert np.all(np.isfinite(vals))
    # B.B == B[B.x,:]  (:9)
    colptr, rowval, nzval = cbo.bspline_eval_csc(t, 3, g["mr"], x)
    Bm = cbo.basis_matrix(t, 3, x, q)
    D = np.zeros_like(Bm)
    for c in range(nf):
        D[rowval[colptr[c] - 1:colptr[c + 1] - 1] - 1, c] = nzval[colptr[c] - 1:colptr[c + 1] - 1]
    assert np.array_equal(D, Bm)


DERIV = ["k3_D", "k3_DD", "k34_D", "k34_DD", "k43_D", "k43_DD",
         "k7_D", "k7_DD", "k7_12", "k7_22", "k7_23", "k7_33"]


@pytest.mark.parametrize("name", DERIV)
def test_derivative_matrices(golden, name):
    g = golden[name]
    kl, kr = g["kl"], g["kr"]
    tt = cbo.linear_knots(g["a0"], g["b0"], g["N"])
    if kl == kr:
        q = cbo.num_quadrature_points(kl, 3)
    else:  # test/bsplines/derivatives.jl:40-45
        q = max(cbo.num_quadrature_points(3, 1), cbo.num_quadrature_points(4, 1))
    tL, tR = cbo.expand_knots(tt, kl), cbo.expand_knots(tt, kr)
    x, w = cbo.lgwt(tL, q)
    x2, w2 = cbo.lgwt(tR, q)
    assert np.array_equal(x, x2) and np.array_equal(w, w2)
    band, l, u = cbo.bspline_assemble(tL, kl, tR, kr, x, w, q, g["a"], g["b"])
    assert (l, u) == (max(kl, kr) - 1, max(kl, kr) - 1)
    ref = frac_matrix(g["matrix"])
    A = cbo.band_to_dense(band, ref.shape[0], ref.shape[1], l, u)
    # `≈` in the reference is norm-wise (rtol sqrt(eps)); the restatement is far tighter
    assert np.max(np.abs(A - ref)) <= 2e-13 * np.max(np.abs(ref))


def test_docs_eval(golden):
    g = golden["docs"]["usage_eval"]
    k = g["k"]
    tt = cbo.linear_knots(g["a"], g["b"], g["N"])
    t = cbo.expand_knots(tt, k)
    nf = len(t) - k
    iv, vals = cbo.bspline_eval_ell(t, k, k, [0.5])
    row = np.zeros(nf)
    row[iv[0] - k:iv[0]] = vals[0]
    np.testing.assert_allclose(row, g["B_0.5_row"], rtol=0, atol=2e-17)
    colptr, rowval, nzval = cbo.bspline_eval_csc(t, k, k, g["x_quarter"])
    entries = [(int(r), c + 1) for c in range(nf) for r in rowval[colptr[c] - 1:colptr[c + 1] - 1]]
    assert entries == [(r, c) for r, c, _ in g["csc_quarter"]]  # 14 stored entries, same order
    np.testing.assert_allclose(nzval, [v for _, _, v in g["csc_quarter"]], rtol=6e-6)
    colptr, rowval, nzval = cbo.bspline_eval_csc(t, k, k, cbo.julia_range(0, 1, 11), 0, 1)
    assert list(rowval) == [1, 2] and list(nzval) == [1.0, 0.125]
    # spline evaluation, docs/src/bsplines/splines.md:54-96
    s = golden["docs"]["spline_eval"]
    c = np.sin(np.arange(1, 9, dtype=np.float64))
    got = [cbo.deboor(t, k, c, x, cbo.find_interval(t, k, x, k)) for x in s["x"]]
    np.testing.assert_allclose(got, s["s"], rtol=4e-16)
    iv, vals = cbo.bspline_eval_ell(t, k, k, g["x_quarter"])
    chic = []
    for p in range(5):
        rowv = np.zeros(nf)
        rowv[iv[p] - k:iv[p]] = vals[p]
        chic.append(rowv @ c)
    np.testing.assert_allclose(chic, s["chi_c_quarter"], rtol=4e-16)
    # B'B k=4, docs/src/bsplines/usage.md:38-49
    q = cbo.num_quadrature_points(k, 3)
    x, w = cbo.lgwt(t, q)
    band, l, u = cbo.bspline_assemble(t, k, t, k, x, w, q)
    S = cbo.band_to_dense(band, nf, nf, l, u)
    np.testing.assert_allclose(S[:4, 0], golden["docs"]["mass_k4"]["col1"], rtol=6e-6)
    np.testing.assert_allclose(S[3, 3], golden["docs"]["mass_k4"]["diag4"], rtol=6e-6)


def test_piecewise_closed_forms():
    # test/bsplines/splines.jl:28-56
    k = 3
    t = cbo.expand_knots(cbo.linear_knots(0, 1, 2), k)
    for x in (cbo.julia_range(0, 1, 50), cbo.julia_range(-0.5, 0.6, 40), cbo.julia_range(0.5, 1.6, 40)):
        colptr, rowval, nzval = cbo.bspline_eval_csc(t, k, k, x)
        chi = np.zeros((len(x), 4))
        for c in range(4):
            chi[rowval[colptr[c] - 1:colptr[c + 1] - 1] - 1, c] = nzval[colptr[c] - 1:colptr[c + 1] - 1]
        Bt = np.zeros_like(chi)
        Bt[:, 0] = (x >= 0) * (x < 0.5) * ((2 * x - 1) ** 2)
        Bt[:, 1] = (x >= 0) * (x < 0.5) * (2 / 3 * (1 - (3 * x - 1) ** 2)) + (x >= 0.5) * (x < 1) * (2 * (x - 1) ** 2)
        Bt[:, 2] = (x >= 0) * (x < 0.5) * (2 * x ** 2) + (x >= 0.5) * (x < 1) * (2 / 3 * (1 - (3 * x - 2) ** 2))
        Bt[:, 3] = (x >= 0.5) * (x <= 1) * ((2 * x - 1) ** 2)
        for j in range(4):
            d = np.sqrt(np.sum((chi[:, j] - Bt[:, j]) ** 2))
            assert d < 1e-15
    _, vals = cbo.bspline_eval_ell(t, k, k, cbo.julia_range(-1, -0.5, 10))
    assert np.all(vals == 0)


def test_restricted_mass_is_subblock():
    # test/bsplines/splines.jl:99-113
    k = 4
    t = cbo.expand_knots(cbo.linear_knots(0, 1, 5), k)
    q = cbo.num_quadrature_points(k, 3)
    x, w = cbo.lgwt(t, q)
    nf = len(t) - k
    band, l, u = cbo.bspline_assemble(t, k, t, k, x, w, q)
    S = cbo.band_to_dense(band, nf, nf, l, u)
    b2, l2, u2 = cbo.bspline_assemble(t, k, t, k, x, w, q, rowL0=1, m=nf - 2, colR0=1, n=nf - 2)
    assert np.array_equal(cbo.band_to_dense(b2, nf - 2, nf - 2, l2, u2), S[1:-1, 1:-1])
    b3, l3, u3 = cbo.bspline_assemble(t, k, t, k, x, w, q, rowL0=1, m=nf - 2, colR0=0, n=nf)
    assert (l3, u3) == (k - 2, k)
    assert np.array_equal(cbo.band_to_dense(b3, nf - 2, nf, l3, u3), S[1:-1, :])


def test_readme_bspline_k5(golden):
    g = golden["docs"]["readme_bspline_k5"]
    k = g["k"]
    t = cbo.expand_knots(cbo.linear_knots(g["a"], g["b"], g["N"]), k)
    q = cbo.num_quadrature_points(k, 3)
    x, w = cbo.lgwt(t, q)
    nf = len(t) - k
    n = nf - 2
    S, l, u = cbo.bspline_assemble(t, k, t, k, x, w, q, rowL0=1, m=n, colR0=1, n=n)
    S = cbo.band_to_dense(S, n, n, l, u)
    np.testing.assert_allclose(S[0, :5], g["mass_row1"], rtol=6e-6)
    T, l, u = cbo.bspline_assemble(t, k, t, k, x, w, q, 1, 1, rowL0=1, m=n, colR0=1, n=n)
    T = cbo.band_to_dense(T, n, n, l, u)
    np.testing.assert_allclose(T[0, :5], g["laplacian_row1"], rtol=6e-6)
    np.testing.assert_allclose(T[2, :], g["laplacian_row3"], rtol=6e-6)


def test_readme_fedvr(golden):
    g = golden["docs"]["readme_fedvr"]
    B = cbo.FEDVROracle(cbo.julia_range(g["a"], g["b"], g["nel"] + 1), g["order"])
    assert B.N == 17
    A = B.dense(B.blocks(2), *g["sel"])
    assert A.shape == (15, 15)
    np.testing.assert_allclose(A[:3, :3], g["block"], rtol=6e-6)
    np.testing.assert_allclose(A[:3, 3], g["east"], rtol=6e-6)
    np.testing.assert_allclose(A[3, 3], g["bridge_diag"], rtol=6e-6)
    np.testing.assert_allclose(A[3, 7], g["bridge_corner"], rtol=6e-6)
    np.testing.assert_allclose(A, A.T, atol=1e-9)
    assert B.block_structure(*g["sel"]) == [3, 1, 3, 1, 3, 1, 3]


def test_fedvr_block_structure():
    # test/fedvr/block_structure.jl:28-80
    t = cbo.julia_range(-10.0, 10.0, 10)
    B = cbo.FEDVROracle(t, [4, 3, 4, 2, 5, 4, 2, 4, 8])
    assert B.block_structure() == [3, 1, 1, 1, 2, 1, 1, 3, 1, 2, 1, 1, 2, 1, 7]
    N = B.N
    a = int(np.floor(0.4 * N + 0.5))
    b = int(np.floor(0.6 * N + 0.5)) + 1
    exp = {(1, b): [3, 1, 1, 1, 2, 1, 1, 3, 1, 2, 1, 1], (2, b): [2, 1, 1, 1, 2, 1, 1, 3, 1, 2, 1, 1],
           (3, b): [1, 1, 1, 1, 2, 1, 1, 3, 1, 2, 1, 1], (4, b): [2, 1, 2, 1, 1, 3, 1, 2, 1, 1],
           (1, b - 1): [3, 1, 1, 1, 2, 1, 1, 3, 1, 3], (1, b - 2): [3, 1, 1, 1, 2, 1, 1, 3, 1, 2],
           (1, b - 3): [3, 1, 1, 1, 2, 1, 1, 3, 1, 1], (1, b - 4): [3, 1, 1, 1, 2, 1, 1, 4],
           (a, N): [3, 1, 2, 1, 1, 2, 1, 7], (a, N - 1): [3, 1, 2, 1, 1, 2, 1, 6],
           (a, N - 7): [3, 1, 2, 1, 1, 3], (a, N - 8): [3, 1, 2, 1, 1, 2]}
    for (lo, hi), bs in exp.items():
        assert B.block_structure(lo, hi) == bs, (lo, hi)
    assert B.elements(1, b) == (1, 8) and B.elements(a, N - 8) == (5, 8)
    B1 = cbo.FEDVROracle(cbo.julia_range(-10.0, 10.0, 2), 8)
    assert B1.block_structure() == [8] and B1.block_structure(2, 7) == [6]
    assert B1.block_bandwidths() == ([0], [0])
    # test/fedvr/block_structure.jl:20-25 (the l,u the skyline matrix must at least have)
    l, u = cbo.FEDVROracle(cbo.julia_range(1.0, 7.0, 7), [2, 2, 3, 4, 2, 4]).block_bandwidths()
    assert l[:-1][:7] == [1, 1, 2, 1, 2, 1, 1][:len(l) - 1][:7]


def test_fedvr_restricted_is_subblock():
    # test/fedvr/derivatives.jl:46-111
    for order, sel, N in [(4, (2, 12), 5), (4, (1, 13), 5), ([5, 4, 2, 2], (5, 9), 5), ([5, 4, 3, 2], (5, 10), 5),
                          (2, (2, 3), 5), (2, (1, 5), 5), (5, (2, 4), 2)]:
        B = cbo.FEDVROracle(cbo.julia_range(0, 20, N), order)
        for nder in (1, 2):
            blk = B.blocks(nder)
            full = B.dense(blk)
            sub = B.dense(blk, *sel)
            assert np.array_equal(full[sel[0] - 1:sel[1], sel[0] - 1:sel[1]], sub)


def test_fd_goldens(golden):
    g = golden["docs"]["fd"]
    # test/fd/derivatives.jl:2-22: FiniteDifferences(5, 1.0)
    al, be, ga, de = cbo.fd_cartesian(5)
    dl, d, du = cbo.fd_gradient(de, ga, 1.0)
    assert np.all(d == 0) and np.all(du == 0.5) and np.all(dl == -0.5)
    dv, ev = cbo.fd_laplacian(al, be, 1.0)
    assert np.all(dv == -2) and np.all(ev == 1)
    dv, ev = cbo.fd_laplacian(*cbo.fd_cartesian(3)[:2], 0.25)
    assert dv[0] == g["cartesian_N3_dx025"]["dv"] and ev[0] == g["cartesian_N3_dx025"]["ev"]
    # README.md:110-114 staggered N=3, rho=0.25
    al, be, ga, de = cbo.fd_staggered_uniform(3)
    dv, ev = cbo.fd_laplacian(al, be, 0.25)
    np.testing.assert_allclose(ev, g["staggered_N3_rho025"]["ev"], rtol=6e-6)
    np.testing.assert_allclose(dv[1:], g["staggered_N3_rho025"]["dv"][1:], rtol=6e-6)
    assert dv[0] == g["staggered_N3_rho025"]["dv_1_uncorrected"]
    np.testing.assert_allclose(cbo.laplacian_correction(dv, 0.25, 1.0, 0)[0], -65.25, rtol=1e-15)
    # uniform grid through the non-uniform formulas agrees with the closed forms
    r = 0.25 * np.arange(1, 21) - 0.125
    a2, b2, _, d2 = cbo.fd_staggered_nonuniform(r)
    a1, b1, _, d1 = cbo.fd_staggered_uniform(20)
    np.testing.assert_allclose(a2 * 0.25 ** 2, a1, rtol=1e-13)
    np.testing.assert_allclose(b2 * 0.25 ** 2, b1, rtol=1e-13)


def test_apply_matches_dense():
    rng = np.random.default_rng(3)
    B = cbo.FEDVROracle(cbo.julia_range(0, 3, 7), [4, 3, 5, 2, 6, 4])
    blk = B.blocks(2)
    X = rng.standard_normal((B.N, 5))
    Y0 = rng.standard_normal((B.N, 5))
    Y = B.apply(blk, X, 0.5, -2.0, Y0.copy(order="F"))
    np.testing.assert_allclose(Y, 0.5 * B.dense(blk) @ X - 2.0 * Y0, rtol=1e-13, atol=1e-10)
    k = 5
    t = cbo.expand_knots(cbo.linear_knots(0, 1, 9), k)
    q = cbo.num_quadrature_points(k)
    x, w = cbo.lgwt(t, q)
    nf = len(t) - k
    band, l, u = cbo.bspline_assemble(t, k, t, k, x, w, q, 0, 1, rowL0=1, m=nf - 3, colR0=0, n=nf)
    A = cbo.band_to_dense(band, nf - 3, nf, l, u)
    X = rng.standard_normal((nf, 4))
    np.testing.assert_allclose(cbo.banded_apply(band, nf - 3, nf, l, u, X), A @ X, rtol=1e-13, atol=1e-13)
    dl, d, du = rng.standard_normal(9), rng.standard_normal(10), rng.standard_normal(9)
    T = np.diag(d) + np.diag(dl, -1) + np.diag(du, 1)
    X = rng.standard_normal((10, 3))
    np.testing.assert_allclose(cbo.tridiag_apply(dl, d, du, X), T @ X, rtol=1e-13, atol=1e-13)
    np.testing.assert_allclose(cbo.tridiag_apply(dl, d, du, X, threads=True), T @ X, rtol=1e-13, atol=1e-13)


def test_eval_literal_equals_linear_cost():
    rng = np.random.default_rng(5)
    tt = cbo.linear_knots(0, 1, 12)
    k = 4
    t = cbo.expand_knots(tt, k)
    x = np.concatenate([rng.uniform(-0.1, 1.1, 200), tt])
    lit = cbo.bspline_eval_literal(tt, k, k, k, x)
    iv, vals = cbo.bspline_eval_ell(t, k, k, x)
    nf = len(t) - k
    D = np.zeros((len(x), nf))
    for p in range(len(x)):
        if iv[p]:
            D[p, iv[p] - k:iv[p]] = vals[p]
    assert np.array_equal(D, lit)


def test_density_pinv_branch_projects_products():
    # src/densities.jl:92-99: with enough nodes the product of two splines of order k
    # projected by least squares reproduces a spline that is itself in the space when
    # one factor is constant.
    k = 4
    t = cbo.expand_knots(cbo.linear_knots(0, 1, 6), k)
    q = cbo.num_quadrature_points(k)
    x, w = cbo.lgwt(t, q)
    V = cbo.basis_matrix(t, k, x, q)
    nf = len(t) - k
    one = np.ones(nf)  # partition of unity: sum_j B_j = 1
    c = np.sin(np.arange(1, nf + 1))
    rho = cbo.density_bspline(V, one, c)
    np.testing.assert_allclose(rho, c, rtol=0, atol=1e-13)


def test_density_docs_transcript(golden):
    """docs/src/densities.md:119-182: cf = R \\ f, cg = R \\ g, rho = Density(R*cf, R*cg) (k = 7, 71 intervals on
    0..10): the printed first / last 16 coefficients and norm(rho - R \\ (f g)).  Pins ``B \\ f`` (bspline_ldiv) and
    the general Density branch end to end; the transcript went through a sparse QR, so agreement is ~1e-10."""
    g = golden["docs"]["density_k7"]
    tt = cbo.linear_knots(g["a"], g["b"], g["N"])
    t, q, x, w = bspline(tt, g["k"])
    V = cbo.basis_matrix(t, g["k"], x, q)
    f, h = np.sin(2 * np.pi * x), x * np.exp(-x)
    cf, cg, ch = cbo.bspline_ldiv(V, f), cbo.bspline_ldiv(V, h), cbo.bspline_ldiv(V, f * h)
    rho = cbo.density_bspline(V, cf, cg)
    assert rho.shape == (77,)
    np.testing.assert_allclose(rho[:16], g["first16"], rtol=0, atol=2e-10)
    np.testing.assert_allclose(rho[-16:], g["last16"], rtol=0, atol=2e-10)
    assert abs(np.linalg.norm(rho - ch) - g["norm_rho_minus_ch"]) <= 1e-9


def test_more_docs_transcripts(golden):
    """Further REPL transcripts of the reference docs: the B-spline matrix of sin(2 pi r) (docs/src/diagonal_operators.md),
    the first-derivative matrix for k = 3 (docs/src/bsplines/operators.md), 17-digit ExpKnotSet knots (docs/src/index.md)."""
    g = golden["docs"]["diag_op_k7"]
    k = g["k"]
    t = cbo.expand_knots(cbo.linear_knots(g["a"], g["b"], g["N"]), k)
    q = cbo.num_quadrature_points(k, 3)
    x, w = cbo.lgwt(t, q)
    nf = len(t) - k
    assert nf == 77
    band, l, u = cbo.bspline_assemble(t, k, t, k, x, w, q, op=np.sin(2 * np.pi * x))
    F = cbo.band_to_dense(band, nf, nf, l, u)
    np.testing.assert_allclose(F[:7, :6], g["top_left"], rtol=6e-6)
    np.testing.assert_allclose(F[73:, 73:], g["bottom_right"], rtol=6e-6)
    g = golden["docs"]["grad_k3"]
    t = cbo.expand_knots(cbo.linear_knots(g["a"], g["b"], g["N"]), g["k"])
    q = cbo.num_quadrature_points(g["k"], 3)
    x, w = cbo.lgwt(t, q)
    band, l, u = cbo.bspline_assemble(t, g["k"], t, g["k"], x, w, q, 0, 1)
    G = cbo.band_to_dense(band, 5, 5, l, u)
    np.testing.assert_allclose(G, g["matrix"], rtol=6e-6, atol=1e-15)      # the printed diagonal is 1e-17 round-off
    g = golden["docs"]["expknots_k5"]
    t = cbo.expand_knots(cbo.exp_knots(g["a"], g["b"], g["N"]), g["k"])
    np.testing.assert_allclose(t, g["knots"], rtol=4e-16, atol=0)


def test_docs_basis_table_and_mass_k7(golden):
    """docs/src/bsplines/inner_products.md:33-90: the basis functions at the quadrature nodes (B.B, 560 stored entries)
    and the mass matrix B.S for k = 7 on 10 intervals (q = 8 nodes per interval)."""
    g = golden["docs"]["basis_table_k7"]
    k = g["k"]
    t = cbo.expand_knots(cbo.linear_knots(g["a"], g["b"], g["N"]), k)
    q = cbo.num_quadrature_points(k, 3)
    assert q == 8
    x, w = cbo.lgwt(t, q)
    V = cbo.basis_matrix(t, k, x, q)
    assert V.shape == (80, 16) and np.count_nonzero(V) == g["nnz"]
    np.testing.assert_allclose(V[:8, 0], g["col1_rows1_8"], rtol=6e-6)
    np.testing.assert_allclose(V[:8, 1], g["col2_rows1_8"], rtol=6e-6)
    np.testing.assert_allclose(V[72:, 15][::-1], g["col1_rows1_8"], rtol=6e-6)      # mirror image
    band, l, u = cbo.bspline_assemble(t, k, t, k, x, w, q)
    S = cbo.band_to_dense(band, 16, 16, l, u)
    np.testing.assert_allclose(S[:8, :6], g["S_top_left"], rtol=6e-6, atol=1e-300)


def test_hydrogen_staggered_fd_levels_oracle():
    """test/fd/derivatives.jl:182-208 on the oracle's stencils: StaggeredFiniteDifferences(N, 0.25), r_max = 300,
    Laplacian with the Coulomb correction at the origin (Z = 1, l = 0), V = -1/r: |E_1 + 1/2| < 3e-5 and the ten
    lowest levels to 1e-3 relative."""
    from scipy.linalg import eigvalsh_tridiagonal
    rmax, rho = 300, 0.25
    N = int(np.ceil(rmax / rho + 0.5))
    al, be, _, _ = cbo.fd_staggered_uniform(N)
    dv, ev = cbo.fd_laplacian(al, be, rho)
    dv = cbo.laplacian_correction(dv, rho, 1.0, 0)
    r = rho * np.arange(1, N + 1) - rho / 2
    vals = eigvalsh_tridiagonal(dv / -2 - 1.0 / r, ev / -2, select="i", select_range=(0, 9))
    lam = -1.0 / (2.0 * np.arange(1, 11) ** 2)
    err = vals - lam
    assert abs(err[0]) < 3e-5
    assert np.all(np.abs(err) / np.abs(1e-10 + np.abs(lam)) < 1e-3)
    # without the correction the ground state is visibly off: the test does pin laplacian_correction!
    dv0, _ = cbo.fd_laplacian(al, be, rho)
    v0 = eigvalsh_tridiagonal(dv0 / -2 - 1.0 / r, ev / -2, select="i", select_range=(0, 0))
    assert abs(v0[0] - lam[0]) > 1e-3


# ----------------------------------------------------------------------------------------------
# FE-DVR element arithmetic pinned against a 50-digit evaluation of the same formulas
# ----------------------------------------------------------------------------------------------
def _mp_gausslobatto(n, mp):
    """Gauss-Lobatto rule with n points in 50-digit arithmetic: -1, the roots of P'_{n-1}, 1;
    w = 2 / (n (n-1) P_{n-1}(x)^2) (what FastGaussQuadrature.gausslobatto tabulates)."""
    N = n - 1
    if n == 2:
        return [mp.mpf(-1), mp.mpf(1)], [mp.mpf(1), mp.mpf(1)]
    guess = np.sort(np.polynomial.legendre.Legendre.basis(N).deriv().roots().real)     # independent of the oracle
    # P'_N = N (P_{N-1} - x P_N) / (1 - x^2): Newton on the numerator from the float64 guesses
    dP = lambda x: mp.legendre(N - 1, x) - x * mp.legendre(N, x)
    inner = [mp.findroot(dP, mp.mpf(float(g))) for g in guess]
    assert all(inner[i] < inner[i + 1] for i in range(len(inner) - 1)) and all(abs(float(r) - g) < 1e-8 for r, g in zip(inner, guess))
    xs = [mp.mpf(-1)] + inner + [mp.mpf(1)]
    ws = [2 / (N * (N + 1) * mp.legendre(N, x) ** 2) for x in xs]
    return xs, ws


@pytest.mark.parametrize("order", list(range(2, 17)))
def test_fedvr_element_matrices_pinned_to_mpmath(order):
    """cbo_lagrangeder / cbo_fedvr_element (src/fedvr_derivatives.jl:15-57) and the Gauss-Lobatto grid
    (src/quadrature.jl:81-85, src/fedvr.jl:25-37) against mpmath at 50 digits: <= 1e-13 relative for
    orders 2..16 (round 1 pinned these to 6 digits through the README printout only)."""
    import mpmath as mp
    mp.mp.dps = 50
    a, b, c = 0.3, 1.7, 2.9                              # two elements of different width: one bridge
    O = cbo.FEDVROracle(np.array([a, b, c]), order)
    xi, wi = _mp_gausslobatto(order, mp)
    xg, wg = cbo.gausslobatto(order)
    assert max(abs(mp.mpf(float(v)) - r) for v, r in zip(xg, xi)) < mp.mpf("2e-16")
    assert max(abs(mp.mpf(float(v)) - r) / r for v, r in zip(wg, wi)) < mp.mpf("1e-14")
    els = []
    for lo, hi in ((a, b), (b, c)):
        lo, hi = mp.mpf(lo), mp.mpf(hi)
        els.append(([lo + (hi - lo) * (x + 1) / 2 for x in xi], [(hi - lo) * w / 2 for w in wi]))
    bridge = 1 / mp.sqrt(els[0][1][-1] + els[1][1][0])
    for e, (xe, we) in enumerate(els):
        ne = [1 / mp.sqrt(w) for w in we]
        if e == 0:
            ne[-1] = bridge
        else:
            ne[0] = bridge
        o = order
        Lp = [[None] * o for _ in range(o)]
        for m in range(o):
            Lp[m][m] = (int(m == o - 1) - int(m == 0)) / (2 * we[m])
            for mp_ in range(o):
                if mp_ == m:
                    continue
                f = mp.mpf(1)
                for kk in range(o):
                    if kk in (m, mp_):
                        continue
                    f *= (xe[mp_] - xe[kk]) / (xe[m] - xe[kk])
                Lp[m][mp_] = f / (xe[m] - xe[mp_])
        for nder in (1, 2):
            D = np.zeros((o, o))
            for m in range(o):
                for mp_ in range(o):
                    s = mp.mpf(0)
                    for qd in range(o):
                        lt = (mp.mpf(1) if m == qd else mp.mpf(0)) if nder == 1 else -Lp[m][qd]
                        s += lt * we[qd] * Lp[mp_][qd]
                    D[m, mp_] = float(ne[m] * ne[mp_] * s)
            blk = O.blocks(nder)[O.boff[e]:O.boff[e] + o * o].reshape(o, o).T     # column-major o x o
            err = np.max(np.abs(blk - D)) / np.max(np.abs(D))
            assert err <= 1e-13, (order, e, nder, err)
        # nodes and normalisation too
        xs = O.x[O.estart[e]:O.estart[e] + o]
        assert max(abs(mp.mpf(float(v)) - r) for v, r in zip(xs, xe)) < mp.mpf("1e-15")
        ns = O.n[O.estart[e]:O.estart[e] + o]
        assert max(abs(mp.mpf(float(v)) - r) / r for v, r in zip(ns, ne)) < mp.mpf("1e-14")


def test_linear_cost_assembly_equals_literal_scan():
    """cbo_bspline_assemble visits only the intervals of the row function; the literal all-interval scan
    (cbo_bspline_assemble_literal) must give bit-identical bands (same terms, same order)."""
    for k, ni, ml, mr in [(4, 30, 4, 4), (7, 25, 7, 7), (3, 12, 1, 3), (5, 20, 2, 5)]:
        tt = np.sort(np.random.default_rng(k).uniform(0, 5, ni + 1))
        tt[3] = tt[4]                                    # an empty interior interval
        t = cbo.expand_knots(tt, k, ml, mr)
        q = cbo.num_quadrature_points(k, 3)
        x, w = cbo.lgwt(t, q)
        op = np.cos(x)
        for a, b, o in [(0, 0, None), (1, 1, None), (0, 1, None), (0, 0, op), (1, 1, op)]:
            b1, l, u = cbo.bspline_assemble(t, k, t, k, x, w, q, a, b, o)
            b2, _, _ = cbo.bspline_assemble_literal(t, k, t, k, x, w, q, a, b, o)
            assert np.array_equal(b1, b2), (k, a, b)
        b1, _, _ = cbo.bspline_assemble(t, k, t, k, x, w, q, 1, 1, None, 1, len(t) - k - 3, 2, len(t) - k - 4)
        b2, _, _ = cbo.bspline_assemble_literal(t, k, t, k, x, w, q, 1, 1, None, 1, len(t) - k - 3, 2, len(t) - k - 4)
        assert np.array_equal(b1, b2)


def test_density_banded_qr_equals_dense_pinv():
    """The linear-cost least-squares oracle used at BASELINE sizes (Givens QR of the banded Vandermonde matrix)
    against the literal dense pinv of src/densities.jl:96, incl. weights, broadcasting and a restriction."""
    rng = np.random.default_rng(7)
    for k, ni, ml, mr in [(4, 30, 4, 4), (7, 25, 7, 7), (3, 12, 1, 3), (8, 40, 8, 8)]:
        tt = np.sort(rng.uniform(0, 5, ni + 1))
        t = cbo.expand_knots(tt, k, ml, mr)
        q = cbo.num_quadrature_points(k, 3)
        x, w = cbo.lgwt(t, q)
        nf = len(t) - k
        V = cbo.basis_matrix(t, k, x, q)
        f, g = rng.standard_normal((2, nf, 3))
        wv = rng.uniform(0.5, 1.5, len(x))
        for ww in (None, wv):
            r1 = cbo.density_bspline(V, f, g, ww)
            r2 = cbo.density_bspline_qr(t, k, x, q, f, g, ww)
            assert np.max(np.abs(r1 - r2)) <= 1e-12 * np.max(np.abs(r1))
        r1 = cbo.density_bspline(V[:, 1:nf - 2], f[:nf - 3], g[:nf - 3, :1])
        r2 = cbo.density_bspline_qr(t, k, x, q, f[:nf - 3], g[:nf - 3, :1], None, 1, nf - 3)
        assert np.max(np.abs(r1 - r2)) <= 1e-12 * np.max(np.abs(r1))


def test_complex_fedvr_restatement_reduces_to_the_real_one():
    """The numpy transcription of the complex-scaled FE-DVR element matrices (parity unpinned in the reference) must
    reproduce the pinned real restatement when the rotation angle vanishes (e^{i phi} = 1, scaling from the first
    element: same offset t0 = t[1] in element_grid)."""
    t = cbo.julia_range(0.0, 7.0, 9)
    for order in (5, np.array([3, 4, 2, 6, 5, 3, 4, 7])):
        fields = cbo.fedvr_complex_fields(t, order, 0.0, 0.0)
        O = cbo.FEDVROracle(t, order)
        for nder in (1, 2):
            A, _ = cbo.fedvr_complex_dense(fields, nder)
            ref = O.dense(O.blocks(nder))
            assert np.max(np.abs(A.imag)) == 0.0
            assert np.max(np.abs(A.real - ref)) <= 1e-12 * np.max(np.abs(ref))
