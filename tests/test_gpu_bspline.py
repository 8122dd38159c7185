"""GPU parity: B-spline evaluation and banded assembly vs the CPU oracle and the
reference's golden matrices, through the C ABI (python host mirror -> ctypes)."""
from fractions import Fraction

import numpy as np
import pytest

from oracle import cbo
from tests.util import RTOL, csc_to_dense, ell_to_dense, oracle_basis, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cb():
    import torch
    assert torch.cuda.is_available()
    import compactbases_jl_b200 as cb
    return cb


def frac_matrix(rows):
    return np.array([[float(Fraction(a, b)) for a, b in r] for r in rows])


KNOTSETS = [
    ("lin_k1", lambda cb: cb.LinearKnotSet(1, 0, 1, 7)),
    ("lin_k2", lambda cb: cb.LinearKnotSet(2, 0, 1, 7)),
    ("lin_k3", lambda cb: cb.LinearKnotSet(3, 0, 1, 2)),
    ("lin_k4", lambda cb: cb.LinearKnotSet(4, 0, 1, 5)),
    ("lin_k7", lambda cb: cb.LinearKnotSet(7, 0, 100, 50)),
    ("lin_k7_single", lambda cb: cb.LinearKnotSet(7, 0, 1, 1)),
    ("exp_k10", lambda cb: cb.ExpKnotSet(10, -3.0, 2.0, 40)),
    ("exp_k5_no0", lambda cb: cb.ExpKnotSet(5, -2.0, 1.0, 12, include0=False)),
    ("arb_repeated", lambda cb: cb.ArbitraryKnotSet(3, [0.0, 1, 1, 3, 4, 6], 1, 3)),
    ("arb_k4_mult", lambda cb: cb.ArbitraryKnotSet(4, [0.0, 0.5, 0.5, 0.5, 1.0, 2.0, 2.0, 3.5], 2, 1)),
    ("lin_k16", lambda cb: cb.LinearKnotSet(16, -1, 1, 9)),
    ("lin_k6_ml1", lambda cb: cb.LinearKnotSet(6, 0, 2, 11, 1, 1)),
]


# knot sets beyond the bank-private replicas (~1100 knots) and beyond the shared-memory table (6144 knots: bucket
# table + global knots)
LARGE_KNOTSETS = [
    ("lin_k5_3000", lambda cb: cb.LinearKnotSet(5, 0, 10, 3000)),
    ("lin_k4_7000", lambda cb: cb.LinearKnotSet(4, -1, 1, 7000)),
    ("exp_k3_9000", lambda cb: cb.ExpKnotSet(3, -3.0, 2.0, 9000)),
]


def edge_points(tt, rng, n=400):
    lo, hi = tt[0], tt[-1]
    span = hi - lo
    pts = [rng.uniform(lo - 0.1 * span, hi + 0.1 * span, n), tt, np.nextafter(tt, -np.inf), np.nextafter(tt, np.inf),
           [lo, hi, np.nan, np.inf, -np.inf]]
    return np.concatenate([np.asarray(p, dtype=np.float64) for p in pts])


@pytest.mark.parametrize("name,make", KNOTSETS + LARGE_KNOTSETS, ids=[k[0] for k in KNOTSETS + LARGE_KNOTSETS])
def test_eval_matches_oracle(cb, name, make):
    t = make(cb)
    B = cb.BSpline(t)
    te = cbo.expand_knots(t.t, t.k, t.ml, t.mr)
    rng = np.random.default_rng(11)
    x = edge_points(t.t, rng)
    for m in range(0, min(t.k, 4)):
        iv, vals = B.eval_ell(x, m)
        iv, vals = iv.cpu().numpy(), vals.cpu().numpy()
        iv_o, vals_o = cbo.bspline_eval_ell(te, t.k, t.mr, x, m)
        assert np.array_equal(iv, iv_o), name                      # interval index: bit-exact
        if m == 0:                                                 # values: element-wise (cancellation-free)
            np.testing.assert_allclose(vals, vals_o, rtol=RTOL, atol=1e-300)
        else:
            assert relerr(vals, vals_o) <= RTOL
        assert np.all(np.isfinite(vals))


@pytest.mark.parametrize("name,make", KNOTSETS, ids=[k[0] for k in KNOTSETS])
def test_eval_csc_structure_exact(cb, name, make):
    t = make(cb)
    B = cb.BSpline(t)
    te = cbo.expand_knots(t.t, t.k, t.ml, t.mr)
    rng = np.random.default_rng(12)
    x = edge_points(t.t, rng, 300)
    x = x[np.isfinite(x)]
    nf = B.nf
    for col0, ncols in [(0, nf), (1, max(nf - 2, 1)) if nf > 2 else (0, nf)]:
        S = B.eval_csc(x, col0, ncols)
        cp, rv, nz = cbo.bspline_eval_csc(te, t.k, t.mr, x, col0, ncols)
        assert np.array_equal(S.colptr.cpu().numpy(), cp)
        assert np.array_equal(S.rowval.cpu().numpy(), rv)
        np.testing.assert_allclose(S.nzval.cpu().numpy(), nz, rtol=RTOL, atol=1e-300)
        assert np.all(S.nzval.cpu().numpy() != 0)                  # exact zeros are not stored


@pytest.mark.parametrize("order", ["random", "sorted", "knots", "clustered"])
def test_eval_csc_counting_sort_large_batches(cb, order):
    """B[x,:] as SparseMatrixCSC through the hand-written counting sort (csrc/csc.cu) on batches of many blocks:
    unsorted points, sorted points (one block's points share a column: long runs), every knot as a point (exact
    zeros are dropped, src/bsplines.jl:217-224) and points clustered in three intervals; full basis and a restriction.
    colptr / rowval bit-exact against the oracle."""
    t = cb.LinearKnotSet(7, 0, 100, 1000)
    B = cb.BSpline(t)
    te = cbo.expand_knots(t.t, 7)
    rng = np.random.default_rng(21)
    if order == "random":
        x = rng.uniform(-1.0, 101.0, 200_003)                        # some points outside the support
    elif order == "sorted":
        x = np.sort(rng.uniform(0.0, 100.0, 150_001))
    elif order == "knots":
        x = np.concatenate([np.asarray(t.t), np.asarray(t.t)[::-1], rng.uniform(0, 100, 5000)])
    else:
        x = np.concatenate([rng.uniform(3.0, 3.1, 40_000), rng.uniform(50.0, 50.2, 40_000), rng.uniform(99.9, 100.0, 20_000)])
        rng.shuffle(x)
    for col0, ncols in [(0, B.nf), (3, B.nf - 10)]:
        S = B.eval_csc(x, col0, ncols)
        cp, rv, nz = cbo.bspline_eval_csc(te, 7, 7, x, col0, ncols)
        assert np.array_equal(S.colptr.cpu().numpy(), cp)
        assert np.array_equal(S.rowval.cpu().numpy(), rv)
        np.testing.assert_allclose(S.nzval.cpu().numpy(), nz, rtol=RTOL, atol=1e-300)
        assert np.all(S.nzval.cpu().numpy() != 0)


def test_docs_golden_eval(cb, golden):
    g = golden["docs"]["usage_eval"]
    B = cb.BSpline(cb.LinearKnotSet(g["k"], g["a"], g["b"], g["N"]))
    row = B[0.5, :]
    np.testing.assert_allclose(row, g["B_0.5_row"], rtol=0, atol=1.2e-16)   # 1 ulp of the 17-digit transcript
    S = B[np.array(g["x_quarter"]), :]
    cp, rv, nz = (v.cpu().numpy() for v in (S.colptr, S.rowval, S.nzval))
    entries = [(int(r), c + 1) for c in range(B.nf) for r in rv[cp[c] - 1:cp[c + 1] - 1]]
    assert entries == [(r, c) for r, c, _ in g["csc_quarter"]]     # 14 stored entries
    np.testing.assert_allclose(nz, [v for _, _, v in g["csc_quarter"]], rtol=6e-6)
    S1 = B[cbo.julia_range(0, 1, 11), 1]
    assert list(S1.rowval.cpu().numpy()) == [1, 2] and list(S1.nzval.cpu().numpy()) == [1.0, 0.125]
    # spline evaluation (B*c)[x], docs/src/bsplines/splines.md:54-96
    s = golden["docs"]["spline_eval"]
    c = np.sin(np.arange(1, 9, dtype=np.float64))
    got = B.evaluate(np.array(s["x"]), cb.ColMajor.from_numpy(c)).numpy()[:, 0]
    np.testing.assert_allclose(got, s["s"], rtol=1e-14)
    chi = S.toarray()
    np.testing.assert_allclose(chi @ c, s["chi_c_quarter"], rtol=1e-14)


def test_eval_edge_cases(cb):
    B = cb.BSpline(cb.LinearKnotSet(3, 0, 1, 2))
    S = B[np.array([]), :]
    assert S.nnz() == 0 and np.all(S.colptr.cpu().numpy() == 1)
    iv, vals = B.eval_ell(cbo.julia_range(-1, -0.5, 10))
    assert np.all(iv.cpu().numpy() == 0) and np.all(vals.cpu().numpy() == 0)
    # piecewise closed forms, test/bsplines/splines.jl:28-56
    for x in (cbo.julia_range(0, 1, 50), cbo.julia_range(-0.5, 0.6, 40), cbo.julia_range(0.5, 1.6, 40)):
        chi = B[x, :].toarray()
        Bt = np.zeros_like(chi)
        Bt[:, 0] = (x >= 0) * (x < 0.5) * ((2 * x - 1) ** 2)
        Bt[:, 1] = (x >= 0) * (x < 0.5) * (2 / 3 * (1 - (3 * x - 1) ** 2)) + (x >= 0.5) * (x < 1) * (2 * (x - 1) ** 2)
        Bt[:, 2] = (x >= 0) * (x < 0.5) * (2 * x ** 2) + (x >= 0.5) * (x < 1) * (2 / 3 * (1 - (3 * x - 2) ** 2))
        Bt[:, 3] = (x >= 0.5) * (x <= 1) * ((2 * x - 1) ** 2)
        assert np.max(np.sqrt(np.sum((chi - Bt) ** 2, axis=0))) < 1e-15


def test_eval_config1_size_properties(cb):
    """BASELINE config 1 at full size: partition of unity + oracle on a sample."""
    t = cb.LinearKnotSet(7, 0, 100, 1000)
    B = cb.BSpline(t)
    rng = np.random.default_rng(1)
    x = rng.uniform(0, 100, 1_000_000)
    iv, vals = B.eval_ell(x)
    s = vals.sum(dim=1).cpu().numpy()
    assert np.max(np.abs(s - 1)) < 1e-14                            # partition of unity (ml = mr = k)
    d1 = B.eval_ell(x, 1)[1].sum(dim=1).cpu().numpy()
    assert np.max(np.abs(d1)) < 1e-9
    te = cbo.expand_knots(t.t, 7)
    sel = rng.integers(0, x.size, 20000)
    iv_o, vals_o = cbo.bspline_eval_ell(te, 7, 7, x[sel])
    assert np.array_equal(iv.cpu().numpy()[sel], iv_o)
    np.testing.assert_allclose(vals.cpu().numpy()[sel], vals_o, rtol=RTOL, atol=1e-300)


def test_quad_table_and_nodes(cb):
    for t in (cb.LinearKnotSet(7, 0, 1, 10), cb.ArbitraryKnotSet(3, [0.0, 1, 1, 3, 4, 6], 1, 3)):
        B = cb.BSpline(t)
        te, q, x, w = oracle_basis(t)
        assert np.array_equal(B.x.cpu().numpy(), x) and np.array_equal(B.w.cpu().numpy(), w)   # lgwt bit-exact
        for m in (0, 1, 2):
            tab = B.basis_functions(m).cpu().numpy()
            assert relerr(tab, cbo.basis_table(te, t.k, x, q, m)) <= RTOL


DERIV = ["k3_D", "k3_DD", "k34_D", "k34_DD", "k43_D", "k43_DD", "k7_D", "k7_DD", "k7_12", "k7_22", "k7_23", "k7_33"]


@pytest.mark.parametrize("name", DERIV)
def test_golden_derivative_matrices(cb, golden, name):
    g = golden[name]
    kl, kr = g["kl"], g["kr"]
    if kl == kr:
        q = cb.num_quadrature_points(kl, 3)
    else:
        q = max(cb.num_quadrature_points(3, 1), cb.num_quadrature_points(4, 1))
    L = cb.BSpline(cb.LinearKnotSet(kl, g["a0"], g["b0"], g["N"]), q)
    R = L if kl == kr else cb.BSpline(cb.LinearKnotSet(kr, g["a0"], g["b0"], g["N"]), q)
    A = cb.diffop(L, R, g["a"], g["b"])
    ref = frac_matrix(g["matrix"])
    assert A.shape == ref.shape and A.bandwidths() == (max(kl, kr) - 1, max(kl, kr) - 1)
    assert relerr(A.dense(), ref) <= RTOL
    if (g["a"], g["b"]) == (0, 1):
        D = cb.Derivative()
        assert np.array_equal((L.T @ D @ R).dense(), A.dense())
    if (g["a"], g["b"]) == (1, 1):
        D = cb.Derivative()
        assert np.array_equal((L.T @ D.T @ D @ R).dense(), A.dense())


def test_golden_mass_discontinuous(cb, golden):
    g = golden["mass_k3_discontinuous"]
    B = cb.BSpline(cb.ArbitraryKnotSet(g["k"], g["tt"], g["ml"], g["mr"]))
    S = (B.T @ B).dense()
    np.testing.assert_allclose(S, frac_matrix(g["matrix"]), atol=1e-15, rtol=0)
    assert np.array_equal(B.S.dense(), S)


@pytest.mark.parametrize("name,make", KNOTSETS, ids=[k[0] for k in KNOTSETS])
def test_assembly_matches_oracle(cb, name, make):
    t = make(cb)
    B = cb.BSpline(t)
    te, q, x, w = oracle_basis(t)
    nf = B.nf
    D = cb.Derivative()
    f = lambda r: np.sin(r) + 2.0
    cases = [("mass", B.T @ B, dict()), ("pot", B.T @ cb.QuasiDiagonal(f) @ B, dict(op=f(x)))]
    if t.k > 1:
        cases.append(("grad", B.T @ D @ B, dict(a=0, b=1)))
        cases.append(("lap", B.T @ D.T @ D @ B, dict(a=1, b=1)))
        cases.append(("wlap", B.T @ D.T @ cb.QuasiDiagonal(f) @ D @ B, dict(a=1, b=1, op=f(x))))
    for label, A, kw in cases:
        band, l, u = cbo.bspline_assemble(te, t.k, te, t.k, x, w, q, **kw)
        ref = cbo.band_to_dense(band, nf, nf, l, u)
        got = A.dense()
        assert got.shape == ref.shape
        assert relerr(got, ref) <= RTOL, (name, label)
        if label in ("mass", "pot", "lap", "wlap"):
            assert relerr(got, got.T) <= 1e-13


def test_restricted_assembly_is_subblock(cb):
    # test/bsplines/splines.jl:99-113
    B = cb.BSpline(cb.LinearKnotSet(4, 0, 1, 5))
    nf = B.nf
    S = (B.T @ B).dense()
    Br = B[:, 2:nf - 1]
    Sr = Br.T @ Br
    assert Sr.bandwidths() == (3, 3) and np.array_equal(Sr.dense(), S[1:-1, 1:-1])
    Sx = Br.T @ B
    assert Sx.bandwidths() == (2, 4) and np.array_equal(Sx.dense(), S[1:-1, :])
    Sy = B.T @ Br
    assert Sy.bandwidths() == (4, 2) and np.array_equal(Sy.dense(), S[:, 1:-1])
    D = cb.Derivative()
    T = (B.T @ D.T @ D @ B).dense()
    assert np.array_equal((Br.T @ D.T @ D @ Br).dense(), T[1:-1, 1:-1])
    B2 = cb.BSpline(cb.LinearKnotSet(2, 0, 1, 6))
    T2 = B2.T @ B2
    assert isinstance(T2, cb.Tridiagonal)                          # k == 2 && m == n
    te, q, x, w = oracle_basis(B2.t)
    band, l, u = cbo.bspline_assemble(te, 2, te, 2, x, w, q)
    assert relerr(T2.dense(), cbo.band_to_dense(band, B2.nf, B2.nf, l, u)) <= RTOL


def test_readme_and_docs_matrices(cb, golden):
    g = golden["docs"]["readme_bspline_k5"]
    B = cb.BSpline(cb.LinearKnotSet(g["k"], g["a"], g["b"], g["N"]))
    Br = B[:, 2:B.nf - 1]
    D = cb.Derivative()
    S = (Br.T @ Br).dense()
    np.testing.assert_allclose(S[0, :5], g["mass_row1"], rtol=6e-6)
    T = (Br.T @ D.T @ D @ Br).dense()
    np.testing.assert_allclose(T[0, :5], g["laplacian_row1"], rtol=6e-6)
    np.testing.assert_allclose(T[2, :], g["laplacian_row3"], rtol=6e-6)
    g4 = golden["docs"]["usage_eval"]
    B4 = cb.BSpline(cb.LinearKnotSet(g4["k"], g4["a"], g4["b"], g4["N"]))
    S4 = (B4.T @ B4).dense()
    np.testing.assert_allclose(S4[:4, 0], golden["docs"]["mass_k4"]["col1"], rtol=6e-6)
    # R'*QuasiDiagonal(sin 2 pi r)*R, docs/src/diagonal_operators.md:100-138
    g = golden["docs"]["diag_op_k7"]
    R = cb.BSpline(cb.LinearKnotSet(g["k"], g["a"], g["b"], g["N"]))
    F = (R.T @ cb.QuasiDiagonal(lambda r: np.sin(2 * np.pi * r)) @ R).dense()
    assert F.shape == (77, 77)
    np.testing.assert_allclose(F[:7, :6], g["top_left"], rtol=6e-6)
    np.testing.assert_allclose(F[73:, 73:], g["bottom_right"], rtol=6e-6)
    # B'D*B for k = 3, docs/src/bsplines/operators.md:31-37
    g = golden["docs"]["grad_k3"]
    B3 = cb.BSpline(cb.LinearKnotSet(g["k"], g["a"], g["b"], g["N"]))
    np.testing.assert_allclose((B3.T @ D @ B3).dense(), g["matrix"], rtol=6e-6, atol=1e-15)
    # B.B (basis functions at the nodes) and B.S for k = 7, 10 intervals, docs/src/bsplines/inner_products.md:33-90
    g = golden["docs"]["basis_table_k7"]
    B7 = cb.BSpline(cb.LinearKnotSet(g["k"], g["a"], g["b"], g["N"]))
    tab = B7.basis_functions(0).cpu().numpy()                       # [interval][node][live function]
    assert tab.shape == (10, 8, 7) and np.count_nonzero(tab) == g["nnz"]
    np.testing.assert_allclose(tab[0, :, 0], g["col1_rows1_8"], rtol=6e-6)
    np.testing.assert_allclose(tab[0, :, 1], g["col2_rows1_8"], rtol=6e-6)
    np.testing.assert_allclose((B7.T @ B7).dense()[:8, :6], g["S_top_left"], rtol=6e-6, atol=1e-300)
    # ExpKnotSet knots to 17 digits, docs/src/index.md:68-92
    g = golden["docs"]["expknots_k5"]
    te = cb.ExpKnotSet(g["k"], g["a"], g["b"], g["N"])
    np.testing.assert_allclose(cbo.expand_knots(te.t, te.k, te.ml, te.mr), g["knots"], rtol=4e-16, atol=0)


def test_coulomb_on_chip_equals_opvals(cb):
    t = cb.ExpKnotSet(10, -3.0, 2.0, 200)
    B = cb.BSpline(t)
    Br = B[:, 2:B.nf - 1]
    V1 = (Br.T @ cb.CoulombPotential(1.0) @ Br).dense()
    te, q, x, w = oracle_basis(t)
    n = B.nf - 2
    band, l, u = cbo.bspline_assemble(te, 10, te, 10, x, w, q, op=-1.0 / x, rowL0=1, m=n, colR0=1, n=n)
    assert relerr(V1, cbo.band_to_dense(band, n, n, l, u)) <= RTOL
    V2 = (Br.T @ cb.QuasiDiagonal(lambda r: -1.0 / r) @ Br).dense()
    assert relerr(V1, V2) <= 1e-14


def test_assemble_part_shards_sum_to_whole(cb):
    """Interval-sharded assembly: partial sums over disjoint interval ranges add up."""
    import ctypes as C
    import torch
    from compactbases_jl_b200 import _lib
    from compactbases_jl_b200.operators import _ptr
    B = cb.BSpline(cb.ExpKnotSet(6, -2.0, 1.0, 57))
    nf, k = B.nf, 6
    full = (B.T @ cb.Derivative().T @ cb.Derivative() @ B)
    W = 2 * k - 1
    acc = torch.zeros((nf, W), dtype=torch.float64, device="cuda")
    nint = B.t.numintervals()
    cuts = [0, 13, 14, 40, nint]
    for s_lo, s_hi in zip(cuts[:-1], cuts[1:]):
        part = torch.zeros((nf, W), dtype=torch.float64, device="cuda")
        _lib.call("cb_bspline_assemble_part", B.handle, B.handle, 1, 1, None, 0, 0.0, 0, nf, 0, nf, k - 1, k - 1,
                  0, nf, s_lo, s_hi, _ptr(part), None)
        acc += part
    assert relerr(acc.cpu().numpy(), full.data.cpu().numpy()) <= 1e-14
    # column sub-range over the full interval range == those columns of the whole (recompute-halo form)
    sub = torch.zeros((20, W), dtype=torch.float64, device="cuda")
    _lib.call("cb_bspline_assemble_part", B.handle, B.handle, 1, 1, None, 0, 0.0, 0, nf, 0, nf, k - 1, k - 1,
              17, 37, 0, nint, _ptr(sub), None)
    assert np.array_equal(sub.cpu().numpy(), full.data.cpu().numpy()[17:37])


def test_errors(cb):
    B3 = cb.BSpline(cb.LinearKnotSet(3, 0, 1, 4))
    B4 = cb.BSpline(cb.LinearKnotSet(4, 0, 1, 5))
    with pytest.raises(ValueError):                                 # ArgumentError, src/bsplines.jl:81
        B3.T @ B4
    with pytest.raises(ValueError):
        cb.LinearKnotSet(3, 0, 1, 4, 4, 3)
    with pytest.raises(NotImplementedError):
        cb.BSpline(cb.LinearKnotSet(17, 0, 1, 4))
    with pytest.raises(IndexError):
        B3[:, 0:3]


@pytest.mark.parametrize("world", [2, 3, 5])
def test_fused_halo_exchange_single_device_emulation(cb, world):
    """cb_bspline_assemble_sharded with halo = exchange (the in-kernel mailbox exchange of csrc/comm.cu): all ranks'
    communicators live on this one GPU (cb_comm_connect_local) and the ranks run one after the other from right to
    left, so that no kernel ever waits for one that has not run (a rank's pushes only go left).  Several rounds
    exercise the double-buffered mailboxes and the ack words.  The real concurrent multi-GPU run is
    scripts/mgpu_check.py (gpurun --gpus N)."""
    import torch
    from compactbases_jl_b200.sharding import Comm, assemble_interval_sharded, interval_shards
    D = cb.Derivative()
    cases = [(cb.BSpline(cb.ExpKnotSet(10, -3.0, 2.0, 3000)), None), (cb.BSpline(cb.LinearKnotSet(7, 0, 10, 400)), (2, 390)),
             (cb.BSpline(cb.LinearKnotSet(4, 0, 1, 64, ml=1, mr=4)), None)]
    for B, sel in cases:
        R = B if sel is None else B[:, sel[0]:sel[1]]
        comms = [Comm(r, world, connect=False) for r in range(world)]
        Comm.connect_local(comms)
        full = {"T": (R.T @ D.T @ D @ R).data, "V": (R.T @ cb.CoulombPotential(1.0) @ R).data}
        shards = interval_shards(B.t.numintervals(), B.k, B.t.ml, R.col0, R.ncols, world)
        for rnd in range(5):
            for name, (a, b_, cz) in (("T", (1, 1, None)), ("V", (0, 0, 1.0))):
                got = torch.full_like(full[name], float("nan"))
                for r in reversed(range(world)):
                    sh = comms[r].shard(R)
                    assert (sh.s_lo, sh.s_hi, sh.own_lo, sh.own_hi, sh.touch_lo) == \
                        (shards[r].s_lo, shards[r].s_hi, shards[r].own_lo, shards[r].own_hi, shards[r].touch_lo)
                    lo, hi, cols = assemble_interval_sharded(R, R, a, b_, r, world, coulomb_Z=cz, halo="fused", comm=comms[r])
                    got[lo:hi] = cols
                torch.cuda.synchronize()
                assert all(c.status() == 0 for c in comms)
                err = float((got - full[name]).abs().max() / full[name].abs().max())
                assert err <= 1e-14, (B.k, world, rnd, name, err)
        # recompute mode through the same entry: bit-identical to the single-GPU band
        got = torch.empty_like(full["T"])
        for r in range(world):
            lo, hi, cols = assemble_interval_sharded(R, R, 1, 1, r, world, halo="recompute", comm=comms[r])
            got[lo:hi] = cols
        assert torch.equal(got, full["T"])


@pytest.mark.parametrize("k,nint,kind", [(2, 9, "lin"), (3, 17, "lin"), (4, 40, "exp"), (7, 300, "lin"), (10, 777, "exp"), (11, 150, "lin"), (12, 60, "lin")])
def test_hamiltonian_pair_matches_separate_assemblies(cb, k, nint, kind):
    """cb_bspline_assemble_pair (values + derivatives from one Cox-de Boor sweep) against the two separate
    materialisations R'*V*R and R'*D'*D*R (bit-identical inside the streaming kernel's range, k <= 11) and against the
    oracle; full basis, a Dirichlet restriction (test/bsplines/operators.jl:15-22), Coulomb / nodal-value / no operator."""
    t = cb.LinearKnotSet(k, 0.5, 7.5, nint) if kind == "lin" else cb.ExpKnotSet(k, -2.0, 1.0, nint)
    B = cb.BSpline(t)
    D = cb.Derivative()
    te = cbo.expand_knots(t.t, k)
    q = cbo.num_quadrature_points(k, 3)
    xq, wq = cbo.lgwt(te, q)
    for R in (B, B[:, 2:B.nf - 1]):
        for op, opf in ((cb.CoulombPotential(2.0), lambda x: -2.0 / x), (cb.QuasiDiagonal(lambda x: np.sin(x) + 2), lambda x: np.sin(x) + 2),
                        (None, None)):
            V, T = cb.hamiltonian_bands(R, op)
            Vs = (R.T @ op @ R) if op is not None else (R.T @ R)
            Ts = R.T @ D.T @ D @ R
            assert V.bandwidths() == (k - 1, k - 1) and T.bandwidths() == (k - 1, k - 1)
            if k == 2:
                pass                                                # Matrix(undef, A, B) is a Tridiagonal for k = 2 (src/bsplines.jl:258-262)
            elif k <= 11:
                assert torch_equal(V.data, Vs.data) and torch_equal(T.data, Ts.data)
            else:
                assert relerr(V.data.cpu().numpy(), Vs.data.cpu().numpy()) <= RTOL and relerr(T.data.cpu().numpy(), Ts.data.cpu().numpy()) <= RTOL
            Vo, _, _ = cbo.bspline_assemble(te, k, te, k, xq, wq, q, 0, 0, None if opf is None else opf(xq), R.col0, R.ncols, R.col0, R.ncols)
            To, _, _ = cbo.bspline_assemble(te, k, te, k, xq, wq, q, 1, 1, None, R.col0, R.ncols, R.col0, R.ncols)
            assert relerr(V.data.cpu().numpy(), Vo) <= RTOL and relerr(T.data.cpu().numpy(), To) <= RTOL


def torch_equal(a, b):
    import torch
    return bool(torch.equal(a, b))
