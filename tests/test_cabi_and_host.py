"""CPU-side checks: the C-ABI library loads and exports every symbol include/*.h declares,
compute entry points fail loudly without a device, and the host logic of the Python mirror
(knot sets, quadrature rules, FE-DVR grid and block structure) agrees with the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import cbo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "compactbases_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from compactbases_jl_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/compactbases_cuda.h but not exported"
    # and the ctypes table covers the header, so the Python mirror cannot silently miss an entry point
    assert set(names) == set(_lib.EXPORTS)
    assert b"sm_100a" in lib.cb_version()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine without a GPU")
    import compactbases_jl_b200 as cb
    from compactbases_jl_b200 import _lib
    with pytest.raises(RuntimeError):
        cb.BSpline(cb.LinearKnotSet(3, 0, 1, 4))
    # straight through the C ABI: status CB_ERR_CUDA, never a CPU result
    tt = np.array([0.0, 0.5, 1.0]); xi = np.array([0.0]); wr = np.array([2.0])
    h = C.c_void_p()
    rc = _lib.load().cb_bspline_create(C.c_void_p(tt.ctypes.data), 3, 2, 2, 2, C.c_void_p(xi.ctypes.data),
                                       C.c_void_p(wr.ctypes.data), 1, C.byref(h))
    assert rc == _lib.CB_ERR_CUDA and h.value is None
    n = C.c_int(-1)
    assert _lib.load().cb_device_count(C.byref(n)) in (_lib.CB_ERR_CUDA, _lib.CB_OK)


def test_argument_errors_before_any_device_work():
    from compactbases_jl_b200 import _lib
    lib = _lib.load()
    h = C.c_void_p()
    tt = np.array([0.0, 1.0])
    rc = lib.cb_bspline_create(C.c_void_p(tt.ctypes.data), 2, 3, 4, 3, None, None, 0, C.byref(h))
    assert rc == _lib.CB_ERR_ARG and b"Left multiplicity" in lib.cb_last_error()
    rc = lib.cb_tridiag_apply(None, None, None, 5, None, 5, 1, 1.0, 0.0, None, 5, None)
    assert rc == _lib.CB_ERR_ARG
    rc = lib.cb_fd_derivative(7, 2, None, None, 5, 1.0, 0, 5, 0.0, 0, None, None, None, None)
    assert rc == _lib.CB_ERR_ARG and b"unknown scheme" in lib.cb_last_error()
    with pytest.raises(ValueError):
        _lib.check(rc)


def test_knot_sets_match_oracle():
    import compactbases_jl_b200 as cb
    t = cb.LinearKnotSet(7, 0, 100, 1000)
    assert np.array_equal(t.t, cbo.linear_knots(0, 100, 1000))
    assert len(t) == 1013 and t.numfunctions() == 1006
    assert np.array_equal(t.collect(), cbo.expand_knots(t.t, 7))
    e = cb.ExpKnotSet(10, -3.0, 2.0, 200)
    assert np.array_equal(e.t, cbo.exp_knots(-3.0, 2.0, 200)) and e.t[0] == 0.0
    a = cb.ArbitraryKnotSet(3, [0.0, 1, 1, 3, 4, 6], 1, 3)
    assert list(a.collect()) == [0.0, 1, 1, 3, 4, 6, 6, 6]           # test/bsplines/knot_sets.jl:47
    assert a[1] == 0.0 and a[8] == 6.0
    with pytest.raises(IndexError):
        a[9]
    assert list(a.nonempty_intervals()) == list(cbo.nonempty_intervals(a.collect()))
    assert cb.LinearKnotSet(2, 0, 1, 2, 1, 1).numfunctions() == 1
    with pytest.raises(ValueError):
        cb.ArbitraryKnotSet(3, [0.0], 1, 1)


def test_quadrature_rules_match_oracle():
    import compactbases_jl_b200 as cb
    from compactbases_jl_b200.quadrature import fma, lerp
    for n in range(1, 21):
        x, w = cb.gausslegendre(n)
        xo, wo = cbo.gausslegendre(n)
        assert np.max(np.abs(x - xo)) <= 2e-16 and np.max(np.abs(w - wo)) <= 4e-16
        assert abs(w.sum() - 2) < 1e-15
    for n in range(2, 21):
        x, w = cb.gausslobatto(n)
        xo, wo = cbo.gausslobatto(n)
        assert np.max(np.abs(x - xo)) <= 2e-16 and np.max(np.abs(w - wo)) <= 4e-16
    # test/bsplines/knot_sets.jl:254-259: 2-point rule
    x, w = cb.gausslegendre(2)
    np.testing.assert_allclose(x, [-1 / np.sqrt(3), 1 / np.sqrt(3)], rtol=1e-13)
    assert [cb.num_quadrature_points(k, 3) for k in (7, 8, 10)] == [8, 9, 11]
    rng = np.random.default_rng(0)
    a, b, c = rng.standard_normal(500), rng.standard_normal(500), rng.standard_normal(500)
    exact = np.array([cbo._fma(x_, y_, z_) for x_, y_, z_ in zip(a, b, c)])
    assert np.array_equal(fma(a, b, c), exact)
    assert np.array_equal(lerp(np.array([1.0]), np.array([3.0]), np.array([0.5])), [2.0])


def test_host_only_fedvr_grid_matches_oracle():
    from compactbases_jl_b200.fedvr import FEDVR
    for t, order in [(cbo.julia_range(-10.0, 10.0, 10), [4, 3, 4, 2, 5, 4, 2, 4, 8]), (cbo.julia_range(0, 1000, 101), 10)]:
        B = FEDVR(t, order, device="host")
        O = cbo.FEDVROracle(t, order)
        assert np.array_equal(B.x, O.x) and np.array_equal(B.n, O.n) and np.array_equal(B.wcat, O.wcat)
        assert np.array_equal(B.estart, O.estart) and np.array_equal(B.boff, O.boff)
        assert B._block_structure() == O.block_structure()
    B = FEDVR(cbo.julia_range(-10.0, 10.0, 10), [4, 3, 4, 2, 5, 4, 2, 4, 8], device="host")
    N = B.N
    a = int(np.floor(0.4 * N + 0.5))
    b = int(np.floor(0.6 * N + 0.5)) + 1
    # test/fedvr/block_structure.jl:28-80
    assert B._block_structure(1, b) == [3, 1, 1, 1, 2, 1, 1, 3, 1, 2, 1, 1]
    assert B._block_structure(4, b) == [2, 1, 2, 1, 1, 3, 1, 2, 1, 1]
    assert B._block_structure(a, N - 8) == [3, 1, 2, 1, 1, 2]
    assert B.elements(1, b) == (1, 8)


def test_band_algebra_host_logic():
    """The band arithmetic of the host mirror (A - sigma*B, widening, symmetry test, Tridiagonal -> band) needs no
    device: checked on CPU tensors against dense numpy."""
    import torch
    import compactbases_jl_b200 as cb
    from compactbases_jl_b200.operators import _band_axpy, _band_is_symmetric

    def band_of(A, l, u):
        n = A.shape[0]
        data = np.zeros((n, l + u + 1))
        for j in range(n):
            for i in range(max(0, j - u), min(n, j + l + 1)):
                data[j, u + i - j] = A[i, j]
        return cb.BandedMatrix(torch.as_tensor(data), n, n, l, u)

    rng = np.random.default_rng(5)
    n = 9
    A = np.triu(np.tril(rng.standard_normal((n, n)), 2), -1)          # l = 1, u = 2
    B = np.triu(np.tril(rng.standard_normal((n, n)), 1), -3)          # l = 3, u = 1
    Ab, Bb = band_of(A, 1, 2), band_of(B, 3, 1)
    assert np.array_equal(Ab.dense(), A) and np.array_equal(Bb.dense(), B)
    C = _band_axpy(Ab, Bb, -0.7)
    assert (C.l, C.u) == (3, 2) and np.allclose(C.dense(), A - 0.7 * B, rtol=0, atol=1e-15)
    assert np.array_equal(Ab.widened(4, 4).dense(), A)
    assert np.allclose((Ab / -2).dense(), A / -2) and np.allclose((3 * Ab).dense(), 3 * A) and np.allclose((-Ab).dense(), -A)
    S = A + A.T
    assert _band_is_symmetric(band_of(S, 2, 2)) and not _band_is_symmetric(band_of(A, 2, 2)) and not _band_is_symmetric(Ab)
    S2 = S.copy(); S2[3, 4] *= 1 + 1e-15                              # rounding-level asymmetry is accepted
    assert _band_is_symmetric(band_of(S2, 2, 2))
    dl, d, du = rng.standard_normal(n - 1), rng.standard_normal(n), rng.standard_normal(n - 1)
    T = cb.Tridiagonal(*(torch.as_tensor(v) for v in (dl, d, du)))
    assert np.array_equal(cb.as_banded(T).dense(), np.diag(d) + np.diag(dl, -1) + np.diag(du, 1))
    assert np.allclose((T / 4).dense(), T.dense() / 4)
    Dg = cb.Diagonal(torch.as_tensor(d))
    assert np.allclose((cb.as_banded(T) + Dg).dense(), T.dense() + np.diag(d))


def test_implicit_stencils_loop_equals_closed_form():
    """The oracle's restatement of the reference's loop (implicit_stencil_common, src/fd_derivatives.jl:105-125) against
    its closed form on the local steps, and the uniform-grid limits alpha = beta = 1/h^2, M = (1, 10, 1)/12.  (The
    product computes these on the device: cb_fd_implicit_stencils, checked bit for bit in tests/test_gpu_next_rows.py.)"""
    j = np.arange(1, 60)
    r = 0.05 * j + 0.002 * j ** 2
    st = cbo.implicit_stencils_nonuniform(r)
    lp = cbo.implicit_stencils_loop(r)
    for name in st:
        np.testing.assert_array_equal(lp[name], st[name])
    h = 0.25
    u = cbo.implicit_stencils_nonuniform(h * np.arange(1, 12))
    np.testing.assert_allclose(u["alpha"], 1 / h ** 2, rtol=1e-14)
    np.testing.assert_allclose(u["alpha_t"], 1 / h ** 2, rtol=1e-14)
    np.testing.assert_allclose(u["beta"], 1 / h ** 2, rtol=1e-14)
    np.testing.assert_allclose(u["m_d"], 10 / 12, rtol=1e-14)
    np.testing.assert_allclose(u["m_dl"], 1 / 12, rtol=1e-13)
    np.testing.assert_allclose(u["m_du"], 1 / 12, rtol=1e-13)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm) prints exactly one JSON line with
    the contract's keys; ranks other than 0 exit quietly."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-500:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "GB/s" and d["higher_is_better"] is True and d["value"] > 0
    for key in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r1 = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
                        capture_output=True, text=True, timeout=120, cwd=root, env=env)
    assert r1.returncode == 0 and r1.stdout.strip() == ""


def test_block_skyline_layout_host_logic():
    """BlockSkylineMatrix index arithmetic (BlockBandedMatrices' bb_blockstarts / bb_blockstrides restated): every
    non-zero of the oracle's dense FE-DVR operator lies inside the skyline of block_structure / block_bandwidths
    (src/fedvr.jl:121-176), panels tile `data` without gaps, and dense() inverts the placement."""
    import compactbases_jl_b200 as cb
    from oracle import cbo
    cases = [([4, 3, 4, 2, 5, 4, 2, 4, 8], None), ([2, 2, 3, 4, 2, 4], None), (4, (2, 12)), ([5, 4, 2, 2], (5, 9)), (10, (3, 40)),
             ([3, 3], None), (6, (6, 6))]
    for orders, sel in cases:
        nel = len(orders) if not np.isscalar(orders) else 5
        t = cbo.julia_range(0.0, 20.0, nel + 1)
        O = cbo.FEDVROracle(t, orders)
        F = cb.FEDVR(t, orders, device="host")
        a, b = (1, O.N) if sel is None else sel
        R = F if sel is None else F.restrict(a, b)
        rows = R.block_structure()
        assert rows == O.block_structure(a, b)
        l, u = R.block_bandwidths(rows)
        S = cb.BlockSkylineMatrix(rows, l, u, "cpu")
        assert S.n == b - a + 1
        # panels are contiguous and cover data exactly
        ends = S.start + S.stride * np.asarray(rows)
        assert S.start[0] == 0 and np.array_equal(S.start[1:], ends[:-1]) and ends[-1] == S.ndata
        A = O.dense(O.blocks(2), a, b)
        S.data.zero_()
        d = S.data.numpy()
        for J in range(len(rows)):
            for K in range(len(rows)):
                blkA = A[S.cum[K]:S.cum[K + 1], S.cum[J]:S.cum[J + 1]]
                st = S.block_start(K, J)
                if st is None:
                    assert not np.any(blkA), (orders, sel, K, J)      # nothing of the operator outside the skyline
                    continue
                for j in range(rows[J]):
                    d[st + j * S.stride[J]:st + j * S.stride[J] + rows[K]] = blkA[:, j]
        assert np.array_equal(S.dense(), A)


def test_batch_range_entry_matches_the_host_mirror():
    """cb_batch_range (host-only entry: contiguous shares of points / vectors / pairs) against sharding.batch_range:
    shares are contiguous, cover [0, n) and differ by at most one."""
    from compactbases_jl_b200 import _lib
    from compactbases_jl_b200.sharding import batch_range
    lib = _lib.load()
    for n in (0, 1, 7, 8, 1000, 4096, 10 ** 7 + 3):
        for world in (1, 2, 3, 8):
            prev = 0
            for rank in range(world):
                lo, hi = C.c_int64(-1), C.c_int64(-1)
                assert lib.cb_batch_range(n, rank, world, C.byref(lo), C.byref(hi)) == 0
                assert (lo.value, hi.value) == batch_range(n, rank, world)
                assert lo.value == prev and hi.value - lo.value in (n // world, n // world + 1)
                prev = hi.value
            assert prev == n
    lo, hi = C.c_int64(0), C.c_int64(0)
    assert lib.cb_batch_range(10, 3, 3, C.byref(lo), C.byref(hi)) != 0          # rank outside the world
