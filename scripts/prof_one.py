"""Runs ONE kernel family a few times (for ncu captures).  usage: prof_one.py {fedvr|eval|asm3|banded|density|tridiag}"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import compactbases_jl_b200 as cb
from compactbases_jl_b200 import _lib
from compactbases_jl_b200.operators import _ptr, _stream
what = sys.argv[1]
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
D = cb.Derivative()
if what == "fedvr":
    F = cb.FEDVR(np.linspace(0, 1000, 10001), 10)
    A = F.T @ D.T @ D @ F
    X = torch.randn((1024, F.N), dtype=torch.float64, device="cuda"); Y = torch.empty_like(X)
    fn = lambda: cb.mul(Y, A, X)
elif what == "fedvr_asm":  # K4 alone: C2 element blocks (order 10, 1e4 elements)
    F = cb.FEDVR(np.linspace(0, 1000, 10001), 10)
    def fn():
        F._ops.clear()
        return F.T @ D.T @ D @ F
elif what == "eval":
    B = cb.BSpline(cb.LinearKnotSet(7, 0, 100, 1000))
    nx = 16_000_000
    x = torch.rand(nx, dtype=torch.float64, device="cuda") * 100
    iv = torch.empty(nx, dtype=torch.int32, device="cuda"); vals = torch.empty((nx, 7), dtype=torch.float64, device="cuda")
    fn = lambda: _lib.call("cb_bspline_eval", B.handle, _ptr(x), nx, 0, _ptr(iv), _ptr(vals), _stream())
elif what == "eval5":      # evaluation on the C5 basis: 5e4 intervals (knots do not fit the shared-memory table), k = 8
    B = cb.BSpline(cb.LinearKnotSet(8, 0, 500, 50000))
    nx = 16_000_000
    x = torch.rand(nx, dtype=torch.float64, device="cuda") * 500
    iv = torch.empty(nx, dtype=torch.int32, device="cuda"); vals = torch.empty((nx, 8), dtype=torch.float64, device="cuda")
    fn = lambda: _lib.call("cb_bspline_eval", B.handle, _ptr(x), nx, 0, _ptr(iv), _ptr(vals), _stream())
elif what == "eval3":      # evaluation on the C3 basis: 2e5 exponential intervals, k = 10
    B = cb.BSpline(cb.ExpKnotSet(10, -3.0, 2.0, 200000))
    nx = 16_000_000
    x = torch.rand(nx, dtype=torch.float64, device="cuda") * 100
    iv = torch.empty(nx, dtype=torch.int32, device="cuda"); vals = torch.empty((nx, 10), dtype=torch.float64, device="cuda")
    fn = lambda: _lib.call("cb_bspline_eval", B.handle, _ptr(x), nx, 0, _ptr(iv), _ptr(vals), _stream())
elif what == "csc1m":      # B[x,:] as SparseMatrixCSC, C1 size
    B = cb.BSpline(cb.LinearKnotSet(7, 0, 100, 1000))
    x = torch.rand(1_000_000, dtype=torch.float64, device="cuda") * 100
    fn = lambda: B.eval_csc(x)
elif what == "eval1m":     # C1 size: 1e6 points
    B = cb.BSpline(cb.LinearKnotSet(7, 0, 100, 1000))
    nx = 1_000_000
    x = torch.rand(nx, dtype=torch.float64, device="cuda") * 100
    iv = torch.empty(nx, dtype=torch.int32, device="cuda"); vals = torch.empty((nx, 7), dtype=torch.float64, device="cuda")
    fn = lambda: _lib.call("cb_bspline_eval", B.handle, _ptr(x), nx, 0, _ptr(iv), _ptr(vals), _stream())
elif what == "asm3":
    B = cb.BSpline(cb.ExpKnotSet(10, -3.0, 2.0, 200000))
    fn = lambda: (B.T @ cb.CoulombPotential(1.0) @ B, B.T @ D.T @ D @ B)
elif what == "asm3k":      # kernel-only loop through the C ABI (no Python operator glue)
    B = cb.BSpline(cb.ExpKnotSet(10, -3.0, 2.0, 200000))
    band = torch.empty((B.nf, 19), dtype=torch.float64, device="cuda")
    def fn():
        for _ in range(10):
            _lib.call("cb_bspline_assemble_coulomb", B.handle, B.handle, 1.0, 0, B.nf, 0, B.nf, 9, 9, _ptr(band), _stream())
            _lib.call("cb_bspline_assemble", B.handle, B.handle, 1, 1, None, 0, B.nf, 0, B.nf, 9, 9, _ptr(band), _stream())
elif what == "banded":
    B = cb.BSpline(cb.LinearKnotSet(7, 0, 100, 1000))
    S = B.T @ D.T @ D @ B
    X = torch.randn((65536, B.nf), dtype=torch.float64, device="cuda"); Y = torch.empty_like(X)
    fn = lambda: cb.mul(Y, S, X)
elif what == "banded3":    # C3 shape: k = 10 band, 200 009 rows, 256 vectors
    B = cb.BSpline(cb.ExpKnotSet(10, -3.0, 2.0, 200000))
    S = B.T @ D.T @ D @ B
    X = torch.randn((256, B.nf), dtype=torch.float64, device="cuda"); Y = torch.empty_like(X)
    fn = lambda: cb.mul(Y, S, X)
elif what == "density":
    B = cb.BSpline(cb.LinearKnotSet(8, 0, 500, 50000))
    P = 1024
    rho = cb.Density(B, nlhs=P, nrhs=P)
    f = torch.randn((P, B.nf), dtype=torch.float64, device="cuda"); g = torch.randn((P, B.nf), dtype=torch.float64, device="cuda")
    o = cb.ColMajor.zeros(B.nf, P)
    fn = lambda: rho.copyto(cb.ColMajor(f), cb.ColMajor(g), o)
elif what == "tridiag":
    R = cb.StaggeredFiniteDifferences(10_000_000, 1e-3)
    Lp = R.T @ D.T @ D @ R
    X = torch.randn((128, 10_000_000), dtype=torch.float64, device="cuda"); Y = torch.empty_like(X)
    fn = lambda: cb.mul(Y, Lp, X)
for _ in range(reps):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
NT = 1 if os.environ.get("PROF_SINGLE") else 10
e0.record()
for _ in range(NT):
    fn()
e1.record(); torch.cuda.synchronize()
per = e0.elapsed_time(e1) / NT / (20 if what == "asm3k" else 1)
print(what, "ms", per, "(mean of back-to-back calls)")
if os.environ.get("PROF_TIME"):
    ts = []
    for _ in range(20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / (20 if what == "asm3k" else 1))
    ts.sort()
    print(f"{what}: mean-b2b {per*1e3:.1f} us  min {ts[0]*1e3:.1f} us  med {ts[len(ts)//2]*1e3:.1f} us")
