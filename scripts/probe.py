"""Quick device-time probe of every kernel at the BASELINE config sizes (development aid)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import compactbases_jl_b200 as cb

def timeit(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))

D = cb.Derivative()
which = sys.argv[1:] or ["c1", "c2", "c3", "c4", "c5"]
if "c1" in which:
    B = cb.BSpline(cb.LinearKnotSet(7, 0, 100, 1000))
    for nx in (1_000_000, 16_000_000):
        x = torch.rand(nx, dtype=torch.float64, device="cuda") * 100
        iv = torch.empty(nx, dtype=torch.int32, device="cuda"); vals = torch.empty((nx, 7), dtype=torch.float64, device="cuda")
        from compactbases_jl_b200 import _lib
        from compactbases_jl_b200.operators import _ptr, _stream
        f = lambda: _lib.call("cb_bspline_eval", B.handle, _ptr(x), nx, 0, _ptr(iv), _ptr(vals), _stream())
        mn, md = timeit(f)
        print(f"C1 eval nx={nx}: min {mn*1e3:.1f} us  med {md*1e3:.1f} us  -> {nx*68/mn/1e6:.1f} GB/s, {nx/mn/1e6:.2f} Gpts/s")
    mn, md = timeit(lambda: (B.T @ B, B.T @ D.T @ D @ B))
    print(f"C1 assemble mass+lap: min {mn*1e3:.1f} us med {md*1e3:.1f} us")
if "c2" in which:
    t0 = time.time(); F = cb.FEDVR(np.linspace(0, 1000, 10001), 10); print("FEDVR host ctor s", time.time() - t0)
    def asm():
        F._ops.clear(); return F.T @ D.T @ D @ F
    mn, md = timeit(asm)
    print(f"C2 assemble: min {mn*1e3:.1f} us med {md*1e3:.1f} us")
    A = asm()
    X = torch.randn((1024, F.N), dtype=torch.float64, device="cuda"); Y = torch.empty_like(X)
    mn, md = timeit(lambda: cb.mul(Y, A, X))
    print(f"C2 apply: min {mn*1e3:.1f} us med {md*1e3:.1f} us -> {1.4826e9/mn/1e6:.1f} GB/s")
    del X, Y
if "c3" in which:
    t0 = time.time(); B = cb.BSpline(cb.ExpKnotSet(10, -3.0, 2.0, 200000)); print("C3 ctor s", time.time() - t0)
    mn, md = timeit(lambda: (B.T @ cb.CoulombPotential(1.0) @ B, B.T @ D.T @ D @ B))
    print(f"C3 assemble V+T: min {mn*1e3:.1f} us med {md*1e3:.1f} us -> {62.4e6/mn/1e6:.1f} GB/s; {1.6e9/mn/1e9:.2f} TFLOP/s")
if "c4" in which:
    N, nv = 10_000_000, 256
    R = cb.StaggeredFiniteDifferences(N, 1e-3)
    Lp = R.T @ D.T @ D @ R
    X = torch.randn((nv, N), dtype=torch.float64, device="cuda"); Y = torch.empty_like(X)
    mn, md = timeit(lambda: cb.mul(Y, Lp, X), n=5)
    print(f"C4 tridiag apply N=1e7 x {nv}: min {mn:.3f} ms med {md:.3f} ms -> {2*8*N*nv/mn/1e6:.1f} GB/s")
    del X, Y
if "c5" in which:
    B = cb.BSpline(cb.LinearKnotSet(8, 0, 500, 50000))
    t0 = time.time(); rho = cb.Density(B, nlhs=2, nrhs=2); torch.cuda.synchronize(); print("C5 plan s", time.time() - t0)
    for P in (1024, 4096):
        f = torch.randn((P, B.nf), dtype=torch.float64, device="cuda"); g = torch.randn((P, B.nf), dtype=torch.float64, device="cuda")
        out = cb.ColMajor.zeros(B.nf, P)
        mn, md = timeit(lambda: rho.copyto(cb.ColMajor(f), cb.ColMajor(g), out), n=5)
        print(f"C5 density P={P}: min {mn:.3f} ms med {md:.3f} ms -> {3*8*B.nf*P/mn/1e6:.1f} GB/s; {2.3e11*P/1e4/mn/1e9:.2f} TFLOP/s")
    B7 = cb.BSpline(cb.LinearKnotSet(7, 0, 100, 1000))
    S = B7.T @ D.T @ D @ B7
    X = torch.randn((65536, B7.nf), dtype=torch.float64, device="cuda"); Y = torch.empty_like(X)
    mn, md = timeit(lambda: cb.mul(Y, S, X))
    print(f"banded apply 1006 x 65536 (l=u=6): min {mn*1e3:.1f} us -> {2*8*B7.nf*65536/mn/1e6:.1f} GB/s")
